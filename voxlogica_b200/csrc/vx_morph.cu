// A4 near / interior and A10 border.
//   near(A)     = A u { x | some a in A is adjacent to x }   closure, Greta.pdf p.24, p.30
//   interior(A) = !near(!A)                                   Greta.pdf p.35
// on a finite grid: voxels outside the image do not exist, so they never dilate in
// and never erode.  Adjacency is 6 (faces) or 26 (faces+edges+corners);
// a 2D image (nz == 1) gets 4 / 8 automatically.
//
// Kernel: each thread owns 16 consecutive voxels in x (one uint4 of 0/1 bytes) and
// marches along z keeping the three most recent planes in registers, so every plane
// costs three row loads (y-1, y, y+1; two of them L1 hits from the neighbouring
// thread rows) instead of nine.  The x-neighbourhood is a byte shift inside the
// 128-bit register, with the one missing byte on each side fetched from the
// neighbouring lane by a width-16 shuffle.  Erosion is the same kernel on the
// complemented input.  HBM: 1 byte read + 1 byte written per voxel.
#include "vx_internal.h"

namespace vx {

constexpr int kMorphZL = 16;  // planes marched by one thread

struct Plane {
    uint4 c;        // FULL: OR of rows y-1,y,y+1;  FACE: row y
    uint4 s;        // FACE only: OR of rows y-1,y+1
    unsigned lb, rb;  // edge bytes (x-1, x+16) of `c`, only meaningful on strip-edge lanes
};

__device__ __forceinline__ uint4 or4(uint4 a, uint4 b) {
    return make_uint4(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w);
}

// lo / hi: optional halo planes [ny][nx] standing for z = -1 and z = nz (the boundary planes of
// the neighbouring slabs of a z-decomposed volume; they may live in a peer GPU's memory, read
// over NVLink).  nullptr = the volume really ends there.
template <bool FULL, bool ERODE>
__device__ __forceinline__ Plane load_plane(const uint8_t* __restrict__ vol, int nx, int ny, int nz,
                                            int xc, int y, int z, bool col_ok, bool need_l,
                                            bool need_r, const uint8_t* __restrict__ lo,
                                            const uint8_t* __restrict__ hi) {
    Plane p;
    p.c = make_uint4(0, 0, 0, 0);
    p.s = make_uint4(0, 0, 0, 0);
    p.lb = p.rb = 0;
    if (!col_ok) return p;
    const uint8_t* plane;
    if (z >= 0 && z < nz) plane = vol + (size_t)z * ny * nx;
    else if (z == -1 && lo != nullptr) plane = lo;
    else if (z == nz && hi != nullptr) plane = hi;
    else return p;
    const unsigned flip = ERODE ? 0x01010101u : 0u;
#pragma unroll
    for (int dy = -1; dy <= 1; dy++) {
        int yy = y + dy;
        if (yy < 0 || yy >= ny) continue;
        const uint8_t* row = plane + (size_t)yy * nx + (size_t)xc * 16;
        uint4 w = __ldg(reinterpret_cast<const uint4*>(row));
        w.x ^= flip; w.y ^= flip; w.z ^= flip; w.w ^= flip;
        if (FULL || dy == 0) {
            p.c = or4(p.c, w);
            if (need_l) p.lb |= (unsigned)(__ldg(row - 1) ^ (ERODE ? 1 : 0));
            if (need_r) p.rb |= (unsigned)(__ldg(row + 16) ^ (ERODE ? 1 : 0));
        } else {
            p.s = or4(p.s, w);
        }
    }
    return p;
}

// grid: x = strips * ceil(ny/16), y = ceil(nz/ZL), z = nb;  block (16,16)
template <bool FULL, bool ERODE>
__global__ void __launch_bounds__(256)
morph16_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int nx, int ny, int nz,
               const uint8_t* __restrict__ lo_all, const uint8_t* __restrict__ hi_all) {
    const int ncx = nx >> 4;
    const int nstrips = (ncx + 15) >> 4;
    const int strip = blockIdx.x % nstrips, yblk = blockIdx.x / nstrips;
    const int lane16 = threadIdx.x;
    const int xc = strip * 16 + lane16;
    const int y = yblk * 16 + threadIdx.y;
    const int z0 = blockIdx.y * kMorphZL;
    const size_t voff = (size_t)blockIdx.z * nx * ny * nz;
    const uint8_t* vol = in + voff;
    uint8_t* ovol = out + voff;
    const bool col_ok = xc < ncx && y < ny;
    const bool need_l = col_ok && lane16 == 0 && xc > 0;           // left neighbour in another strip
    const bool need_r = col_ok && lane16 == 15 && xc + 1 < ncx;    // right neighbour in another strip
    const uint8_t* lo = lo_all ? lo_all + (size_t)blockIdx.z * nx * ny : nullptr;
    const uint8_t* hi = hi_all ? hi_all + (size_t)blockIdx.z * nx * ny : nullptr;

    Plane prev = load_plane<FULL, ERODE>(vol, nx, ny, nz, xc, y, z0 - 1, col_ok, need_l, need_r, lo, hi);
    Plane cur = load_plane<FULL, ERODE>(vol, nx, ny, nz, xc, y, z0, col_ok, need_l, need_r, lo, hi);
    const int zend = min(z0 + kMorphZL, nz);
    for (int z = z0; z < zend; z++) {
        Plane next = load_plane<FULL, ERODE>(vol, nx, ny, nz, xc, y, z + 1, col_ok, need_l, need_r, lo, hi);
        uint4 C, E;
        unsigned lb, rb;
        if (FULL) {
            C = or4(or4(prev.c, cur.c), next.c);
            E = make_uint4(0, 0, 0, 0);
            lb = prev.lb | cur.lb | next.lb;
            rb = prev.rb | cur.rb | next.rb;
        } else {
            C = cur.c;
            E = or4(or4(prev.c, next.c), cur.s);
            lb = cur.lb;
            rb = cur.rb;
        }
        // bytes just outside this thread's 16 voxels, from the neighbouring lanes
        unsigned from_left = __shfl_up_sync(0xffffffffu, C.w >> 24, 1, 16);
        unsigned from_right = __shfl_down_sync(0xffffffffu, C.x & 0xffu, 1, 16);
        unsigned L = lane16 == 0 ? lb : from_left;
        unsigned R = lane16 == 15 ? rb : from_right;
        if (xc + 1 >= ncx) R = 0;  // last chunk of the row: nothing to the right
        uint4 r;
        r.x = C.x | (C.x << 8) | L | (C.x >> 8) | (C.y << 24);
        r.y = C.y | (C.y << 8) | (C.x >> 24) | (C.y >> 8) | (C.z << 24);
        r.z = C.z | (C.z << 8) | (C.y >> 24) | (C.z >> 8) | (C.w << 24);
        r.w = C.w | (C.w << 8) | (C.z >> 24) | (C.w >> 8) | (R << 24);
        r = or4(r, E);
        if (ERODE) { r.x ^= 0x01010101u; r.y ^= 0x01010101u; r.z ^= 0x01010101u; r.w ^= 0x01010101u; }
        if (col_ok)
            *reinterpret_cast<uint4*>(ovol + ((size_t)z * ny + y) * nx + (size_t)xc * 16) = r;
        prev = cur;
        cur = next;
    }
}

// any nx: one thread per voxel
template <bool FULL, bool ERODE>
__global__ void __launch_bounds__(256)
morph_generic_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int nx, int ny,
                     int nz, size_t total, const uint8_t* __restrict__ lo_all,
                     const uint8_t* __restrict__ hi_all) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int x = (int)(i % nx);
    size_t t = i / nx;
    int y = (int)(t % ny);
    t /= ny;
    int z = (int)(t % nz);
    const size_t v_idx = t / nz;
    const uint8_t* vol = in + v_idx * (size_t)nx * ny * nz;
    unsigned acc = 0;
    for (int dz = -1; dz <= 1; dz++) {
        int zz = z + dz;
        const uint8_t* plane;
        if (zz >= 0 && zz < nz) plane = vol + (size_t)zz * ny * nx;
        else if (zz == -1 && lo_all != nullptr) plane = lo_all + v_idx * (size_t)nx * ny;
        else if (zz == nz && hi_all != nullptr) plane = hi_all + v_idx * (size_t)nx * ny;
        else continue;
        for (int dy = -1; dy <= 1; dy++) {
            int yy = y + dy;
            if (yy < 0 || yy >= ny) continue;
            for (int dx = -1; dx <= 1; dx++) {
                int xx = x + dx;
                if (xx < 0 || xx >= nx) continue;
                if (!FULL && (dx != 0) + (dy != 0) + (dz != 0) > 1) continue;
                unsigned v = plane[(size_t)yy * nx + xx] != 0;
                acc |= ERODE ? (v ^ 1u) : v;
            }
        }
    }
    out[i] = (uint8_t)(ERODE ? (acc ^ 1u) : acc);
}

__global__ void __launch_bounds__(256)
border_kernel(uint8_t* __restrict__ out, int nx, int ny, int nz, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int x = (int)(i % nx);
    size_t t = i / nx;
    int y = (int)(t % ny);
    int z = (int)((t / ny) % nz);
    // an axis of extent 1 has no faces (a 2D image is not all border)
    bool b = (nx > 1 && (x == 0 || x == nx - 1)) || (ny > 1 && (y == 0 || y == ny - 1)) ||
             (nz > 1 && (z == 0 || z == nz - 1));
    out[i] = b ? 1 : 0;
}

// nx % 16 == 0: one thread per 16 voxels, one 128-bit store
__global__ void __launch_bounds__(256)
border16_kernel(uint4* __restrict__ out, int ncx, int nx, int ny, int nz, size_t nchunks) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nchunks) return;
    const int xc = (int)(i % ncx);
    size_t t = i / ncx;
    const int y = (int)(t % ny);
    const int z = (int)((t / ny) % nz);
    const bool face = (ny > 1 && (y == 0 || y == ny - 1)) || (nz > 1 && (z == 0 || z == nz - 1));
    unsigned w = face ? 0x01010101u : 0u;
    uint4 v = make_uint4(w, w, w, w);
    if (nx > 1 && xc == 0) v.x |= 1u;
    if (nx > 1 && xc == ncx - 1) v.w |= 0x01000000u;
    out[i] = v;
}

static int morph(vx_ctx ctx, vx_img a, int conn, bool erode, const void* lo_plane, const void* hi_plane,
                 vx_img* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    VX_REQUIRE(conn == VX_CONN_FACE || conn == VX_CONN_FULL, VX_E_INVALID_ARG,
               "adjacency must be 6 or 26, got %d", conn);
    OpGuard g;
    VX_TRY(g.begin(ctx));
    Image* A = nullptr;
    VX_TRY(g.input(a, &A, VX_U8));
    Image* O = nullptr;
    VX_TRY(new_image(g.c, VX_U8, A->nx, A->ny, A->nz, A->nb, A->sp, &O));
    const uint8_t* in = (const uint8_t*)A->d;
    uint8_t* o = (uint8_t*)O->d;
    const uint8_t* lo = (const uint8_t*)lo_plane;
    const uint8_t* hi = (const uint8_t*)hi_plane;
    const bool full = conn == VX_CONN_FULL;
    const char* name;
    if ((A->nx & 15) == 0) {
        name = "morph16_kernel";
        int ncx = A->nx >> 4, nstrips = (ncx + 15) >> 4;
        dim3 grid(nstrips * ((A->ny + 15) / 16), (A->nz + kMorphZL - 1) / kMorphZL, A->nb);
        dim3 block(16, 16);
        VX_LAUNCH(g.c, g.s, name);
        if (full && !erode) morph16_kernel<true, false><<<grid, block, 0, g.st>>>(in, o, A->nx, A->ny, A->nz, lo, hi);
        else if (full) morph16_kernel<true, true><<<grid, block, 0, g.st>>>(in, o, A->nx, A->ny, A->nz, lo, hi);
        else if (!erode) morph16_kernel<false, false><<<grid, block, 0, g.st>>>(in, o, A->nx, A->ny, A->nz, lo, hi);
        else morph16_kernel<false, true><<<grid, block, 0, g.st>>>(in, o, A->nx, A->ny, A->nz, lo, hi);
    } else {
        name = "morph_generic_kernel";
        size_t total = A->count();
        unsigned grid = (unsigned)((total + 255) / 256);
        VX_LAUNCH(g.c, g.s, name);
        if (full && !erode) morph_generic_kernel<true, false><<<grid, 256, 0, g.st>>>(in, o, A->nx, A->ny, A->nz, total, lo, hi);
        else if (full) morph_generic_kernel<true, true><<<grid, 256, 0, g.st>>>(in, o, A->nx, A->ny, A->nz, total, lo, hi);
        else if (!erode) morph_generic_kernel<false, false><<<grid, 256, 0, g.st>>>(in, o, A->nx, A->ny, A->nz, total, lo, hi);
        else morph_generic_kernel<false, true><<<grid, 256, 0, g.st>>>(in, o, A->nx, A->ny, A->nz, total, lo, hi);
    }
    int r = check_launch(name);
    if (r != VX_OK) { drop(O); return r; }
    return g.finish(O, out);
}

}  // namespace vx

using namespace vx;

extern "C" {

int32_t vx_near(vx_ctx ctx, vx_img a, int32_t conn, vx_img* out) { return morph(ctx, a, conn, false, nullptr, nullptr, out); }
int32_t vx_interior(vx_ctx ctx, vx_img a, int32_t conn, vx_img* out) { return morph(ctx, a, conn, true, nullptr, nullptr, out); }
// near / interior of one z-slab of a larger volume, fused with the halo: the planes at z = -1
// and z = nz are read straight from `lo_plane` / `hi_plane` (nb * ny * nx bytes each, NULL at
// the real end of the volume), which may be peer-GPU memory mapped with vx_ipc_open -- the
// kernel then pulls the neighbour's boundary plane over NVLink while it computes, and no
// extended copy of the slab is ever built.  The caller orders the planes' producers before
// this call (they belong to another stream or process).
int32_t vx_near_halo(vx_ctx ctx, vx_img a, int32_t conn, const void* lo_plane, const void* hi_plane, vx_img* out) {
    return morph(ctx, a, conn, false, lo_plane, hi_plane, out);
}
int32_t vx_interior_halo(vx_ctx ctx, vx_img a, int32_t conn, const void* lo_plane, const void* hi_plane, vx_img* out) {
    return morph(ctx, a, conn, true, lo_plane, hi_plane, out);
}

int32_t vx_border(vx_ctx ctx, vx_img like, vx_img* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    OpGuard g;
    VX_TRY(g.begin(ctx));
    Image* A = nullptr;
    VX_TRY(g.input(like, &A, 0));
    Image* O = nullptr;
    VX_TRY(new_image(g.c, VX_U8, A->nx, A->ny, A->nz, A->nb, A->sp, &O));
    size_t total = A->count();
    if ((A->nx & 15) == 0) {
        size_t nchunks = total / 16;
        VX_LAUNCH(g.c, g.s, "border_kernel");
        border16_kernel<<<(unsigned)((nchunks + 255) / 256), 256, 0, g.st>>>((uint4*)O->d, A->nx / 16, A->nx, A->ny, A->nz, nchunks);
    } else {
        VX_LAUNCH(g.c, g.s, "border_kernel");
        border_kernel<<<(unsigned)((total + 255) / 256), 256, 0, g.st>>>((uint8_t*)O->d, A->nx, A->ny, A->nz, total);
    }
    int r = check_launch("border_kernel");
    if (r != VX_OK) { drop(O); return r; }
    return g.finish(O, out);
}

}  // extern "C"
