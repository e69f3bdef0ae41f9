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
#include "vx_bits.h"
#include "vx_internal.h"

namespace vx {

constexpr int kMorphZL = 16;  // planes marched by one thread

struct Plane {
    uint4 c;        // FULL: rows y-1,y,y+1 combined;  FACE: row y
    uint4 s;        // FACE only: rows y-1,y+1 combined
    unsigned lb, rb;  // edge bytes (x-1, x+16) of `c`, only meaningful on strip-edge lanes
};

// Dilation combines with OR, erosion with AND; a voxel outside the image is the identity of
// the combine (0 resp. 1), so it never dilates in and never erodes.
template <bool ERODE>
__device__ __forceinline__ unsigned comb(unsigned a, unsigned b) { return ERODE ? (a & b) : (a | b); }
template <bool ERODE>
__device__ __forceinline__ uint4 comb4(uint4 a, uint4 b) {
    return make_uint4(comb<ERODE>(a.x, b.x), comb<ERODE>(a.y, b.y), comb<ERODE>(a.z, b.z), comb<ERODE>(a.w, b.w));
}

// rows y-1, y, y+1 of one plane at this thread's 16 voxels; p = address of the centre row's
// chunk, nullptr when the plane does not exist.  Everything that does not depend on z (row
// validity, strip-edge flags, the row pitch) is computed once by the caller: the z loop is
// three 128-bit loads and a dozen logic instructions (the first version re-derived every
// address with 64-bit multiplies: 190 instructions per plane, issue-bound at 69 %).
template <bool FULL, bool ERODE>
__device__ __forceinline__ Plane load_plane(const uint8_t* __restrict__ p, int nx, bool up, bool dn, bool need_l,
                                            bool need_r) {
    constexpr unsigned ID = ERODE ? 0x01010101u : 0u;
    Plane r;
    r.c = r.s = make_uint4(ID, ID, ID, ID);
    r.lb = r.rb = ID & 1u;
    if (p == nullptr) return r;
    const uint4 wc = __ldg(reinterpret_cast<const uint4*>(p));
    uint4 wu = make_uint4(ID, ID, ID, ID), wd = wu;
    if (up) wu = __ldg(reinterpret_cast<const uint4*>(p - nx));
    if (dn) wd = __ldg(reinterpret_cast<const uint4*>(p + nx));
    if (FULL) {
        r.c = comb4<ERODE>(comb4<ERODE>(wc, wu), wd);
        if (need_l) {
            r.lb = __ldg(p - 1);
            if (up) r.lb = comb<ERODE>(r.lb, (unsigned)__ldg(p - nx - 1));
            if (dn) r.lb = comb<ERODE>(r.lb, (unsigned)__ldg(p + nx - 1));
        }
        if (need_r) {
            r.rb = __ldg(p + 16);
            if (up) r.rb = comb<ERODE>(r.rb, (unsigned)__ldg(p - nx + 16));
            if (dn) r.rb = comb<ERODE>(r.rb, (unsigned)__ldg(p + nx + 16));
        }
    } else {
        r.c = wc;
        r.s = comb4<ERODE>(wu, wd);
        if (need_l) r.lb = __ldg(p - 1);
        if (need_r) r.rb = __ldg(p + 16);
    }
    return r;
}

// grid: x = strips * ceil(ny/16), y = ceil(nz/ZL), z = nb;  block (16,16).
// lo / hi: optional halo planes [ny][nx] standing for z = -1 and z = nz (the boundary planes of
// the neighbouring slabs of a z-decomposed volume; they may live in a peer GPU's memory, read
// over NVLink).  nullptr = the volume really ends there.
template <bool FULL, bool ERODE>
__global__ void __launch_bounds__(256)
morph16_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int nx, int ny, int nz,
               const uint8_t* __restrict__ lo_all, const uint8_t* __restrict__ hi_all) {
    constexpr unsigned IDB = ERODE ? 1u : 0u;
    const int ncx = nx >> 4;
    const int nstrips = (ncx + 15) >> 4;
    const int strip = blockIdx.x % nstrips, yblk = blockIdx.x / nstrips;
    const int lane16 = threadIdx.x;
    const int xc = strip * 16 + lane16;
    const int y = yblk * 16 + threadIdx.y;
    const int z0 = blockIdx.y * kMorphZL;
    const size_t ps = (size_t)nx * ny;
    const size_t voff = (size_t)blockIdx.z * ps * nz;
    const bool col_ok = xc < ncx && y < ny;
    const bool need_l = col_ok && lane16 == 0 && xc > 0;           // left neighbour in another strip
    const bool need_r = col_ok && lane16 == 15 && xc + 1 < ncx;    // right neighbour in another strip
    const bool up = y > 0, dn = y + 1 < ny;
    const size_t roff = (size_t)y * nx + (size_t)xc * 16;          // this thread's chunk inside a plane
    const uint8_t* lo = (lo_all && col_ok) ? lo_all + (size_t)blockIdx.z * ps + roff : nullptr;
    const uint8_t* hi = (hi_all && col_ok) ? hi_all + (size_t)blockIdx.z * ps + roff : nullptr;
    const uint8_t* pz = in + voff + (size_t)z0 * ps + roff;        // plane z0
    uint8_t* oz = out + voff + (size_t)z0 * ps + roff;

    Plane prev = load_plane<FULL, ERODE>(!col_ok ? nullptr : (z0 > 0 ? pz - ps : lo), nx, up, dn, need_l, need_r);
    Plane cur = load_plane<FULL, ERODE>(col_ok ? pz : nullptr, nx, up, dn, need_l, need_r);
    const int zend = min(z0 + kMorphZL, nz);
    for (int z = z0; z < zend; z++) {
        pz += ps;
        Plane next = load_plane<FULL, ERODE>(!col_ok ? nullptr : (z + 1 < nz ? pz : hi), nx, up, dn, need_l, need_r);
        uint4 C, E;
        unsigned lb, rb;
        if (FULL) {
            C = comb4<ERODE>(comb4<ERODE>(prev.c, cur.c), next.c);
            E = C;   // unused
            lb = comb<ERODE>(comb<ERODE>(prev.lb, cur.lb), next.lb);
            rb = comb<ERODE>(comb<ERODE>(prev.rb, cur.rb), next.rb);
        } else {
            C = cur.c;
            E = comb4<ERODE>(comb4<ERODE>(prev.c, next.c), cur.s);
            lb = cur.lb;
            rb = cur.rb;
        }
        // bytes just outside this thread's 16 voxels, from the neighbouring lanes
        const unsigned from_left = __shfl_up_sync(0xffffffffu, C.w >> 24, 1, 16);
        const unsigned from_right = __shfl_down_sync(0xffffffffu, C.x & 0xffu, 1, 16);
        const unsigned L = lane16 == 0 ? lb : from_left;
        unsigned R = lane16 == 15 ? rb : from_right;
        if (xc + 1 >= ncx) R = IDB;  // last chunk of the row: nothing to the right
        uint4 r;
        r.x = comb<ERODE>(comb<ERODE>(C.x, (C.x << 8) | L), __funnelshift_r(C.x, C.y, 8));
        r.y = comb<ERODE>(comb<ERODE>(C.y, __funnelshift_l(C.x, C.y, 8)), __funnelshift_r(C.y, C.z, 8));
        r.z = comb<ERODE>(comb<ERODE>(C.z, __funnelshift_l(C.y, C.z, 8)), __funnelshift_r(C.z, C.w, 8));
        r.w = comb<ERODE>(comb<ERODE>(C.w, __funnelshift_l(C.z, C.w, 8)), (C.w >> 8) | (R << 24));
        if (!FULL) r = comb4<ERODE>(r, E);
        if (col_ok) *reinterpret_cast<uint4*>(oz) = r;
        oz += ps;
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

// ---- near / interior on the bit form -----------------------------------------------------------
// One thread per 32-voxel row word and chunk of kBitZL planes, marching along z.  A plane
// contributes  A = its three rows (y-1, y, y+1) spread by one voxel in x  and  C = what it gives to
// the planes above and below; for 26-adjacency C = A, for 6 only the centre row is spread,
// the rows y+-1 enter A unspread and C is the bare centre row.  out(z) = A(z) | C(z-1) | C(z+1):
// three row loads per output word instead of nine.  Erosion is the dilation of the complement
// inside the image (voxels outside do not exist), complemented.  Traffic: 1/8 byte in and out per
// voxel; the rows of a neighbourhood are L1/L2 hits.
constexpr int kBitZL = 8;

// word w of the row at stream position p: its bits c and the bits spread by one voxel in x
template <bool ERODE>
__device__ __forceinline__ void load_row_bits(const unsigned* __restrict__ fb, unsigned long long p, unsigned vmask,
                                              bool has_l, bool has_r, unsigned& c, unsigned& xs) {
    const unsigned long long k = p >> 5;
    const unsigned sh = (unsigned)p & 31u;
    const unsigned b = __ldg(fb + k);
    unsigned l = 0, r = 0;
    if (sh == 0) {
        c = b;
        if (has_l) l = __ldg(fb + k - 1) >> 31;
        if (has_r) r = __ldg(fb + k + 1) & 1u;
    } else {
        const unsigned hi = __ldg(fb + k + 1);
        c = __funnelshift_r(b, hi, sh);
        if (has_l) l = (b >> (sh - 1)) & 1u;
        if (has_r) r = (hi >> sh) & 1u;
    }
    c &= vmask;
    if (ERODE) {
        c = ~c & vmask;
        l = has_l ? l ^ 1u : 0u;
        r = has_r ? r ^ 1u : 0u;
    }
    xs = c | (c << 1) | (c >> 1) | l | (r << 31);
}

template <bool FULL, bool ERODE>
__global__ void __launch_bounds__(256)
morph_bits_kernel(const unsigned* __restrict__ in, unsigned* __restrict__ out, int nx, int ny, int nz, int nwx,
                  int nzc, unsigned nthreads, int align16) {
    unsigned t = blockIdx.x * 256u + threadIdx.x;
    if (t >= nthreads) return;
    const int w = (int)(t % (unsigned)nwx); t /= (unsigned)nwx;
    const int y = (int)(t % (unsigned)ny); t /= (unsigned)ny;
    const int zc = (int)(t % (unsigned)nzc);
    const unsigned vol = t / (unsigned)nzc;
    const int z0 = zc * kBitZL, z1 = min(nz, z0 + kBitZL);
    const int valid = min(32, nx - w * 32);
    const unsigned vmask = valid >= 32 ? 0xffffffffu : ((1u << valid) - 1u);
    const bool has_l = w > 0, has_r = (w + 1) * 32 < nx;
    const bool up = y > 0, dn = y + 1 < ny;
    const unsigned long long plane = (unsigned long long)nx * ny;
    // stream position of word w of row (y, z0) of this volume; a plane further is + nx * ny
    unsigned long long pz = ((unsigned long long)vol * nz + z0) * plane + (unsigned long long)y * nx + (unsigned)w * 32u;
    auto load_plane = [&](bool exists, unsigned long long p, unsigned& A, unsigned& C) {
        A = C = 0u;
        if (!exists) return;
        unsigned c, xs;
        load_row_bits<ERODE>(in, p, vmask, has_l, has_r, c, xs);
        A = xs;
        C = FULL ? xs : c;
        if (up) {
            load_row_bits<ERODE>(in, p - nx, vmask, has_l, has_r, c, xs);
            A |= FULL ? xs : c;
        }
        if (dn) {
            load_row_bits<ERODE>(in, p + nx, vmask, has_l, has_r, c, xs);
            A |= FULL ? xs : c;
        }
        if (FULL) C = A;
    };
    unsigned Am, Cm, A0, C0, A1, C1;
    load_plane(z0 > 0, pz - plane, Am, Cm);
    load_plane(true, pz, A0, C0);
    // (fully unrolled: the row loads of the planes ahead are issued before the stores of the planes behind)
#pragma unroll
    for (int i = 0; i < kBitZL; i++) {
        const int z = z0 + i;
        if (z >= z1) break;
        load_plane(z + 1 < nz, pz + plane, A1, C1);
        unsigned acc = A0 | Cm | C1;
        if (ERODE) acc = ~acc;
        flat_store32(out, pz, valid, acc & vmask, align16 != 0);
        Cm = C0; A0 = A1; C0 = C1;
        pz += plane;
    }
}

// faces of every volume, as bits
__global__ void __launch_bounds__(256)
border_bits_kernel(unsigned* __restrict__ out, int nx, int ny, int nz, int nwx, unsigned nwords, int align16) {
    const unsigned wid = blockIdx.x * 256u + threadIdx.x;
    if (wid >= nwords) return;
    const unsigned row = wid / (unsigned)nwx;
    const int w = (int)(wid - row * (unsigned)nwx);
    const unsigned pz = row / (unsigned)ny;
    const int y = (int)(row - pz * (unsigned)ny);
    const int z = (int)(pz % (unsigned)nz);
    const int valid = min(32, nx - w * 32);
    const unsigned vmask = valid >= 32 ? 0xffffffffu : ((1u << valid) - 1u);
    // an axis of extent 1 has no faces (a 2D image is not all border)
    const bool face = (ny > 1 && (y == 0 || y == ny - 1)) || (nz > 1 && (z == 0 || z == nz - 1));
    unsigned v = face ? vmask : 0u;
    if (nx > 1 && w == 0) v |= 1u;
    if (nx > 1 && (w + 1) * 32 >= nx) v |= 1u << (valid - 1);
    flat_store32(out, (unsigned long long)row * nx + (unsigned)w * 32u, valid, v, align16 != 0);
}

static int morph_bits(vx_ctx ctx, vx_img a, int conn, bool erode, vx_img* out) {
    OpGuard g;
    VX_TRY(g.begin(ctx));
    Image* A = nullptr;
    const unsigned* Ab = nullptr;
    VX_TRY(g.input_bits(a, &A, &Ab));
    Image* O = nullptr;
    VX_TRY(new_bits_image(g.c, A->nx, A->ny, A->nz, A->nb, A->sp, &O));
    const int align16 = (A->nx % 16) == 0;
    if (!align16) {
        cudaError_t ce = cudaMemsetAsync(O->bits, 0, O->bit_words() * 4, g.st);
        if (ce != cudaSuccess) { drop(O); return fail(VX_E_CUDA, "memset: %s", cudaGetErrorString(ce)); }
    }
    const int nwx = (A->nx + 31) / 32;
    const int nzc = (A->nz + kBitZL - 1) / kBitZL;
    const unsigned long long nth = (unsigned long long)A->nb * nzc * A->ny * nwx;
    if (nth >= 0xffffffffull) { drop(O); return fail(VX_E_UNSUPPORTED, "batch too large for one launch"); }
    const unsigned grid = (unsigned)((nth + 255) / 256);
    const bool full = conn == VX_CONN_FULL;
    {
        VX_LAUNCH(g.c, g.s, "morph_bits_kernel");
        if (full && !erode) morph_bits_kernel<true, false><<<grid, 256, 0, g.st>>>(Ab, O->bits, A->nx, A->ny, A->nz, nwx, nzc, (unsigned)nth, align16);
        else if (full) morph_bits_kernel<true, true><<<grid, 256, 0, g.st>>>(Ab, O->bits, A->nx, A->ny, A->nz, nwx, nzc, (unsigned)nth, align16);
        else if (!erode) morph_bits_kernel<false, false><<<grid, 256, 0, g.st>>>(Ab, O->bits, A->nx, A->ny, A->nz, nwx, nzc, (unsigned)nth, align16);
        else morph_bits_kernel<false, true><<<grid, 256, 0, g.st>>>(Ab, O->bits, A->nx, A->ny, A->nz, nwx, nzc, (unsigned)nth, align16);
    }
    int r = check_launch("morph_bits_kernel");
    if (r != VX_OK) { drop(O); return r; }
    return g.finish(O, out);
}

static int morph(vx_ctx ctx, vx_img a, int conn, bool erode, const void* lo_plane, const void* hi_plane,
                 vx_img* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    VX_REQUIRE(conn == VX_CONN_FACE || conn == VX_CONN_FULL, VX_E_INVALID_ARG,
               "adjacency must be 6 or 26, got %d", conn);
    // without halo planes the operator runs on the bit form; the halo variants read the neighbours'
    // boundary planes as bytes from peer memory and keep the byte kernel
    if (lo_plane == nullptr && hi_plane == nullptr) return morph_bits(ctx, a, conn, erode, out);
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
    VX_TRY(g.input_meta(like, &A));   // only the shape is used
    Image* O = nullptr;
    VX_TRY(new_bits_image(g.c, A->nx, A->ny, A->nz, A->nb, A->sp, &O));
    const int align16 = (A->nx % 16) == 0;
    if (!align16) {
        cudaError_t ce = cudaMemsetAsync(O->bits, 0, O->bit_words() * 4, g.st);
        if (ce != cudaSuccess) { drop(O); return fail(VX_E_CUDA, "memset: %s", cudaGetErrorString(ce)); }
    }
    const int nwx = (A->nx + 31) / 32;
    const unsigned long long nwords = (unsigned long long)A->nb * A->nz * A->ny * nwx;
    if (nwords >= 0xffffffffull) { drop(O); return fail(VX_E_UNSUPPORTED, "batch too large for one launch"); }
    {
        VX_LAUNCH(g.c, g.s, "border_kernel");
        border_bits_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, g.st>>>(O->bits, A->nx, A->ny, A->nz, nwx, (unsigned)nwords, align16);
    }
    int r = check_launch("border_kernel");
    if (r != VX_OK) { drop(O); return r; }
    return g.finish(O, out);
}

}  // extern "C"
