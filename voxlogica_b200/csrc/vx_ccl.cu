// A3 through / reachability and A8 maxvol via union-find connected-component
// labelling (Greta.pdf p.32-38 semantics, p.78 "Reachability via
// Connected Components").
//
//   through(a, b) = union of the connected components of b that contain a voxel of a.
//
// Algorithm (B200-first; not the OpenCL branch's ~29 relabelling sweeps, p.78):
//   1. init    one warp per 32-voxel row segment: ballot the foreground into a bit
//              mask and label every voxel with the start of its x-run inside the
//              segment, so whole runs are already merged before any atomic is issued;
//   2. merge   per voxel, link to the backward neighbour rows ((y-1,z),(y-1,z-1),
//              (y,z-1),(y+1,z-1) for 26-adjacency; (y-1,z),(y,z-1) for 6) using the bit
//              masks; a voxel only issues a union where a NEW neighbour run begins,
//              i.e. O(#run contacts) atomicMin unions, not O(#voxels x 13);
//   3. select  roots are the minimum linear index of their component (unions always
//              hang the larger root under the smaller), so labels are canonical.
//              through: seeds OR a flag bit into their root, then every voxel of b
//              looks its root up.  maxvol: run lengths are accumulated on roots.
// Labels are u32 linear indices inside their own volume (< 2^31), bit 31 is the flag.
#include "vx_internal.h"

namespace vx {

constexpr unsigned kFlag = 0x80000000u;
constexpr unsigned kIdx = 0x7fffffffu;
constexpr int kCclThreads = 256;

struct Geo {
    int nx, ny, nz, nb, nwx;  // nwx = 32-voxel words per row
    unsigned long long rows;  // nb*nz*ny
    unsigned long long nvox;
};

__device__ __forceinline__ unsigned find_root(const unsigned* L, unsigned x) {
    unsigned p = __ldcg(L + x) & kIdx;
    while (p != x) {
        x = p;
        p = __ldcg(L + x) & kIdx;
    }
    return x;
}

__device__ __forceinline__ void unite(unsigned* L, unsigned a, unsigned b) {
    bool done = false;
    do {
        a = find_root(L, a);
        b = find_root(L, b);
        if (a < b) {
            unsigned old = atomicMin(L + b, a);
            done = (old == b);
            b = old;
        } else if (b < a) {
            unsigned old = atomicMin(L + a, b);
            done = (old == a);
            a = old;
        } else {
            done = true;
        }
    } while (!done);
}

// decode the (row, word) handled by this warp
struct WarpPos {
    bool ok;
    int x, y, z, w;
    unsigned long long row;  // global row index (over the batch)
    unsigned vol;            // volume index
    unsigned lin;            // linear index of voxel x inside its volume
    int lane;
};
__device__ __forceinline__ WarpPos warp_pos(const Geo& g) {
    WarpPos p;
    unsigned long long wid = ((unsigned long long)blockIdx.x * kCclThreads + threadIdx.x) >> 5;
    p.lane = threadIdx.x & 31;
    p.ok = wid < g.rows * g.nwx;
    if (!p.ok) wid = 0;
    p.row = wid / g.nwx;
    p.w = (int)(wid % g.nwx);
    p.y = (int)(p.row % g.ny);
    unsigned long long t = p.row / g.ny;
    p.z = (int)(t % g.nz);
    p.vol = (unsigned)(t / g.nz);
    p.x = p.w * 32 + p.lane;
    p.lin = (unsigned)(((unsigned long long)p.z * g.ny + p.y) * g.nx + p.x);
    return p;
}

__global__ void __launch_bounds__(kCclThreads)
ccl_init_kernel(const uint8_t* __restrict__ img, unsigned* __restrict__ L,
                unsigned* __restrict__ bits, Geo g) {
    WarpPos p = warp_pos(g);
    if (!p.ok) return;  // whole warp
    const size_t gi = (size_t)p.vol * g.nvox + p.lin;
    bool fg = p.x < g.nx && img[gi] != 0;
    unsigned m = __ballot_sync(0xffffffffu, fg);
    if (p.lane == 0) bits[p.row * g.nwx + p.w] = m;
    if (fg) {
        unsigned zeros_below = ~m & ((1u << p.lane) - 1u);
        int start = zeros_below ? 32 - __clz(zeros_below) : 0;
        L[gi] = p.lin - (unsigned)(p.lane - start);
    }
}

template <bool FULL>
__global__ void __launch_bounds__(kCclThreads)
ccl_merge_kernel(unsigned* __restrict__ Lall, const unsigned* __restrict__ bits, Geo g) {
    WarpPos p = warp_pos(g);
    if (!p.ok) return;
    const unsigned* brow = bits + p.row * g.nwx;
    const unsigned m = __ldg(brow + p.w);
    if (m == 0) return;  // warp-uniform
    const bool fg = (m >> p.lane) & 1u;
    unsigned* L = Lall + (size_t)p.vol * g.nvox;
    const bool left_in_word = p.lane > 0 && ((m >> (p.lane - 1)) & 1u);
    // run continues from the previous 32-voxel segment
    if (p.lane == 0 && fg && p.w > 0 && (__ldg(brow + p.w - 1) >> 31)) unite(L, p.lin, p.lin - 1);

    // backward neighbour rows
    const int ndir = FULL ? 4 : 2;
    const int dys[4] = {-1, 0, -1, 1};
    const int dzs[4] = {0, -1, -1, -1};
#pragma unroll
    for (int d = 0; d < ndir; d++) {
        const int yy = p.y + dys[d], zz = p.z + dzs[d];
        if (yy < 0 || yy >= g.ny || zz < 0) continue;  // warp-uniform
        const long long drow = (long long)dzs[d] * g.ny + dys[d];
        const unsigned* nrow = brow + drow * g.nwx;
        const unsigned c = __ldg(nrow + p.w);
        unsigned long long W = (unsigned long long)c << 1;
        if (FULL) {
            if (p.w > 0) W |= (unsigned long long)(__ldg(nrow + p.w - 1) >> 31);
            if (p.w + 1 < g.nwx) W |= (unsigned long long)(__ldg(nrow + p.w + 1) & 1u) << 33;
        }
        if (W == 0 || !fg) continue;
        const unsigned t = (unsigned)(W >> p.lane) & 7u;  // bit0 n(x-1), bit1 n(x), bit2 n(x+1)
        const unsigned nlin = (unsigned)((long long)p.lin + drow * g.nx);  // neighbour at same x
        if (FULL) {
            if ((t & 4u) && !(t & 2u)) unite(L, p.lin, nlin + 1);  // a new neighbour run starts at x+1
            if (!left_in_word) {
                if (t & 2u) unite(L, p.lin, nlin);
                else if (t & 1u) unite(L, p.lin, nlin - 1);
            }
        } else {
            if ((t & 2u) && !(left_in_word && (t & 1u))) unite(L, p.lin, nlin);
        }
    }
}

// seeds: voxels of a & b flag the root of their component
__global__ void __launch_bounds__(kCclThreads)
ccl_seed_kernel(const uint8_t* __restrict__ seed, unsigned* __restrict__ Lall,
                const unsigned* __restrict__ bits, Geo g) {
    WarpPos p = warp_pos(g);
    if (!p.ok) return;
    const unsigned m = __ldg(bits + p.row * g.nwx + p.w);
    if (m == 0) return;
    if (!((m >> p.lane) & 1u)) return;
    const size_t gi = (size_t)p.vol * g.nvox + p.lin;
    if (seed[gi] == 0) return;
    unsigned* L = Lall + (size_t)p.vol * g.nvox;
    unsigned r = find_root(L, p.lin);
    if (!(__ldcg(L + r) & kFlag)) atomicOr(L + r, kFlag);
}

enum SelMode { SEL_THROUGH = 0, SEL_LABELS = 1, SEL_MAXVOL = 2 };

// final pass: resolve every voxel's root and emit the requested image
template <int MODE>
__global__ void __launch_bounds__(kCclThreads)
ccl_select_kernel(const unsigned* __restrict__ Lall, const unsigned* __restrict__ bits, void* out_,
                  const unsigned* __restrict__ cnt, const unsigned* __restrict__ vmax, Geo g) {
    WarpPos p = warp_pos(g);
    if (!p.ok || p.x >= g.nx) return;
    const unsigned m = __ldg(bits + p.row * g.nwx + p.w);
    const bool fg = (m >> p.lane) & 1u;
    const size_t gi = (size_t)p.vol * g.nvox + p.lin;
    const unsigned* L = Lall + (size_t)p.vol * g.nvox;
    unsigned r = 0;
    if (fg) r = find_root(L, p.lin);
    if (MODE == SEL_THROUGH) {
        reinterpret_cast<uint8_t*>(out_)[gi] = (fg && (L[r] & kFlag)) ? 1 : 0;
    } else if (MODE == SEL_LABELS) {
        reinterpret_cast<unsigned*>(out_)[gi] = fg ? r + 1u : 0u;
    } else {
        reinterpret_cast<uint8_t*>(out_)[gi] =
            (fg && cnt[(size_t)p.vol * g.nvox + r] == vmax[p.vol]) ? 1 : 0;
    }
}

// maxvol step 1: zero the counters of the roots
__global__ void __launch_bounds__(kCclThreads)
ccl_zero_roots_kernel(const unsigned* __restrict__ Lall, const unsigned* __restrict__ bits,
                      unsigned* __restrict__ cnt, Geo g) {
    WarpPos p = warp_pos(g);
    if (!p.ok) return;
    const unsigned m = __ldg(bits + p.row * g.nwx + p.w);
    if (!((m >> p.lane) & 1u)) return;
    const size_t gi = (size_t)p.vol * g.nvox + p.lin;
    if ((Lall[gi] & kIdx) == p.lin) cnt[gi] = 0;
}
// maxvol step 2: one atomicAdd per x-run segment (all its voxels share a root)
__global__ void __launch_bounds__(kCclThreads)
ccl_count_kernel(const unsigned* __restrict__ Lall, const unsigned* __restrict__ bits,
                 unsigned* __restrict__ cnt, Geo g) {
    WarpPos p = warp_pos(g);
    if (!p.ok) return;
    const unsigned m = __ldg(bits + p.row * g.nwx + p.w);
    if (!((m >> p.lane) & 1u)) return;
    const bool run_start = p.lane == 0 || !((m >> (p.lane - 1)) & 1u);
    if (!run_start) return;
    // length of the run starting at this lane, inside the 32-voxel segment
    unsigned above = ~(m >> p.lane);  // first zero above
    int len = above ? __ffs(above) - 1 : 32 - p.lane;
    if (len > 32 - p.lane) len = 32 - p.lane;
    const unsigned* L = Lall + (size_t)p.vol * g.nvox;
    unsigned r = find_root(L, p.lin);
    atomicAdd(cnt + (size_t)p.vol * g.nvox + r, (unsigned)len);
}
// maxvol step 3: per-volume maximum over roots
__global__ void __launch_bounds__(kCclThreads)
ccl_max_kernel(const unsigned* __restrict__ Lall, const unsigned* __restrict__ bits,
               const unsigned* __restrict__ cnt, unsigned* __restrict__ vmax, Geo g) {
    WarpPos p = warp_pos(g);
    if (!p.ok) return;
    const unsigned m = __ldg(bits + p.row * g.nwx + p.w);
    unsigned v = 0;
    if ((m >> p.lane) & 1u) {
        const size_t gi = (size_t)p.vol * g.nvox + p.lin;
        if ((Lall[gi] & kIdx) == p.lin) v = cnt[gi];
    }
    v = __reduce_max_sync(0xffffffffu, v);
    if (p.lane == 0 && v) atomicMax(vmax + p.vol, v);
}

static Geo make_geo(const Image* b) {
    Geo g;
    g.nx = b->nx; g.ny = b->ny; g.nz = b->nz; g.nb = b->nb;
    g.nwx = (b->nx + 31) / 32;
    g.rows = (unsigned long long)b->nb * b->nz * b->ny;
    g.nvox = b->nvox();
    return g;
}

// shared front end: labels + bit masks of image B on the op's stream
static int ccl_label(OpGuard& og, const Image* B, bool full, unsigned** L_out, unsigned** bits_out,
                     Geo* geo, unsigned* grid_out) {
    Geo g = make_geo(B);
    unsigned long long warps = g.rows * g.nwx;
    unsigned long long blocks = (warps * 32 + kCclThreads - 1) / kCclThreads;
    VX_REQUIRE(blocks < 0x7fffffffull, VX_E_UNSUPPORTED, "batch too large for one launch");
    unsigned grid = (unsigned)blocks;
    unsigned *L = nullptr, *bits = nullptr;
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * B->count(), (void**)&L));
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * warps, (void**)&bits));
    {
        VX_LAUNCH(og.c, og.s, "ccl_init_kernel");
        ccl_init_kernel<<<grid, kCclThreads, 0, og.st>>>((const uint8_t*)B->d, L, bits, g);
    }
    VX_TRY(check_launch("ccl_init_kernel"));
    {
        VX_LAUNCH(og.c, og.s, "ccl_merge_kernel");
        if (full) ccl_merge_kernel<true><<<grid, kCclThreads, 0, og.st>>>(L, bits, g);
        else ccl_merge_kernel<false><<<grid, kCclThreads, 0, og.st>>>(L, bits, g);
    }
    VX_TRY(check_launch("ccl_merge_kernel"));
    *L_out = L; *bits_out = bits; *geo = g; *grid_out = grid;
    return VX_OK;
}

}  // namespace vx

using namespace vx;

extern "C" {

int32_t vx_through(vx_ctx ctx, vx_img a, vx_img b, int32_t conn, vx_img* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    VX_REQUIRE(conn == VX_CONN_FACE || conn == VX_CONN_FULL, VX_E_INVALID_ARG,
               "adjacency must be 6 or 26, got %d", conn);
    OpGuard og;
    VX_TRY(og.begin(ctx));
    Image *A = nullptr, *B = nullptr;
    VX_TRY(og.input(a, &A, VX_U8));
    VX_TRY(og.input(b, &B, VX_U8));
    VX_TRY(same_shape(A, B));
    unsigned *L = nullptr, *bits = nullptr, grid = 0;
    Geo g;
    VX_TRY(ccl_label(og, B, conn == VX_CONN_FULL, &L, &bits, &g, &grid));
    {
        VX_LAUNCH(og.c, og.s, "ccl_seed_kernel");
        ccl_seed_kernel<<<grid, kCclThreads, 0, og.st>>>((const uint8_t*)A->d, L, bits, g);
    }
    VX_TRY(check_launch("ccl_seed_kernel"));
    Image* O = nullptr;
    VX_TRY(new_image(og.c, VX_U8, B->nx, B->ny, B->nz, B->nb, B->sp, &O));
    {
        VX_LAUNCH(og.c, og.s, "ccl_select_kernel");
        ccl_select_kernel<SEL_THROUGH><<<grid, kCclThreads, 0, og.st>>>(L, bits, O->d, nullptr, nullptr, g);
    }
    int r = check_launch("ccl_select_kernel");
    if (r == VX_OK) r = scratch_free(og.c, og.s, L);
    if (r == VX_OK) r = scratch_free(og.c, og.s, bits);
    if (r != VX_OK) { drop(O); return r; }
    return og.finish(O, out);
}

int32_t vx_components(vx_ctx ctx, vx_img a, int32_t conn, vx_img* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    VX_REQUIRE(conn == VX_CONN_FACE || conn == VX_CONN_FULL, VX_E_INVALID_ARG,
               "adjacency must be 6 or 26, got %d", conn);
    OpGuard og;
    VX_TRY(og.begin(ctx));
    Image* A = nullptr;
    VX_TRY(og.input(a, &A, VX_U8));
    unsigned *L = nullptr, *bits = nullptr, grid = 0;
    Geo g;
    VX_TRY(ccl_label(og, A, conn == VX_CONN_FULL, &L, &bits, &g, &grid));
    Image* O = nullptr;
    VX_TRY(new_image(og.c, VX_U32, A->nx, A->ny, A->nz, A->nb, A->sp, &O));
    {
        VX_LAUNCH(og.c, og.s, "ccl_select_kernel");
        ccl_select_kernel<SEL_LABELS><<<grid, kCclThreads, 0, og.st>>>(L, bits, O->d, nullptr, nullptr, g);
    }
    int r = check_launch("ccl_select_kernel");
    if (r == VX_OK) r = scratch_free(og.c, og.s, L);
    if (r == VX_OK) r = scratch_free(og.c, og.s, bits);
    if (r != VX_OK) { drop(O); return r; }
    return og.finish(O, out);
}

int32_t vx_maxvol(vx_ctx ctx, vx_img a, int32_t conn, vx_img* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    VX_REQUIRE(conn == VX_CONN_FACE || conn == VX_CONN_FULL, VX_E_INVALID_ARG,
               "adjacency must be 6 or 26, got %d", conn);
    OpGuard og;
    VX_TRY(og.begin(ctx));
    Image* A = nullptr;
    VX_TRY(og.input(a, &A, VX_U8));
    unsigned *L = nullptr, *bits = nullptr, grid = 0;
    Geo g;
    VX_TRY(ccl_label(og, A, conn == VX_CONN_FULL, &L, &bits, &g, &grid));
    unsigned *cnt = nullptr, *vmax = nullptr;
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * A->count(), (void**)&cnt));
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * A->nb, (void**)&vmax));
    VX_CUDA(cudaMemsetAsync(vmax, 0, sizeof(unsigned) * A->nb, og.st));
    {
        VX_LAUNCH(og.c, og.s, "ccl_zero_roots_kernel");
        ccl_zero_roots_kernel<<<grid, kCclThreads, 0, og.st>>>(L, bits, cnt, g);
    }
    VX_TRY(check_launch("ccl_zero_roots_kernel"));
    {
        VX_LAUNCH(og.c, og.s, "ccl_count_kernel");
        ccl_count_kernel<<<grid, kCclThreads, 0, og.st>>>(L, bits, cnt, g);
    }
    VX_TRY(check_launch("ccl_count_kernel"));
    {
        VX_LAUNCH(og.c, og.s, "ccl_max_kernel");
        ccl_max_kernel<<<grid, kCclThreads, 0, og.st>>>(L, bits, cnt, vmax, g);
    }
    VX_TRY(check_launch("ccl_max_kernel"));
    Image* O = nullptr;
    VX_TRY(new_image(og.c, VX_U8, A->nx, A->ny, A->nz, A->nb, A->sp, &O));
    {
        VX_LAUNCH(og.c, og.s, "ccl_select_kernel");
        ccl_select_kernel<SEL_MAXVOL><<<grid, kCclThreads, 0, og.st>>>(L, bits, O->d, cnt, vmax, g);
    }
    int r = check_launch("ccl_select_kernel");
    if (r == VX_OK) r = scratch_free(og.c, og.s, L);
    if (r == VX_OK) r = scratch_free(og.c, og.s, bits);
    if (r == VX_OK) r = scratch_free(og.c, og.s, cnt);
    if (r == VX_OK) r = scratch_free(og.c, og.s, vmax);
    if (r != VX_OK) { drop(O); return r; }
    return og.finish(O, out);
}

}  // extern "C"
