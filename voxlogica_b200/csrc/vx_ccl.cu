// A3 through / reachability and A8 maxvol via union-find connected-component
// labelling (Greta.pdf p.32-38 semantics, p.78 "Reachability via
// Connected Components").
//
//   through(a, b) = union of the connected components of b that contain a voxel of a.
//
// Algorithm (B200-first; not the OpenCL branch's ~29 relabelling sweeps, p.78).
// Everything after the first pass works on 32-voxel bit words, one THREAD per word:
//   1. init     pack the foreground into bit words (128-bit loads) and create the "anchors"
//               of every x-run: the run start (parent = itself) and, where a run crosses a
//               word boundary, the first voxel of the next word (parent = the anchor of the
//               run's segment in the previous word; lanes trade their words by shuffle).
//               Only anchors ever get a parent entry, so the label array is written
//               sparsely (a few entries per word instead of 32);
//   2. merge    per word, contacts with the backward neighbour rows ((y-1,z), (y-1,z-1),
//               (y,z-1), (y+1,z-1) for 26-adjacency; (y-1,z), (y,z-1) for 6) are found
//               with bit operations on the neighbour words; one union per place where a
//               NEW neighbour run begins next to a run of this word.  Unions hang the
//               larger root under the smaller (atomicMin), finds compress paths with
//               atomicMin (this can only replace a parent by an ancestor, never lose a
//               link);
//   3. flatten  every anchor is pointed straight at its root;
//   4. select   per word and per run segment: anchor -> root -> test (seed flag /
//               component size) and write 32 output voxels with two 128-bit stores.
// Roots are the minimum linear index of their component, so labels are canonical.
// Labels are u32 linear indices inside their own volume (< 2^31); bit 31 is the flag.
#include "vx_internal.h"

namespace vx {

constexpr unsigned kFlag = 0x80000000u;
constexpr unsigned kIdx = 0x7fffffffu;
constexpr int kCclThreads = 256;

struct Geo {
    int nx, ny, nz, nb, nwx;    // nwx = 32-voxel words per row
    unsigned long long rows;    // nb*nz*ny
    unsigned long long nwords;  // rows*nwx
    unsigned long long nvox;    // voxels per volume
};

__device__ __forceinline__ unsigned find_root(const unsigned* L, unsigned x) {
    unsigned p = __ldcg(L + x) & kIdx;
    while (p != x) {
        x = p;
        p = __ldcg(L + x) & kIdx;
    }
    return x;
}

// find with path splitting: every node visited is re-pointed at its grandparent (atomicMin can
// only replace a parent by an ancestor, so concurrent unions never lose a link).  Splitting
// halves the path for every later find through these nodes; with min-index linking the
// trees of a percolating component get deep, and compressing only the end points left
// finds walking hundreds of L2-latency hops (ncu: long-scoreboard bound).
__device__ __forceinline__ unsigned find_compress(unsigned* L, unsigned x) {
    unsigned p = __ldcg(L + x) & kIdx;
    while (p != x) {
        const unsigned gp = __ldcg(L + p) & kIdx;
        if (gp == p) return p;
        atomicMin(L + x, gp);
        x = p;
        p = gp;
    }
    return x;
}

__device__ __forceinline__ void unite(unsigned* L, unsigned a, unsigned b) {
    bool done = false;
    do {
        a = find_compress(L, a);
        b = find_compress(L, b);
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

// ---- bit helpers on one 32-voxel word -----------------------------------------------------
// first bit of the run segment (inside the word) that contains set bit j
__device__ __forceinline__ int seg_start(unsigned m, int j) {
    const unsigned zeros_below = ~m & ((1u << j) - 1u);
    return zeros_below ? 32 - __clz(zeros_below) : 0;
}
// mask of the segment that starts at set bit p
__device__ __forceinline__ unsigned seg_mask(unsigned m, int p) {
    const unsigned above = ~(m >> p);
    int len = above ? __ffs(above) - 1 : 32;
    if (len > 32 - p) len = 32 - p;
    return (len >= 32 ? 0xffffffffu : ((1u << len) - 1u)) << p;
}

// the word handled by this thread
struct WordPos {
    bool ok;
    int w, y, z;
    unsigned long long wid, row;
    unsigned vol;
    unsigned lin0;  // linear index (inside the volume) of the word's first voxel
};
__device__ __forceinline__ WordPos word_pos(const Geo& g) {
    WordPos p;
    p.wid = (unsigned long long)blockIdx.x * kCclThreads + threadIdx.x;
    p.ok = p.wid < g.nwords;
    const unsigned long long wid = p.ok ? p.wid : 0;
    p.row = wid / g.nwx;
    p.w = (int)(wid % g.nwx);
    p.y = (int)(p.row % g.ny);
    const unsigned long long t = p.row / g.ny;
    p.z = (int)(t % g.nz);
    p.vol = (unsigned)(t / g.nz);
    p.lin0 = (unsigned)(((unsigned long long)p.z * g.ny + p.y) * g.nx + (unsigned)p.w * 32u);
    return p;
}

// ---- 1. init: one thread per word ----------------------------------------------------------
// 32 voxels are read with two 128-bit loads and packed to a bit word; the anchors of the word
// get their first parents: a run start points at itself, bit 0 of a run that continues from
// the previous word points at that word's last segment (a chain of at most nwx links along a
// row, removed by the first find).  The previous word's bits come from the neighbouring lane.
__device__ __forceinline__ unsigned load_bits32(const uint8_t* __restrict__ src, int nx, int x0);

__global__ void __launch_bounds__(kCclThreads)
ccl_init_kernel(const uint8_t* __restrict__ img, unsigned* __restrict__ Lall,
                unsigned* __restrict__ bits, Geo g) {
    const WordPos p = word_pos(g);
    unsigned m = 0;
    if (p.ok) {
        m = load_bits32(img + p.row * g.nx + (size_t)p.w * 32, g.nx, p.w * 32);
        bits[p.wid] = m;
    }
    unsigned prev_m = __shfl_up_sync(0xffffffffu, m, 1);
    if (!p.ok) return;
    if (p.w == 0) prev_m = 0;
    else if ((threadIdx.x & 31) == 0)
        prev_m = load_bits32(img + p.row * g.nx + (size_t)(p.w - 1) * 32, g.nx, (p.w - 1) * 32);
    if (m == 0) return;
    unsigned* L = Lall + (size_t)p.vol * g.nvox;
    unsigned starts = m & ~((m << 1) | (prev_m >> 31));
    while (starts) {
        const int j = __ffs(starts) - 1;
        starts &= starts - 1;
        L[p.lin0 + j] = p.lin0 + j;
    }
    if ((m & 1u) && (prev_m >> 31)) L[p.lin0] = p.lin0 - 32 + seg_start(prev_m, 31);
}

// ---- 2. merge: one thread per word finds the contacts, the block unites them together -------
// Finding contacts is bit arithmetic on a word and its backward neighbour rows; uniting is
// pointer chasing through L2.  On noisy masks a word has tens of contacts: if each thread
// united its own, a warp would crawl through them with 2-3 live lanes (ncu: 2.7 active
// threads per instruction, long-scoreboard bound).  So contacts are first written as (a, b)
// pairs into a block queue in shared memory, then all 256 threads drain the queue, one
// union per lane: every lane of a warp has an independent chain of loads in flight.
constexpr int kQCap = 4096;   // pairs per block and drain round (32 KB); overflow unites directly

template <bool FULL>
__global__ void __launch_bounds__(kCclThreads)
ccl_merge_kernel(unsigned* __restrict__ Lall, const unsigned* __restrict__ bits, Geo g) {
    __shared__ uint2 q[kQCap];
    __shared__ unsigned qn;
    const WordPos p = word_pos(g);
    // volume of the block's first word; a block spans at most two volumes (bit 31 of a pair's
    // first index says "the second one"; indices inside a volume are < 2^31)
    const unsigned long long words_per_vol = (unsigned long long)g.nwx * g.ny * g.nz;
    const unsigned vol0 = (unsigned)(((unsigned long long)blockIdx.x * kCclThreads) / words_per_vol);
    if (threadIdx.x == 0) qn = 0;
    __syncthreads();
    const unsigned* brow = bits + p.row * g.nwx;
    const unsigned m = p.ok ? __ldg(brow + p.w) : 0u;
    const unsigned prev_m = (p.ok && p.w > 0) ? __ldg(brow + p.w - 1) : 0u;
    const unsigned left = (m << 1) | (prev_m >> 31);  // bit j: voxel j-1 of the same row is set
    const unsigned vtag = (p.vol != vol0) ? kFlag : 0u;
    unsigned* L = Lall + (size_t)p.vol * g.nvox;
    const bool far_vol = p.vol > vol0 + 1;   // cannot happen while a volume has >= 256 words; then unite directly

    const int ndir = FULL ? 4 : 2;
    const int dys[4] = {-1, 0, -1, 1};
    const int dzs[4] = {0, -1, -1, -1};
    // bits of the row (y+dy, z+dz) relative to voxel j of this word: r0 = same x, rm = x-1,
    // rp = x+1; all zero when the row is outside the volume
    auto row_bits = [&](int dy, int dz, unsigned& r0, unsigned& rm, unsigned& rp, unsigned& pw) {
        r0 = rm = rp = pw = 0u;
        const int yy = p.y + dy, zz = p.z + dz;
        if (m == 0 || yy < 0 || yy >= g.ny || zz < 0 || zz >= g.nz) return false;
        const unsigned* nrow = brow + ((long long)dz * g.ny + dy) * g.nwx;
        const unsigned nc = __ldg(nrow + p.w);
        pw = p.w > 0 ? __ldg(nrow + p.w - 1) : 0u;
        const unsigned nn = (FULL && p.w + 1 < g.nwx) ? __ldg(nrow + p.w + 1) : 0u;
        r0 = nc;
        rm = (nc << 1) | (pw >> 31);
        rp = (nc >> 1) | (nn << 31);
        return true;
    };
    // A contact of a run of this word with a run of row (y-1,z-1) [(y+1,z-1)] at voxels x / x'
    // needs no union of its own when a voxel at x or x' of row (y-1,z) [(y+1,z)] or of row
    // (y,z-1) is set: that voxel is adjacent to both runs and is united with each of them by a
    // face-direction step (dy or dz alone) of this pass, which is always made.  On solid
    // components this removes nearly every diagonal union, i.e. half of the pointer chasing.
    // (Skipping the z-face unions that are implied through row y-1 as well was measured: fewer
    // unions, but deeper trees -- the flatten pass and the merge itself got slower on blobs.)
    unsigned u0 = 0, um = 0, up = 0;   // (y-1,z)
    unsigned v0 = 0, vm = 0, vp = 0;   // (y,z-1)
    for (int d0 = 0; d0 < ndir; d0 += 2) {
#pragma unroll
        for (int dd = 0; dd < 2; dd++) {
            const int d = d0 + dd;
            unsigned N0, Nm, Np, np;
            if (!row_bits(dys[d], dzs[d], N0, Nm, Np, np)) continue;
            const unsigned nc = N0;
            const long long drow = (long long)dzs[d] * g.ny + dys[d];
            const unsigned nlin0 = (unsigned)((long long)p.lin0 + drow * g.nx);  // neighbour word, bit 0
            unsigned first, second;
            if (FULL) {
                first = m & Np & ~N0;             // a neighbour run begins at x+1 next to voxel x
                second = m & ~left & (N0 | Nm);   // this run begins at x and touches x or x-1 below
                if (d == 0) { u0 = N0; um = Nm; up = Np; }
                if (d == 1) { v0 = N0; vm = Nm; vp = Np; }
                if (d >= 2) {
                    unsigned c0 = u0 | v0, cm = um | vm, cp = up | vp;
                    if (d == 3) {   // (y+1,z) takes the place of (y-1,z)
                        unsigned w0, wm, wp, wpw;
                        row_bits(1, 0, w0, wm, wp, wpw);
                        c0 = v0 | w0; cm = vm | wm; cp = vp | wp;
                    }
                    first &= ~(c0 | cp);
                    second &= ~(c0 | (~N0 & cm));
                }
            } else {
                first = 0;
                second = m & N0 & ~(left & Nm);   // face contact that the voxel to the left did not see
            }
            const unsigned cnt = __popc(first) + __popc(second);
            if (cnt == 0) continue;
            unsigned slot = far_vol ? kQCap : atomicAdd(&qn, cnt);
            while (first) {
                const int j = __ffs(first) - 1;
                first &= first - 1;
                // the neighbour voxel x+1 starts its run, so it is an anchor itself (bit 0 of the
                // next word when j == 31)
                const unsigned a = p.lin0 + seg_start(m, j), b = nlin0 + j + 1;
                if (slot < kQCap) q[slot++] = make_uint2(a | vtag, b);
                else unite(L, a, b);
            }
            while (second) {
                const int j = __ffs(second) - 1;
                second &= second - 1;
                unsigned b;
                if ((N0 >> j) & 1u) b = nlin0 + seg_start(nc, j);
                else if (j > 0) b = nlin0 + seg_start(nc, j - 1);
                else b = nlin0 - 32 + seg_start(np, 31);
                const unsigned a = p.lin0 + seg_start(m, j);
                if (slot < kQCap) q[slot++] = make_uint2(a | vtag, b);
                else unite(L, a, b);
            }
        }
        __syncthreads();
        const unsigned n = qn < (unsigned)kQCap ? qn : (unsigned)kQCap;
        for (unsigned i = threadIdx.x; i < n; i += kCclThreads) {
            const uint2 e = q[i];
            unite(Lall + (size_t)(vol0 + (e.x >> 31)) * g.nvox, e.x & kIdx, e.y);
        }
        __syncthreads();
        if (threadIdx.x == 0) qn = 0;
        __syncthreads();
    }
}

// anchors of a word: run starts plus bit 0 when the run continues from the previous word
__device__ __forceinline__ unsigned anchor_mask(unsigned m, unsigned prev_m) {
    return (m & ~((m << 1) | (prev_m >> 31))) | (m & 1u & (prev_m >> 31));
}

// ---- 3. flatten ----------------------------------------------------------------------------
__global__ void __launch_bounds__(kCclThreads)
ccl_flatten_kernel(unsigned* __restrict__ Lall, const unsigned* __restrict__ bits, Geo g) {
    const WordPos p = word_pos(g);
    if (!p.ok) return;
    const unsigned* brow = bits + p.row * g.nwx;
    const unsigned m = __ldg(brow + p.w);
    if (m == 0) return;
    const unsigned prev_m = p.w > 0 ? __ldg(brow + p.w - 1) : 0u;
    unsigned* L = Lall + (size_t)p.vol * g.nvox;
    unsigned anchors = anchor_mask(m, prev_m);
    while (anchors) {
        const int j = __ffs(anchors) - 1;
        anchors &= anchors - 1;
        const unsigned a = p.lin0 + j;
        const unsigned r = find_root(L, a);
        if (r != a) L[a] = r;
    }
}

// 32 boolean voxels of `img` at this word as a bit mask
__device__ __forceinline__ unsigned nib4(unsigned w) {
    w = __vcmpne4(w, 0u) & 0x01010101u;
    return ((w * 0x01020408u) >> 24) & 0xfu;
}
__device__ __forceinline__ unsigned load_bits32(const uint8_t* __restrict__ src, int nx, int x0) {
    unsigned v = 0;
    if ((nx & 15) == 0) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(src));
        v = nib4(a.x) | (nib4(a.y) << 4) | (nib4(a.z) << 8) | (nib4(a.w) << 12);
        if (x0 + 16 < nx) {
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(src + 16));
            v |= (nib4(b.x) | (nib4(b.y) << 4) | (nib4(b.z) << 8) | (nib4(b.w) << 12)) << 16;
        }
    } else {
        const int lim = min(32, nx - x0);
        for (int b = 0; b < lim; b++) v |= (unsigned)(src[b] != 0) << b;
    }
    return v;
}
__device__ __forceinline__ unsigned bytes4(unsigned nib) { return (nib * 0x00204081u) & 0x01010101u; }
__device__ __forceinline__ void store_bits32(uint8_t* __restrict__ dst, int nx, int x0, unsigned v) {
    if ((nx & 15) == 0) {
        *reinterpret_cast<uint4*>(dst) = make_uint4(bytes4(v & 0xfu), bytes4((v >> 4) & 0xfu),
                                                    bytes4((v >> 8) & 0xfu), bytes4((v >> 12) & 0xfu));
        if (x0 + 16 < nx)
            *reinterpret_cast<uint4*>(dst + 16) = make_uint4(bytes4((v >> 16) & 0xfu), bytes4((v >> 20) & 0xfu),
                                                             bytes4((v >> 24) & 0xfu), bytes4(v >> 28));
    } else {
        const int lim = min(32, nx - x0);
        for (int b = 0; b < lim; b++) dst[b] = (uint8_t)((v >> b) & 1u);
    }
}

// ---- seeds: every run segment of b that contains a voxel of a flags its root ----------------
__global__ void __launch_bounds__(kCclThreads)
ccl_seed_kernel(const uint8_t* __restrict__ seed, unsigned* __restrict__ Lall,
                const unsigned* __restrict__ bits, Geo g) {
    const WordPos p = word_pos(g);
    if (!p.ok) return;
    const unsigned m = __ldg(bits + p.wid);
    if (m == 0) return;
    unsigned s = load_bits32(seed + p.row * g.nx + (size_t)p.w * 32, g.nx, p.w * 32) & m;
    unsigned* L = Lall + (size_t)p.vol * g.nvox;
    while (s) {
        const int j = __ffs(s) - 1;
        const int st = seg_start(m, j);
        s &= ~seg_mask(m, st);
        const unsigned r = find_root(L, p.lin0 + st);
        if (!(__ldcg(L + r) & kFlag)) atomicOr(L + r, kFlag);
    }
}

enum SelMode { SEL_THROUGH = 0, SEL_LABELS = 1, SEL_MAXVOL = 2 };

// ---- 4. select -----------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kCclThreads)
ccl_select_kernel(const unsigned* __restrict__ Lall, const unsigned* __restrict__ bits, void* out_,
                  const unsigned* __restrict__ cnt, const unsigned* __restrict__ vmax, Geo g) {
    const WordPos p = word_pos(g);
    if (!p.ok) return;
    const unsigned m = __ldg(bits + p.wid);
    const unsigned* L = Lall + (size_t)p.vol * g.nvox;
    const int x0 = p.w * 32;
    const size_t goff = p.row * g.nx + (size_t)x0;
    unsigned rest = m, outbits = 0;
    if (MODE == SEL_LABELS) {
        unsigned* dst = reinterpret_cast<unsigned*>(out_) + goff;
        const int lim = min(32, g.nx - x0);
        unsigned lab[32];
#pragma unroll
        for (int b = 0; b < 32; b++) lab[b] = 0;
        while (rest) {
            const int st = __ffs(rest) - 1;
            const unsigned sm = seg_mask(m, st);
            rest &= ~sm;
            const unsigned r = find_root(L, p.lin0 + st) + 1u;
#pragma unroll
            for (int b = 0; b < 32; b++)
                if ((sm >> b) & 1u) lab[b] = r;
        }
        if ((g.nx & 3) == 0) {
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (4 * q < lim)
                    reinterpret_cast<uint4*>(dst)[q] =
                        make_uint4(lab[4 * q], lab[4 * q + 1], lab[4 * q + 2], lab[4 * q + 3]);
        } else {
#pragma unroll
            for (int b = 0; b < 32; b++)
                if (b < lim) dst[b] = lab[b];
        }
        return;
    }
    while (rest) {
        const int st = __ffs(rest) - 1;
        const unsigned sm = seg_mask(m, st);
        rest &= ~sm;
        const unsigned r = find_root(L, p.lin0 + st);
        bool keep;
        if (MODE == SEL_THROUGH) keep = (__ldcg(L + r) & kFlag) != 0;
        else keep = cnt[(size_t)p.vol * g.nvox + r] == vmax[p.vol];
        if (keep) outbits |= sm;
    }
    store_bits32(reinterpret_cast<uint8_t*>(out_) + goff, g.nx, x0, outbits);
}

// ---- maxvol --------------------------------------------------------------------------------
// step 1: zero the counters of the roots (run starts whose parent is themselves)
__global__ void __launch_bounds__(kCclThreads)
ccl_zero_roots_kernel(const unsigned* __restrict__ Lall, const unsigned* __restrict__ bits,
                      unsigned* __restrict__ cnt, Geo g) {
    const WordPos p = word_pos(g);
    if (!p.ok) return;
    const unsigned* brow = bits + p.row * g.nwx;
    const unsigned m = __ldg(brow + p.w);
    if (m == 0) return;
    const unsigned prev_m = p.w > 0 ? __ldg(brow + p.w - 1) : 0u;
    unsigned starts = m & ~((m << 1) | (prev_m >> 31));
    const size_t vbase = (size_t)p.vol * g.nvox;
    while (starts) {
        const int j = __ffs(starts) - 1;
        starts &= starts - 1;
        if ((Lall[vbase + p.lin0 + j] & kIdx) == p.lin0 + j) cnt[vbase + p.lin0 + j] = 0;
    }
}
// step 2: every run segment adds its length to its root's counter.  Lanes of a warp whose
// segments share a root (the normal case inside a large component: one hot counter) first add
// up among themselves and issue ONE atomicAdd; a per-segment atomic serialised on that counter
// in L2 and made this the longest kernel of maxvol.
__global__ void __launch_bounds__(kCclThreads)
ccl_count_kernel(const unsigned* __restrict__ Lall, const unsigned* __restrict__ bits,
                 unsigned* __restrict__ cnt, Geo g) {
    const WordPos p = word_pos(g);
    const unsigned m = p.ok ? __ldg(bits + p.wid) : 0u;
    const unsigned* L = Lall + (size_t)p.vol * g.nvox;
    const int lane = threadIdx.x & 31;
    unsigned rest = m;
    while (__any_sync(0xffffffffu, rest != 0u)) {
        const bool has = rest != 0u;
        unsigned r = 0, c = 0;
        if (has) {
            const int st = __ffs(rest) - 1;
            const unsigned sm = seg_mask(m, st);
            rest &= ~sm;
            r = find_root(L, p.lin0 + st);
            c = (unsigned)__popc(sm);
        }
        const unsigned long long key = has ? (((unsigned long long)p.vol << 32) | r) : ~0ull;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const unsigned total = __reduce_add_sync(peers, c);
        if (has && lane == __ffs(peers) - 1) atomicAdd(cnt + (size_t)p.vol * g.nvox + r, total);
    }
}
// step 3: per-volume maximum over roots
__global__ void __launch_bounds__(kCclThreads)
ccl_max_kernel(const unsigned* __restrict__ Lall, const unsigned* __restrict__ bits,
               const unsigned* __restrict__ cnt, unsigned* __restrict__ vmax, Geo g) {
    const WordPos p = word_pos(g);
    if (!p.ok) return;
    const unsigned* brow = bits + p.row * g.nwx;
    const unsigned m = __ldg(brow + p.w);
    if (m == 0) return;
    const unsigned prev_m = p.w > 0 ? __ldg(brow + p.w - 1) : 0u;
    unsigned starts = m & ~((m << 1) | (prev_m >> 31));
    const size_t vbase = (size_t)p.vol * g.nvox;
    unsigned v = 0;
    while (starts) {
        const int j = __ffs(starts) - 1;
        starts &= starts - 1;
        if ((Lall[vbase + p.lin0 + j] & kIdx) == p.lin0 + j) v = max(v, cnt[vbase + p.lin0 + j]);
    }
    if (v) atomicMax(vmax + p.vol, v);
}

// ---- slab support: labels / reached flags of one z plane, flagging roots by label -----------
// one thread per word of plane z of every volume
__global__ void __launch_bounds__(kCclThreads)
ccl_plane_kernel(const unsigned* __restrict__ Lall, const unsigned* __restrict__ bits,
                 unsigned* __restrict__ labels, uint8_t* __restrict__ reached, Geo g, int z) {
    const unsigned long long words_per_plane = (unsigned long long)g.nwx * g.ny;
    const unsigned long long t = (unsigned long long)blockIdx.x * kCclThreads + threadIdx.x;
    if (t >= words_per_plane * g.nb) return;
    const unsigned vol = (unsigned)(t / words_per_plane);
    const unsigned long long wp = t % words_per_plane;
    const int y = (int)(wp / g.nwx), w = (int)(wp % g.nwx);
    const unsigned long long wid = ((unsigned long long)vol * g.nz + z) * words_per_plane + wp;
    const unsigned m = __ldg(bits + wid);
    const unsigned* L = Lall + (size_t)vol * g.nvox;
    const unsigned lin0 = (unsigned)(((unsigned long long)z * g.ny + y) * g.nx + (unsigned)w * 32u);
    const int x0 = w * 32, lim = min(32, g.nx - x0);
    const size_t o = ((size_t)vol * g.ny + y) * g.nx + x0;
    unsigned rest = m;
    for (int b = 0; b < lim; b++) { labels[o + b] = 0u; reached[o + b] = 0; }
    while (rest) {
        const int st = __ffs(rest) - 1;
        const unsigned sm = seg_mask(m, st);
        rest &= ~sm;
        const unsigned r = find_root(L, lin0 + st);
        const uint8_t hit = (__ldcg(L + r) & kFlag) ? 1 : 0;
        for (int b = st; b < lim && ((sm >> b) & 1u); b++) { labels[o + b] = r + 1u; reached[o + b] = hit; }
    }
}

__global__ void ccl_flag_labels_kernel(unsigned* __restrict__ L, const unsigned* __restrict__ labels, size_t n,
                                       unsigned nvox) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned lab = labels[i];
    if (lab == 0 || lab > nvox) return;
    const unsigned r = find_root(L, lab - 1u);
    atomicOr(L + r, kFlag);
}

// sizes[i] = voxel count of the component whose label is labels[i] (0 for label 0 / invalid)
__global__ void ccl_sizes_kernel(const unsigned* __restrict__ L, const unsigned* __restrict__ cnt,
                                 const unsigned* __restrict__ labels, size_t n, unsigned nvox,
                                 unsigned* __restrict__ sizes) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned lab = labels[i];
    sizes[i] = (lab == 0 || lab > nvox) ? 0u : cnt[find_root(L, lab - 1u)];
}

// flag every root whose component holds exactly `size` voxels (one thread per word; roots are run starts)
__global__ void __launch_bounds__(kCclThreads)
ccl_flag_size_kernel(unsigned* __restrict__ Lall, const unsigned* __restrict__ bits,
                     const unsigned* __restrict__ cnt, unsigned size, Geo g) {
    const WordPos p = word_pos(g);
    if (!p.ok) return;
    const unsigned* brow = bits + p.row * g.nwx;
    const unsigned m = __ldg(brow + p.w);
    if (m == 0) return;
    const unsigned prev_m = p.w > 0 ? __ldg(brow + p.w - 1) : 0u;
    unsigned starts = m & ~((m << 1) | (prev_m >> 31));
    const size_t vbase = (size_t)p.vol * g.nvox;
    while (starts) {
        const int j = __ffs(starts) - 1;
        starts &= starts - 1;
        const size_t a = vbase + p.lin0 + j;
        if ((Lall[a] & kIdx) == p.lin0 + j && cnt[a] == size) atomicOr(Lall + a, kFlag);
    }
}

static Geo make_geo(const Image* b) {
    Geo g;
    g.nx = b->nx; g.ny = b->ny; g.nz = b->nz; g.nb = b->nb;
    g.nwx = (b->nx + 31) / 32;
    g.rows = (unsigned long long)b->nb * b->nz * b->ny;
    g.nwords = g.rows * g.nwx;
    g.nvox = b->nvox();
    return g;
}

// shared front end: parents + bit words of image B on the op's stream
static int ccl_label(OpGuard& og, const Image* B, bool full, unsigned** L_out, unsigned** bits_out,
                     Geo* geo, unsigned* grid_out, bool flatten = true) {
    Geo g = make_geo(B);
    unsigned long long blocks = (g.nwords + kCclThreads - 1) / kCclThreads;
    VX_REQUIRE(blocks < 0x7fffffffull, VX_E_UNSUPPORTED, "batch too large for one launch");
    unsigned grid = (unsigned)blocks;
    unsigned *L = nullptr, *bits = nullptr;
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * B->count(), (void**)&L));
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * g.nwords, (void**)&bits));
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
    if (flatten) {
        {
            VX_LAUNCH(og.c, og.s, "ccl_flatten_kernel");
            ccl_flatten_kernel<<<grid, kCclThreads, 0, og.st>>>(L, bits, g);
        }
        VX_TRY(check_launch("ccl_flatten_kernel"));
    }
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
    // (select without the flatten pass was measured: -7 % on blobs and sparse masks, +23 % on one
    // giant component at 512^3, where every run then walks the whole chain: flatten stays)
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

// ---------------------------------------------------------------- CCL state (slab decomposition)
namespace vx {
struct CclState {
    Ctx* c = nullptr;
    unsigned *L = nullptr, *bits = nullptr;
    unsigned* cnt = nullptr;    // per-root voxel counts, allocated by vx_ccl_count
    Geo g;
    unsigned grid = 0;
    int nx = 0, ny = 0, nz = 0, nb = 0;
    float sp[3] = {1.f, 1.f, 1.f};
    cudaEvent_t ev = nullptr;   // last operation enqueued on the state
    int home = 0;               // stream the scratch memory belongs to
};
static std::mutex g_ccl_mu;
static std::vector<CclState*> g_ccl(1, nullptr);

static CclState* get_state(vx_ccl h) {
    std::lock_guard<std::mutex> lk(g_ccl_mu);
    if (h == 0 || h >= g_ccl.size()) return nullptr;
    return g_ccl[h];
}
// order this thread's stream after the last operation on the state
static int state_enter(OpGuard& og, vx_ctx ctx, vx_ccl h, CclState** st) {
    VX_TRY(og.begin(ctx));
    *st = get_state(h);
    VX_REQUIRE(*st != nullptr && (*st)->c == og.c, VX_E_INVALID_ARG, "invalid CCL state handle");
    VX_CUDA(cudaStreamWaitEvent(og.st, (*st)->ev, 0));
    return VX_OK;
}
static int state_leave(OpGuard& og, CclState* st) {
    VX_CUDA(cudaEventRecord(st->ev, og.st));
    return VX_OK;
}
}  // namespace vx

extern "C" {

int32_t vx_ccl_create(vx_ctx ctx, vx_img b, int32_t conn, vx_ccl* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    VX_REQUIRE(conn == VX_CONN_FACE || conn == VX_CONN_FULL, VX_E_INVALID_ARG,
               "adjacency must be 6 or 26, got %d", conn);
    OpGuard og;
    VX_TRY(og.begin(ctx));
    Image* B = nullptr;
    VX_TRY(og.input(b, &B, VX_U8));
    CclState* st = new CclState();
    st->c = og.c; st->home = og.s;
    st->nx = B->nx; st->ny = B->ny; st->nz = B->nz; st->nb = B->nb;
    st->sp[0] = B->sp[0]; st->sp[1] = B->sp[1]; st->sp[2] = B->sp[2];
    int rc = ccl_label(og, B, conn == VX_CONN_FULL, &st->L, &st->bits, &st->g, &st->grid);
    if (rc == VX_OK && cudaEventCreateWithFlags(&st->ev, cudaEventDisableTiming) != cudaSuccess)
        rc = fail(VX_E_CUDA, "event creation failed");
    if (rc != VX_OK) { delete st; return rc; }
    VX_TRY(state_leave(og, st));
    VX_TRY(og.finish_scalar());
    std::lock_guard<std::mutex> lk(g_ccl_mu);
    g_ccl.push_back(st);
    *out = g_ccl.size() - 1;
    return VX_OK;
}

int32_t vx_ccl_seed(vx_ctx ctx, vx_ccl state, vx_img a) {
    OpGuard og;
    CclState* st = nullptr;
    VX_TRY(state_enter(og, ctx, state, &st));
    Image* A = nullptr;
    VX_TRY(og.input(a, &A, VX_U8));
    VX_REQUIRE(A->nx == st->nx && A->ny == st->ny && A->nz == st->nz && A->nb == st->nb, VX_E_SHAPE_MISMATCH,
               "seed image does not have the shape of the labelled image");
    {
        VX_LAUNCH(og.c, og.s, "ccl_seed_kernel");
        ccl_seed_kernel<<<st->grid, kCclThreads, 0, og.st>>>((const uint8_t*)A->d, st->L, st->bits, st->g);
    }
    VX_TRY(check_launch("ccl_seed_kernel"));
    VX_TRY(state_leave(og, st));
    return og.finish_scalar();
}

int32_t vx_ccl_plane(vx_ctx ctx, vx_ccl state, int32_t z, vx_img* labels, vx_img* reached) {
    VX_REQUIRE(labels != nullptr && reached != nullptr, VX_E_INVALID_ARG, "NULL output");
    OpGuard og;
    CclState* st = nullptr;
    VX_TRY(state_enter(og, ctx, state, &st));
    VX_REQUIRE(z >= 0 && z < st->nz, VX_E_INVALID_ARG, "plane %d outside [0,%d)", z, st->nz);
    Image *Lb = nullptr, *Rc = nullptr;
    VX_TRY(new_image(og.c, VX_U32, st->nx, st->ny, 1, st->nb, st->sp, &Lb));
    int rc = new_image(og.c, VX_U8, st->nx, st->ny, 1, st->nb, st->sp, &Rc);
    if (rc != VX_OK) { drop(Lb); return rc; }
    const unsigned long long words = (unsigned long long)st->g.nwx * st->ny * st->nb;
    {
        VX_LAUNCH(og.c, og.s, "ccl_plane_kernel");
        ccl_plane_kernel<<<(unsigned)((words + kCclThreads - 1) / kCclThreads), kCclThreads, 0, og.st>>>(
            st->L, st->bits, (unsigned*)Lb->d, (uint8_t*)Rc->d, st->g, z);
    }
    rc = check_launch("ccl_plane_kernel");
    if (rc == VX_OK) rc = state_leave(og, st);
    if (rc == VX_OK) rc = publish(og.c, Lb, og.s);
    if (rc == VX_OK) rc = publish(og.c, Rc, og.s);
    if (rc != VX_OK) { drop(Lb); drop(Rc); return rc; }
    *labels = Lb->handle;
    *reached = Rc->handle;
    return VX_OK;
}

int32_t vx_ccl_flag(vx_ctx ctx, vx_ccl state, const uint32_t* labels, uint64_t n) {
    OpGuard og;
    CclState* st = nullptr;
    VX_TRY(state_enter(og, ctx, state, &st));
    VX_REQUIRE(st->nb == 1, VX_E_UNSUPPORTED, "vx_ccl_flag takes a single-volume state (labels are per volume)");
    if (n == 0) return VX_OK;
    VX_REQUIRE(labels != nullptr, VX_E_INVALID_ARG, "labels is NULL");
    unsigned* d = nullptr;
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * n, (void**)&d));
    VX_CUDA(cudaMemcpyAsync(d, labels, sizeof(unsigned) * n, cudaMemcpyDefault, og.st));
    {
        VX_LAUNCH(og.c, og.s, "ccl_flag_labels_kernel");
        ccl_flag_labels_kernel<<<(unsigned)((n + 255) / 256), 256, 0, og.st>>>(st->L, d, (size_t)n, (unsigned)st->g.nvox);
    }
    VX_TRY(check_launch("ccl_flag_labels_kernel"));
    VX_TRY(scratch_free(og.c, og.s, d));
    return state_leave(og, st);
}

int32_t vx_ccl_select(vx_ctx ctx, vx_ccl state, vx_img* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    OpGuard og;
    CclState* st = nullptr;
    VX_TRY(state_enter(og, ctx, state, &st));
    Image* O = nullptr;
    VX_TRY(new_image(og.c, VX_U8, st->nx, st->ny, st->nz, st->nb, st->sp, &O));
    {
        VX_LAUNCH(og.c, og.s, "ccl_select_kernel");
        ccl_select_kernel<SEL_THROUGH><<<st->grid, kCclThreads, 0, og.st>>>(st->L, st->bits, O->d, nullptr, nullptr, st->g);
    }
    int rc = check_launch("ccl_select_kernel");
    if (rc == VX_OK) rc = state_leave(og, st);
    if (rc != VX_OK) { drop(O); return rc; }
    return og.finish(O, out);
}

// ---- component sizes on a kept state: the local part of a slab-decomposed maxvol -------------
int32_t vx_ccl_count(vx_ctx ctx, vx_ccl state, uint64_t* local_max) {
    VX_REQUIRE(local_max != nullptr, VX_E_INVALID_ARG, "local_max is NULL");
    OpGuard og;
    CclState* st = nullptr;
    VX_TRY(state_enter(og, ctx, state, &st));
    VX_REQUIRE(st->nb == 1, VX_E_UNSUPPORTED, "vx_ccl_count takes a single-volume state");
    unsigned* vmax = nullptr;
    if (st->cnt == nullptr) VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * st->g.nvox, (void**)&st->cnt));
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned), (void**)&vmax));
    VX_CUDA(cudaMemsetAsync(vmax, 0, sizeof(unsigned), og.st));
    {
        VX_LAUNCH(og.c, og.s, "ccl_zero_roots_kernel");
        ccl_zero_roots_kernel<<<st->grid, kCclThreads, 0, og.st>>>(st->L, st->bits, st->cnt, st->g);
    }
    VX_TRY(check_launch("ccl_zero_roots_kernel"));
    {
        VX_LAUNCH(og.c, og.s, "ccl_count_kernel");
        ccl_count_kernel<<<st->grid, kCclThreads, 0, og.st>>>(st->L, st->bits, st->cnt, st->g);
    }
    VX_TRY(check_launch("ccl_count_kernel"));
    {
        VX_LAUNCH(og.c, og.s, "ccl_max_kernel");
        ccl_max_kernel<<<st->grid, kCclThreads, 0, og.st>>>(st->L, st->bits, st->cnt, vmax, st->g);
    }
    VX_TRY(check_launch("ccl_max_kernel"));
    unsigned h = 0;
    VX_CUDA(cudaMemcpyAsync(&h, vmax, sizeof(unsigned), cudaMemcpyDeviceToHost, og.st));
    VX_TRY(scratch_free(og.c, og.s, vmax));
    VX_TRY(state_leave(og, st));
    VX_CUDA(cudaStreamSynchronize(og.st));
    *local_max = h;
    return og.finish_scalar();
}

int32_t vx_ccl_sizes(vx_ctx ctx, vx_ccl state, const uint32_t* labels, uint64_t n, uint32_t* sizes) {
    OpGuard og;
    CclState* st = nullptr;
    VX_TRY(state_enter(og, ctx, state, &st));
    VX_REQUIRE(st->nb == 1, VX_E_UNSUPPORTED, "vx_ccl_sizes takes a single-volume state");
    VX_REQUIRE(st->cnt != nullptr, VX_E_INVALID_ARG, "call vx_ccl_count first");
    if (n == 0) return og.finish_scalar();
    VX_REQUIRE(labels != nullptr && sizes != nullptr, VX_E_INVALID_ARG, "NULL argument");
    unsigned *d = nullptr, *o = nullptr;
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * n, (void**)&d));
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * n, (void**)&o));
    VX_CUDA(cudaMemcpyAsync(d, labels, sizeof(unsigned) * n, cudaMemcpyDefault, og.st));
    {
        VX_LAUNCH(og.c, og.s, "ccl_sizes_kernel");
        ccl_sizes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, og.st>>>(st->L, st->cnt, d, (size_t)n, (unsigned)st->g.nvox, o);
    }
    VX_TRY(check_launch("ccl_sizes_kernel"));
    VX_CUDA(cudaMemcpyAsync(sizes, o, sizeof(unsigned) * n, cudaMemcpyDefault, og.st));
    VX_TRY(scratch_free(og.c, og.s, d));
    VX_TRY(scratch_free(og.c, og.s, o));
    VX_TRY(state_leave(og, st));
    VX_CUDA(cudaStreamSynchronize(og.st));
    return og.finish_scalar();
}

int32_t vx_ccl_flag_size(vx_ctx ctx, vx_ccl state, uint32_t size) {
    OpGuard og;
    CclState* st = nullptr;
    VX_TRY(state_enter(og, ctx, state, &st));
    VX_REQUIRE(st->cnt != nullptr, VX_E_INVALID_ARG, "call vx_ccl_count first");
    if (size != 0) {
        VX_LAUNCH(og.c, og.s, "ccl_flag_size_kernel");
        ccl_flag_size_kernel<<<st->grid, kCclThreads, 0, og.st>>>(st->L, st->bits, st->cnt, size, st->g);
    }
    VX_TRY(check_launch("ccl_flag_size_kernel"));
    VX_TRY(state_leave(og, st));
    return og.finish_scalar();
}

int32_t vx_ccl_destroy(vx_ctx ctx, vx_ccl state) {
    OpGuard og;
    CclState* st = nullptr;
    VX_TRY(state_enter(og, ctx, state, &st));
    {
        std::lock_guard<std::mutex> lk(g_ccl_mu);
        g_ccl[state] = nullptr;
    }
    // the scratch goes back to the free list of the calling stream, which is ordered after the
    // last operation on the state
    scratch_free(og.c, og.s, st->L);
    scratch_free(og.c, og.s, st->bits);
    if (st->cnt) scratch_free(og.c, og.s, st->cnt);
    cudaEventDestroy(st->ev);
    delete st;
    return VX_OK;
}

}  // extern "C"
