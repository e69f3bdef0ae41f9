// A3 through / reachability and A8 maxvol via union-find connected-component
// labelling (Greta.pdf p.32-38 semantics, p.78 "Reachability via
// Connected Components").
//
//   through(a, b) = union of the connected components of b that contain a voxel of a.
//
// Algorithm (B200-first; not the OpenCL branch's ~29 relabelling sweeps, p.78).  Everything works
// on 32-voxel bit words of the rows, one THREAD per word, and on a DENSE numbering of the runs:
//   0. nodes    a node is a run SEGMENT: a maximal run of set voxels inside one row word.  Nodes
//               are numbered in raster order by a prefix sum over the words (count per block,
//               one-block scan of the block totals, block scan), so node ids are dense: the parent
//               array has one u32 per node -- a few MB that stay in L2 -- instead of one per voxel,
//               and the first node of a component in raster order has its smallest id;
//   1. init     parent[id] = id, except for a segment that continues a run of the previous word of
//               its row: it points at that word's last segment (id - 1);
//   2. merge    per word, contacts with the backward neighbour rows ((y-1,z), (y-1,z-1),
//               (y,z-1), (y+1,z-1) for 26-adjacency; (y-1,z), (y,z-1) for 6) are found
//               with bit operations on the neighbour words; one union per place where a
//               NEW neighbour run begins next to a run of this word.  Unions hang the
//               larger root under the smaller (atomicMin), finds compress paths with
//               atomicMin (this can only replace a parent by an ancestor, never lose a
//               link);
//   3. flatten  every node is pointed straight at its root (one thread per NODE: coalesced);
//   4. select   per word and per segment: node -> root -> test (seed flag / component size) and
//               write the 32 output voxels as one word of the result's bit form.
// Roots are the smallest node id of their component = its first voxel in raster order, so labels
// (linear index of that voxel + 1) are canonical.  Bit 31 of a parent entry is the flag.
// Inputs and boolean results are bit-form images (Image::bits): this operator never touches a
// byte per voxel.
#include "vx_bits.h"
#include "vx_internal.h"

namespace vx {

constexpr unsigned kFlag = 0x80000000u;
constexpr unsigned kIdx = 0x7fffffffu;
constexpr int kCclThreads = 256;

struct Geo {
    int align16;                // nx % 16 == 0: row words of a flat bit image are written with plain stores
    int nx, ny, nz, nb, nwx;    // nwx = 32-voxel words per row
    unsigned long long rows;    // nb*nz*ny
    unsigned long long nwords;  // rows*nwx
    unsigned long long nvox;    // voxels per volume
    unsigned nblocks;           // blocks of kCclThreads words
    unsigned nscan;             // blocks of the node numbering pass (kCclThreads x kNodeWords words each)
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
// segments of a word: the bits where a segment starts, and how many there are
__device__ __forceinline__ unsigned seg_starts(unsigned m) { return m & ~(m << 1); }
// node id of the segment that contains set bit j of a word whose first segment is node `base`
__device__ __forceinline__ unsigned node_of(unsigned base, unsigned m, int j) {
    return base + __popc(seg_starts(m) & (0xffffffffu >> (31 - j))) - 1u;
}

// the word handled by this thread
struct WordPos {
    bool ok;
    int w, y, z;
    unsigned long long wid, row;
    unsigned vol;
    unsigned lin0;  // linear index (inside the volume) of the word's first voxel
};
__device__ __forceinline__ WordPos word_pos(const Geo& g, unsigned tile) {
    WordPos p;
    p.wid = (unsigned long long)tile * kCclThreads + threadIdx.x;
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
__device__ __forceinline__ WordPos word_pos(const Geo& g) { return word_pos(g, blockIdx.x); }

// block-wide exclusive scan of one value per thread (kCclThreads threads); *total = block sum
__device__ __forceinline__ unsigned block_exscan(unsigned v, unsigned* total) {
    __shared__ unsigned wsum[kCclThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kCclThreads / 32; w++) {
        const unsigned t = wsum[w];
        if (w < warp) base += t;
        tot += t;
    }
    *total = tot;
    return base + incl - v;
}

// ---- 0 + 1. nodes: bit words, node ids and first parents in ONE pass ------------------------------
// `img` is the image's flat bit form; `bits` receives the row-aligned words the later passes index,
// wbase[word] the id of the word's first node, P[id] the first parents: P[id] = id, except for a
// segment at bit 0 that continues a run of the previous word of its row -- it hangs under that
// word's last segment, the node just before this word's first.
// Node ids are a prefix sum of the segment counts over all words.  Single pass with a decoupled
// look-back (three launches -- count, scan of the block totals, init -- in round 1): a block takes a
// ticket, scans its 256 counts, publishes its total, and its first warp walks back over the status
// words of the blocks before it (32 at a time) until it meets one whose inclusive prefix is known.
// Tickets, not blockIdx, order the blocks: whoever a block waits for is already running.
// status word: state << 32 | value;  state 0 = nothing yet, 1 = block total, 2 = inclusive prefix.
constexpr unsigned long long kStAgg = 1ull << 32, kStIncl = 2ull << 32;
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// status: [nscan] status words (zeroed), then the ticket counter (zeroed), then the node total (output).
// A thread takes kNodeWords consecutive words (a block 1024): the ticket, the two barriers and the look-back
// are paid once per 4 KB of image instead of once per KB -- with one word per thread the pass was all
// overhead (0.15 ms for the 18 MB of 16 BraTS volumes).
constexpr int kNodeWords = 4;
__global__ void __launch_bounds__(kCclThreads)
ccl_nodes_kernel(const unsigned* __restrict__ img, unsigned* __restrict__ bits, unsigned* __restrict__ wbase,
                 unsigned* __restrict__ P, unsigned long long* __restrict__ status, Geo g) {
    __shared__ unsigned s_bid, s_prefix;
    unsigned* ticket = reinterpret_cast<unsigned*>(status + g.nscan);
    unsigned* ntotal = reinterpret_cast<unsigned*>(status + g.nscan + 1);
    if (threadIdx.x == 0) s_bid = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned bid = s_bid;
    const unsigned long long wid0 = ((unsigned long long)bid * kCclThreads + threadIdx.x) * kNodeWords;
    unsigned m[kNodeWords], cont = 0, count = 0;   // cont bit i: word i continues a run of the word before it
    {
        unsigned long long row = (wid0 < g.nwords ? wid0 : 0) / g.nwx;
        int w = (int)((wid0 < g.nwords ? wid0 : 0) % g.nwx);
        unsigned prev = (wid0 < g.nwords && w > 0) ? flat_row_word(img, row, w - 1, g.nx) : 0u;
#pragma unroll
        for (int i = 0; i < kNodeWords; i++) {
            m[i] = wid0 + i < g.nwords ? flat_row_word(img, row, w, g.nx) : 0u;
            if ((m[i] & 1u) && w > 0 && (prev >> 31)) cont |= 1u << i;
            count += __popc(seg_starts(m[i]));
            prev = m[i];
            if (++w == g.nwx) { w = 0; row++; }
        }
    }
    unsigned total;
    const unsigned excl = block_exscan(count, &total);
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        unsigned prefix = 0;
        if (bid > 0) {
            if (lane == 0) st_status(status + bid, kStAgg | total);
            long long look = (long long)bid - 1;
            while (true) {
                const long long idx = look - lane;
                unsigned long long v = idx >= 0 ? ld_status(status + idx) : kStIncl;
                while (__any_sync(0xffffffffu, (v >> 32) == 0ull)) {
                    if ((v >> 32) == 0ull) v = ld_status(status + idx);
                }
                const unsigned incl = __ballot_sync(0xffffffffu, (v >> 32) == 2ull);
                // lanes up to the first one that knows its inclusive prefix contribute
                const int last = incl ? __ffs(incl) - 1 : 31;
                unsigned part = lane <= last ? (unsigned)v : 0u;
#pragma unroll
                for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                prefix += part;
                if (incl) break;
                look -= 32;
            }
        }
        if (lane == 0) {
            st_status(status + bid, kStIncl | (unsigned long long)(prefix + total));
            s_prefix = prefix;
            if (bid == g.nscan - 1) *ntotal = prefix + total;
        }
    }
    __syncthreads();
    if (wid0 >= g.nwords) return;
    unsigned id = s_prefix + excl, base[kNodeWords];
#pragma unroll
    for (int i = 0; i < kNodeWords; i++) {
        base[i] = id;
        unsigned starts = seg_starts(m[i]);
        while (starts) {
            const int j = __ffs(starts) - 1;
            starts &= starts - 1;
            P[id] = (j == 0 && ((cont >> i) & 1u)) ? id - 1u : id;
            id++;
        }
    }
    if (wid0 + kNodeWords <= g.nwords) {   // wid0 is a multiple of 4: aligned 16-byte stores
        *reinterpret_cast<uint4*>(bits + wid0) = make_uint4(m[0], m[1], m[2], m[3]);
        *reinterpret_cast<uint4*>(wbase + wid0) = make_uint4(base[0], base[1], base[2], base[3]);
    } else {
#pragma unroll
        for (int i = 0; i < kNodeWords; i++)
            if (wid0 + i < g.nwords) { bits[wid0 + i] = m[i]; wbase[wid0 + i] = base[i]; }
    }
}

// ---- 2. merge: one thread per word finds the contacts, the block unites them together -------
// Finding contacts is bit arithmetic on a word and its backward neighbour rows; uniting is
// pointer chasing through L2.  On noisy masks a word has tens of contacts: if each thread
// united its own, a warp would crawl through them with 2-3 live lanes (ncu: 2.7 active
// threads per instruction, long-scoreboard bound).  So contacts are first written as (a, b)
// pairs of node ids into a block queue in shared memory, then all 256 threads drain the queue,
// one union per lane: every lane of a warp has an independent chain of loads in flight.
constexpr int kMergeTiles = 4;   // tiles a block takes from the counter at a time
constexpr int kQCap = 4096;   // pairs per block and drain round (32 KB); overflow unites directly

template <bool FULL>
__global__ void __launch_bounds__(kCclThreads)
ccl_merge_kernel(unsigned* __restrict__ P, const unsigned* __restrict__ bits, const unsigned* __restrict__ wbase,
                 unsigned* __restrict__ next_tile, Geo g) {
    __shared__ uint2 q[kQCap];
    __shared__ unsigned qn, s_tile;
    // A block walks over tiles of 256 words, kMergeTiles at a time from a shared counter: with 32 KB of queue a
    // block is expensive to start and the tiles of a mask differ widely in work (one tile per block: through26 on
    // blobs 0.90 ms per 16 BraTS volumes, 0.61 ms this way; a fixed assignment of tiles to blocks lost 10 % on
    // percolating noise, where the blocks that drew the long union chains finished last).
    for (;;) {
    __syncthreads();   // everybody has read s_tile of the previous round
    if (threadIdx.x == 0) s_tile = atomicAdd(next_tile, (unsigned)kMergeTiles);
    __syncthreads();
    const unsigned tile0 = s_tile;
    if (tile0 >= g.nblocks) break;
    for (unsigned tile = tile0; tile < min(tile0 + (unsigned)kMergeTiles, g.nblocks); tile++) {
    const WordPos p = word_pos(g, tile);
    const unsigned* brow = bits + p.row * g.nwx;
    const unsigned* wrow = wbase + p.row * g.nwx;
    const unsigned m = p.ok ? __ldg(brow + p.w) : 0u;
    // a block whose 256 words hold no voxel has nothing to unite (contacts are found from the word that owns
    // the later voxel): one barrier instead of seven.  The masks of the GTV spec after the first are lesions:
    // 99 % of the blocks leave here.
    if (threadIdx.x == 0) qn = 0;
    if (__syncthreads_or(m != 0u) == 0) continue;
    const unsigned prev_m = (p.ok && p.w > 0) ? __ldg(brow + p.w - 1) : 0u;
    const unsigned left = (m << 1) | (prev_m >> 31);  // bit j: voxel j-1 of the same row is set
    const unsigned mybase = m ? __ldg(wrow + p.w) : 0u;

    const int ndir = FULL ? 4 : 2;
    const int dys[4] = {-1, 0, -1, 1};
    const int dzs[4] = {0, -1, -1, -1};
    // bits of the row (y+dy, z+dz) relative to voxel j of this word: r0 = same x, rm = x-1,
    // rp = x+1; all zero when the row is outside the volume
    auto row_bits = [&](int dy, int dz, unsigned& r0, unsigned& rm, unsigned& rp, unsigned& pw, unsigned& nn) {
        r0 = rm = rp = pw = nn = 0u;
        const int yy = p.y + dy, zz = p.z + dz;
        if (m == 0 || yy < 0 || yy >= g.ny || zz < 0 || zz >= g.nz) return false;
        const unsigned* nrow = brow + ((long long)dz * g.ny + dy) * g.nwx;
        const unsigned nc = __ldg(nrow + p.w);
        pw = p.w > 0 ? __ldg(nrow + p.w - 1) : 0u;
        nn = (FULL && p.w + 1 < g.nwx) ? __ldg(nrow + p.w + 1) : 0u;
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
            unsigned N0, Nm, Np, np, nn;
            if (!row_bits(dys[d], dzs[d], N0, Nm, Np, np, nn)) continue;
            const unsigned nc = N0;
            const unsigned* nwrow = wrow + ((long long)dzs[d] * g.ny + dys[d]) * g.nwx;   // node bases of the neighbour row
            unsigned first, second;
            if (FULL) {
                first = m & Np & ~N0;             // a neighbour run begins at x+1 next to voxel x
                second = m & ~left & (N0 | Nm);   // this run begins at x and touches x or x-1 below
                if (d == 0) { u0 = N0; um = Nm; up = Np; }
                if (d == 1) { v0 = N0; vm = Nm; vp = Np; }
                if (d >= 2) {
                    unsigned c0 = u0 | v0, cm = um | vm, cp = up | vp;
                    if (d == 3) {   // (y+1,z) takes the place of (y-1,z)
                        unsigned w0, wm, wp, wpw, wnn;
                        row_bits(1, 0, w0, wm, wp, wpw, wnn);
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
            const unsigned nbase = __ldg(nwrow + p.w);
            unsigned slot = atomicAdd(&qn, cnt);
            while (first) {
                const int j = __ffs(first) - 1;
                first &= first - 1;
                // the neighbour voxel x+1 starts its run (bit 0 of the next word when j == 31)
                const unsigned a = node_of(mybase, m, j);
                const unsigned b = j < 31 ? node_of(nbase, nc, j + 1) : __ldg(nwrow + p.w + 1);
                if (slot < kQCap) q[slot++] = make_uint2(a, b);
                else unite(P, a, b);
            }
            while (second) {
                const int j = __ffs(second) - 1;
                second &= second - 1;
                unsigned b;
                if ((N0 >> j) & 1u) b = node_of(nbase, nc, j);
                else if (j > 0) b = node_of(nbase, nc, j - 1);
                else b = nbase - 1u;   // voxel 31 of the previous word: that word's last segment
                const unsigned a = node_of(mybase, m, j);
                if (slot < kQCap) q[slot++] = make_uint2(a, b);
                else unite(P, a, b);
            }
        }
        __syncthreads();
        const unsigned n = qn < (unsigned)kQCap ? qn : (unsigned)kQCap;
        for (unsigned i = threadIdx.x; i < n; i += kCclThreads) {
            const uint2 e = q[i];
            unite(P, e.x, e.y);
        }
        __syncthreads();
        if (threadIdx.x == 0) qn = 0;
        __syncthreads();
    }
    }
    }
}

// ---- 3. flatten: one thread per node ----------------------------------------------------------
__global__ void __launch_bounds__(kCclThreads)
ccl_flatten_kernel(unsigned* __restrict__ P, const unsigned* __restrict__ ntotal) {
    const unsigned n = __ldg(ntotal);
    for (unsigned id = blockIdx.x * kCclThreads + threadIdx.x; id < n; id += gridDim.x * kCclThreads) {
        const unsigned r = find_root(P, id);
        if (r != id) P[id] = r;
    }
}

// linear index (inside its volume) of every node's first voxel: what labels are made of
__global__ void __launch_bounds__(kCclThreads)
ccl_npos_kernel(const unsigned* __restrict__ bits, const unsigned* __restrict__ wbase, unsigned* __restrict__ npos, Geo g) {
    const WordPos p = word_pos(g);
    if (!p.ok) return;
    const unsigned m = __ldg(bits + p.wid);
    if (m == 0) return;
    unsigned starts = seg_starts(m), id = __ldg(wbase + p.wid);
    while (starts) {
        const int j = __ffs(starts) - 1;
        starts &= starts - 1;
        npos[id++] = p.lin0 + j;
    }
}

// ---- seeds: every run segment of b that contains a voxel of a flags its root ----------------
// four consecutive words of the row-aligned bit image (one 16-byte load; the passes over the words of a
// sparse mask are all launch and address arithmetic with one word per thread)
__device__ __forceinline__ void load_words4(const unsigned* __restrict__ bits, unsigned long long wid0,
                                            unsigned long long nwords, unsigned (&m)[4]) {
    if (wid0 + 4 <= nwords) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(bits + wid0));
        m[0] = v.x; m[1] = v.y; m[2] = v.z; m[3] = v.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) m[i] = wid0 + i < nwords ? __ldg(bits + wid0 + i) : 0u;
    }
}
constexpr int kWordsPerThread = 4;
static unsigned grid_words4(const Geo& g) {
    return (unsigned)((g.nwords + (unsigned long long)kCclThreads * kWordsPerThread - 1) / ((unsigned long long)kCclThreads * kWordsPerThread));
}

__global__ void __launch_bounds__(kCclThreads)
ccl_seed_kernel(const unsigned* __restrict__ seed, unsigned* __restrict__ P, const unsigned* __restrict__ bits,
                const unsigned* __restrict__ wbase, Geo g) {
    const unsigned long long wid0 = ((unsigned long long)blockIdx.x * kCclThreads + threadIdx.x) * kWordsPerThread;
    if (wid0 >= g.nwords) return;
    unsigned m4[4];
    load_words4(bits, wid0, g.nwords, m4);
    if ((m4[0] | m4[1] | m4[2] | m4[3]) == 0u) return;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const unsigned m = m4[i];
        if (m == 0u) continue;
        const unsigned long long wid = wid0 + i, row = wid / g.nwx;
        unsigned s = flat_row_word(seed, row, (int)(wid - row * g.nwx), g.nx) & m;
        if (s == 0u) continue;
        const unsigned base = __ldg(wbase + wid);
        while (s) {
            const int j = __ffs(s) - 1;
            const int st = seg_start(m, j);
            s &= ~seg_mask(m, st);
            const unsigned r = find_root(P, node_of(base, m, st));
            if (!(__ldcg(P + r) & kFlag)) atomicOr(P + r, kFlag);
        }
    }
}

enum SelMode { SEL_THROUGH = 0, SEL_LABELS = 1, SEL_MAXVOL = 2 };

// ---- 4. select -----------------------------------------------------------------------------
// u32 labels (first voxel of the root + 1): one word of the mask per thread, 128 bytes out
__global__ void __launch_bounds__(kCclThreads)
ccl_labels_kernel(const unsigned* __restrict__ P, const unsigned* __restrict__ bits, const unsigned* __restrict__ wbase,
                  unsigned* __restrict__ out, const unsigned* __restrict__ npos, Geo g) {
    const WordPos p = word_pos(g);
    if (!p.ok) return;
    const unsigned m = __ldg(bits + p.wid);
    unsigned id = m ? __ldg(wbase + p.wid) : 0u;
    const int x0 = p.w * 32;
    unsigned rest = m;
    unsigned* dst = out + p.row * g.nx + (size_t)x0;
    const int lim = min(32, g.nx - x0);
    unsigned lab[32];
#pragma unroll
    for (int b = 0; b < 32; b++) lab[b] = 0;
    while (rest) {
        const int st = __ffs(rest) - 1;
        const unsigned sm = seg_mask(m, st);
        rest &= ~sm;
        const unsigned r = __ldg(npos + find_root(P, id++)) + 1u;
#pragma unroll
        for (int b = 0; b < 32; b++)
            if ((sm >> b) & 1u) lab[b] = r;
    }
    if ((g.nx & 3) == 0) {
#pragma unroll
        for (int q = 0; q < 8; q++)
            if (4 * q < lim)
                reinterpret_cast<uint4*>(dst)[q] = make_uint4(lab[4 * q], lab[4 * q + 1], lab[4 * q + 2], lab[4 * q + 3]);
    } else {
#pragma unroll
        for (int b = 0; b < 32; b++)
            if (b < lim) dst[b] = lab[b];
    }
}

// boolean results (through, maxvol): four words per thread
template <int MODE>
__global__ void __launch_bounds__(kCclThreads)
ccl_select4_kernel(const unsigned* __restrict__ P, const unsigned* __restrict__ bits, const unsigned* __restrict__ wbase,
                   unsigned* __restrict__ out, const unsigned* __restrict__ cnt, const unsigned* __restrict__ vmax, Geo g) {
    const unsigned long long wid0 = ((unsigned long long)blockIdx.x * kCclThreads + threadIdx.x) * kWordsPerThread;
    if (wid0 >= g.nwords) return;
    unsigned m4[4];
    load_words4(bits, wid0, g.nwords, m4);
    unsigned long long row = wid0 / g.nwx;
    int w = (int)(wid0 - row * g.nwx);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (wid0 + i >= g.nwords) break;
        const unsigned m = m4[i];
        unsigned outbits = 0;
        if (m != 0u) {
            unsigned id = __ldg(wbase + wid0 + i), rest = m;
            const unsigned vol = (unsigned)(row / ((unsigned long long)g.ny * g.nz));
            while (rest) {
                const int st = __ffs(rest) - 1;
                const unsigned sm = seg_mask(m, st);
                rest &= ~sm;
                const unsigned r = find_root(P, id++);
                bool keep;
                if (MODE == SEL_THROUGH) keep = (__ldcg(P + r) & kFlag) != 0;
                else keep = cnt[r] == vmax[vol];
                if (keep) outbits |= sm;
            }
        }
        flat_store_row_word(out, row, w, g.nx, outbits, g.align16 != 0);
        if (++w == g.nwx) { w = 0; row++; }
    }
}

// ---- maxvol --------------------------------------------------------------------------------
// step 1: zero the counters (one per node)
__global__ void __launch_bounds__(kCclThreads)
ccl_zero_counts_kernel(unsigned* __restrict__ cnt, const unsigned* __restrict__ ntotal) {
    const unsigned n = __ldg(ntotal);
    for (unsigned id = blockIdx.x * kCclThreads + threadIdx.x; id < n; id += gridDim.x * kCclThreads) cnt[id] = 0u;
}
// step 2: every run segment adds its length to its root's counter.  Lanes of a warp whose
// segments share a root (the normal case inside a large component: one hot counter) first add
// up among themselves and issue ONE atomicAdd; a per-segment atomic serialised on that counter
// in L2 and made this the longest kernel of maxvol.
__global__ void __launch_bounds__(kCclThreads)
ccl_count_kernel(const unsigned* __restrict__ P, const unsigned* __restrict__ bits, const unsigned* __restrict__ wbase,
                 unsigned* __restrict__ cnt, Geo g) {
    const WordPos p = word_pos(g);
    const unsigned m = p.ok ? __ldg(bits + p.wid) : 0u;
    unsigned id = m ? __ldg(wbase + p.wid) : 0u;
    const int lane = threadIdx.x & 31;
    unsigned rest = m;
    while (__any_sync(0xffffffffu, rest != 0u)) {
        const bool has = rest != 0u;
        unsigned r = 0xffffffffu, c = 0;
        if (has) {
            const int st = __ffs(rest) - 1;
            const unsigned sm = seg_mask(m, st);
            rest &= ~sm;
            r = find_root(P, id++);
            c = (unsigned)__popc(sm);
        }
        const unsigned peers = __match_any_sync(0xffffffffu, r);
        const unsigned total = __reduce_add_sync(peers, c);
        if (has && lane == __ffs(peers) - 1) atomicAdd(cnt + r, total);
    }
}
// step 3: per-volume maximum over roots
__global__ void __launch_bounds__(kCclThreads)
ccl_max_kernel(const unsigned* __restrict__ P, const unsigned* __restrict__ bits, const unsigned* __restrict__ wbase,
               const unsigned* __restrict__ cnt, unsigned* __restrict__ vmax, Geo g) {
    const WordPos p = word_pos(g);
    if (!p.ok) return;
    const unsigned m = __ldg(bits + p.wid);
    if (m == 0) return;
    const unsigned base = __ldg(wbase + p.wid);
    const int nseg = __popc(seg_starts(m));
    unsigned v = 0;
    for (int i = 0; i < nseg; i++)
        if ((P[base + i] & kIdx) == base + i) v = max(v, cnt[base + i]);
    if (v) atomicMax(vmax + p.vol, v);
}

// ---- slab support: labels / reached flags of one z plane, flagging roots by label -----------
// one thread per word of plane z of every volume
__global__ void __launch_bounds__(kCclThreads)
ccl_plane_kernel(const unsigned* __restrict__ P, const unsigned* __restrict__ bits, const unsigned* __restrict__ wbase,
                 const unsigned* __restrict__ npos, unsigned* __restrict__ labels, uint8_t* __restrict__ reached, Geo g,
                 int z) {
    const unsigned long long words_per_plane = (unsigned long long)g.nwx * g.ny;
    const unsigned long long t = (unsigned long long)blockIdx.x * kCclThreads + threadIdx.x;
    if (t >= words_per_plane * g.nb) return;
    const unsigned vol = (unsigned)(t / words_per_plane);
    const unsigned long long wp = t % words_per_plane;
    const int y = (int)(wp / g.nwx), w = (int)(wp % g.nwx);
    const unsigned long long wid = ((unsigned long long)vol * g.nz + z) * words_per_plane + wp;
    const unsigned m = __ldg(bits + wid);
    unsigned id = m ? __ldg(wbase + wid) : 0u;
    const int x0 = w * 32, lim = min(32, g.nx - x0);
    const size_t o = ((size_t)vol * g.ny + y) * g.nx + x0;
    unsigned rest = m;
    for (int b = 0; b < lim; b++) { labels[o + b] = 0u; reached[o + b] = 0; }
    while (rest) {
        const int st = __ffs(rest) - 1;
        const unsigned sm = seg_mask(m, st);
        rest &= ~sm;
        const unsigned r = find_root(P, id++);
        const uint8_t hit = (__ldcg(P + r) & kFlag) ? 1 : 0;
        const unsigned lab = __ldg(npos + r) + 1u;
        for (int b = st; b < lim && ((sm >> b) & 1u); b++) { labels[o + b] = lab; reached[o + b] = hit; }
    }
}

// node of the run that starts at voxel `lin` of a single-volume state (a label - 1); ~0 when no run starts there
__device__ __forceinline__ unsigned node_at(const unsigned* __restrict__ bits, const unsigned* __restrict__ wbase,
                                            unsigned lin, const Geo& g) {
    const unsigned row = lin / (unsigned)g.nx;
    const unsigned x = lin - row * (unsigned)g.nx;
    const unsigned long long wid = (unsigned long long)row * g.nwx + (x >> 5);
    const unsigned m = __ldg(bits + wid);
    if (!((m >> (x & 31u)) & 1u)) return 0xffffffffu;
    return node_of(__ldg(wbase + wid), m, (int)(x & 31u));
}

__global__ void ccl_flag_labels_kernel(unsigned* __restrict__ P, const unsigned* __restrict__ bits,
                                       const unsigned* __restrict__ wbase, const unsigned* __restrict__ labels, size_t n,
                                       Geo g) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned lab = labels[i];
    if (lab == 0 || lab > (unsigned)g.nvox) return;
    const unsigned id = node_at(bits, wbase, lab - 1u, g);
    if (id == 0xffffffffu) return;
    atomicOr(P + find_root(P, id), kFlag);
}

// sizes[i] = voxel count of the component whose label is labels[i] (0 for label 0 / invalid)
__global__ void ccl_sizes_kernel(const unsigned* __restrict__ P, const unsigned* __restrict__ bits,
                                 const unsigned* __restrict__ wbase, const unsigned* __restrict__ cnt,
                                 const unsigned* __restrict__ labels, size_t n, Geo g, unsigned* __restrict__ sizes) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned lab = labels[i];
    unsigned id = 0xffffffffu;
    if (lab != 0 && lab <= (unsigned)g.nvox) id = node_at(bits, wbase, lab - 1u, g);
    sizes[i] = id == 0xffffffffu ? 0u : cnt[find_root(P, id)];
}

// flag every root whose component holds exactly `size` voxels (one thread per node)
__global__ void __launch_bounds__(kCclThreads)
ccl_flag_size_kernel(unsigned* __restrict__ P, const unsigned* __restrict__ cnt, unsigned size,
                     const unsigned* __restrict__ ntotal) {
    const unsigned n = __ldg(ntotal);
    for (unsigned id = blockIdx.x * kCclThreads + threadIdx.x; id < n; id += gridDim.x * kCclThreads)
        if ((P[id] & kIdx) == id && cnt[id] == size) atomicOr(P + id, kFlag);
}

static Geo make_geo(const Image* b) {
    Geo g;
    g.nx = b->nx; g.ny = b->ny; g.nz = b->nz; g.nb = b->nb;
    g.nwx = (b->nx + 31) / 32;
    g.align16 = (b->nx % 16) == 0;
    g.rows = (unsigned long long)b->nb * b->nz * b->ny;
    g.nwords = g.rows * g.nwx;
    g.nvox = b->nvox();
    g.nblocks = (unsigned)((g.nwords + kCclThreads - 1) / kCclThreads);
    g.nscan = (unsigned)((g.nwords + kCclThreads * kNodeWords - 1) / (kCclThreads * kNodeWords));
    return g;
}

// a new boolean image in bit form for a kernel that writes row words (flat_store_row_word): rows
// that do not start on half words are merged with atomicOr into a zeroed stream
static int new_bits_out(OpGuard& og, int nx, int ny, int nz, int nb, const float* sp, Image** O) {
    VX_TRY(new_bits_image(og.c, nx, ny, nz, nb, sp, O));
    if (nx % 16) {
        cudaError_t e = cudaMemsetAsync((*O)->bits, 0, (*O)->bit_words() * 4, og.st);
        if (e != cudaSuccess) { drop(*O); return fail(VX_E_CUDA, "memset: %s", cudaGetErrorString(e)); }
    }
    return VX_OK;
}

// Device state of one labelling: the row-aligned bit words, the id of every word's first node,
// the parents (one per node; capacity = the most segments the rows can hold) and the scan of the
// block totals, whose last entry is the number of nodes.
struct CclBufs {
    unsigned *bits = nullptr, *wbase = nullptr, *P = nullptr;
    unsigned long long* status = nullptr;   // look-back status words, the ticket, the node total
    size_t cap = 0;   // nodes P can hold
    Geo g;
    unsigned grid = 0;
    const unsigned* ntotal() const { return reinterpret_cast<const unsigned*>(status + g.nscan + 1); }
    unsigned* next_tile() const { return reinterpret_cast<unsigned*>(status + g.nscan + 2); }   // merge pass: tile counter
};
constexpr int kNodeGrid = kNumSMs * 8;   // blocks of the one-thread-per-node kernels (grid stride)

static void ccl_free(Ctx* c, int s, CclBufs& b) {
    if (b.bits) scratch_free(c, s, b.bits);
    if (b.wbase) scratch_free(c, s, b.wbase);
    if (b.P) scratch_free(c, s, b.P);
    if (b.status) scratch_free(c, s, b.status);
    b = CclBufs();
}

// shared front end: nodes, parents and bit words of image B (bit form Bbits) on the op's stream.
// On failure everything allocated here is released.
static int ccl_label(OpGuard& og, const Image* B, const unsigned* Bbits, bool full, CclBufs* out) {
    CclBufs b;
    b.g = make_geo(B);
    const Geo& g = b.g;
    VX_REQUIRE(g.nblocks < 0x7fffffffu && g.nwords < 0xffffffffull, VX_E_UNSUPPORTED, "batch too large for one launch");
    b.grid = g.nblocks;
    // a row of nx voxels holds at most ceil(nx / 2) runs; a run cut by word boundaries adds one node per cut
    const unsigned long long cap = g.rows * (unsigned long long)((g.nx + 1) / 2 + g.nwx);
    VX_REQUIRE(cap < 0x7fffffffull, VX_E_UNSUPPORTED, "too many possible runs (%llu) for 31-bit node ids", cap);
    b.cap = (size_t)cap;
    int rc = scratch_alloc(og.c, og.s, sizeof(unsigned) * g.nwords, (void**)&b.bits);
    if (rc == VX_OK) rc = scratch_alloc(og.c, og.s, sizeof(unsigned) * g.nwords, (void**)&b.wbase);
    if (rc == VX_OK) rc = scratch_alloc(og.c, og.s, sizeof(unsigned long long) * ((size_t)g.nscan + 3), (void**)&b.status);
    if (rc == VX_OK) rc = scratch_alloc(og.c, og.s, sizeof(unsigned) * b.cap, (void**)&b.P);
    auto step = [&](const char* name) { if (rc == VX_OK) rc = check_launch(name); };
    if (rc == VX_OK) {
        cudaError_t ce = cudaMemsetAsync(b.status, 0, sizeof(unsigned long long) * ((size_t)g.nscan + 3), og.st);
        if (ce != cudaSuccess) rc = fail(VX_E_CUDA, "memset: %s", cudaGetErrorString(ce));
    }
    if (rc == VX_OK) {
        VX_LAUNCH(og.c, og.s, "ccl_nodes_kernel");
        ccl_nodes_kernel<<<g.nscan, kCclThreads, 0, og.st>>>(Bbits, b.bits, b.wbase, b.P, b.status, g);
    }
    step("ccl_nodes_kernel");
    if (rc == VX_OK) {
        VX_LAUNCH(og.c, og.s, "ccl_merge_kernel");
        const unsigned mgrid = std::min<unsigned>(b.grid, kNumSMs * 6);   // 6 blocks of 32 KB queue + registers per SM
        if (full) ccl_merge_kernel<true><<<mgrid, kCclThreads, 0, og.st>>>(b.P, b.bits, b.wbase, b.next_tile(), g);
        else ccl_merge_kernel<false><<<mgrid, kCclThreads, 0, og.st>>>(b.P, b.bits, b.wbase, b.next_tile(), g);
    }
    step("ccl_merge_kernel");
    // (select without the flatten pass was measured: -7 % on blobs and sparse masks, +23 % on one
    // giant component at 512^3, where every run then walks the whole chain: flatten stays)
    if (rc == VX_OK) {
        VX_LAUNCH(og.c, og.s, "ccl_flatten_kernel");
        ccl_flatten_kernel<<<kNodeGrid, kCclThreads, 0, og.st>>>(b.P, b.ntotal());
    }
    step("ccl_flatten_kernel");
    if (rc != VX_OK) { ccl_free(og.c, og.s, b); return rc; }
    *out = b;
    return VX_OK;
}

// first-voxel positions of the nodes (labels are made of them): scratch of the caller
static int ccl_npos(OpGuard& og, const CclBufs& b, unsigned** npos) {
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * b.cap, (void**)npos));
    {
        VX_LAUNCH(og.c, og.s, "ccl_npos_kernel");
        ccl_npos_kernel<<<b.grid, kCclThreads, 0, og.st>>>(b.bits, b.wbase, *npos, b.g);
    }
    return check_launch("ccl_npos_kernel");
}

// component sizes: cnt (one counter per node, caller's scratch of b.cap entries) and the per-volume maximum
static int ccl_sizes_of(OpGuard& og, const CclBufs& b, unsigned* cnt, unsigned* vmax) {
    VX_CUDA(cudaMemsetAsync(vmax, 0, sizeof(unsigned) * b.g.nb, og.st));
    {
        VX_LAUNCH(og.c, og.s, "ccl_zero_counts_kernel");
        ccl_zero_counts_kernel<<<kNodeGrid, kCclThreads, 0, og.st>>>(cnt, b.ntotal());
    }
    VX_TRY(check_launch("ccl_zero_counts_kernel"));
    {
        VX_LAUNCH(og.c, og.s, "ccl_count_kernel");
        ccl_count_kernel<<<b.grid, kCclThreads, 0, og.st>>>(b.P, b.bits, b.wbase, cnt, b.g);
    }
    VX_TRY(check_launch("ccl_count_kernel"));
    {
        VX_LAUNCH(og.c, og.s, "ccl_max_kernel");
        ccl_max_kernel<<<b.grid, kCclThreads, 0, og.st>>>(b.P, b.bits, b.wbase, cnt, vmax, b.g);
    }
    return check_launch("ccl_max_kernel");
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
    const unsigned *Ab = nullptr, *Bb = nullptr;
    VX_TRY(og.input_bits(a, &A, &Ab));
    VX_TRY(og.input_bits(b, &B, &Bb));
    VX_TRY(same_shape(A, B));
    CclBufs cb;
    VX_TRY(ccl_label(og, B, Bb, conn == VX_CONN_FULL, &cb));
    Image* O = nullptr;
    int r = VX_OK;
    {
        VX_LAUNCH(og.c, og.s, "ccl_seed_kernel");
        ccl_seed_kernel<<<grid_words4(cb.g), kCclThreads, 0, og.st>>>(Ab, cb.P, cb.bits, cb.wbase, cb.g);
    }
    r = check_launch("ccl_seed_kernel");
    if (r == VX_OK) r = new_bits_out(og, B->nx, B->ny, B->nz, B->nb, B->sp, &O);
    if (r == VX_OK) {
        {
            VX_LAUNCH(og.c, og.s, "ccl_select_kernel");
            ccl_select4_kernel<SEL_THROUGH><<<grid_words4(cb.g), kCclThreads, 0, og.st>>>(cb.P, cb.bits, cb.wbase, O->bits,
                                                                                         nullptr, nullptr, cb.g);
        }
        r = check_launch("ccl_select_kernel");
    }
    ccl_free(og.c, og.s, cb);
    if (r != VX_OK) { if (O) drop(O); return r; }
    return og.finish(O, out);
}

int32_t vx_components(vx_ctx ctx, vx_img a, int32_t conn, vx_img* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    VX_REQUIRE(conn == VX_CONN_FACE || conn == VX_CONN_FULL, VX_E_INVALID_ARG,
               "adjacency must be 6 or 26, got %d", conn);
    OpGuard og;
    VX_TRY(og.begin(ctx));
    Image* A = nullptr;
    const unsigned* Ab = nullptr;
    VX_TRY(og.input_bits(a, &A, &Ab));
    CclBufs cb;
    VX_TRY(ccl_label(og, A, Ab, conn == VX_CONN_FULL, &cb));
    Image* O = nullptr;
    unsigned* npos = nullptr;
    int r = ccl_npos(og, cb, &npos);
    if (r == VX_OK) r = new_image(og.c, VX_U32, A->nx, A->ny, A->nz, A->nb, A->sp, &O);
    if (r == VX_OK) {
        {
            VX_LAUNCH(og.c, og.s, "ccl_select_kernel");
            ccl_labels_kernel<<<cb.grid, kCclThreads, 0, og.st>>>(cb.P, cb.bits, cb.wbase, reinterpret_cast<unsigned*>(O->d), npos, cb.g);
        }
        r = check_launch("ccl_select_kernel");
    }
    if (npos) scratch_free(og.c, og.s, npos);
    ccl_free(og.c, og.s, cb);
    if (r != VX_OK) { if (O) drop(O); return r; }
    return og.finish(O, out);
}

int32_t vx_maxvol(vx_ctx ctx, vx_img a, int32_t conn, vx_img* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    VX_REQUIRE(conn == VX_CONN_FACE || conn == VX_CONN_FULL, VX_E_INVALID_ARG,
               "adjacency must be 6 or 26, got %d", conn);
    OpGuard og;
    VX_TRY(og.begin(ctx));
    Image* A = nullptr;
    const unsigned* Ab = nullptr;
    VX_TRY(og.input_bits(a, &A, &Ab));
    CclBufs cb;
    VX_TRY(ccl_label(og, A, Ab, conn == VX_CONN_FULL, &cb));
    unsigned *cnt = nullptr, *vmax = nullptr;
    Image* O = nullptr;
    int r = scratch_alloc(og.c, og.s, sizeof(unsigned) * cb.cap, (void**)&cnt);
    if (r == VX_OK) r = scratch_alloc(og.c, og.s, sizeof(unsigned) * A->nb, (void**)&vmax);
    if (r == VX_OK) r = ccl_sizes_of(og, cb, cnt, vmax);
    if (r == VX_OK) r = new_bits_out(og, A->nx, A->ny, A->nz, A->nb, A->sp, &O);
    if (r == VX_OK) {
        {
            VX_LAUNCH(og.c, og.s, "ccl_select_kernel");
            ccl_select4_kernel<SEL_MAXVOL><<<grid_words4(cb.g), kCclThreads, 0, og.st>>>(cb.P, cb.bits, cb.wbase, O->bits, cnt,
                                                                                        vmax, cb.g);
        }
        r = check_launch("ccl_select_kernel");
    }
    if (cnt) scratch_free(og.c, og.s, cnt);
    if (vmax) scratch_free(og.c, og.s, vmax);
    ccl_free(og.c, og.s, cb);
    if (r != VX_OK) { if (O) drop(O); return r; }
    return og.finish(O, out);
}

}  // extern "C"

// ---------------------------------------------------------------- CCL state (slab decomposition)
namespace vx {
struct CclState {
    Ctx* c = nullptr;
    CclBufs b;
    unsigned* npos = nullptr;   // first-voxel positions of the nodes, made by the first vx_ccl_plane
    unsigned* cnt = nullptr;    // per-root voxel counts, allocated by vx_ccl_count
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
    const unsigned* Bb = nullptr;
    VX_TRY(og.input_bits(b, &B, &Bb));
    CclState* st = new CclState();
    st->c = og.c; st->home = og.s;
    st->nx = B->nx; st->ny = B->ny; st->nz = B->nz; st->nb = B->nb;
    st->sp[0] = B->sp[0]; st->sp[1] = B->sp[1]; st->sp[2] = B->sp[2];
    int rc = ccl_label(og, B, Bb, conn == VX_CONN_FULL, &st->b);
    if (rc == VX_OK && cudaEventCreateWithFlags(&st->ev, cudaEventDisableTiming) != cudaSuccess) {
        ccl_free(og.c, og.s, st->b);
        rc = fail(VX_E_CUDA, "event creation failed");
    }
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
    const unsigned* Ab = nullptr;
    VX_TRY(og.input_bits(a, &A, &Ab));
    VX_REQUIRE(A->nx == st->nx && A->ny == st->ny && A->nz == st->nz && A->nb == st->nb, VX_E_SHAPE_MISMATCH,
               "seed image does not have the shape of the labelled image");
    {
        VX_LAUNCH(og.c, og.s, "ccl_seed_kernel");
        ccl_seed_kernel<<<grid_words4(st->b.g), kCclThreads, 0, og.st>>>(Ab, st->b.P, st->b.bits, st->b.wbase, st->b.g);
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
    if (st->npos == nullptr) VX_TRY(ccl_npos(og, st->b, &st->npos));
    Image *Lb = nullptr, *Rc = nullptr;
    VX_TRY(new_image(og.c, VX_U32, st->nx, st->ny, 1, st->nb, st->sp, &Lb));
    int rc = new_image(og.c, VX_U8, st->nx, st->ny, 1, st->nb, st->sp, &Rc);
    if (rc != VX_OK) { drop(Lb); return rc; }
    const unsigned long long words = (unsigned long long)st->b.g.nwx * st->ny * st->nb;
    {
        VX_LAUNCH(og.c, og.s, "ccl_plane_kernel");
        ccl_plane_kernel<<<(unsigned)((words + kCclThreads - 1) / kCclThreads), kCclThreads, 0, og.st>>>(
            st->b.P, st->b.bits, st->b.wbase, st->npos, (unsigned*)Lb->d, (uint8_t*)Rc->d, st->b.g, z);
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
        ccl_flag_labels_kernel<<<(unsigned)((n + 255) / 256), 256, 0, og.st>>>(st->b.P, st->b.bits, st->b.wbase, d, (size_t)n,
                                                                              st->b.g);
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
    VX_TRY(new_bits_out(og, st->nx, st->ny, st->nz, st->nb, st->sp, &O));
    {
        VX_LAUNCH(og.c, og.s, "ccl_select_kernel");
        ccl_select4_kernel<SEL_THROUGH><<<grid_words4(st->b.g), kCclThreads, 0, og.st>>>(st->b.P, st->b.bits, st->b.wbase, O->bits,
                                                                                        nullptr, nullptr, st->b.g);
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
    if (st->cnt == nullptr) VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * st->b.cap, (void**)&st->cnt));
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned), (void**)&vmax));
    int rc = ccl_sizes_of(og, st->b, st->cnt, vmax);
    unsigned h = 0;
    if (rc == VX_OK && cudaMemcpyAsync(&h, vmax, sizeof(unsigned), cudaMemcpyDeviceToHost, og.st) != cudaSuccess)
        rc = fail(VX_E_CUDA, "copy of the maximum failed");
    scratch_free(og.c, og.s, vmax);
    VX_TRY(rc);
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
    Scratch sc(og.c, og.s);
    unsigned *d = nullptr, *o = nullptr;
    VX_TRY(sc.alloc(sizeof(unsigned) * n, &d));
    VX_TRY(sc.alloc(sizeof(unsigned) * n, &o));
    VX_CUDA(cudaMemcpyAsync(d, labels, sizeof(unsigned) * n, cudaMemcpyDefault, og.st));
    {
        VX_LAUNCH(og.c, og.s, "ccl_sizes_kernel");
        ccl_sizes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, og.st>>>(st->b.P, st->b.bits, st->b.wbase, st->cnt, d, (size_t)n,
                                                                        st->b.g, o);
    }
    VX_TRY(check_launch("ccl_sizes_kernel"));
    VX_CUDA(cudaMemcpyAsync(sizes, o, sizeof(unsigned) * n, cudaMemcpyDefault, og.st));
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
        ccl_flag_size_kernel<<<kNodeGrid, kCclThreads, 0, og.st>>>(st->b.P, st->cnt, size, st->b.ntotal());
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
    ccl_free(og.c, og.s, st->b);
    if (st->npos) scratch_free(og.c, og.s, st->npos);
    if (st->cnt) scratch_free(og.c, og.s, st->cnt);
    cudaEventDestroy(st->ev);
    delete st;
    return VX_OK;
}

}  // extern "C"
