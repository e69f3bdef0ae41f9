// A9 percentiles(img, mask, correction) (Greta.pdf p.43
// "Normalization", p.52 percentiles(intensity(flair), brain)).
//
//   for x in mask:  ( #{y in mask : img(y) < img(x)} + correction * #{y in mask : img(y) = img(x)} ) / |mask|
//   outside mask :  0
// The rank is an exact integer; only the final division rounds (binary64, then to f32).
//
// Device algorithm, per volume of the batch and with no host round trip:
//   1. compact the masked voxels into packed (sortable key << 32 | voxel index) pairs
//      (block-aggregated: one atomicAdd per 4096 voxels);
//   2. stable LSD radix sort of the pairs with 9-bit digits.  Only the bits in which the
//      keys of the batch actually differ are sorted: the compaction records min and max key,
//      a one-thread kernel turns their xor into the number of passes (3 for intensities
//      within a few octaves, 4 at most), and the kernels of the remaining passes exit at
//      once -- all decided on the device, the host never waits.  The unit of work is a
//      block-owned tile of 4096 keys: digits are ranked inside each warp with ballots
//      (running counts in warp-private shared memory, no atomics), the tile is reordered in
//      shared memory and written out as runs of consecutive keys per digit;
//   3. every sorted position finds the start of its run of equal keys (its rank) and
//      scatters the normalised rank back to its voxel.
#include <algorithm>
#include <mutex>
#include <type_traits>

#include "vx_bits.h"
#include "vx_internal.h"

namespace vx {

constexpr int kPcThreads = 256;
constexpr int kRadixBits = 9;
constexpr int kDigits = 1 << kRadixBits;         // 512
constexpr int kMaxPasses = (32 + kRadixBits - 1) / kRadixBits;   // 4
// plan[0] = number of passes to run, plan[1] = min key, plan[2] = max key (whole batch)

// Lanes of `active` whose digit equals this lane's digit.  match.any does this in one instruction
// but serialises over the distinct values in the warp (up to 32 with 512 digits; ncu: 70% of
// the scatter kernel's stalls were waiting for it); kRadixBits ballots take constant time.
__device__ __forceinline__ unsigned digit_peers(unsigned active, unsigned d, int nbits) {
    unsigned peers = active;
#pragma unroll
    for (int b = 0; b < 9; b++) {
        if (b < nbits) {
            const int m = (int)(d << (31 - b)) >> 31;      // 0 or ~0: bit b of the digit, as a mask
            const unsigned bal = __ballot_sync(active, m < 0);
            peers &= ~(bal ^ (unsigned)m);
        }
    }
    return peers;
}

__device__ __forceinline__ unsigned sortable_key(float v) {
    v = v + 0.0f;  // -0.0 -> +0.0 so that both zeros get the same rank
    unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ---- 1. compaction -------------------------------------------------------------------
// Pairs are packed as (key << 32) | voxel index so that every pass moves ONE 8-byte word per
// element (a scattered 8-byte store costs one L2 transaction, two 4-byte stores cost two).
// grid (blocks, nb).  Each block walks tiles of 256*16 voxels of its volume.
__global__ void __launch_bounds__(kPcThreads)
pc_compact_kernel(const float* __restrict__ img, const uint8_t* __restrict__ mask, size_t nvox,
                  unsigned long long* __restrict__ pairs, unsigned* __restrict__ count,
                  unsigned* __restrict__ plan) {
    __shared__ unsigned warp_tot[kPcThreads / 32];
    __shared__ unsigned tile_base;
    // the pairs of a tile are collected in shared memory and leave it as one contiguous run:
    // storing them straight from the lanes meant 32 partial-sector writes per store instruction
    __shared__ unsigned long long stage[kPcThreads * 16];
    unsigned kmin = 0xffffffffu, kmax = 0u;
    const size_t vol = blockIdx.y;
    const float* a = img + vol * nvox;
    const uint8_t* m = mask + vol * nvox;
    unsigned long long* po = pairs + vol * nvox;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t tile = (size_t)kPcThreads * 16;
    const bool aligned = (nvox & 15) == 0;  // every volume of the batch starts 16-aligned
    for (size_t t0 = (size_t)blockIdx.x * tile; t0 < nvox; t0 += (size_t)gridDim.x * tile) {
        const size_t first = t0 + (size_t)threadIdx.x * 16;
        unsigned bits = 0;
        float v[16];
        if (aligned && first + 16 <= nvox) {
            const uint4 mw = __ldg(reinterpret_cast<const uint4*>(m + first));
            const unsigned w[4] = {mw.x, mw.y, mw.z, mw.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const unsigned t = __vcmpne4(w[q], 0u) & 0x01010101u;
                bits |= (((t * 0x01020408u) >> 24) & 0xfu) << (4 * q);
                const float4 f = __ldg(reinterpret_cast<const float4*>(a + first) + q);
                v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const size_t i = first + e;
                const bool in = i < nvox;
                v[e] = in ? a[i] : 0.f;
                if (in && m[i] != 0) bits |= 1u << e;
            }
        }
        const unsigned c = __popc(bits);
        // exclusive scan over the block
        unsigned incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        unsigned wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kPcThreads / 32; w++) {
            unsigned t = warp_tot[w];
            if (w < warp) wbase += t;
            total += t;
        }
        if (threadIdx.x == 0) tile_base = total ? atomicAdd(count + vol, total) : 0u;
        unsigned pos = wbase + incl - c;    // position inside the tile
#pragma unroll
        for (int e = 0; e < 16; e++) {
            if ((bits >> e) & 1u) {
                const unsigned key = sortable_key(v[e]);
                kmin = min(kmin, key);
                kmax = max(kmax, key);
                stage[pos] = ((unsigned long long)key << 32) | (unsigned)(first + e);
                pos++;
            }
        }
        __syncthreads();
        const unsigned tb = tile_base;
        for (unsigned i = threadIdx.x; i < total; i += kPcThreads) po[tb + i] = stage[i];
        __syncthreads();  // tile_base / warp_tot / stage reuse
    }
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    if (lane == 0 && kmin <= kmax) {
        atomicMin(plan + 1, kmin);
        atomicMax(plan + 2, kmax);
    }
}

__global__ void rs_plan_kernel(unsigned* __restrict__ plan) {
    const unsigned kmin = plan[1], kmax = plan[2];
    unsigned bits = 0;
    if (kmin <= kmax && kmin != kmax) bits = 32u - (unsigned)__clz(kmin ^ kmax);
    plan[0] = (bits + kRadixBits - 1) / kRadixBits;   // 0 when every key is the same
}

// ---- 2. radix sort passes --------------------------------------------------------------
// The unit of work is a block-owned TILE of kTile keys; warp w owns keys [w*512, (w+1)*512) of
// the tile, so position order = (warp, round, lane) and the ranking is stable.
// table layout: [vol][digit][tile]  (tile-major inside a digit so the scan is one sweep)
constexpr int kTile = 4096;
constexpr int kTileWarps = kPcThreads / 32;        // 8
constexpr int kWarpKeys = kTile / kTileWarps;      // 512
constexpr int kWarpRounds = kWarpKeys / 32;        // 16

__global__ void __launch_bounds__(kPcThreads, 6)
rs_hist_kernel(const unsigned long long* __restrict__ buf0, const unsigned long long* __restrict__ buf1,
               const unsigned* __restrict__ count, size_t nvox, unsigned* __restrict__ table, int ntiles_max,
               int pass, const unsigned* __restrict__ plan) {
    // warp-private counters, plain read-modify-write by the leader lane of every digit group
    // (no shared-memory atomics: they cost 2 cycles per lane)
    __shared__ unsigned short cnt[kTileWarps][kDigits];
    if ((unsigned)pass >= plan[0]) return;
    const unsigned long long* pairs = (pass & 1) ? buf1 : buf0;
    const int shift = pass * kRadixBits;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t vol = blockIdx.y;
    const unsigned n = count[vol];
    // the high words (keys) only: 16 independent loads per lane in flight
    const unsigned* khi = reinterpret_cast<const unsigned*>(pairs + vol * nvox) + 1;
    unsigned* tab = table + (vol * kDigits) * (size_t)ntiles_max;
    const unsigned ntiles = (n + kTile - 1) / kTile;
    // a block walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...: the grid is sized for the SMs,
    // not for the volume (the host does not know how many voxels the mask holds)
    for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const unsigned base = tile * kTile + warp * kWarpKeys;
        unsigned key[kWarpRounds];
#pragma unroll
        for (int r = 0; r < kWarpRounds; r++) {
            const unsigned p = base + r * 32 + lane;
            key[r] = p < n ? __ldcs(khi + 2 * (size_t)p) : 0u;
        }
        for (int d = lane; d < kDigits; d += 32) cnt[warp][d] = 0;
        __syncwarp();
        if ((unsigned long long)(tile + 1) * kTile <= n) {   // a full tile (all but the last of a volume): no bounds work
#pragma unroll
            for (int r = 0; r < kWarpRounds; r++) {
                const unsigned d = (key[r] >> shift) & (kDigits - 1);
                const unsigned peers = digit_peers(0xffffffffu, d, kRadixBits);
                if ((peers & ((1u << lane) - 1u)) == 0) cnt[warp][d] = (unsigned short)(cnt[warp][d] + __popc(peers));
                __syncwarp();
            }
        } else {
#pragma unroll
            for (int r = 0; r < kWarpRounds; r++) {
                const unsigned p = base + r * 32 + lane;
                const bool ok = p < n;
                const unsigned active = __ballot_sync(0xffffffffu, ok);
                if (ok) {
                    const unsigned d = (key[r] >> shift) & (kDigits - 1);
                    const unsigned peers = digit_peers(active, d, kRadixBits);
                    if ((peers & ((1u << lane) - 1u)) == 0) cnt[warp][d] = (unsigned short)(cnt[warp][d] + __popc(peers));
                }
                __syncwarp();
            }
        }
        __syncthreads();
        for (int d = threadIdx.x; d < kDigits; d += kPcThreads) {
            unsigned t = 0;
#pragma unroll
            for (int w = 0; w < kTileWarps; w++) t += cnt[w][d];
            tab[(size_t)d * ntiles_max + tile] = t;
        }
        __syncthreads();   // the counters are zeroed again by the next tile
    }
}

// grid (kDigits, nb): block (d, vol) turns row d of the table (one count per tile) into an
// exclusive prefix over tiles and records the row total; the digit bases are then a
// 512-entry scan that every scatter block does for itself.
__global__ void __launch_bounds__(256)
rs_rowscan_kernel(const unsigned* __restrict__ count, unsigned* __restrict__ table,
                  unsigned* __restrict__ totals, int ntiles_max, int pass, const unsigned* __restrict__ plan) {
    __shared__ unsigned wsum[8];
    __shared__ unsigned carry_s;
    if ((unsigned)pass >= plan[0]) return;
    const size_t vol = blockIdx.y;
    const unsigned d = blockIdx.x;
    const unsigned n = count[vol];
    const unsigned ntiles = (n + kTile - 1) / kTile;
    unsigned* row = table + (vol * kDigits + d) * (size_t)ntiles_max;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (unsigned base = 0; base < ntiles; base += 256) {
        const unsigned i = base + threadIdx.x;
        const unsigned v = i < ntiles ? row[i] : 0u;
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        unsigned wbase = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            unsigned t = wsum[w];
            if (w < warp) wbase += t;
            tot += t;
        }
        const unsigned carry = carry_s;
        if (i < ntiles) row[i] = carry + wbase + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[vol * kDigits + d] = carry_s;
}

// exclusive scan over the block of the per-thread pair (a, b) = values of digits 2t, 2t+1;
// returns the exclusive prefix of digit 2t (digit 2t+1 starts at the return value + a)
__device__ __forceinline__ unsigned block_excl_scan_pairs(unsigned a, unsigned b, unsigned* wsum) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned sum = a + b;
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned wbase = 0;
#pragma unroll
    for (int w = 0; w < kTileWarps; w++)
        if (w < warp) wbase += wsum[w];
    __syncthreads();   // wsum may be reused by the next scan
    return wbase + incl - sum;
}

// Scatter of one tile.  Keys are ranked inside the block, written to shared memory in digit
// order and leave it as runs of consecutive keys of one digit going to consecutive addresses
// (8 keys = 64 B on average).  Scattering each key straight from its lane costs one partial
// 32-byte sector write per 8-byte key: ncu showed 5x DRAM write amplification plus the sector
// fills (654 MB read + 761 MB written for a 148 MB pass) and 8 % issue activity.
__global__ void __launch_bounds__(kPcThreads, 4)
rs_scatter_kernel(unsigned long long* __restrict__ buf0, unsigned long long* __restrict__ buf1,
                  const unsigned* __restrict__ count, size_t nvox, const unsigned* __restrict__ table,
                  const unsigned* __restrict__ totals, int ntiles_max, int pass, const unsigned* __restrict__ plan) {
    __shared__ unsigned long long stage[kTile];             // 32 KB
    __shared__ unsigned short wcnt[kTileWarps][kDigits];    // per-warp digit counts, then warp offsets
    __shared__ unsigned dstart[kDigits];                    // first tile-local position of digit d
    __shared__ unsigned gbase[kDigits];                     // global position of this tile's first key of digit d
    __shared__ unsigned wsum[kTileWarps];
    if ((unsigned)pass >= plan[0]) return;
    const unsigned long long* pairs_in = (pass & 1) ? buf1 : buf0;
    unsigned long long* pairs_out = (pass & 1) ? buf0 : buf1;
    const int shift = pass * kRadixBits;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t vol = blockIdx.y;
    const unsigned n = count[vol];
    const unsigned ntiles = (n + kTile - 1) / kTile;
    if (blockIdx.x >= ntiles) return;
    unsigned long long* po = pairs_out + vol * nvox;
    const unsigned* tab = table + (vol * kDigits) * (size_t)ntiles_max;
    // global start of every digit (exclusive scan of the row totals): once per block
    const int d0 = 2 * threadIdx.x, d1 = d0 + 1;
    unsigned dig0, dig1;
    {
        const unsigned* tot = totals + vol * kDigits;
        const unsigned t0 = tot[d0], t1 = tot[d1];
        dig0 = block_excl_scan_pairs(t0, t1, wsum);
        dig1 = dig0 + t0;
    }
    for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const unsigned tile_n = min((unsigned)kTile, n - tile * kTile);
        const unsigned long long* pi = pairs_in + vol * nvox + (size_t)tile * kTile;
        // 1. loads first (16 independent 8-byte loads per lane), counters zeroed meanwhile
        unsigned long long kv[kWarpRounds];
#pragma unroll
        for (int r = 0; r < kWarpRounds; r++) {
            const unsigned p = warp * kWarpKeys + r * 32 + lane;
            kv[r] = p < tile_n ? __ldcs(pi + p) : 0ull;
        }
        const unsigned tb0 = tab[(size_t)d0 * ntiles_max + tile], tb1 = tab[(size_t)d1 * ntiles_max + tile];
        for (int d = lane; d < kDigits; d += 32) wcnt[warp][d] = 0;
        __syncwarp();
        // 2. rank of every key among the keys of its digit inside its warp
        unsigned short lrank[kWarpRounds];
        if (tile_n == (unsigned)kTile) {   // a full tile: no bounds work in the rounds
#pragma unroll
            for (int r = 0; r < kWarpRounds; r++) {
                const unsigned d = ((unsigned)(kv[r] >> 32) >> shift) & (kDigits - 1);
                const unsigned peers = digit_peers(0xffffffffu, d, kRadixBits);
                const unsigned rank = __popc(peers & ((1u << lane) - 1u));
                const int leader = __ffs(peers) - 1;
                unsigned prev = 0;
                if (lane == leader) {
                    prev = wcnt[warp][d];
                    wcnt[warp][d] = (unsigned short)(prev + __popc(peers));
                }
                prev = __shfl_sync(0xffffffffu, prev, leader);
                lrank[r] = (unsigned short)(prev + rank);
                __syncwarp();
            }
        } else {
#pragma unroll
            for (int r = 0; r < kWarpRounds; r++) {
                const unsigned p = warp * kWarpKeys + r * 32 + lane;
                const bool ok = p < tile_n;
                const unsigned active = __ballot_sync(0xffffffffu, ok);
                unsigned prev = 0, rank = 0;
                int leader = 0;
                if (ok) {
                    const unsigned d = ((unsigned)(kv[r] >> 32) >> shift) & (kDigits - 1);
                    const unsigned peers = digit_peers(active, d, kRadixBits);
                    rank = __popc(peers & ((1u << lane) - 1u));
                    leader = __ffs(peers) - 1;
                    if (lane == leader) {
                        prev = wcnt[warp][d];
                        wcnt[warp][d] = (unsigned short)(prev + __popc(peers));
                    }
                }
                prev = __shfl_sync(0xffffffffu, prev, leader);   // all lanes take part; one instruction
                lrank[r] = (unsigned short)(prev + rank);
                __syncwarp();
            }
        }
        __syncthreads();
        // 3. per digit: offsets of the warps inside the digit, tile-local start, global base
        {
            unsigned c0 = 0, c1 = 0;
#pragma unroll
            for (int w = 0; w < kTileWarps; w++) {
                const unsigned a = wcnt[w][d0], b = wcnt[w][d1];
                wcnt[w][d0] = (unsigned short)c0;
                wcnt[w][d1] = (unsigned short)c1;
                c0 += a;
                c1 += b;
            }
            const unsigned s0 = block_excl_scan_pairs(c0, c1, wsum);
            dstart[d0] = s0;
            dstart[d1] = s0 + c0;
            gbase[d0] = dig0 + tb0;
            gbase[d1] = dig1 + tb1;
        }
        __syncthreads();
        // 4. keys into shared memory in digit order
#pragma unroll
        for (int r = 0; r < kWarpRounds; r++) {
            const unsigned p = warp * kWarpKeys + r * 32 + lane;
            if (p < tile_n) {
                const unsigned d = ((unsigned)(kv[r] >> 32) >> shift) & (kDigits - 1);
                stage[dstart[d] + wcnt[warp][d] + lrank[r]] = kv[r];
            }
        }
        __syncthreads();
        // 5. out: consecutive threads write consecutive keys of a run
#pragma unroll
        for (int i = 0; i < kTile / kPcThreads; i++) {
            const unsigned p = i * kPcThreads + threadIdx.x;
            if (p < tile_n) {
                const unsigned long long v = stage[p];
                const unsigned d = ((unsigned)(v >> 32) >> shift) & (kDigits - 1);
                po[gbase[d] + (p - dstart[d])] = v;
            }
        }
        __syncthreads();   // stage / wcnt / dstart / gbase are rewritten by the next tile
    }
}

// ---- 3. ranks ----------------------------------------------------------------------------
// Ranks of the sorted pairs k[0, n): dst[payload] = (less_before + #keys below + correction *
// #keys equal) / n_total.  A warp owns 32 consecutive sorted positions: the start of a run of
// equal keys (= the rank of all its members) is found with one ballot of the run heads.  With f32
// intensities a few per cent of the keys have a twin, so nearly every warp holds a tie somewhere:
// a per-lane binary search made the whole warp walk ~20 dependent loads for it.
__device__ __forceinline__ void rank_sorted_pairs(const unsigned long long* __restrict__ k, unsigned n,
                                                  unsigned long long less_before, double n_total,
                                                  double correction, float* __restrict__ dst) {
    auto key_at = [&](unsigned i) { return (unsigned)(k[i] >> 32); };
    const int lane = threadIdx.x & 31;
    for (unsigned pb = blockIdx.x * kPcThreads + (threadIdx.x & ~31u); pb < n; pb += gridDim.x * kPcThreads) {
        const unsigned p = pb + lane;
        const bool ok = p < n;
        const unsigned long long kv = ok ? k[p] : 0ull;
        const unsigned key = (unsigned)(kv >> 32);
        unsigned prev = __shfl_up_sync(0xffffffffu, key, 1);
        if (lane == 0 && pb > 0) prev = key_at(pb - 1);
        const bool head = ok && (p == 0 || key != prev);
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        unsigned tails = 0;
        if (correction != 0.0) {   // kernel-uniform: run ends are only needed for the tie correction
            const unsigned next = __shfl_down_sync(0xffffffffu, key, 1);
            const bool tail = ok && (p + 1 >= n || (lane == 31 ? key_at(p + 1) != key : next != key));
            tails = __ballot_sync(0xffffffffu, tail);
        }
        if (!ok) continue;
        const unsigned below = heads & (0xffffffffu >> (31 - lane));
        unsigned less;
        if (below) {
            less = pb + (31 - __clz(below));
        } else {   // the run began before this warp's first position: lower bound
            unsigned lo = 0, hi = pb;
            while (lo < hi) {
                unsigned mid = (lo + hi) >> 1;
                if (key_at(mid) < key) lo = mid + 1; else hi = mid;
            }
            less = lo;
        }
        double num = (double)(less_before + less);
        if (correction != 0.0) {
            const unsigned above = tails & (0xffffffffu << lane);
            unsigned up;   // one past the last member of the run
            if (above) {
                up = pb + (__ffs(above) - 1) + 1;
            } else {   // the run continues beyond this warp: upper bound
                unsigned lo = pb + 32, hi = n;
                while (lo < hi) {
                    unsigned mid = (lo + hi) >> 1;
                    if (key_at(mid) <= key) lo = mid + 1; else hi = mid;
                }
                up = lo;
            }
            num = __dadd_rn(num, __dmul_rn(correction, (double)(up - less)));
        }
        dst[(unsigned)kv] = (float)__ddiv_rn(num, n_total);
    }
}

// the batch: grid (blocks, nb); the payload is the voxel index, so the ranks land in the image
__global__ void __launch_bounds__(kPcThreads)
pc_rank_kernel(const unsigned long long* __restrict__ buf0, const unsigned long long* __restrict__ buf1,
               const unsigned* __restrict__ count, size_t nvox, float* __restrict__ out, double correction,
               const unsigned* __restrict__ plan) {
    const unsigned long long* pairs = (plan[0] & 1u) ? buf1 : buf0;   // where the last pass wrote
    const size_t vol = blockIdx.y;
    const unsigned n = count[vol];
    rank_sorted_pairs(pairs + vol * nvox, n, 0ull, (double)n, correction, out + vol * nvox);
}

// ---- building blocks of a slab-decomposed percentiles (single volume, caller-owned buffers) ----
// ranks by payload: the pair at sorted position p holds key and a payload (an index into the
// caller's receive order); dst[payload] = (less_before + #local keys below + correction * #equal) / n_total
__global__ void __launch_bounds__(kPcThreads)
pc_rank_payload_kernel(const unsigned long long* __restrict__ k, unsigned n, unsigned long long less_before,
                       double n_total, double correction, float* __restrict__ dst) {
    rank_sorted_pairs(k, n, less_before, n_total, correction, dst);
}

// out[payload(pairs[i])] = values[i]
__global__ void __launch_bounds__(kPcThreads)
pc_scatter_payload_kernel(const unsigned long long* __restrict__ pairs, const float* __restrict__ values,
                          unsigned long long n, float* __restrict__ out) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * kPcThreads + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * kPcThreads)
        out[(unsigned)pairs[i]] = values[i];
}

// ---- 4. quantised images: ranks by counting ------------------------------------------------
// An image that carries the quantisation tag (vx_internal.h: built by vx_upload_scaled from 16-bit
// voxels, value strictly monotone in the code) has at most 65536 distinct values, in the order of
// their codes.  Its percentiles need no sort: a histogram of the masked codes, one scan per volume,
// one table look-up per voxel -- 3 + 7 bytes per voxel of traffic instead of ~100 per masked voxel.
// Bit for bit the result of the sorting path: the same exact integer counts enter the same
// binary64 expression.
constexpr int kPqCodes = 65536;
constexpr int kPqWin = 8192;        // codes of the batch's range kept in shared memory (32 KB)

template <typename T>
__device__ __forceinline__ unsigned pq_ucode(unsigned h) {   // 16 raw bits -> order-preserving unsigned code
    return std::is_signed<T>::value ? (h ^ 0x8000u) : h;
}

// grid (blocks, nb), dynamic shared memory kPqWin * 4.  hist is [nb][65536], zeroed by the caller.
template <typename T>
__global__ void __launch_bounds__(256)
pq_hist_kernel(const T* __restrict__ q, const unsigned* __restrict__ mask, size_t nvox, const int* __restrict__ qrange,
               unsigned* __restrict__ hist) {
    extern __shared__ unsigned pq_sh[];
    const size_t vol = blockIdx.y;
    const int ulo = qrange[2 * vol], uhi = qrange[2 * vol + 1];
    const int span = uhi - ulo + 1;
    const bool win = span <= kPqWin;            // the usual case: 12-bit scanner data spans 4096 codes
    unsigned* hv = hist + vol * kPqCodes;
    if (win)
        for (int i = threadIdx.x; i < span; i += 256) pq_sh[i] = 0u;
    __syncthreads();
    const T* qv = q + vol * nvox;
    // the mask is read in its bit form: the 16 voxels of group g are half word g of the volume
    const unsigned short* mh = reinterpret_cast<const unsigned short*>(mask) + (vol * nvox) / 16;
    auto add = [&](unsigned h) {
        const int u = (int)pq_ucode<T>(h);
        if (win) atomicAdd(&pq_sh[u - ulo], 1u); else atomicAdd(&hv[u], 1u);
    };
    // 16 voxels per thread and trip when the volume starts 16-byte aligned in both arrays
    if ((nvox & 15) == 0) {
        const size_t ng = nvox >> 4;
        for (size_t g = (size_t)blockIdx.x * 256 + threadIdx.x; g < ng; g += (size_t)gridDim.x * 256) {
            // codes and mask are asked for together (one memory round trip per group, not two in a row)
            const unsigned m = __ldg(mh + g);
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(qv) + 2 * g);
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(qv) + 2 * g + 1);
            if (m == 0u) continue;
            const unsigned qw[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 16; i++)
                if ((m >> i) & 1u) add((qw[i >> 1] >> (16 * (i & 1))) & 0xffffu);
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvox; i += (size_t)gridDim.x * 256)
            if (flat_bit(mask, vol * nvox + i)) add((unsigned)(unsigned short)qv[i]);
    }
    if (win) {
        __syncthreads();
        for (int i = threadIdx.x; i < span; i += 256) {
            const unsigned c = pq_sh[i];
            if (c) atomicAdd(&hv[ulo + i], c);
        }
    }
}

// One block of 1024 threads per volume: counts -> table[u] = (float)((#below + correction * #equal) / n),
// "below" in VALUE order (descending codes when the scaling is decreasing).  Only the codes of the
// batch's range are visited.
__global__ void __launch_bounds__(1024)
pq_table_kernel(const unsigned* __restrict__ hist, const int* __restrict__ qrange, float* __restrict__ table,
                double correction, int increasing) {
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry_s;
    const size_t vol = blockIdx.x;
    const int ulo = qrange[2 * vol], uhi = qrange[2 * vol + 1];
    const int span = uhi - ulo + 1;
    const unsigned* hv = hist + vol * kPqCodes;
    float* tv = table + vol * kPqCodes;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // position p = 0 .. span-1 in value order; code u(p)
    auto code_at = [&](int p) { return increasing ? ulo + p : uhi - p; };
    // pass 1: total
    unsigned long long tot = 0;
    for (int p = threadIdx.x; p < span; p += 1024) tot += hv[code_at(p)];
#pragma unroll
    for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0) wsum[warp] = tot;
    __syncthreads();
    unsigned long long n = 0;
    for (int w = 0; w < 32; w++) n += wsum[w];
    __syncthreads();
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    if (n == 0) return;    // empty mask: the output stays 0 and no table entry is read
    const double nd = (double)n;
    // pass 2: chunks of 1024 positions, exclusive scan inside the chunk + running carry
    for (int p0 = 0; p0 < span; p0 += 1024) {
        const int p = p0 + threadIdx.x;
        const unsigned c = p < span ? hv[code_at(p)] : 0u;
        unsigned long long incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        unsigned long long wbase = 0, chunk = 0;
        for (int w = 0; w < 32; w++) {
            const unsigned long long t = wsum[w];
            if (w < warp) wbase += t;
            chunk += t;
        }
        const unsigned long long less = carry_s + wbase + incl - c;
        // (codes that do not occur get their entry too -- nobody looks them up, but with them the table
        // is non-decreasing in value order, which lets a threshold of the result be decided on codes)
        if (p < span) {
            double num = (double)less;
            if (correction != 0.0) num = __dadd_rn(num, __dmul_rn(correction, (double)c));
            tv[code_at(p)] = (float)__ddiv_rn(num, nd);
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s += chunk;
        __syncthreads();
    }
}

// out = mask ? table[code] : 0.  grid (blocks, nb), dynamic shared memory kPqWin * 4.
template <typename T>
__global__ void __launch_bounds__(256)
pq_apply_kernel(const T* __restrict__ q, const unsigned* __restrict__ mask, size_t nvox, const int* __restrict__ qrange,
                const float* __restrict__ table, float* __restrict__ out) {
    extern __shared__ unsigned pq_sh[];
    float* tab = reinterpret_cast<float*>(pq_sh);
    const size_t vol = blockIdx.y;
    const int ulo = qrange[2 * vol], uhi = qrange[2 * vol + 1];
    const int span = uhi - ulo + 1;
    const bool win = span <= kPqWin;
    const float* tv = table + vol * kPqCodes;
    if (win)
        for (int i = threadIdx.x; i < span; i += 256) tab[i] = tv[ulo + i];
    __syncthreads();
    const T* qv = q + vol * nvox;
    const unsigned short* mh = reinterpret_cast<const unsigned short*>(mask) + (vol * nvox) / 16;
    float* ov = out + vol * nvox;
    auto look = [&](unsigned h) {
        const int u = (int)pq_ucode<T>(h);
        return win ? tab[u - ulo] : __ldg(tv + u);
    };
    if ((nvox & 15) == 0) {
        const size_t ng = nvox >> 4;
        for (size_t g = (size_t)blockIdx.x * 256 + threadIdx.x; g < ng; g += (size_t)gridDim.x * 256) {
            const unsigned m = __ldg(mh + g);
            float4* o = reinterpret_cast<float4*>(ov) + 4 * g;
            if (m == 0u) {
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                o[0] = z; o[1] = z; o[2] = z; o[3] = z;
                continue;
            }
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(qv) + 2 * g);
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(qv) + 2 * g + 1);
            const unsigned qw[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            float r[16];
#pragma unroll
            for (int i = 0; i < 16; i++)
                r[i] = ((m >> i) & 1u) ? look((qw[i >> 1] >> (16 * (i & 1))) & 0xffffu) : 0.f;
#pragma unroll
            for (int j = 0; j < 4; j++) o[j] = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvox; i += (size_t)gridDim.x * 256)
            ov[i] = flat_bit(mask, vol * nvox + i) ? look((unsigned)(unsigned short)qv[i]) : 0.f;
    }
}

__global__ void pc_set_plan_kernel(unsigned* __restrict__ count, unsigned* __restrict__ plan, unsigned n, unsigned passes) {
    count[0] = n;
    plan[0] = passes;
}

}  // namespace vx

using namespace vx;

static void pq_smem_attrs() {
    static std::once_flag once;
    std::call_once(once, [] {
        cudaFuncSetAttribute(pq_hist_kernel<short>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPqWin * 4);
        cudaFuncSetAttribute(pq_hist_kernel<unsigned short>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPqWin * 4);
        cudaFuncSetAttribute(pq_apply_kernel<short>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPqWin * 4);
        cudaFuncSetAttribute(pq_apply_kernel<unsigned short>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPqWin * 4);
    });
}
// grid: a few blocks per SM in total, every block owns one volume's histogram window
static int pq_blocks_per_volume(size_t nvox, int nb) {
    return std::max(1, std::min((int)((nvox / 16 + 255) / 256), (kNumSMs * 6 + nb - 1) / nb));
}

// the f32 array of a deferred percentile image (Image::lazy): out = mask ? table[volume][u(q)] : 0
int vx::pq_apply_launch(Ctx* c, int s, const Image* src, const unsigned* maskbits, const float* table, float* out) {
    pq_smem_attrs();
    const size_t nvox = src->nvox();
    const int nb = src->nb;
    const dim3 grid(pq_blocks_per_volume(nvox, nb), nb);
    { VX_LAUNCH(c, s, "pq_apply_kernel");
      if (src->q_dtype == VX_I16)
          pq_apply_kernel<short><<<grid, 256, kPqWin * 4, c->streams[s]>>>((const short*)src->q16, maskbits, nvox, src->q_range(), table, out);
      else
          pq_apply_kernel<unsigned short><<<grid, 256, kPqWin * 4, c->streams[s]>>>((const unsigned short*)src->q16, maskbits, nvox,
                                                                                    src->q_range(), table, out); }
    return check_launch("pq_apply_kernel");
}

// percentiles of a quantised image (the tag of vx_internal.h) by counting: the table of every volume
template <typename T>
static int percentiles_table(OpGuard& og, Image* A, const unsigned* Mbits, double correction, float* table) {
    const size_t nvox = A->nvox();
    const int nb = A->nb;
    Scratch sc(og.c, og.s);
    unsigned* hist = nullptr;
    VX_TRY(sc.alloc(sizeof(unsigned) * (size_t)nb * kPqCodes, &hist));
    VX_CUDA(cudaMemsetAsync(hist, 0, sizeof(unsigned) * (size_t)nb * kPqCodes, og.st));
    pq_smem_attrs();
    const T* q = static_cast<const T*>(A->q16);
    const int* qrange = A->q_range();
    const int per_vol = pq_blocks_per_volume(nvox, nb);
    { VX_LAUNCH(og.c, og.s, "pq_hist_kernel");
      pq_hist_kernel<T><<<dim3(per_vol, nb), 256, kPqWin * 4, og.st>>>(q, Mbits, nvox, qrange, hist); }
    VX_TRY(check_launch("pq_hist_kernel"));
    { VX_LAUNCH(og.c, og.s, "pq_table_kernel");
      pq_table_kernel<<<nb, 1024, 0, og.st>>>(hist, qrange, table, correction, A->q_increasing ? 1 : 0); }
    return check_launch("pq_table_kernel");
}

extern "C" int32_t vx_percentiles(vx_ctx ctx, vx_img a, vx_img mask, double correction, vx_img* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    OpGuard og;
    VX_TRY(og.begin(ctx));
    Image *A = nullptr, *M = nullptr;
    VX_TRY(og.input(a, &A, VX_F32));
    if (A->has_codes() && og.c->quantised_paths.load()) {   // at most 65536 distinct values: count, do not sort
        const unsigned* Mbits = nullptr;
        VX_TRY(og.input_bits(mask, &M, &Mbits));
        VX_TRY(same_shape(A, M));
        // The result is DEFERRED (Image::lazy): the tables are made now, the f32 array only if
        // somebody asks for it.  The pointwise programs (thresholds of the GTV spec) read codes, mask
        // and table instead.  A mask whose bytes the caller can rewrite (vx_device_ptr) has no
        // lasting bit form: the array is made at once from this call's private copy of the bits.
        float* table = nullptr;
        VX_TRY(dev_alloc(og.c, og.s, sizeof(float) * (size_t)A->nb * kPqCodes, (void**)&table));
        int rc = A->q_dtype == VX_I16 ? percentiles_table<short>(og, A, Mbits, correction, table)
                                      : percentiles_table<unsigned short>(og, A, Mbits, correction, table);
        Image* O = nullptr;
        const bool defer = Mbits == M->bits;
        if (rc == VX_OK)
            rc = defer ? new_deferred_image(og.c, A->nx, A->ny, A->nz, A->nb, A->sp, &O)
                       : new_image(og.c, VX_F32, A->nx, A->ny, A->nz, A->nb, A->sp, &O);
        if (rc != VX_OK) { dev_free(og.c, og.s, table); return rc; }
        if (defer) {
            O->lazy = new Image::Lazy();
            O->lazy->src = A; A->rc.fetch_add(1); A->part_users.fetch_add(1);
            O->lazy->mask = M; M->rc.fetch_add(1); M->part_users.fetch_add(1);
            O->lazy->table = table;   // freed with the image, on its home stream (= this one)
            O->lazy->monotone = correction >= 0.0 && correction <= 1.0;
        } else {
            rc = pq_apply_launch(og.c, og.s, A, Mbits, table, (float*)O->d);
            dev_free(og.c, og.s, table);
            if (rc != VX_OK) { drop(O); return rc; }
        }
        return og.finish(O, out);
    }
    VX_TRY(og.input(mask, &M, VX_U8));
    VX_TRY(same_shape(A, M));
    const size_t nvox = A->nvox(), total = A->count();
    const int nb = A->nb;
    const int ntiles_max = (int)((nvox + kTile - 1) / kTile);
    unsigned long long* buf = nullptr;
    unsigned *count = nullptr, *table = nullptr, *totals = nullptr, *plan = nullptr;
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned long long) * total * 2, (void**)&buf));
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * (nb + 4), (void**)&count));
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * (size_t)nb * kDigits * ntiles_max, (void**)&table));
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * (size_t)nb * kDigits, (void**)&totals));
    unsigned long long *k0 = buf, *k1 = buf + total;
    plan = count + nb;   // {passes, min key, max key}
    VX_CUDA(cudaMemsetAsync(count, 0, sizeof(unsigned) * (nb + 4), og.st));
    VX_CUDA(cudaMemsetAsync(plan + 1, 0xff, sizeof(unsigned), og.st));
    Image* O = nullptr;
    VX_TRY(new_image(og.c, VX_F32, A->nx, A->ny, A->nz, A->nb, A->sp, &O));
    int rc = VX_OK;
    cudaError_t ce = cudaMemsetAsync(O->d, 0, O->bytes, og.st);
    if (ce != cudaSuccess) rc = fail(VX_E_CUDA, "memset: %s", cudaGetErrorString(ce));
    if (rc == VX_OK) {
        int per_vol = (int)((nvox + (size_t)kPcThreads * 16 - 1) / ((size_t)kPcThreads * 16));
        int cap = (kNumSMs * 8 + nb - 1) / nb;
        dim3 grid(per_vol < cap ? per_vol : cap, nb);
        { VX_LAUNCH(og.c, og.s, "pc_compact_kernel");
          pc_compact_kernel<<<grid, kPcThreads, 0, og.st>>>((const float*)A->d, (const uint8_t*)M->d, nvox, k0, count, plan); }
        rc = check_launch("pc_compact_kernel");
    }
    if (rc == VX_OK) {
        { VX_LAUNCH(og.c, og.s, "rs_plan_kernel");
          rs_plan_kernel<<<1, 1, 0, og.st>>>(plan); }
        rc = check_launch("rs_plan_kernel");
    }
    // Blocks walk the tiles with a grid stride.  The grid is sized for the largest mask that
    // keeps one tile per block up to a few thousand blocks per volume: an SM-sized grid was
    // measured 15 % slower (fewer blocks in flight), one block per POSSIBLE tile launches
    // thousands of blocks that only learn that the mask's count ends before their tile.
    const int gx = std::min(ntiles_max, std::max(1024, 8192 / nb));
    const int hx = gx, sx = gx;
    dim3 hgrid(hx, nb), sgrid(sx, nb);
    for (int pass = 0; pass < kMaxPasses && rc == VX_OK; pass++) {
        { VX_LAUNCH(og.c, og.s, "rs_hist_kernel");
          rs_hist_kernel<<<hgrid, kPcThreads, 0, og.st>>>(k0, k1, count, nvox, table, ntiles_max, pass, plan); }
        rc = check_launch("rs_hist_kernel");
        if (rc != VX_OK) break;
        { VX_LAUNCH(og.c, og.s, "rs_rowscan_kernel");
          rs_rowscan_kernel<<<dim3(kDigits, nb), 256, 0, og.st>>>(count, table, totals, ntiles_max, pass, plan); }
        rc = check_launch("rs_rowscan_kernel");
        if (rc != VX_OK) break;
        { VX_LAUNCH(og.c, og.s, "rs_scatter_kernel");
          rs_scatter_kernel<<<sgrid, kPcThreads, 0, og.st>>>(k0, k1, count, nvox, table, totals, ntiles_max, pass, plan); }
        rc = check_launch("rs_scatter_kernel");
    }
    if (rc == VX_OK) {
        // one block per 256 sorted positions, volume-major block order: the scattered 4-byte
        // stores of one volume land while its 36 MB output is L2-resident, instead of
        // spreading over the whole batch (write amplification at batch 16: 2x slower)
        // (a bounded number of blocks per volume, each walking the sorted positions with a grid stride: one
        // block per 256 voxels of the whole volume meant 400 k blocks per launch that only found
        // out that their positions lie beyond the mask's count)
        int per_vol = (int)std::min<size_t>((nvox + kPcThreads - 1) / kPcThreads, (size_t)std::max(4096, 131072 / nb));
        dim3 grid(per_vol, nb);
        { VX_LAUNCH(og.c, og.s, "pc_rank_kernel");
          pc_rank_kernel<<<grid, kPcThreads, 0, og.st>>>(k0, k1, count, nvox, (float*)O->d, correction, plan); }
        rc = check_launch("pc_rank_kernel");
    }
    if (rc == VX_OK) rc = scratch_free(og.c, og.s, buf);
    if (rc == VX_OK) rc = scratch_free(og.c, og.s, count);
    if (rc == VX_OK) rc = scratch_free(og.c, og.s, table);
    if (rc == VX_OK) rc = scratch_free(og.c, og.s, totals);
    if (rc != VX_OK) { drop(O); return rc; }
    return og.finish(O, out);
}


// ------------------------------------------------------------------------------------------
// Building blocks of a slab-decomposed percentiles.  The caller (one rank
// of voxlogica_b200/slab.py) owns the pair buffers in DEVICE memory; a pair is
// (sortable key << 32) | payload.  Single-volume images only.
extern "C" {

int32_t vx_pc_pairs(vx_ctx ctx, vx_img a, vx_img mask, uint64_t* pairs_dev, uint64_t capacity, uint64_t* n_out) {
    VX_REQUIRE(n_out != nullptr, VX_E_INVALID_ARG, "n_out is NULL");
    OpGuard og;
    VX_TRY(og.begin(ctx));
    Image *A = nullptr, *M = nullptr;
    VX_TRY(og.input(a, &A, VX_F32));
    VX_TRY(og.input(mask, &M, VX_U8));
    VX_TRY(same_shape(A, M));
    VX_REQUIRE(A->nb == 1, VX_E_UNSUPPORTED, "vx_pc_pairs takes a single-volume image");
    const size_t nvox = A->nvox();
    VX_REQUIRE(pairs_dev != nullptr && capacity >= nvox, VX_E_INVALID_ARG,
               "the pair buffer must hold one entry per voxel (%zu)", nvox);
    unsigned* count = nullptr;
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * 8, (void**)&count));
    VX_CUDA(cudaMemsetAsync(count, 0, sizeof(unsigned) * 8, og.st));
    int per_vol = (int)((nvox + (size_t)kPcThreads * 16 - 1) / ((size_t)kPcThreads * 16));
    dim3 grid(per_vol < kNumSMs * 8 ? per_vol : kNumSMs * 8, 1);
    { VX_LAUNCH(og.c, og.s, "pc_compact_kernel");
      pc_compact_kernel<<<grid, kPcThreads, 0, og.st>>>((const float*)A->d, (const uint8_t*)M->d, nvox,
                                                       (unsigned long long*)pairs_dev, count, count + 4); }
    VX_TRY(check_launch("pc_compact_kernel"));
    unsigned h = 0;
    VX_CUDA(cudaMemcpyAsync(&h, count, sizeof(unsigned), cudaMemcpyDeviceToHost, og.st));
    VX_TRY(scratch_free(og.c, og.s, count));
    VX_CUDA(cudaStreamSynchronize(og.st));
    *n_out = h;
    return og.finish_scalar();
}

// stable LSD sort of n pairs by their keys (all 32 key bits: four passes, the result ends in
// pairs_dev); scratch_dev: n entries.  n < 2^32.
int32_t vx_pc_sort(vx_ctx ctx, uint64_t* pairs_dev, uint64_t n, uint64_t* scratch_dev) {
    OpGuard og;
    VX_TRY(og.begin(ctx));
    if (n == 0) return og.finish_scalar();
    VX_REQUIRE(pairs_dev != nullptr && scratch_dev != nullptr, VX_E_INVALID_ARG, "NULL buffer");
    VX_REQUIRE(n < 0xffffffffull, VX_E_UNSUPPORTED, "more than 2^32 - 1 pairs");
    const int ntiles = (int)((n + kTile - 1) / kTile);
    unsigned *count = nullptr, *table = nullptr, *totals = nullptr;
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * 8, (void**)&count));
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * (size_t)kDigits * ntiles, (void**)&table));
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * kDigits, (void**)&totals));
    unsigned* plan = count + 4;
    pc_set_plan_kernel<<<1, 1, 0, og.st>>>(count, plan, (unsigned)n, (unsigned)kMaxPasses);
    VX_TRY(check_launch("pc_set_plan_kernel"));
    unsigned long long *k0 = (unsigned long long*)pairs_dev, *k1 = (unsigned long long*)scratch_dev;
    const int gx = std::min(ntiles, 8192);
    for (int pass = 0; pass < kMaxPasses; pass++) {
        { VX_LAUNCH(og.c, og.s, "rs_hist_kernel");
          rs_hist_kernel<<<dim3(gx, 1), kPcThreads, 0, og.st>>>(k0, k1, count, (size_t)n, table, ntiles, pass, plan); }
        VX_TRY(check_launch("rs_hist_kernel"));
        { VX_LAUNCH(og.c, og.s, "rs_rowscan_kernel");
          rs_rowscan_kernel<<<dim3(kDigits, 1), 256, 0, og.st>>>(count, table, totals, ntiles, pass, plan); }
        VX_TRY(check_launch("rs_rowscan_kernel"));
        { VX_LAUNCH(og.c, og.s, "rs_scatter_kernel");
          rs_scatter_kernel<<<dim3(gx, 1), kPcThreads, 0, og.st>>>(k0, k1, count, (size_t)n, table, totals, ntiles, pass, plan); }
        VX_TRY(check_launch("rs_scatter_kernel"));
    }
    static_assert(kMaxPasses % 2 == 0, "an even number of passes leaves the result in the first buffer");
    VX_TRY(scratch_free(og.c, og.s, count));
    VX_TRY(scratch_free(og.c, og.s, table));
    VX_TRY(scratch_free(og.c, og.s, totals));
    VX_CUDA(cudaStreamSynchronize(og.st));
    return og.finish_scalar();
}

// ranks of n SORTED pairs, written at dst_dev[payload]:
//   (less_before + #{local keys below} + correction * #{local keys equal}) / n_total
int32_t vx_pc_rank(vx_ctx ctx, const uint64_t* sorted_dev, uint64_t n, uint64_t less_before, uint64_t n_total,
                   double correction, float* dst_dev) {
    OpGuard og;
    VX_TRY(og.begin(ctx));
    if (n == 0) return og.finish_scalar();
    VX_REQUIRE(sorted_dev != nullptr && dst_dev != nullptr && n_total > 0 && n < 0xffffffffull, VX_E_INVALID_ARG,
               "invalid arguments");
    const unsigned grid = (unsigned)std::min<uint64_t>((n + kPcThreads - 1) / kPcThreads, 131072);
    { VX_LAUNCH(og.c, og.s, "pc_rank_payload_kernel");
      pc_rank_payload_kernel<<<grid, kPcThreads, 0, og.st>>>((const unsigned long long*)sorted_dev, (unsigned)n,
                                                            (unsigned long long)less_before, (double)n_total, correction, dst_dev); }
    VX_TRY(check_launch("pc_rank_payload_kernel"));
    VX_CUDA(cudaStreamSynchronize(og.st));
    return og.finish_scalar();
}

// new f32 image shaped like `like`: out[payload(pairs[i])] = values[i], 0 elsewhere
int32_t vx_pc_scatter(vx_ctx ctx, const uint64_t* pairs_dev, const float* values_dev, uint64_t n, vx_img like, vx_img* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    OpGuard og;
    VX_TRY(og.begin(ctx));
    Image* L = nullptr;
    VX_TRY(og.input(like, &L, 0));
    VX_REQUIRE(L->nb == 1, VX_E_UNSUPPORTED, "vx_pc_scatter takes a single-volume image");
    VX_REQUIRE(n <= L->nvox() && (n == 0 || (pairs_dev != nullptr && values_dev != nullptr)), VX_E_INVALID_ARG,
               "invalid arguments");
    Image* O = nullptr;
    VX_TRY(new_image(og.c, VX_F32, L->nx, L->ny, L->nz, 1, L->sp, &O));
    int rc = VX_OK;
    if (cudaMemsetAsync(O->d, 0, O->bytes, og.st) != cudaSuccess) rc = fail(VX_E_CUDA, "memset failed");
    if (rc == VX_OK && n > 0) {
        const unsigned grid = (unsigned)std::min<uint64_t>((n + kPcThreads - 1) / kPcThreads, 131072);
        { VX_LAUNCH(og.c, og.s, "pc_scatter_payload_kernel");
          pc_scatter_payload_kernel<<<grid, kPcThreads, 0, og.st>>>((const unsigned long long*)pairs_dev, values_dev,
                                                                   (unsigned long long)n, (float*)O->d); }
        rc = check_launch("pc_scatter_payload_kernel");
    }
    if (rc != VX_OK) { drop(O); return rc; }
    return og.finish(O, out);
}

}  // extern "C"
