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
//      warp-owned chunk of 2048 keys: digits are ranked inside the warp with match.any,
//      running offsets live in warp-private shared memory, so no block barrier and no
//      atomics are needed inside a pass;
//   3. every sorted position finds the start of its run of equal keys (its rank) and
//      scatters the normalised rank back to its voxel.
#include "vx_internal.h"

namespace vx {

constexpr int kPcThreads = 256;
constexpr int kWarpChunk = 2048;                 // keys owned by one warp in a sort pass
constexpr int kRounds = kWarpChunk / 32;
constexpr int kPrefetch = 8;                     // rounds whose global loads are issued together
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
            const bool bit = (d >> b) & 1u;
            const unsigned bal = __ballot_sync(active, bit);
            peers &= bit ? bal : ~bal;
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
        __syncthreads();
        unsigned pos = tile_base + wbase + incl - c;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            if ((bits >> e) & 1u) {
                const unsigned key = sortable_key(v[e]);
                kmin = min(kmin, key);
                kmax = max(kmax, key);
                po[pos] = ((unsigned long long)key << 32) | (unsigned)(first + e);
                pos++;
            }
        }
        __syncthreads();  // tile_base / warp_tot reuse
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
// table layout: [vol][digit][chunk]  (chunk-major inside a digit so the scan is one sweep)
__global__ void __launch_bounds__(kPcThreads)
rs_hist_kernel(const unsigned long long* __restrict__ buf0, const unsigned long long* __restrict__ buf1,
               const unsigned* __restrict__ count, size_t nvox, unsigned* __restrict__ table, int nchunks_max,
               int pass, const unsigned* __restrict__ plan) {
    __shared__ unsigned cnt[kPcThreads / 32][kDigits];
    if ((unsigned)pass >= plan[0]) return;
    const unsigned long long* pairs = (pass & 1) ? buf1 : buf0;
    const int shift = pass * kRadixBits;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t vol = blockIdx.y;
    const unsigned n = count[vol];
    const unsigned chunk = blockIdx.x * (kPcThreads / 32) + warp;
    const unsigned nchunks = (n + kWarpChunk - 1) / kWarpChunk;
    if (chunk >= nchunks) return;
    for (int d = lane; d < kDigits; d += 32) cnt[warp][d] = 0;
    __syncwarp();
    const unsigned long long* k = pairs + vol * nvox;
    const unsigned base = chunk * kWarpChunk;
    for (int r0 = 0; r0 < kRounds; r0 += kPrefetch) {
        unsigned keyhi[kPrefetch];   // the loads of kPrefetch rounds are in flight together
#pragma unroll
        for (int j = 0; j < kPrefetch; j++) {
            const unsigned p = base + (r0 + j) * 32 + lane;
            keyhi[j] = p < n ? (unsigned)(__ldcs(k + p) >> 32) : 0u;
        }
#pragma unroll
        for (int j = 0; j < kPrefetch; j++) {
            const unsigned p = base + (r0 + j) * 32 + lane;
            const bool ok = p < n;
            const unsigned active = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const unsigned d = (keyhi[j] >> shift) & (kDigits - 1);
                const unsigned peers = digit_peers(active, d, kRadixBits);
                if ((peers & ((1u << lane) - 1u)) == 0) cnt[warp][d] += __popc(peers);
            }
            __syncwarp();
        }
    }
    unsigned* tab = table + (vol * kDigits) * (size_t)nchunks_max;
    for (int d = lane; d < kDigits; d += 32) tab[(size_t)d * nchunks_max + chunk] = cnt[warp][d];
}

// grid (kDigits, nb): block (d, vol) turns row d of the table (one count per chunk) into an
// exclusive prefix over chunks and records the row total; the digit bases are then a
// 256-entry scan that every scatter warp does for itself.
__global__ void __launch_bounds__(256)
rs_rowscan_kernel(const unsigned* __restrict__ count, unsigned* __restrict__ table,
                  unsigned* __restrict__ totals, int nchunks_max, int pass, const unsigned* __restrict__ plan) {
    __shared__ unsigned wsum[8];
    __shared__ unsigned carry_s;
    if ((unsigned)pass >= plan[0]) return;
    const size_t vol = blockIdx.y;
    const unsigned d = blockIdx.x;
    const unsigned n = count[vol];
    const unsigned nchunks = (n + kWarpChunk - 1) / kWarpChunk;
    unsigned* row = table + (vol * kDigits + d) * (size_t)nchunks_max;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (unsigned base = 0; base < nchunks; base += 256) {
        const unsigned i = base + threadIdx.x;
        const unsigned v = i < nchunks ? row[i] : 0u;
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
        if (i < nchunks) row[i] = carry + wbase + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[vol * kDigits + d] = carry_s;
}

__global__ void __launch_bounds__(kPcThreads)
rs_scatter_kernel(unsigned long long* __restrict__ buf0, unsigned long long* __restrict__ buf1,
                  const unsigned* __restrict__ count, size_t nvox, const unsigned* __restrict__ table,
                  const unsigned* __restrict__ totals, int nchunks_max, int pass, const unsigned* __restrict__ plan) {
    __shared__ unsigned off[kPcThreads / 32][kDigits];
    if ((unsigned)pass >= plan[0]) return;
    const unsigned long long* pairs_in = (pass & 1) ? buf1 : buf0;
    unsigned long long* pairs_out = (pass & 1) ? buf0 : buf1;
    const int shift = pass * kRadixBits;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t vol = blockIdx.y;
    const unsigned n = count[vol];
    const unsigned chunk = blockIdx.x * (kPcThreads / 32) + warp;
    const unsigned nchunks = (n + kWarpChunk - 1) / kWarpChunk;
    if (chunk >= nchunks) return;
    const unsigned* tab = table + (vol * kDigits) * (size_t)nchunks_max;
    {   // digit bases: exclusive scan of the row totals, kDigits/32 digits per lane
        constexpr int PL = kDigits / 32;
        const unsigned* tot = totals + vol * kDigits;
        unsigned t[PL], sum = 0;
#pragma unroll
        for (int j = 0; j < PL; j++) { t[j] = tot[lane * PL + j]; sum += t[j]; }
        unsigned incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        unsigned base = incl - sum;
#pragma unroll
        for (int j = 0; j < PL; j++) {
            const int d = lane * PL + j;
            off[warp][d] = base + tab[(size_t)d * nchunks_max + chunk];
            base += t[j];
        }
    }
    __syncwarp();
    const unsigned long long* pi = pairs_in + vol * nvox;
    unsigned long long* po = pairs_out + vol * nvox;
    const unsigned base = chunk * kWarpChunk;
    for (int r0 = 0; r0 < kRounds; r0 += kPrefetch) {
        unsigned long long kvs[kPrefetch];   // kPrefetch independent loads in flight per lane
#pragma unroll
        for (int j = 0; j < kPrefetch; j++) {
            const unsigned p = base + (r0 + j) * 32 + lane;
            kvs[j] = p < n ? __ldcs(pi + p) : 0ull;
        }
#pragma unroll
        for (int j = 0; j < kPrefetch; j++) {
            const unsigned p = base + (r0 + j) * 32 + lane;
            const bool ok = p < n;
            const unsigned active = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const unsigned long long kv = kvs[j];
                const unsigned d = ((unsigned)(kv >> 32) >> shift) & (kDigits - 1);
                const unsigned peers = digit_peers(active, d, kRadixBits);
                const unsigned rank = __popc(peers & ((1u << lane) - 1u));
                const int leader = __ffs(peers) - 1;
                unsigned start = 0;
                if (lane == leader) {
                    start = off[warp][d];
                    off[warp][d] = start + __popc(peers);
                }
                start = __shfl_sync(peers, start, leader);
                po[start + rank] = kv;
            }
            __syncwarp();
        }
    }
}

// ---- 3. ranks ----------------------------------------------------------------------------
__global__ void __launch_bounds__(kPcThreads)
pc_rank_kernel(const unsigned long long* __restrict__ buf0, const unsigned long long* __restrict__ buf1,
               const unsigned* __restrict__ count, size_t nvox, float* __restrict__ out, double correction,
               const unsigned* __restrict__ plan) {
    const unsigned long long* pairs = (plan[0] & 1u) ? buf1 : buf0;   // where the last pass wrote
    const size_t vol = blockIdx.y;
    const unsigned n = count[vol];
    const unsigned long long* k = pairs + vol * nvox;
    float* o = out + vol * nvox;
    auto key_at = [&](unsigned i) { return (unsigned)(k[i] >> 32); };
    for (unsigned p = blockIdx.x * kPcThreads + threadIdx.x; p < n; p += gridDim.x * kPcThreads) {
        const unsigned long long kv = k[p];
        const unsigned key = (unsigned)(kv >> 32);
        unsigned less = p;
        if (p > 0 && key_at(p - 1) == key) {  // inside a run of equal keys: lower bound
            unsigned lo = 0, hi = p;
            while (lo < hi) {
                unsigned mid = (lo + hi) >> 1;
                if (key_at(mid) < key) lo = mid + 1; else hi = mid;
            }
            less = lo;
        }
        double num = (double)less;
        if (correction != 0.0) {
            unsigned up = p + 1;
            if (up < n && key_at(up) == key) {  // upper bound
                unsigned lo = up, hi = n;
                while (lo < hi) {
                    unsigned mid = (lo + hi) >> 1;
                    if (key_at(mid) <= key) lo = mid + 1; else hi = mid;
                }
                up = lo;
            }
            num = __dadd_rn(num, __dmul_rn(correction, (double)(up - less)));
        }
        o[(unsigned)kv] = (float)__ddiv_rn(num, (double)n);
    }
}

}  // namespace vx

using namespace vx;

extern "C" int32_t vx_percentiles(vx_ctx ctx, vx_img a, vx_img mask, double correction, vx_img* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    OpGuard og;
    VX_TRY(og.begin(ctx));
    Image *A = nullptr, *M = nullptr;
    VX_TRY(og.input(a, &A, VX_F32));
    VX_TRY(og.input(mask, &M, VX_U8));
    VX_TRY(same_shape(A, M));
    const size_t nvox = A->nvox(), total = A->count();
    const int nb = A->nb;
    const int nchunks_max = (int)((nvox + kWarpChunk - 1) / kWarpChunk);
    unsigned long long* buf = nullptr;
    unsigned *count = nullptr, *table = nullptr, *totals = nullptr, *plan = nullptr;
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned long long) * total * 2, (void**)&buf));
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * (nb + 4), (void**)&count));
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * (size_t)nb * kDigits * nchunks_max, (void**)&table));
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
    dim3 sgrid((nchunks_max + kPcThreads / 32 - 1) / (kPcThreads / 32), nb);
    for (int pass = 0; pass < kMaxPasses && rc == VX_OK; pass++) {
        { VX_LAUNCH(og.c, og.s, "rs_hist_kernel");
          rs_hist_kernel<<<sgrid, kPcThreads, 0, og.st>>>(k0, k1, count, nvox, table, nchunks_max, pass, plan); }
        rc = check_launch("rs_hist_kernel");
        if (rc != VX_OK) break;
        { VX_LAUNCH(og.c, og.s, "rs_rowscan_kernel");
          rs_rowscan_kernel<<<dim3(kDigits, nb), 256, 0, og.st>>>(count, table, totals, nchunks_max, pass, plan); }
        rc = check_launch("rs_rowscan_kernel");
        if (rc != VX_OK) break;
        { VX_LAUNCH(og.c, og.s, "rs_scatter_kernel");
          rs_scatter_kernel<<<sgrid, kPcThreads, 0, og.st>>>(k0, k1, count, nvox, table, totals, nchunks_max, pass, plan); }
        rc = check_launch("rs_scatter_kernel");
    }
    if (rc == VX_OK) {
        // one block per 256 sorted positions, volume-major block order: the scattered 4-byte
        // stores of one volume land while its 36 MB output is L2-resident, instead of
        // spreading over the whole batch (write amplification at batch 16: 2x slower)
        int per_vol = (int)((nvox + kPcThreads - 1) / kPcThreads);
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
