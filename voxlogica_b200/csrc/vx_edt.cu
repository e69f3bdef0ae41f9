// A5 distance operators (Greta.pdf p.37 "distance modality",
// p.43 D^I): exact Euclidean distance transform in physical units and the fused
// distance-interval tests behind distlt / distleq / distgeq / distgt.
//
// Arithmetic contract (identical in oracle/vx_oracle.c, so results are bit-exact):
// the squared distance is built separably in IEEE binary32, one rounding per step,
//      g1 = fl(fl(dx*sx)^2),  g2 = min_j fl(g1[j] + fl(fl(dy*sy)^2)),
//      g3 = min_j fl(g2[j] + fl(fl(dz*sz)^2)),  d = sqrt_rn(g3)
// (exact integers for unit spacing below 2^24).  fl(. + t) is monotone, so g3 equals
// the minimum over all feature voxels of the same expression evaluated per offset.
//
// Two device paths:
//  * vx_edt: three axis passes.  The x pass works on ballot bit masks; the y and z
//    passes scan outwards from each voxel with exact early termination and visit only
//    rows/planes that contain a feature at all (tracked as bit sets), so sparse masks
//    cost O(#non-empty rows) and dense masks O(sqrt(result)).
//  * vx_dist_*: "sqrt(g3) OP r" is a monotone step function of g3, hence equivalent to
//    "g3 < T" for a threshold T found on the host; that in turn is a binary dilation
//    by the exact discrete ball {offset : g3(offset) < T}, which is evaluated in the
//    bit domain: pack 32 voxels per word, pre-dilate rows in x for every extent that
//    occurs, then OR the <= 441 (dy,dz) rows per output word.  1 byte read + 1 byte
//    written per voxel; the f32 map is never materialised.
#include <math.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "vx_bits.h"
#include "vx_internal.h"

namespace vx {

constexpr int kEdtThreads = 256;
constexpr int kMaxRowWords = 128;  // x pass supports nx <= 4096
constexpr int kMaxBallEntries = 441;
constexpr int kFeatSlots = 64, kFeatStride = 4;   // feature counters: 64 u64, 32 bytes apart

// ------------------------------------------------------------------ bit helpers
__device__ __forceinline__ int prev_set(const unsigned* __restrict__ fl, int pos) {
    if (pos < 0) return -1;
    int w = pos >> 5;
    unsigned m = __ldg(fl + w) & (0xffffffffu >> (31 - (pos & 31)));
    while (!m) {
        if (--w < 0) return -1;
        m = __ldg(fl + w);
    }
    return w * 32 + 31 - __clz(m);
}
__device__ __forceinline__ int next_set(const unsigned* __restrict__ fl, int pos, int n, int nw) {
    if (pos >= n) return n;
    int w = pos >> 5;
    unsigned m = __ldg(fl + w) & (0xffffffffu << (pos & 31));
    while (!m) {
        if (++w >= nw) return n;
        m = __ldg(fl + w);
    }
    return w * 32 + __ffs(m) - 1;
}

// ------------------------------------------------------------------ x pass
// one warp per row.  g1 = (dx*sx)^2 to the nearest feature in the row, +inf if none.
__global__ void __launch_bounds__(kEdtThreads)
edt_x_kernel(const unsigned* __restrict__ img, float* __restrict__ g1, unsigned* __restrict__ rowflag,
             int nx, int ny, int nz, unsigned long long rows, float sx, unsigned long long* __restrict__ nfeat) {
    __shared__ unsigned sm[kEdtThreads / 32][kMaxRowWords];
    __shared__ int sl_[kEdtThreads / 32][kMaxRowWords], sr_[kEdtThreads / 32][kMaxRowWords];
    __shared__ unsigned s_cnt;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned long long row = (unsigned long long)blockIdx.x * (kEdtThreads / 32) + warp;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const bool ok = row < rows;
    const int nwx = (nx + 31) >> 5;
    // the row's words straight from the image's bit form, one per lane
    unsigned any = 0, cnt = 0;
    for (int c0 = 0; ok && c0 < nwx; c0 += 32) {
        const int c = c0 + lane;
        const unsigned m = c < nwx ? flat_row_word(img, row, c, nx) : 0u;
        if (c < nwx) sm[warp][c] = m;
        any |= m;
        cnt += __popc(m);
    }
    any = __reduce_or_sync(0xffffffffu, any);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    // Feature voxels of the batch: one atomic per BLOCK, spread over kFeatSlots counters in different
    // sectors; the plane flags are made from the row flags afterwards (edt_mode_kernel).  Anything that
    // every row sends to ONE address -- the counter, the 16 plane words of a 512^3 volume, even a plain
    // L2 load of them -- queues up in one L2 slice: that, not the arithmetic, made this pass 0.36 ms
    // (512^3) / 1.04 ms (16 BraTS volumes) long.
    if (lane == 0 && cnt && nfeat != nullptr) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt && nfeat != nullptr)
        atomicAdd(nfeat + (blockIdx.x % kFeatSlots) * kFeatStride, (unsigned long long)s_cnt);
    if (!ok) return;
    if (lane == 0 && any) {
        const int y = (int)(row % ny);
        const unsigned long long pz = row / ny;  // vol*nz + z
        atomicOr(rowflag + pz * ((ny + 31) >> 5) + (y >> 5), 1u << (y & 31));
    }
    float* dst = g1 + row * nx;
    if (!any) {   // no feature in the row
        for (int x = lane; x < nx; x += 32) dst[x] = INFINITY;
        return;
    }
    // Nearest feature to the left / right of every WORD, by a max / min scan over the words (one word per
    // lane, 32 words per round): sl[c] = position of the last feature before word c (-1: none), sr[c] =
    // position of the first feature after word c (-1: none).  A voxel then needs one clz / ffs on its own
    // word and these two numbers -- no search loop (the loops over the row words cost 50 instructions per
    // voxel and made this pass 1.0 ms per 16 BraTS volumes).
    const unsigned* bw = sm[warp];
    int* sl = sl_[warp];
    int* sr = sr_[warp];
    int carry = -1;
    for (int c0 = 0; c0 < nwx; c0 += 32) {
        const int c = c0 + lane;
        const unsigned m = c < nwx ? bw[c] : 0u;
        int v = m ? c * 32 + 31 - __clz(m) : -1;   // last feature of this word
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v = max(v, t);
        }
        int excl = __shfl_up_sync(0xffffffffu, v, 1);
        if (lane == 0) excl = -1;
        if (c < nwx) sl[c] = max(carry, excl);
        carry = max(carry, __shfl_sync(0xffffffffu, v, 31));
    }
    carry = 0x7fffffff;
    for (int c0 = ((nwx - 1) / 32) * 32; c0 >= 0; c0 -= 32) {
        const int c = c0 + lane;
        const unsigned m = c < nwx ? bw[c] : 0u;
        int v = m ? c * 32 + __ffs(m) - 1 : 0x7fffffff;   // first feature of this word
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_down_sync(0xffffffffu, v, o);
            if (lane + o < 32) v = min(v, t);
        }
        int excl = __shfl_down_sync(0xffffffffu, v, 1);
        if (lane == 31) excl = 0x7fffffff;
        const int r = min(carry, excl);
        if (c < nwx) sr[c] = r == 0x7fffffff ? -1 : r;
        carry = min(carry, __shfl_sync(0xffffffffu, v, 0));
    }
    __syncwarp();
    // four voxels per thread: the distance to the left at the first and to the right at the last come from
    // the word (one clz / ffs) or the scans, the ones in between by +1 / reset; one 16-byte store.
    // (One voxel per thread and step was 20-30 instructions per 4 bytes written: issue-bound at 1.5 TB/s.)
    constexpr int BIG = 0x3fffffff;
    const bool vec = (nx & 3) == 0;
    for (int t = lane; 4 * t < nx; t += 32) {
        const int c = t >> 3, b0 = (t & 7) * 4, x0 = 4 * t;
        const unsigned w = bw[c];
        const unsigned ml = w & (0xffffffffu >> (31 - b0)), mr = w & (0xffffffffu << (b0 + 3));
        const int lpos = ml ? c * 32 + 31 - __clz(ml) : sl[c];
        const int rpos = mr ? c * 32 + __ffs(mr) - 1 : sr[c];
        const unsigned nib = w >> b0;
        const int l0 = lpos >= 0 ? x0 - lpos : BIG;
        const int l1 = (nib & 2u) ? 0 : l0 + 1, l2 = (nib & 4u) ? 0 : l1 + 1, l3 = (nib & 8u) ? 0 : l2 + 1;
        const int r3 = rpos >= 0 ? rpos - (x0 + 3) : BIG;
        const int r2 = (nib & 4u) ? 0 : r3 + 1, r1 = (nib & 2u) ? 0 : r2 + 1, r0 = (nib & 1u) ? 0 : r1 + 1;
        float4 v;
        float q;
        q = __fmul_rn((float)min(l0, r0), sx); v.x = __fmul_rn(q, q);
        q = __fmul_rn((float)min(l1, r1), sx); v.y = __fmul_rn(q, q);
        q = __fmul_rn((float)min(l2, r2), sx); v.z = __fmul_rn(q, q);
        q = __fmul_rn((float)min(l3, r3), sx); v.w = __fmul_rn(q, q);
        if (vec) {
            *reinterpret_cast<float4*>(dst + x0) = v;   // row * nx and x0 are multiples of 4
        } else {
            dst[x0] = v.x;
            if (x0 + 1 < nx) dst[x0 + 1] = v.y;
            if (x0 + 2 < nx) dst[x0 + 2] = v.z;
            if (x0 + 3 < nx) dst[x0 + 3] = v.w;
        }
    }
}

// ------------------------------------------------------------------ y / z passes
enum { EDT_MID = 0, EDT_SQRT = 1, EDT_THRESH = 2 };

// AXIS 1: along y, flags = rows of plane (vol,z).  AXIS 2: along z, flags = planes of vol.
template <int AXIS, int MODE>
__global__ void __launch_bounds__(kEdtThreads)
edt_axis_kernel(const float* __restrict__ gin, void* __restrict__ gout_,
                const unsigned* __restrict__ flags, int nx, int ny, int nz, size_t total, float s,
                float T, int invert, const uint8_t* __restrict__ only, const int* __restrict__ windowed) {
    const size_t idx = (size_t)blockIdx.x * kEdtThreads + threadIdx.x;
    if (idx >= total) return;
    const int x = (int)(idx % nx);
    size_t t0 = idx / nx;
    const int y = (int)(t0 % ny);
    t0 /= ny;
    const int z = (int)(t0 % nz);
    const size_t vol = t0 / nz;
    // `only`: the windowed pass settled every column whose flag is 0 (column numbering as in edt_env_kernel)
    if (only != nullptr && *windowed && !only[AXIS == 1 ? (vol * nz + z) * nx + x : (vol * ny + y) * (size_t)nx + x]) return;
    int i, n;
    size_t stride;
    const float* col;
    const unsigned* fl;
    if (AXIS == 1) {
        i = y; n = ny; stride = nx;
        col = gin + ((vol * nz + z) * ny) * (size_t)nx + x;
        fl = flags + (vol * nz + z) * ((ny + 31) >> 5);
    } else {
        i = z; n = nz; stride = (size_t)nx * ny;
        col = gin + (vol * nz) * stride + (size_t)y * nx + x;
        fl = flags + vol * ((nz + 31) >> 5);
    }
    const int nw = (n + 31) >> 5;
    float best = __ldg(col + (size_t)i * stride);
    int jl = prev_set(fl, i - 1), jh = next_set(fl, i + 1, n, nw);
    while (true) {
        const int dl = jl >= 0 ? i - jl : 0x7fffffff;
        const int dh = jh < n ? jh - i : 0x7fffffff;
        const bool low = dl <= dh;
        const int k = low ? dl : dh;
        if (k == 0x7fffffff) break;
        float t = __fmul_rn((float)k, s);
        t = __fmul_rn(t, t);
        if (!(t < best)) break;
        const int j = low ? jl : jh;
        best = fminf(best, __fadd_rn(__ldg(col + (size_t)j * stride), t));
        if (low) jl = prev_set(fl, jl - 1);
        else jh = next_set(fl, jh + 1, n, nw);
    }
    if (MODE == EDT_MID) reinterpret_cast<float*>(gout_)[idx] = best;
    else if (MODE == EDT_SQRT) reinterpret_cast<float*>(gout_)[idx] = __fsqrt_rn(best);
    else reinterpret_cast<uint8_t*>(gout_)[idx] = (uint8_t)((best < T ? 1 : 0) ^ invert);
}

// ------------------------------------------------------------------ lower-envelope y / z passes
// Maurer / Meijster style separable pass: out[i] = min_j ( g[j] + s2 * (i - j)^2 ) computed in
// O(n) per column with a stack of the parabolas on the lower envelope (partial Voronoi
// diagram of the column), one THREAD per column so that a warp reads/writes 32 consecutive x.
// Only used when the arithmetic is exact in integers (integer spacings, every squared
// distance below 2^24): then binary32 is exact too and the result equals the arithmetic
// contract in the header bit for bit.  The stack lives in a scratch array laid out
// [depth][column] (coalesced while neighbouring columns have similar depths); its top entry
// stays in registers, so a step that neither pops nor pushes touches no memory but g[u].
//   entry = (s: position of the parabola, t: first index where it is the minimum, g: its height)
__device__ __forceinline__ uint2 env_pack(int s_, int t_, int g_) { return make_uint2((unsigned)s_ | ((unsigned)t_ << 16), (unsigned)g_); }

template <int AXIS, int MODE>
__global__ void __launch_bounds__(128)
edt_env_kernel(const float* __restrict__ gin, void* __restrict__ gout_, uint2* __restrict__ stack,
               int nx, int ny, int nz, size_t ncols, int s2, float T, int invert, const uint8_t* __restrict__ only,
               const int* __restrict__ windowed) {
    const size_t col = (size_t)blockIdx.x * 128 + threadIdx.x;
    if (col >= ncols) return;
    if (only != nullptr && *windowed && !only[col]) return;   // the windowed pass settled this column
    size_t base, stride;
    int n;
    if (AXIS == 1) {   // col = (vol*nz + z) * nx + x
        const size_t pz = col / nx;
        base = pz * (size_t)ny * nx + (col - pz * nx);
        stride = nx; n = ny;
    } else {           // col = vol * (ny*nx) + y*nx + x
        const size_t plane = (size_t)ny * nx;
        const size_t vol = col / plane;
        base = vol * plane * nz + (col - vol * plane);
        stride = plane; n = nz;
    }
    uint2* st = stack + col;
    // Entry q (the top) lives in registers, and so do copies of the three entries below it
    // (w0 = q-1, w1 = q-2, w2 = q-3): a pop takes its new top from w0 at once and issues the
    // load that refills w2 three pops before it is needed, so runs of pops overlap their L2
    // latencies instead of paying them one after the other (ncu: the dependent entry loads
    // were 75% of all stall samples).  Every entry is also written to the scratch when it is
    // pushed, so the scratch always holds entries [0, q).
    int q = -1, ts = 0, tt = 0, tg = 0;
    uint2 w0 = make_uint2(0, 0), w1 = w0, w2 = w0;
    constexpr int PF = 8;
    const float* gp = gin + base;
    for (int u0 = 0; u0 < n; u0 += PF) {
        float gv[PF];
#pragma unroll
        for (int e = 0; e < PF; e++) gv[e] = (u0 + e < n) ? __ldg(gp + (size_t)e * stride) : INFINITY;
        gp += (size_t)PF * stride;
#pragma unroll
        for (int e = 0; e < PF; e++) {
            if (!(gv[e] < INFINITY)) continue;   // no feature reaches this cell from the previous axes
            const int u = u0 + e, gu = (int)gv[e];
            if (q < 0) { q = 0; ts = u; tt = 0; tg = gu; continue; }
            // drop the parabolas that the new one beats at the start of their own interval
            bool only = false;
            while (true) {
                const int d1 = tt - ts, d2 = tt - u;
                if (tg + s2 * d1 * d1 > gu + s2 * d2 * d2) {
                    if (q == 0) { only = true; break; }
                    q--;
                    ts = (int)(w0.x & 0xffffu); tt = (int)(w0.x >> 16); tg = (int)w0.y;
                    w0 = w1; w1 = w2;
                    if (q >= 3) w2 = st[(size_t)(q - 3) * ncols];
                } else break;
            }
            if (only) { ts = u; tt = 0; tg = gu; continue; }   // the new parabola is the only one left
            // first index where parabola u is strictly below parabola ts
            // floor(num / den) without the ~40-instruction integer division: binary32 estimate,
            // then an exact integer correction (|num| < 2^31, den > 0)
            const int num = s2 * (u * u - ts * ts) + gu - tg, den = 2 * s2 * (u - ts);
            int sep = __float2int_rd(__fdividef((float)num, (float)den));
            while (sep * den > num) sep--;
            while ((sep + 1) * den <= num) sep++;
            const int w = sep + 1;
            if (w < n) {
                const uint2 en = env_pack(ts, tt, tg);
                st[(size_t)q * ncols] = en;
                w2 = w1; w1 = w0; w0 = en;
                q++; ts = u; tt = w; tg = gu;
            }
        }
    }
    float* of = reinterpret_cast<float*>(gout_);
    uint8_t* ob = reinterpret_cast<uint8_t*>(gout_);
    // Read-out: the entries are sorted by the start of their interval, so the column is swept
    // upwards with the NEXT four entries already in flight (their addresses do not depend on
    // data) -- no dependent load on the critical path, and all lanes write row u together.
    if (q >= 0) st[(size_t)q * ncols] = env_pack(ts, tt, tg);
    const uint2 none = make_uint2(0xffff0000u, 0u);   // interval start 65535: never reached (n <= 65535)
    uint2 cur = q >= 0 ? st[0] : none;
    uint2 f0 = q >= 1 ? st[(size_t)1 * ncols] : none, f1 = q >= 2 ? st[(size_t)2 * ncols] : none;
    uint2 f2 = q >= 3 ? st[(size_t)3 * ncols] : none, f3 = q >= 4 ? st[(size_t)4 * ncols] : none;
    int k = 0;
    size_t o = base;
    for (int u = 0; u < n; u++) {
        if ((unsigned)u == (f0.x >> 16)) {
            cur = f0; f0 = f1; f1 = f2; f2 = f3;
            k++;
            f3 = (k + 4 <= q) ? st[(size_t)(k + 4) * ncols] : none;
        }
        float best = INFINITY;
        if (q >= 0) {
            const int d = u - (int)(cur.x & 0xffffu);
            best = (float)((int)cur.y + s2 * d * d);
        }
        if (MODE == EDT_MID) of[o] = best;
        else if (MODE == EDT_SQRT) of[o] = __fsqrt_rn(best);
        else ob[o] = (uint8_t)((best < T ? 1 : 0) ^ invert);
        o += stride;
    }
}

// ------------------------------------------------------------------ windowed y / z pass
// out[i] = min_j fl(g[j] + fl(fl((i-j) s)^2)) only needs the rows within sqrt(out[i]) / s of i, and
// g[i] itself is an upper bound: next to a dense feature set (tissue masks) the window is a few
// rows wide.  A block stages a tile of 32 columns x kWinT positions plus H halo rows on either
// side in shared memory (coalesced 128-byte row segments) and every thread scans outwards from
// its positions until the next offset alone is no smaller than the best value so far -- the same
// termination rule and the same arithmetic as edt_axis_kernel, so the same bits.  Traffic is the
// column itself (+ halo): no stack, no scratch.
// Two levels: H = 16 on every tile, then H = 64 on the tiles that held a position whose scan
// reached the end of the halo undecided (`tileflags`).  A tile whose smallest staged value already
// exceeds what a scan of H rows can settle is not scanned at all (features far away: the normal
// case far from the features).  What the second level cannot settle marks its column
// `unsettled`; those columns are redone by the envelope / outward-scan kernels.
// The scan pays off where features are dense: measured at 512^3, blobs (60 % set) 4.74 -> 3.3 ms
// and Bernoulli 5 % 3.2 -> 1.7 ms for the whole transform, but 1e-4 density 1.3 -> 5.3 ms (every
// tile is scanned to the end of both halos and still unsettled).  So the x pass counts the feature
// voxels of the batch and a one-thread kernel turns that into `windowed[0]`: below 2 % the windowed
// kernels exit at once and the envelope kernels do every column -- decided on the device.
// also: the plane flags (a plane holds a feature) from the row flags, one thread per plane
__global__ void __launch_bounds__(256)
edt_mode_kernel(const unsigned long long* __restrict__ nfeat, unsigned long long total, int* __restrict__ windowed,
                const unsigned* __restrict__ rowflag, unsigned* __restrict__ planeflag, int ny, int nz, int nb) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t == 0) {
        unsigned long long n = 0;
        for (int i = 0; i < kFeatSlots; i++) n += nfeat[i * kFeatStride];
        windowed[0] = n * 50ull >= total ? 1 : 0;
    }
    if (t >= nb * nz) return;
    const int nyw = (ny + 31) >> 5;
    unsigned any = 0;
    for (int w = 0; w < nyw; w++) any |= rowflag[(size_t)t * nyw + w];
    if (any) {
        const int vol = t / nz, z = t - vol * nz;
        atomicOr(planeflag + (size_t)vol * ((nz + 31) >> 5) + (z >> 5), 1u << (z & 31));
    }
}

constexpr int kWinT = 64;    // positions per tile
constexpr int kWinWarps = 8;

// grid (x tiles, chunks along the axis, planes): AXIS 1: plane = vol * nz + z, axis y; AXIS 2: plane = vol * ny + y, axis z
// LEVEL 0: every tile, sets tileflags.  LEVEL 1: flagged tiles only, sets the column flags.
template <int AXIS, int MODE, int H, int LEVEL>
__global__ void __launch_bounds__(32 * kWinWarps)
edt_window_kernel(const float* __restrict__ gin, void* __restrict__ gout_, uint8_t* __restrict__ tileflags,
                  uint8_t* __restrict__ unsettled, int nx, int ny, int nz, float s, float T, int invert,
                  const int* __restrict__ windowed) {
    __shared__ float tile[kWinT + 2 * H][32];
    __shared__ float tk[H + 2];
    __shared__ float wmin[kWinWarps];
    if (!*windowed) return;
    const size_t tile_id = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (LEVEL == 1 && !tileflags[tile_id]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 32 + lane;
    const int c0 = blockIdx.y * kWinT;
    const size_t plane = blockIdx.z;
    size_t base, stride;
    int n;
    if (AXIS == 1) {
        base = plane * (size_t)ny * nx + x; stride = nx; n = ny;
    } else {
        const size_t vol = plane / ny, y = plane - vol * ny;
        base = (vol * nz * (size_t)ny + y) * nx + x; stride = (size_t)nx * ny; n = nz;
    }
    const size_t col = plane * nx + x;
    const bool live = x < nx;
    if (threadIdx.x <= H + 1) {   // squared offsets, rounded as the contract says
        const float t = __fmul_rn((float)threadIdx.x, s);
        tk[threadIdx.x] = __fmul_rn(t, t);
    }
    // stage rows c0 - H .. c0 + T + H - 1 (rows outside the column hold +inf: they never win)
    float lo = INFINITY;
    for (int r = warp; r < kWinT + 2 * H; r += kWinWarps) {
        const int pos = c0 - H + r;
        const float v = (live && pos >= 0 && pos < n) ? __ldg(gin + base + (size_t)pos * stride) : INFINITY;
        tile[r][lane] = v;
        lo = fminf(lo, v);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    if (lane == 0) wmin[warp] = lo;
    __syncthreads();
    lo = wmin[0];
#pragma unroll
    for (int w = 1; w < kWinWarps; w++) lo = fminf(lo, wmin[w]);
    const bool whole = c0 - H <= 0 && c0 + kWinT + H >= n;   // the tile holds the whole column
    // nothing in reach can give a value that a scan of H rows settles: leave the tile to the next level
    if (!whole && !(lo <= tk[H + 1])) {
        if (LEVEL == 0) { if (threadIdx.x == 0) tileflags[tile_id] = 1; }
        else if (live) unsettled[col] = 1;
        return;
    }
    float* of = reinterpret_cast<float*>(gout_);
    uint8_t* ob = reinterpret_cast<uint8_t*>(gout_);
    bool unsure = false;
    constexpr int PER = kWinT / kWinWarps;
    if (live) {
#pragma unroll 1
        for (int e = 0; e < PER; e++) {
            const int li = warp * PER + e, pos = c0 + li;
            if (pos >= n) break;
            const int r = li + H;
            float best = tile[r][lane];
            int k = 1;
            for (; k <= H; k++) {
                const float t = tk[k];
                if (!(t < best)) break;
                best = fminf(best, fminf(__fadd_rn(tile[r - k][lane], t), __fadd_rn(tile[r + k][lane], t)));
            }
            // halo exhausted with rows left unseen and the next offset still below the best value
            if (k > H && (pos - H > 0 || pos + H < n - 1) && tk[H + 1] < best) unsure = true;
            const size_t o = base + (size_t)pos * stride;
            if (MODE == EDT_MID) of[o] = best;
            else if (MODE == EDT_SQRT) of[o] = __fsqrt_rn(best);
            else ob[o] = (uint8_t)((best < T ? 1 : 0) ^ invert);
        }
    }
    if (LEVEL == 0) {
        if (__syncthreads_or(unsure) && threadIdx.x == 0) tileflags[tile_id] = 1;
    } else if (unsure) {
        unsettled[col] = 1;
    }
}

// ------------------------------------------------------------------ bit-domain ball dilation
// pack: one thread per 32-voxel word of a row
__device__ __forceinline__ unsigned nib_of(unsigned w) {
    w = __vcmpne4(w, 0u) & 0x01010101u;
    return ((w * 0x01020408u) >> 24) & 0xfu;
}
__device__ __forceinline__ unsigned bits16(uint4 v) {
    return nib_of(v.x) | (nib_of(v.y) << 4) | (nib_of(v.z) << 8) | (nib_of(v.w) << 12);
}
__device__ __forceinline__ unsigned pack_word(const uint8_t* __restrict__ img, unsigned long long row, int w, int nx) {
    const int x0 = w * 32;
    const uint8_t* src = img + row * nx + x0;
    unsigned m = 0;
    if ((nx & 15) == 0) {
        m = bits16(__ldg(reinterpret_cast<const uint4*>(src)));
        if (x0 + 16 < nx) m |= bits16(__ldg(reinterpret_cast<const uint4*>(src + 16))) << 16;
    } else {
        const int lim = min(32, nx - x0);
        for (int b = 0; b < lim; b++) m |= (unsigned)(src[b] != 0) << b;
    }
    return m;
}

// One thread per 32-voxel word: D[0][word] = the packed bits, D[e][word] = the row dilated by
// +-e voxels in x for e = 1..emax (the neighbouring words come from the neighbouring lanes;
// only the first and last lane of a warp pack theirs again), rowany[row] = 1 when the row
// holds a set voxel.
__global__ void __launch_bounds__(kEdtThreads)
pack_bits_kernel(const unsigned* __restrict__ img, unsigned* __restrict__ D, int nx, int nwx,
                 unsigned long long nwords, int emax, uint8_t* __restrict__ rowany) {
    const unsigned long long wid = (unsigned long long)blockIdx.x * kEdtThreads + threadIdx.x;
    const bool ok = wid < nwords;
    const unsigned long long row = (ok ? wid : 0) / nwx;
    const int w = (int)((ok ? wid : 0) % nwx);
    const unsigned cur = ok ? flat_row_word(img, row, w, nx) : 0u;
    unsigned prev = __shfl_up_sync(0xffffffffu, cur, 1), next = __shfl_down_sync(0xffffffffu, cur, 1);
    if (!ok) return;
    D[wid] = cur;
    if (cur != 0u && rowany != nullptr) rowany[row] = 1;   // zeroed by the caller; every writer stores 1
    if (emax <= 0) return;
    const int lane = threadIdx.x & 31;
    if (w == 0) prev = 0u;
    else if (lane == 0) prev = flat_row_word(img, row, w - 1, nx);
    if (w + 1 >= nwx) next = 0u;
    else if (lane == 31) next = flat_row_word(img, row, w + 1, nx);
    unsigned acc = cur;
    for (int e = 1; e <= emax; e++) {
        acc |= (cur << e) | (prev >> (32 - e)) | (cur >> e) | (next << (32 - e));
        D[(size_t)e * nwords + wid] = acc;
    }
}

struct Ball {
    int n;
    int8_t dy[kMaxBallEntries], dz[kMaxBallEntries], e[kMaxBallEntries];
};

__device__ __forceinline__ unsigned bytes_of(unsigned nib) {
    return (nib * 0x00204081u) & 0x01010101u;
}

// one thread per output word: OR the pre-dilated neighbour rows, unpack to 0/1 bytes.
// The kernel is issue-bound (ncu: 78% issue active, DRAM 4%), so the inner loop is kept to
// load-offset / add / load / or: entry offsets (e*nwords + (dz*ny+dy)*nwx) are precomputed on
// the host, and words whose whole ball lies inside the volume skip the bounds tests.
__global__ void __launch_bounds__(kEdtThreads)
ball_or_kernel(const unsigned* __restrict__ D, unsigned* __restrict__ out, int nx, int ny, int nz,
               int nwx, unsigned long long nwords, const long long* __restrict__ boff,
               const int* __restrict__ bdyz, int nball, int wy, int wz, int invert,
               const uint8_t* __restrict__ rowany, int align16) {
    // Rows of the input that hold no set voxel are flagged by the packing kernel.  When no row
    // that any word of this block can reach is occupied, the block writes its (constant) output
    // without touching the bit rows: the dilation half of `filter` runs on what an erosion left
    // over, which is empty almost everywhere.  The reachable rows are tested conservatively as
    // intervals of the linear row index (row + dz*ny + dy), which may include a few rows of the
    // neighbouring planes.
    bool block_empty = false;
    if (rowany != nullptr) {
        const unsigned long long w_first = (unsigned long long)blockIdx.x * kEdtThreads;
        const unsigned long long w_last = min(w_first + kEdtThreads, nwords) - 1;
        const long long r0 = (long long)(w_first / nwx) - wy, r1 = (long long)(w_last / nwx) + wy;
        const long long rows_total = (long long)(nwords / nwx);
        const int span = (int)(r1 - r0 + 1);
        int found = 0;
        for (int i = threadIdx.x; i < span * (2 * wz + 1); i += kEdtThreads) {
            const long long r = r0 + (i % span) + (long long)(i / span - wz) * ny;
            if (r >= 0 && r < rows_total && rowany[r]) found = 1;
        }
        block_empty = !__syncthreads_or(found);
    }
    const unsigned long long wid = (unsigned long long)blockIdx.x * kEdtThreads + threadIdx.x;
    if (wid >= nwords) return;
    const unsigned long long row = wid / nwx;
    const int w = (int)(wid % nwx);
    const int y = (int)(row % ny);
    const int z = (int)((row / ny) % nz);
    const unsigned* base = D + wid;
    // bits beyond the end of the row count as set for the saturation test (they are never
    // stored): otherwise the last word of every row -- four lanes of each warp at nx = 240 --
    // can never fill up and keeps its whole warp in the loop
    const unsigned pad = (w + 1) * 32 > nx ? 0xffffffffu << (nx - w * 32) : 0u;
    unsigned acc = pad;
    if (block_empty) {
        // nothing within reach
    } else if (y >= wy && y + wy < ny && z >= wz && z + wz < nz) {
        int t = 0;
        // entries come centre-first: next to a dense feature set (the erosion half of
        // `filter`: distgeq(r, !a)) the word is full after a few of them and the rest is skipped
        for (; t + 4 <= nball && acc != 0xffffffffu; t += 4) {
            const unsigned a0 = __ldg(base + __ldg(boff + t)), a1 = __ldg(base + __ldg(boff + t + 1));
            const unsigned a2 = __ldg(base + __ldg(boff + t + 2)), a3 = __ldg(base + __ldg(boff + t + 3));
            acc |= (a0 | a1) | (a2 | a3);
        }
        for (; t < nball && acc != 0xffffffffu; t++) acc |= __ldg(base + __ldg(boff + t));
    } else {
        for (int t = 0; t < nball && acc != 0xffffffffu; t++) {
            const int dyz = __ldg(bdyz + t);
            const int yy = y + (int)(signed char)(dyz & 0xff), zz = z + (int)(signed char)((dyz >> 8) & 0xff);
            if (yy < 0 || yy >= ny || zz < 0 || zz >= nz) continue;
            acc |= __ldg(base + __ldg(boff + t));
        }
    }
    acc &= ~pad;
    if (invert) acc = ~acc & ~pad;
    flat_store_row_word(out, row, w, nx, acc, align16 != 0);   // the result stays in bit form
}

// ------------------------------------------------------------------ host: thresholds and the ball
// smallest binary32 T such that sqrtf(T) >= r  (strict = false)  or  sqrtf(T) > r  (strict = true):
//   d <  r  <=>  g3 < T(strict=false);     d <= r  <=>  g3 < T(strict=true)
static float threshold_for(float r, bool strict) {
    auto ok = [&](float t) { float s = sqrtf(t); return strict ? s > r : s >= r; };
    if (ok(0.0f)) return 0.0f;                 // r <= 0 (or r < 0 when strict)
    // beyond sqrt(FLT_MAX) every finite squared distance passes (an empty feature set,
    // d = +inf, is reported as "not within r" even for r = +inf)
    if (!ok(INFINITY) || r > 1.8e19f) return INFINITY;
    double rr = (double)r * (double)r;
    float t = (float)rr;
    if (!(t > 0.0f)) t = nextafterf(0.0f, 1.0f);
    while (!ok(t)) t = nextafterf(t, INFINITY);
    while (t > 0.0f && ok(nextafterf(t, 0.0f))) t = nextafterf(t, 0.0f);
    return t;
}

static inline float sq_step(int d, float s) {
    volatile float t = (float)d * s;  // volatile: one rounding per operation, no contraction
    volatile float q = t * t;
    return q;
}
static inline float add_step(float a, float b) {
    volatile float r = a + b;
    return r;
}

// exact discrete ball {(dx,dy,dz) : g3(offset) < T} as per-(dy,dz) x extents
static bool build_ball(const float sp[3], float T, int nx, int ny, int nz, Ball* ball, int* emax) {
    ball->n = 0;
    *emax = 0;
    if (!(T > 0.0f)) return true;  // empty ball: nothing is closer than 0
    auto val = [&](int dx, int dy, int dz) {
        return add_step(add_step(sq_step(dx, sp[0]), sq_step(dy, sp[1])), sq_step(dz, sp[2]));
    };
    int wy = 0, wz = 0;
    while (wy + 1 < ny && val(0, wy + 1, 0) < T) { if (++wy > 120) return false; }
    while (wz + 1 < nz && val(0, 0, wz + 1) < T) { if (++wz > 120) return false; }
    for (int dz = -wz; dz <= wz; dz++)
        for (int dy = -wy; dy <= wy; dy++) {
            if (!(val(0, dy, dz) < T)) continue;
            int e = 0;
            while (e + 1 < nx && val(e + 1, dy, dz) < T) {
                if (++e > 31) return false;
            }
            if (ball->n >= kMaxBallEntries) return false;
            ball->dy[ball->n] = (int8_t)dy; ball->dz[ball->n] = (int8_t)dz; ball->e[ball->n] = (int8_t)e;
            ball->n++;
            if (e > *emax) *emax = e;
        }
    return true;
}

// One pass along y (AXIS 1) or z (AXIS 2): the windowed kernel first, then -- only on the columns
// it left unsettled -- the exact-integer envelope kernel (env_ok) or the outward scan.
// colflags: ncols bytes of scratch; stack: one uint2 per voxel when env_ok; rowflags: the emptiness
// flags the outward scan uses (rows of a plane for y, planes of a volume for z).
template <int AXIS, int MODE>
static int axis_pass(OpGuard& og, const float* in, void* out, const Image* A, bool env_ok, uint2* stack,
                     uint8_t* colflags, const unsigned* rowflags, float T, int invert, const int* mode) {
    const int nx = A->nx, ny = A->ny, nz = A->nz;
    const size_t total = A->count();
    const float s = AXIS == 1 ? A->sp[1] : A->sp[2];
    const int n = AXIS == 1 ? ny : nz;
    const size_t planes = AXIS == 1 ? (size_t)A->nb * nz : (size_t)A->nb * ny;
    const size_t ncols = planes * nx;
    const unsigned chunks = (unsigned)((n + kWinT - 1) / kWinT);
    // (grid.z holds at most 65535 planes: a batch beyond that goes to the fallback kernels whole)
    const bool windowed = mode != nullptr && planes <= 65535u;
    const uint8_t* only = windowed ? colflags : nullptr;
    if (windowed) VX_CUDA(cudaMemsetAsync(colflags, 0, ncols, og.st));
    if (windowed) {
        const dim3 grid((nx + 31) / 32, chunks, (unsigned)planes);
        uint8_t* tileflags = nullptr;
        Scratch sc(og.c, og.s);
        VX_TRY(sc.alloc((size_t)grid.x * grid.y * grid.z, &tileflags));
        VX_CUDA(cudaMemsetAsync(tileflags, 0, (size_t)grid.x * grid.y * grid.z, og.st));
        {
            VX_LAUNCH(og.c, og.s, AXIS == 1 ? "edt_window_kernel_y" : "edt_window_kernel_z");
            edt_window_kernel<AXIS, MODE, 16, 0><<<grid, 32 * kWinWarps, 0, og.st>>>(in, out, tileflags, colflags, nx, ny, nz, s, T, invert, mode);
        }
        VX_TRY(check_launch("edt_window_kernel"));
        {
            VX_LAUNCH(og.c, og.s, AXIS == 1 ? "edt_window2_kernel_y" : "edt_window2_kernel_z");
            edt_window_kernel<AXIS, MODE, 64, 1><<<grid, 32 * kWinWarps, 0, og.st>>>(in, out, tileflags, colflags, nx, ny, nz, s, T, invert, mode);
        }
    }
    VX_TRY(check_launch("edt_window_kernel"));
    if (env_ok) {
        VX_LAUNCH(og.c, og.s, AXIS == 1 ? "edt_env_kernel_y" : "edt_env_kernel_z");
        edt_env_kernel<AXIS, MODE><<<(unsigned)((ncols + 127) / 128), 128, 0, og.st>>>(in, out, stack, nx, ny, nz, ncols,
                                                                                     (int)(s * s), T, invert, only, mode);
    } else {
        VX_LAUNCH(og.c, og.s, AXIS == 1 ? "edt_axis_kernel_y" : "edt_axis_kernel_z");
        edt_axis_kernel<AXIS, MODE><<<(unsigned)((total + kEdtThreads - 1) / kEdtThreads), kEdtThreads, 0, og.st>>>(
            in, out, rowflags, nx, ny, nz, total, s, T, invert, only, mode);
    }
    return check_launch("edt fallback pass");
}

// squared-distance passes shared by vx_edt and the large-radius fallback of vx_dist_*
// x pass + flags: g1, the row flags, the plane flags and the dense / sparse decision.  `flags` is one scratch
// block: row flags | plane flags | (pad) | mode (2 words) | feature counters.
struct EdtFlags {
    unsigned* base = nullptr;
    unsigned *rowflag = nullptr, *planeflag = nullptr;
    int* mode = nullptr;
};
static int edt_x_pass(OpGuard& og, const Image* A, const unsigned* Abits, float* g1, EdtFlags* f) {
    const unsigned long long rows = (unsigned long long)A->nb * A->nz * A->ny;
    const size_t n_rowflag = (size_t)A->nb * A->nz * ((A->ny + 31) / 32);
    const size_t n_planeflag = (size_t)A->nb * ((A->nz + 31) / 32);
    const size_t n_flagwords = (n_rowflag + n_planeflag + 1) / 2 * 2;   // keeps the counters behind them 8-byte aligned
    const size_t words = n_flagwords + 2 + 2 * (size_t)kFeatSlots * kFeatStride;
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * words, (void**)&f->base));
    VX_CUDA(cudaMemsetAsync(f->base, 0, sizeof(unsigned) * words, og.st));
    f->rowflag = f->base;
    f->planeflag = f->base + n_rowflag;
    f->mode = reinterpret_cast<int*>(f->base + n_flagwords);
    unsigned long long* nfeat = reinterpret_cast<unsigned long long*>(f->base + n_flagwords + 2);
    {
        VX_LAUNCH(og.c, og.s, "edt_x_kernel");
        unsigned grid = (unsigned)((rows + kEdtThreads / 32 - 1) / (kEdtThreads / 32));
        edt_x_kernel<<<grid, kEdtThreads, 0, og.st>>>(Abits, g1, f->rowflag, A->nx, A->ny, A->nz, rows, A->sp[0], nfeat);
    }
    VX_TRY(check_launch("edt_x_kernel"));
    edt_mode_kernel<<<(A->nb * A->nz + 255) / 256, 256, 0, og.st>>>(nfeat, (unsigned long long)A->count(), f->mode, f->rowflag,
                                                                  f->planeflag, A->ny, A->nz, A->nb);
    return check_launch("edt_mode_kernel");
}

static int edt_passes(OpGuard& og, const Image* A, const unsigned* Abits, void* out, int final_mode, float T, int invert) {
    const size_t total = A->count();
    const unsigned long long rows = (unsigned long long)A->nb * A->nz * A->ny;
    VX_REQUIRE(A->nx <= kMaxRowWords * 32, VX_E_UNSUPPORTED, "vx_edt supports nx <= %d",
               kMaxRowWords * 32);
    float *g1 = nullptr, *g2 = nullptr;
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(float) * total, (void**)&g1));
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(float) * total, (void**)&g2));
    EdtFlags fl;
    VX_TRY(edt_x_pass(og, A, Abits, g1, &fl));
    unsigned* flags = fl.base;
    unsigned *rowflag = fl.rowflag, *planeflag = fl.planeflag;
    int* mode = fl.mode;
    // exact-integer fast path per axis: every spacing up to this axis is a small integer and no
    // squared distance can reach 2^24, so binary32 arithmetic is exact and the O(n)
    // lower-envelope pass yields the contract's value bit for bit
    auto int_spacing = [](float sp) { return sp >= 1.0f && sp <= 64.0f && sp == floorf(sp); };
    auto span2 = [](float sp, int n) { double d = (double)sp * (n - 1); return d * d; };
    const bool ix = int_spacing(A->sp[0]), iy = int_spacing(A->sp[1]), iz = int_spacing(A->sp[2]);
    const double bound_y = span2(A->sp[0], A->nx) + span2(A->sp[1], A->ny);
    const double bound_z = bound_y + span2(A->sp[2], A->nz);
    const bool env_y = ix && iy && bound_y < 16777216.0 && A->ny <= 65535 && A->ny > 1;
    const bool env_z = ix && iy && iz && bound_z < 16777216.0 && A->nz <= 65535 && A->nz > 1;
    uint2* stack = nullptr;
    uint8_t* colflags = nullptr;
    if (env_y || env_z) VX_TRY(scratch_alloc(og.c, og.s, sizeof(uint2) * total, (void**)&stack));
    VX_TRY(scratch_alloc(og.c, og.s, (size_t)A->nb * std::max(A->nz, A->ny) * A->nx, (void**)&colflags));
    VX_TRY((axis_pass<1, EDT_MID>(og, g1, g2, A, env_y, stack, colflags, rowflag, 0.f, 0, mode)));
    if (final_mode == EDT_SQRT) VX_TRY((axis_pass<2, EDT_SQRT>(og, g2, out, A, env_z, stack, colflags, planeflag, 0.f, 0, mode)));
    else VX_TRY((axis_pass<2, EDT_THRESH>(og, g2, out, A, env_z, stack, colflags, planeflag, T, invert, mode)));
    VX_TRY(scratch_free(og.c, og.s, colflags));
    if (stack) VX_TRY(scratch_free(og.c, og.s, stack));
    VX_TRY(scratch_free(og.c, og.s, g1));
    VX_TRY(scratch_free(og.c, og.s, g2));
    return scratch_free(og.c, og.s, flags);
}

// ---- the two halves of the transform, for a caller that owns one z-slab ----
// x and y passes only: squared distance inside each plane (+inf where the plane has no feature)
static int edt_xy_only(OpGuard& og, const Image* A, const unsigned* Abits, float* out) {
    const size_t total = A->count();
    const unsigned long long rows = (unsigned long long)A->nb * A->nz * A->ny;
    VX_REQUIRE(A->nx <= kMaxRowWords * 32, VX_E_UNSUPPORTED, "vx_edt supports nx <= %d", kMaxRowWords * 32);
    float* g1 = nullptr;
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(float) * total, (void**)&g1));
    EdtFlags fl;
    VX_TRY(edt_x_pass(og, A, Abits, g1, &fl));
    unsigned* flags = fl.base;
    int* mode = fl.mode;
    auto int_spacing = [](float sp) { return sp >= 1.0f && sp <= 64.0f && sp == floorf(sp); };
    auto span2 = [](float sp, int n) { double d = (double)sp * (n - 1); return d * d; };
    const bool env_y = int_spacing(A->sp[0]) && int_spacing(A->sp[1]) &&
                       span2(A->sp[0], A->nx) + span2(A->sp[1], A->ny) < 16777216.0 && A->ny <= 65535 && A->ny > 1;
    {
        Scratch sc(og.c, og.s);
        uint2* stack = nullptr;
        uint8_t* colflags = nullptr;
        if (env_y) VX_TRY(sc.alloc(sizeof(uint2) * total, &stack));
        VX_TRY(sc.alloc((size_t)A->nb * A->nz * A->nx, &colflags));
        VX_TRY((axis_pass<1, EDT_MID>(og, g1, out, A, env_y, stack, colflags, flags, 0.f, 0, mode)));
    }
    VX_TRY(scratch_free(og.c, og.s, g1));
    return scratch_free(og.c, og.s, flags);
}

// z pass only, on a squared in-plane distance map; max_in bounds its finite values
static int edt_z_only(OpGuard& og, const Image* S, float* out, double max_in) {
    const size_t total = S->count();
    auto int_spacing = [](float sp) { return sp >= 1.0f && sp <= 64.0f && sp == floorf(sp); };
    const double span = (double)S->sp[2] * (S->nz - 1);
    const bool env_z = int_spacing(S->sp[0]) && int_spacing(S->sp[1]) && int_spacing(S->sp[2]) &&
                       max_in + span * span < 16777216.0 && S->nz <= 65535 && S->nz > 1;
    Scratch sc(og.c, og.s);
    uint2* stack = nullptr;
    uint8_t* colflags = nullptr;
    unsigned* planeflag = nullptr;
    if (env_z) VX_TRY(sc.alloc(sizeof(uint2) * total, &stack));
    VX_TRY(sc.alloc((size_t)S->nb * S->ny * S->nx, &colflags));
    // the outward scan gets no emptiness information for a map that came from elsewhere: every plane counts
    const size_t n_planeflag = (size_t)S->nb * ((S->nz + 31) / 32);
    VX_TRY(sc.alloc(sizeof(unsigned) * n_planeflag, &planeflag));
    VX_CUDA(cudaMemsetAsync(planeflag, 0xff, sizeof(unsigned) * n_planeflag, og.st));
    // (a map that came from elsewhere carries no feature count: the envelope / scan kernels do every column)
    return axis_pass<2, EDT_SQRT>(og, (const float*)S->d, out, S, env_z, stack, colflags, planeflag, 0.f, 0, nullptr);
}

static int dist_threshold(vx_ctx ctx, vx_img a, float r, bool strict, int invert, vx_img* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    VX_REQUIRE(r == r, VX_E_INVALID_ARG, "radius is NaN");
    OpGuard og;
    VX_TRY(og.begin(ctx));
    Image* A = nullptr;
    const unsigned* Abits = nullptr;
    VX_TRY(og.input_bits(a, &A, &Abits));
    const float T = threshold_for(r, strict);
    Image* O = nullptr;
    Ball ball;
    int emax = 0;
    int rc = VX_OK;
    const bool by_ball = build_ball(A->sp, T, A->nx, A->ny, A->nz, &ball, &emax);
    // the ball dilation works on bits and leaves bits; the large-radius fallback writes bytes
    if (by_ball) VX_TRY(new_bits_image(og.c, A->nx, A->ny, A->nz, A->nb, A->sp, &O));
    else VX_TRY(new_image(og.c, VX_U8, A->nx, A->ny, A->nz, A->nb, A->sp, &O));
    const int align16 = (A->nx % 16) == 0;
    if (by_ball && !align16) {
        cudaError_t ce = cudaMemsetAsync(O->bits, 0, O->bit_words() * 4, og.st);
        if (ce != cudaSuccess) { drop(O); return fail(VX_E_CUDA, "memset: %s", cudaGetErrorString(ce)); }
    }
    if (by_ball) {
        const int nwx = (A->nx + 31) / 32;
        const unsigned long long nwords = (unsigned long long)A->nb * A->nz * A->ny * nwx;
        unsigned* D = nullptr;
        uint8_t* rowany = nullptr;
        const size_t nrows = (size_t)(nwords / nwx);
        rc = scratch_alloc(og.c, og.s, sizeof(unsigned) * nwords * (emax + 1), (void**)&D);
        if (rc == VX_OK) rc = scratch_alloc(og.c, og.s, nrows, (void**)&rowany);
        if (rc == VX_OK) {
            cudaError_t ce = cudaMemsetAsync(rowany, 0, nrows, og.st);
            if (ce != cudaSuccess) rc = fail(VX_E_CUDA, "memset: %s", cudaGetErrorString(ce));
        }
        const unsigned grid = (unsigned)((nwords + kEdtThreads - 1) / kEdtThreads);
        if (rc == VX_OK) {
            { VX_LAUNCH(og.c, og.s, "pack_bits_kernel");
              pack_bits_kernel<<<grid, kEdtThreads, 0, og.st>>>(Abits, D, A->nx, nwx, nwords, emax, rowany); }
            rc = check_launch("pack_bits_kernel");
        }
        long long* boff = nullptr;
        if (rc == VX_OK && ball.n > 0) {
            // entry tables: word offset into D (level, row shift) and the packed (dy, dz)
            std::vector<long long> tab((size_t)ball.n + ((size_t)ball.n + 1) / 2);
            int* dyz = reinterpret_cast<int*>(tab.data() + ball.n);
            int wy = 0, wz = 0;
            std::vector<int> order(ball.n);
            for (int t = 0; t < ball.n; t++) order[t] = t;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) {   // centre first
                return ball.dy[a] * ball.dy[a] + ball.dz[a] * ball.dz[a] < ball.dy[b] * ball.dy[b] + ball.dz[b] * ball.dz[b];
            });
            for (int i = 0; i < ball.n; i++) {
                const int t = order[i];
                tab[i] = (long long)ball.e[t] * (long long)nwords + ((long long)ball.dz[t] * A->ny + ball.dy[t]) * nwx;
                dyz[i] = (int)(uint8_t)ball.dy[t] | ((int)(uint8_t)ball.dz[t] << 8);
                wy = std::max(wy, std::abs((int)ball.dy[t]));
                wz = std::max(wz, std::abs((int)ball.dz[t]));
            }
            rc = scratch_alloc(og.c, og.s, sizeof(long long) * tab.size(), (void**)&boff);
            if (rc == VX_OK) {
                cudaError_t ce = cudaMemcpyAsync(boff, tab.data(), sizeof(long long) * tab.size(), cudaMemcpyHostToDevice, og.st);
                if (ce != cudaSuccess) rc = fail(VX_E_CUDA, "ball table upload: %s", cudaGetErrorString(ce));
            }
            if (rc == VX_OK) {
                { VX_LAUNCH(og.c, og.s, "ball_or_kernel");
                  ball_or_kernel<<<grid, kEdtThreads, 0, og.st>>>(D, O->bits, A->nx, A->ny, A->nz, nwx, nwords, boff,
                                                                 reinterpret_cast<const int*>(boff + ball.n), ball.n, wy, wz, invert, rowany, align16); }
                rc = check_launch("ball_or_kernel");
            }
            if (boff) scratch_free(og.c, og.s, boff);
        } else if (rc == VX_OK) {   // empty ball: nothing is within the radius
            const size_t n = O->count();
            cudaError_t ce = cudaMemsetAsync(O->bits, invert ? 0xff : 0, O->bit_words() * 4, og.st);
            if (ce == cudaSuccess && invert && (n & 31)) {   // bits past the batch stay 0
                const unsigned last = (1u << (n & 31)) - 1u;
                ce = cudaMemcpyAsync(O->bits + (n >> 5), &last, 4, cudaMemcpyHostToDevice, og.st);
            }
            if (ce != cudaSuccess) rc = fail(VX_E_CUDA, "memset: %s", cudaGetErrorString(ce));
        }
        if (rc == VX_OK) rc = scratch_free(og.c, og.s, D);
        if (rc == VX_OK && rowany) rc = scratch_free(og.c, og.s, rowany);
    } else {
        rc = edt_passes(og, A, Abits, O->d, EDT_THRESH, T, invert);
    }
    if (rc != VX_OK) { drop(O); return rc; }
    return og.finish(O, out);
}

}  // namespace vx

using namespace vx;

extern "C" {

int32_t vx_edt(vx_ctx ctx, vx_img a, vx_img* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    OpGuard og;
    VX_TRY(og.begin(ctx));
    Image* A = nullptr;
    const unsigned* Abits = nullptr;
    VX_TRY(og.input_bits(a, &A, &Abits));
    Image* O = nullptr;
    VX_TRY(new_image(og.c, VX_F32, A->nx, A->ny, A->nz, A->nb, A->sp, &O));
    int rc = edt_passes(og, A, Abits, O->d, EDT_SQRT, 0.f, 0);
    if (rc != VX_OK) { drop(O); return rc; }
    return og.finish(O, out);
}

int32_t vx_edt_sq_xy(vx_ctx ctx, vx_img a, vx_img* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    OpGuard og;
    VX_TRY(og.begin(ctx));
    Image* A = nullptr;
    const unsigned* Abits = nullptr;
    VX_TRY(og.input_bits(a, &A, &Abits));
    Image* O = nullptr;
    VX_TRY(new_image(og.c, VX_F32, A->nx, A->ny, A->nz, A->nb, A->sp, &O));
    int rc = edt_xy_only(og, A, Abits, (float*)O->d);
    if (rc != VX_OK) { drop(O); return rc; }
    return og.finish(O, out);
}

int32_t vx_edt_z_from_sq(vx_ctx ctx, vx_img sq, double max_in, vx_img* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    OpGuard og;
    VX_TRY(og.begin(ctx));
    Image* S = nullptr;
    VX_TRY(og.input(sq, &S, VX_F32));
    Image* O = nullptr;
    VX_TRY(new_image(og.c, VX_F32, S->nx, S->ny, S->nz, S->nb, S->sp, &O));
    int rc = edt_z_only(og, S, (float*)O->d, max_in);
    if (rc != VX_OK) { drop(O); return rc; }
    return og.finish(O, out);
}

int32_t vx_edt_signed(vx_ctx ctx, vx_img a, int32_t conn, vx_img* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    // boundary = a & near(!a);  signed = a ? -edt(boundary) : edt(a)
    vx_img na = 0, nna = 0, bnd = 0, dout = 0, din = 0, neg = 0, res = 0;
    int rc = vx_not(ctx, a, &na);
    if (rc == VX_OK) rc = vx_near(ctx, na, conn, &nna);
    if (rc == VX_OK) rc = vx_and(ctx, a, nna, &bnd);
    if (rc == VX_OK) rc = vx_edt(ctx, a, &dout);
    if (rc == VX_OK) rc = vx_edt(ctx, bnd, &din);
    if (rc == VX_OK) rc = vx_arith_scalar(ctx, din, VX_RSUB, 0.0f, &neg);
    if (rc == VX_OK) {
        // res = a ? neg : dout   ==  select(a, neg, 0) + select(!a, dout, 0)
        vx_op p[7] = {{VXP_LOAD, 0, 0, 0, 0, 0.f}, {VXP_LOAD, 1, 1, 0, 0, 0.f}, {VXP_LOAD, 2, 2, 0, 0, 0.f},
                      {VXP_SELECT, 3, 0, 1, 0, 0.f}, {VXP_NOT, 4, 0, 0, 0, 0.f},
                      {VXP_SELECT, 5, 4, 2, 0, 0.f}, {VXP_ARITH, 6, 3, 5, VX_ADD, 0.f}};
        vx_img leaves[3] = {a, neg, dout};
        rc = vx_fused_pointwise(ctx, p, 7, leaves, 3, nullptr, 0, &res);
    }
    vx_img tmp[6] = {na, nna, bnd, dout, din, neg};
    for (vx_img t : tmp)
        if (t) vx_release(t);
    if (rc != VX_OK) return rc;
    *out = res;
    return VX_OK;
}

int32_t vx_dist_lt(vx_ctx ctx, vx_img a, float r, vx_img* out) { return dist_threshold(ctx, a, r, false, 0, out); }
int32_t vx_dist_leq(vx_ctx ctx, vx_img a, float r, vx_img* out) { return dist_threshold(ctx, a, r, true, 0, out); }
int32_t vx_dist_geq(vx_ctx ctx, vx_img a, float r, vx_img* out) { return dist_threshold(ctx, a, r, false, 1, out); }
int32_t vx_dist_gt(vx_ctx ctx, vx_img a, float r, vx_img* out) { return dist_threshold(ctx, a, r, true, 1, out); }

}  // extern "C"
