// A6 crossCorrelation(rho, a, b, region, m, M, k) -- texture similarity
// (Greta.pdf p.43 "Texture analysis", p.53, p.48 similarTo).
//
// For every voxel x:  h1 = k-bin histogram of a over {y : |y-x| <= rho} (physical
// units, clipped to the image),  h2 = k-bin histogram of b over `region`;  bin i is
// [m + i*(M-m)/k, m + (i+1)*(M-m)/k), values outside [m, M) are not counted; the result
// is the Pearson correlation of the two k-vectors.  Written with exact integers:
//     r = (k*P - n1*n2) / sqrt( (k*S11 - n1^2) * (k*S22 - n2^2) ),
//     P = sum h1_i h2_i,  S11 = sum h1_i^2,  S22 = sum h2_i^2,  n = sum h
// so the only roundings are two integer->binary64 conversions, one product, one sqrt and
// one division -- the same IEEE operations the oracle performs, hence bit-exact.
//
// Pipeline: xc_bin_kernel turns the f32 image into u8 CODES (code = bin + 1, 0 = not
// counted: outside [m, M), NaN, and -- through the TMA zero fill -- outside the image);
// xc_region_hist_kernel + xc_stat_kernel build the region statistics; the sliding-ball kernel
// (xcorr_tma_kernel, below) produces the coefficients.  This operator is NOT memory-bound: a
// radius-5 ball has 515 voxels and the work is counter updates in shared memory.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <cstdlib>
#include <utility>
#include <vector>

#include "vx_bits.h"
#include "vx_internal.h"
#include "vx_tma.h"

namespace vx {

constexpr int kXcThreads = 128;
constexpr int kXcWarps = kXcThreads / 32;
constexpr int kXcMaxBins = 254;  // codes are u8: 1..k, 0 = not counted

__host__ __device__ __forceinline__ int xc_bin_of(float v, double m, double M, int k) {
    double dv = (double)v;
    if (!(dv >= m) || !(dv < M)) return k;  // not counted
#ifdef __CUDA_ARCH__
    double t = __ddiv_rn(__dmul_rn(__dsub_rn(dv, m), (double)k), __dsub_rn(M, m));
#else
    double t = ((dv - m) * (double)k) / (M - m);
#endif
    int i = (int)t;
    if (i >= k) i = k - 1;
    return i;
}

// Binning without a binary64 division per voxel.  bin_of is monotone in v, so it is fully
// described by its k+1 break points edge[i] = smallest binary32 whose bin is >= i (edge[0]:
// smallest value >= m, edge[k]: smallest value >= M).  The host finds them with xc_bin_of
// itself (a few evaluations per edge), the kernels then need one float guess and a couple
// of float compares per voxel -- and produce exactly xc_bin_of(v) for every v.
constexpr int kEdgeStride = 256;   // floats per volume in the edge table (k + 1 <= 255 used)

// code of a value: bin + 1, or 0 when the value is not counted
__device__ __forceinline__ unsigned code_from_edges(float v, const float* __restrict__ e, int k, float inv_w) {
    if (!(v >= e[0]) || !(v < e[k])) return 0u;   // outside [m, M) or NaN
    int g = (int)((v - e[0]) * inv_w);
    g = g < 0 ? 0 : (g > k - 1 ? k - 1 : g);
    while (g > 0 && v < e[g]) g--;
    while (g < k - 1 && v >= e[g + 1]) g++;
    return (unsigned)g + 1u;
}

// The code image has its own row pitch (a multiple of 16 bytes, what a TMA tensor map needs);
// bytes between nx and the pitch are never read as data.
// grid (blocks, nb): 16 voxels per thread (4 x float4 in, one uint4 out) when pitch == nx and
// nvox % 16 == 0, else one voxel per thread and step.
__global__ void __launch_bounds__(256)
xc_bin_kernel(const float* __restrict__ a, uint8_t* __restrict__ codes, size_t nvox, int nx, int pitch,
              const float* __restrict__ edges, int k) {
    __shared__ float e[kEdgeStride];
    const size_t vol = blockIdx.y;
    for (int i = threadIdx.x; i <= k; i += 256) e[i] = edges[vol * kEdgeStride + i];
    __syncthreads();
    const float inv_w = (e[k] > e[0]) ? (float)k / (e[k] - e[0]) : 0.0f;
    const float* src = a + vol * nvox;
    if (pitch == nx && (nvox & 15) == 0) {
        uint8_t* dst = codes + vol * nvox;
        const size_t ng = nvox >> 4;
        for (size_t g = (size_t)blockIdx.x * 256 + threadIdx.x; g < ng; g += (size_t)gridDim.x * 256) {
            const float4* p = reinterpret_cast<const float4*>(src) + g * 4;
            unsigned w[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float4 f = __ldg(p + q);
                w[q] = code_from_edges(f.x, e, k, inv_w) | (code_from_edges(f.y, e, k, inv_w) << 8) |
                       (code_from_edges(f.z, e, k, inv_w) << 16) | (code_from_edges(f.w, e, k, inv_w) << 24);
            }
            reinterpret_cast<uint4*>(dst)[g] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    } else {
        uint8_t* dst = codes + vol * (nvox / nx) * (size_t)pitch;
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvox; i += (size_t)gridDim.x * 256) {
            const size_t row = i / nx;
            dst[row * pitch + (i - row * nx)] = (uint8_t)code_from_edges(src[i], e, k, inv_w);
        }
    }
}

// The same codes for a quantised image (the tag of vx_internal.h: 16-bit codes q kept with the
// image, value = fl(fl(q * slope) + inter)): 2 bytes read per voxel instead of 4, and when the codes
// of the volume span at most kXqWin values the bin of every possible value is tabulated once per
// block in shared memory, so a voxel costs one table look-up.  Same bins by construction: the table
// entry is code_from_edges of the very float the image holds.
constexpr int kXqWin = 8192;
template <typename T>
__global__ void __launch_bounds__(256)
xc_bin_q16_kernel(const T* __restrict__ q_all, uint8_t* __restrict__ codes, size_t nvox, int nx, int pitch,
                  const float* __restrict__ edges, int k, const int* __restrict__ qrange, float slope, float inter) {
    __shared__ float e[kEdgeStride];
    __shared__ uint8_t tab[kXqWin];
    constexpr int bias = std::is_signed<T>::value ? 32768 : 0;
    const size_t vol = blockIdx.y;
    for (int i = threadIdx.x; i <= k; i += 256) e[i] = edges[vol * kEdgeStride + i];
    __syncthreads();
    const float inv_w = (e[k] > e[0]) ? (float)k / (e[k] - e[0]) : 0.0f;
    const int ulo = qrange[2 * vol], span = qrange[2 * vol + 1] - ulo + 1;
    const bool win = span <= kXqWin;
    auto code_of_u = [&](int u) { return code_from_edges(__fadd_rn(__fmul_rn((float)(u - bias), slope), inter), e, k, inv_w); };
    if (win)
        for (int i = threadIdx.x; i < span; i += 256) tab[i] = (uint8_t)code_of_u(ulo + i);
    __syncthreads();
    auto code_of = [&](unsigned h) {   // h: the 16 raw bits of a voxel
        const int u = (int)(std::is_signed<T>::value ? (h ^ 0x8000u) : h);
        return win ? (unsigned)tab[u - ulo] : code_of_u(u);
    };
    const T* src = q_all + vol * nvox;
    if (pitch == nx && (nvox & 15) == 0) {
        uint8_t* dst = codes + vol * nvox;
        const size_t ng = nvox >> 4;
        for (size_t g = (size_t)blockIdx.x * 256 + threadIdx.x; g < ng; g += (size_t)gridDim.x * 256) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(src) + 2 * g);
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(src) + 2 * g + 1);
            const unsigned qw[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            unsigned w[4];
#pragma unroll
            for (int i = 0; i < 4; i++)
                w[i] = code_of(qw[2 * i] & 0xffffu) | (code_of(qw[2 * i] >> 16) << 8) | (code_of(qw[2 * i + 1] & 0xffffu) << 16) |
                       (code_of(qw[2 * i + 1] >> 16) << 24);
            reinterpret_cast<uint4*>(dst)[g] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    } else {
        uint8_t* dst = codes + vol * (nvox / nx) * (size_t)pitch;
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvox; i += (size_t)gridDim.x * 256) {
            const size_t row = i / nx;
            dst[row * pitch + (i - row * nx)] = (uint8_t)code_of((unsigned)(unsigned short)src[i]);
        }
    }
}

// histogram of b over region: grid (blocks, nb), h2 is [nb][256] (zeroed by the caller).
// CODED: `b` is the code image of the sliding pass (b is the image that was binned for it).
template <bool CODED>
__global__ void __launch_bounds__(256)
xc_region_hist_kernel(const void* __restrict__ b_, const unsigned* __restrict__ region, size_t nvox, int nx, int pitch,
                      const float* __restrict__ edges, int k, unsigned* __restrict__ h2) {
    __shared__ unsigned sh[256];
    __shared__ float e[kEdgeStride];
    sh[threadIdx.x] = 0;
    const size_t vol = blockIdx.y;
    for (int i = threadIdx.x; i <= k; i += 256) e[i] = edges[vol * kEdgeStride + i];
    __syncthreads();
    const float inv_w = (e[k] > e[0]) ? (float)k / (e[k] - e[0]) : 0.0f;
    // the region is read in its bit form: the 16 voxels of group g are half word g of the volume
    const unsigned short* reg = reinterpret_cast<const unsigned short*>(region) + (vol * nvox) / 16;
    auto code_at = [&](size_t i) -> unsigned {
        if (CODED) {
            const uint8_t* cb = reinterpret_cast<const uint8_t*>(b_);
            if (pitch == nx) return cb[vol * nvox + i];
            const size_t row = i / nx;
            return cb[(vol * (nvox / nx) + row) * (size_t)pitch + (i - row * nx)];
        }
        return code_from_edges(reinterpret_cast<const float*>(b_)[vol * nvox + i], e, k, inv_w);
    };
    if ((nvox & 31) == 0) {
        // one word (32 voxels) per lane; a word that holds region voxels is handed round the warp and
        // every lane takes one of its bits, so a compact region does not leave 31 lanes waiting for one
        const unsigned* regw = region + (vol * nvox) / 32;
        const size_t nw = nvox >> 5;
        const int lane = threadIdx.x & 31;
        for (size_t w0 = ((size_t)blockIdx.x * 256 + (threadIdx.x & ~31u)); w0 < nw; w0 += (size_t)gridDim.x * 256) {
            const size_t w = w0 + lane;
            const unsigned r = w < nw ? __ldg(regw + w) : 0u;
            unsigned busy = __ballot_sync(0xffffffffu, r != 0u);
            while (busy) {
                const int src = __ffs(busy) - 1;
                busy &= busy - 1;
                const unsigned rr = __shfl_sync(0xffffffffu, r, src);
                if ((rr >> lane) & 1u) {
                    const unsigned code = code_at(((w0 + src) << 5) + lane);
                    if (code) atomicAdd(&sh[code - 1], 1u);
                }
            }
        }
    } else if ((nvox & 15) == 0) {
        const size_t ng = nvox >> 4;
        for (size_t g = (size_t)blockIdx.x * 256 + threadIdx.x; g < ng; g += (size_t)gridDim.x * 256) {
            unsigned r = __ldg(reg + g);
            while (r) {
                const int j = __ffs(r) - 1;
                r &= r - 1;
                const unsigned code = code_at((g << 4) + j);
                if (code) atomicAdd(&sh[code - 1], 1u);
            }
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvox; i += (size_t)gridDim.x * 256) {
            if (!flat_bit(region, vol * nvox + i)) continue;
            const unsigned code = code_at(i);
            if (code) atomicAdd(&sh[code - 1], 1u);
        }
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(h2 + vol * 256 + threadIdx.x, sh[threadIdx.x]);
}

// Statistics of one region histogram, shared by every voxel of the volume.
struct XcStat {
    long long n2;     // sum of h2
    double d2d;       // (double)(k*S22 - n2^2), the exact integer correctly rounded
    int d2pos;        // k*S22 - n2^2 > 0
    int clo, chi;     // codes (bin + 1) of the first and last non-empty bin; clo > chi: region empty
    int pad;
};

// k*S22 - n2^2 needs more than 64 bits once a bin holds > 4.3e8 voxels (k = 100): it is formed
// in 128 bits and rounded to binary64 by hand (round to nearest even), which is what the
// oracle's (double)(__int128) conversion produces.
__host__ __device__ inline double u128_to_double(unsigned long long hi, unsigned long long lo) {
    if (hi == 0) return (double)lo;   // exact conversion rule of the language (RN)
    int lz = 0;
    for (unsigned long long t = hi; !(t >> 63); t <<= 1) lz++;
    const int nbits = 128 - lz;            // > 64
    const int shift = nbits - 53;          // 12 .. 75
    unsigned long long mant, rem_hi, rem_lo;   // value = mant * 2^shift + rem
    if (shift >= 64) {
        mant = hi >> (shift - 64);
        rem_hi = (shift == 64) ? 0 : (hi & ((1ull << (shift - 64)) - 1));
        rem_lo = lo;
    } else {
        mant = (hi << (64 - shift)) | (lo >> shift);
        rem_hi = 0;
        rem_lo = lo & ((1ull << shift) - 1);
    }
    // half = 2^(shift-1)
    const unsigned long long half_hi = shift > 64 ? (1ull << (shift - 65)) : 0;
    const unsigned long long half_lo = shift > 64 ? 0 : (1ull << (shift - 1));
    const bool gt = rem_hi > half_hi || (rem_hi == half_hi && rem_lo > half_lo);
    const bool eq = rem_hi == half_hi && rem_lo == half_lo;
    if (gt || (eq && (mant & 1ull))) mant++;
    return ldexp((double)mant, shift);     // mant <= 2^53: exact
}

__host__ __device__ inline XcStat xc_region_stat(const unsigned* h, int k) {
    XcStat s;
    s.n2 = 0; s.clo = 1; s.chi = 0; s.pad = 0;
    unsigned long long s_hi = 0, s_lo = 0;   // S22 in 128 bits
    bool any = false;
    for (int i = 0; i < k; i++) {
        const unsigned long long v = h[i];
        if (v) { if (!any) s.clo = i + 1; s.chi = i + 1; any = true; }
        s.n2 += (long long)v;
        const unsigned long long sq = v * v;   // < 2^64
        const unsigned long long nlo = s_lo + sq;
        s_hi += nlo < s_lo;
        s_lo = nlo;
    }
    // k * S22 (k < 2^8, S22 < 2^72): schoolbook on 32-bit limbs
    const unsigned long long kk = (unsigned long long)k;
    const unsigned long long p0 = (s_lo & 0xffffffffull) * kk, p1 = (s_lo >> 32) * kk;
    unsigned long long lo = p0 + (p1 << 32);
    unsigned long long hi = s_hi * kk + (p1 >> 32) + (lo < p0);
    // n2^2 (n2 < 2^40)
    const unsigned long long n = (unsigned long long)s.n2, nl = n & 0xffffffffull, nh = n >> 32;
    const unsigned long long ll = nl * nl, lh = nl * nh, hh = nh * nh;
    unsigned long long q_lo = ll + (lh << 33);
    unsigned long long q_hi = hh + (lh >> 31) + (q_lo < ll);
    // difference (>= 0 by Cauchy-Schwarz)
    const unsigned long long d_lo = lo - q_lo;
    const unsigned long long d_hi = hi - q_hi - (lo < q_lo);
    s.d2pos = (d_hi | d_lo) != 0;
    s.d2d = u128_to_double(d_hi, d_lo);
    return s;
}

__global__ void xc_stat_kernel(const unsigned* __restrict__ h2, int k, int nb, XcStat* __restrict__ vstat) {
    const int vol = blockIdx.x * blockDim.x + threadIdx.x;
    if (vol < nb) vstat[vol] = xc_region_stat(h2 + (size_t)vol * 256, k);
}

struct XcGeo {
    int nx, ny, nz, k, ncols, seg, nseg, nw, pitch;  // nw = histogram words per thread = ceil(k/2)
};

__device__ __forceinline__ void hist_add(unsigned char* hb, unsigned bin, unsigned& S2, unsigned& n1) {
    unsigned short* p = reinterpret_cast<unsigned short*>(hb + (bin >> 1) * (kXcThreads * 4) + ((bin & 1u) << 1));
    unsigned h = *p;
    S2 += 2u * h + 1u;
    n1 += 1u;
    *p = (unsigned short)(h + 1u);
}
__device__ __forceinline__ void hist_sub(unsigned char* hb, unsigned bin, unsigned& S2, unsigned& n1) {
    unsigned short* p = reinterpret_cast<unsigned short*>(hb + (bin >> 1) * (kXcThreads * 4) + ((bin & 1u) << 1));
    unsigned h = *p;
    S2 -= 2u * h - 1u;
    n1 -= 1u;
    *p = (unsigned short)(h - 1u);
}

// Fallback for balls the TMA kernel cannot hold (more than 8 voxels wide in x, or a histogram
// that does not fit beside the ring): scattered global byte loads, u16 counters, one voxel
// column per thread.  grid (ceil(nx/32), ceil(nz*nseg/4), nb), block 128.  Dynamic shared memory:
//   hist  [nw][128] u32   private histograms (two u16 bins per word)
//   h2s   [2*nw]    u32   region histogram of this volume (zero padded)
//   cols  [ncols]   u32   ball columns: dx | dz<<8 | ey<<16 (dx,dz as signed bytes)
__global__ void __launch_bounds__(kXcThreads)
xcorr_direct_kernel(const uint8_t* __restrict__ codes, float* __restrict__ out,
             const unsigned* __restrict__ cols_g, const unsigned* __restrict__ h2,
             const XcStat* __restrict__ vstat, XcGeo g) {
    extern __shared__ __align__(16) unsigned smem[];
    unsigned* hist = smem;
    unsigned* h2s = smem + (size_t)g.nw * kXcThreads;
    unsigned* cols = h2s + 2 * g.nw;
    const int vol = blockIdx.z;
    for (int i = threadIdx.x; i < 2 * g.nw; i += kXcThreads) h2s[i] = i < g.k ? h2[(size_t)vol * 256 + i] : 0u;
    for (int i = threadIdx.x; i < g.ncols; i += kXcThreads) cols[i] = cols_g[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 32 + lane;
    const int r = blockIdx.y * kXcWarps + warp;
    const int z = r / g.nseg, sg = r % g.nseg;
    if (x >= g.nx || z >= g.nz) return;
    const int y0 = sg * g.seg;
    const int y1 = min(g.ny, y0 + g.seg);
    if (y0 >= y1) return;
    const size_t nvox = (size_t)g.nx * g.ny * g.nz;
    const uint8_t* bv = codes + (size_t)vol * g.pitch * g.ny * g.nz;
    float* ov = out + (size_t)vol * nvox;
    const XcStat vs = vstat[vol];
    const long long n2 = vs.n2;
    const unsigned K = (unsigned)g.k;
    unsigned char* hb = reinterpret_cast<unsigned char*>(hist) + threadIdx.x * 4;

    for (int w = 0; w < g.nw; w++) hist[w * kXcThreads + threadIdx.x] = 0u;
    unsigned S2 = 0, n1 = 0;
    // full ball around (x, y0, z)
    for (int c = 0; c < g.ncols; c++) {
        const unsigned cw = cols[c];
        const int xx = x + (int)(signed char)(cw & 0xffu);
        const int zz = z + (int)(signed char)((cw >> 8) & 0xffu);
        const int ey = (int)(cw >> 16);
        if (xx < 0 || xx >= g.nx || zz < 0 || zz >= g.nz) continue;
        const int ya = max(0, y0 - ey), yb = min(g.ny - 1, y0 + ey);
        const uint8_t* colp = bv + ((size_t)zz * g.ny) * g.pitch + xx;
        for (int yy = ya; yy <= yb; yy++) {
            unsigned b = __ldg(colp + (size_t)yy * g.pitch);
            if (b) hist_add(hb, b - 1u, S2, n1);
        }
    }
    for (int y = y0; y < y1; y++) {
        // ---- emit the coefficient at (x, y, z)
        unsigned long long P = 0;
        for (int w = 0; w < g.nw; w++) {
            const unsigned v = hist[w * kXcThreads + threadIdx.x];
            P += (unsigned long long)(v & 0xffffu) * h2s[2 * w];
            P += (unsigned long long)(v >> 16) * h2s[2 * w + 1];
        }
        const long long num = (long long)K * (long long)P - (long long)n1 * n2;
        const long long d1 = (long long)K * (long long)S2 - (long long)n1 * (long long)n1;
        float res = 0.0f;
        if (d1 > 0 && vs.d2pos)
            res = (float)__ddiv_rn((double)num, __dsqrt_rn(__dmul_rn((double)d1, vs.d2d)));
        ov[((size_t)z * g.ny + y) * g.nx + x] = res;
        if (y + 1 >= y1) break;
        // ---- slide the ball to y+1
        for (int c = 0; c < g.ncols; c++) {
            const unsigned cw = cols[c];
            const int xx = x + (int)(signed char)(cw & 0xffu);
            const int zz = z + (int)(signed char)((cw >> 8) & 0xffu);
            const int ey = (int)(cw >> 16);
            if (xx < 0 || xx >= g.nx || zz < 0 || zz >= g.nz) continue;
            const uint8_t* colp = bv + ((size_t)zz * g.ny) * g.pitch + xx;
            const int yin = y + 1 + ey, yout = y - ey;
            if (yin < g.ny) {
                unsigned b = __ldg(colp + (size_t)yin * g.pitch);
                if (b) hist_add(hb, b - 1u, S2, n1);
            }
            if (yout >= 0) {
                unsigned b = __ldg(colp + (size_t)yout * g.pitch);
                if (b) hist_sub(hb, b - 1u, S2, n1);
            }
        }
    }
}


// ---------------------------------------------------------------------------------------
// TMA-fed sliding-ball kernel (the default).  Design, with the B200 measurements behind it
// (scripts/ubench/smem_ops.cu, profiles/r02_ubench_smem.txt):
//
//  * the operator is bound by shared-memory operations, not by HBM: every voxel step moves
//    2 x #columns counters (162 for rho = 5) of a PRIVATE k-bin histogram.  On B200 a
//    conflict-free ATOMS costs as much as one LDS (0.6-0.9 SM cycles per warp instruction),
//    so a counter move is one `atom.shared.add.u32` instead of LDS + IADD + STS, and the
//    value it returns is exactly what the incremental sum of squares needs:
//        S11' - S11 = sum_in (2h+1) - sum_out (2h-1) = 2 (sum_in h - sum_out h) + 2 #columns
//    (h = counter value before the move; the identity telescopes for ANY order in which the
//    atomics of one thread are performed, because u32 counters may wrap transiently).
//    No per-voxel sweep over the histogram is left except the dot product with the region
//    histogram, which only visits the bins where the region histogram is not zero.
//  * counters are u32, hist[code][thread]: 32 lanes hit 32 banks, never a conflict.
//    (k+1) x 4 bytes per thread bounds the block to 6 warps x 2 blocks per SM for k = 100.
//  * a block = 6 warps = 6 consecutive z planes of one 32-voxel x chunk, marching along a y
//    segment.  The u8 codes it needs (x0-16 .. x0+47, 6 + 2*wz planes) arrive in shared memory
//    through TMA: a producer warp issues one `cp.async.bulk.tensor.4d` per chunk of 4 rows
//    (box 64 x 4 x planes x 1, zero fill outside the image = the "not counted" code 0) into a
//    ring of chunks guarded by mbarriers (full: bytes landed, ready: flat flags computed,
//    empty: all 6 consumer warps are done with the chunk).  No __syncthreads in the loop.
//  * flat-region shortcut per WARP: the producer marks every (row, plane) segment whose
//    in-image bytes all equal one reference code; when all rows and planes a warp's slide
//    touches are flat its histograms cannot change and the step is a plain store.  Rows above
//    and below the image count as flat too (they hold nothing but the not-counted code 0): a
//    column that starts at the top of the image builds its first ball from row counts (81
//    counter adds instead of 515 for rho = 5), and a slide whose flat window reaches over the
//    top or bottom moves a KNOWN number of voxels between code 0 and the reference code -- two
//    counter updates and a closed-form change of the sum of squares.  (Before, every column
//    paid a full build and ten full slides at its two ends, background or not: 13 of the ~110
//    busy-step equivalents of an average BraTS column.)
constexpr int kXtWarps = 6;                  // consumer warps = z planes per block.  (12 per block, one block per SM, the same
                                             // 12 warps per SM on one ring: 7 % faster where every step is busy -- noise 9.43 ->
                                             // 8.80 ms per 16 volumes -- but BraTS volumes 3.63 -> 4.05 ms: a block of 12 planes
                                             // is busy where any of them is, and background columns cost 0.69 instead of 0.54)
constexpr int kXtCons = kXtWarps * 32;       // histogram owners
constexpr int kXtThreads = kXtCons + 64;     // + the copy warp and the flag warp
constexpr int kXtCH = 4;                     // rows per TMA chunk
constexpr int kXtRowB = 64;                  // bytes per window row: 16 + 32 + 16 voxels (the copy engine wants
                                             // the innermost tile coordinate 16-byte aligned: measured, scripts/ubench/tma_probe.cu)
constexpr int kXtHalo = 16;                  // voxels of x halo on either side
constexpr int kXtMaxChunks = 16;
constexpr int kXtLook = 3;                   // ring chunks beyond the live ones: loads in flight / landed ahead
constexpr unsigned kXtHS = kXtCons * 4;      // byte stride between two codes of one thread's histogram

#ifdef XT_TRACE
__device__ long long xt_trace[1024];
#define XT_T(cond, idx) do { if ((cond) && xt_bx == 3 && xt_by == 20 && vol == 1) xt_trace[(idx)] = clock64(); } while (0)
#else
#define XT_T(cond, idx) do { } while (0)
#endif

struct XtGeo {
    int nx, ny, nz, nb, k, seg, nseg, ball;
    int ncols, nreal;          // length of the column table (with alignment filler) / real columns
    int wx, wy, wz, P, nchunk, pitch;
    int gstart[17], gend[17];  // columns with y half-extent e are [gstart[e], gend[e]), gstart 4-aligned
};

__device__ __forceinline__ unsigned lds_u8(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
template <int OFF>
__device__ __forceinline__ unsigned lds_u8_imm(unsigned addr) {   // LDS.U8 R, [R + imm]
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}
__device__ __forceinline__ unsigned atoms_add(unsigned addr, unsigned delta) {
    unsigned v;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(v) : "r"(addr), "r"(delta) : "memory");
    return v;
}

// ---- balls known at compile time (unit spacing, integer rho^2): the column list as constant
// expressions, so that the slide of the specialised kernel is straight-line code whose ring
// reads carry their column offset as an immediate (7 instructions per counter pair: 2 LDS.U8,
// 2 IMAD, 2 ATOMS, 1 IADD3) instead of a table-driven loop (11.5).
constexpr int cisqrt(int v) {
    int r = 0;
    while ((r + 1) * (r + 1) <= v) r++;
    return r;
}
template <int R2>
struct BallCols {
    static constexpr int W = cisqrt(R2);
    static constexpr int MAXC = (2 * W + 1) * (2 * W + 1);
    int n;
    int dx[MAXC], dz[MAXC], e[MAXC];
    int cnt[W + 1];   // columns per y half-extent
    constexpr BallCols() : n(0), dx{}, dz{}, e{}, cnt{} {
        for (int z = -W; z <= W; z++)
            for (int x = -W; x <= W; x++) {
                const int rem = R2 - x * x - z * z;
                if (rem < 0) continue;
                dx[n] = x; dz[n] = z; e[n] = cisqrt(rem);
                cnt[e[n]]++;
                n++;
            }
    }
};
template <int R2>
struct Ball {
    static constexpr BallCols<R2> c{};
    static constexpr int W = BallCols<R2>::W;
};
template <int R2>
constexpr BallCols<R2> Ball<R2>::c;

// number of columns with y half-extent e (a compare chain over constants: folds when e is known)
template <int R2, int... E>
__device__ __forceinline__ int ball_cnt_impl(int e, std::integer_sequence<int, E...>) {
    int r = 0;
    ((r = (e == E) ? std::integral_constant<int, Ball<R2>::c.cnt[E]>::value : r), ...);
    return r;
}
template <int R2>
__device__ __forceinline__ int ball_cnt(int e) {
    return ball_cnt_impl<R2>(e, std::make_integer_sequence<int, Ball<R2>::W + 1>{});
}

// sum over the columns that lie inside the image in x and z of sgn[e(column)] (this lane's x, this warp's z)
template <int R2, int... I>
__device__ __forceinline__ int ball_signed_cols_impl(int x, int z, int nx, int nz, const int* sgn, std::integer_sequence<int, I...>) {
    int d = 0;
    ((d += ((unsigned)(x + std::integral_constant<int, Ball<R2>::c.dx[I]>::value) < (unsigned)nx &&
            (unsigned)(z + std::integral_constant<int, Ball<R2>::c.dz[I]>::value) < (unsigned)nz)
               ? sgn[std::integral_constant<int, Ball<R2>::c.e[I]>::value]
               : 0),
     ...);
    return d;
}
template <int R2>
__device__ __forceinline__ int ball_signed_cols(int x, int z, int nx, int nz, const int* sgn) {
    return ball_signed_cols_impl<R2>(x, z, nx, nz, sgn, std::make_integer_sequence<int, Ball<R2>::c.n>{});
}

constexpr int kXtBatch = 14;   // counter pairs whose shared-memory operations are in flight together

template <int R2, int I>
__device__ __forceinline__ unsigned xt_ld(const unsigned* prow) {   // prow[e]: ring row of this column's voxel
    constexpr int e = Ball<R2>::c.e[I];
    constexpr int off = (Ball<R2>::W + Ball<R2>::c.dz[I]) * (kXtCH * kXtRowB) + kXtHalo + Ball<R2>::c.dx[I];
    return lds_u8_imm<off>(prow[e]);
}
// one batch: all ring reads, then the counter addresses, then the atomics (+1 for the voxel that
// enters the ball, -1 for the one that leaves); ri/ro receive the counters' previous values
template <int R2, int BASE, int... J>
__device__ __forceinline__ void xt_batch(std::integer_sequence<int, J...>, unsigned hbase, const unsigned* pin,
                                         const unsigned* pout, unsigned* ri, unsigned* ro) {
    ((ri[J] = xt_ld<R2, BASE + J>(pin)), ...);
    ((ro[J] = xt_ld<R2, BASE + J>(pout)), ...);
    ((ri[J] = hbase + ri[J] * kXtHS), ...);
    ((ro[J] = hbase + ro[J] * kXtHS), ...);
    ((ri[J] = atoms_add(ri[J], 1u)), ...);
    ((ro[J] = atoms_add(ro[J], 0xffffffffu)), ...);
}
// all columns of the ball in batches; the previous batch's returns are summed after the next
// batch's atomics were issued, so nothing waits for an atomic's round trip
template <int R2, int BASE>
__device__ __forceinline__ void xt_slide(unsigned hbase, const unsigned* pin, const unsigned* pout, unsigned& A,
                                         const unsigned* pri, const unsigned* pro, int pn) {
    constexpr int N = Ball<R2>::c.n;
    if constexpr (BASE < N) {
        constexpr int CNT = (N - BASE) < kXtBatch ? (N - BASE) : kXtBatch;
        unsigned ri[kXtBatch], ro[kXtBatch];
        xt_batch<R2, BASE>(std::make_integer_sequence<int, CNT>{}, hbase, pin, pout, ri, ro);
#pragma unroll
        for (int j = 0; j < kXtBatch; j++)
            if (j < pn) A += pri[j] - pro[j];
        xt_slide<R2, BASE + CNT>(hbase, pin, pout, A, ri, ro, CNT);
    } else {
#pragma unroll
        for (int j = 0; j < kXtBatch; j++)
            if (j < pn) A += pri[j] - pro[j];
    }
}

// one row of the initial ball: columns whose half-extent reaches |dy| add their voxel of ring row
// `prow`; with `nin` (a flat neighbourhood) every column adds all its voxels that lie inside the image
// at once -- nin[e] of the 2e+1 rows of a column with half-extent e -- since they hold one value.
// Batched like the slide: the ring reads of a batch are issued before its atomics.
template <int R2, int BASE, int... J>
__device__ __forceinline__ void xt_init_batch(std::integer_sequence<int, J...>, unsigned hbase, unsigned prow, int ady,
                                              const unsigned* nin) {
    unsigned v[sizeof...(J)];
    ((v[J] = (Ball<R2>::c.e[BASE + J] >= ady)
                 ? lds_u8_imm<(Ball<R2>::W + Ball<R2>::c.dz[BASE + J]) * (kXtCH * kXtRowB) + kXtHalo + Ball<R2>::c.dx[BASE + J]>(prow)
                 : 0u),
     ...);
    (((Ball<R2>::c.e[BASE + J] >= ady)
          ? (void)atoms_add(hbase + v[J] * kXtHS, nin ? nin[std::integral_constant<int, Ball<R2>::c.e[BASE + J]>::value] : 1u)
          : (void)0),
     ...);
}
template <int R2, int BASE>
__device__ __forceinline__ void xt_init_row(unsigned hbase, unsigned prow, int ady, const unsigned* nin) {
    constexpr int N = Ball<R2>::c.n;
    if constexpr (BASE < N) {
        constexpr int CNT = (N - BASE) < kXtBatch ? (N - BASE) : kXtBatch;
        xt_init_batch<R2, BASE>(std::make_integer_sequence<int, CNT>{}, hbase, prow, ady, nin);
        xt_init_row<R2, BASE + CNT>(hbase, prow, ady, nin);
    }
}

// grid: (units * 4 * nb) blocks in ONE dimension.  A unit is 2 x chunks by 2 plane blocks (of one y segment);
// the host sorts the units from the centre of the volume outwards and the blocks run unit by unit, inside a
// unit volume by volume, the (up to) four places of the unit together: block i works on volume (i / 4) % nb
// and place order[(i / (4 nb)) * 4 + i % 4] (0xffffffff: the unit has no such place, the block leaves).
// Why: blocks in the middle of a head slide through texture for their whole column (700 k cycles), blocks in
// the background skip most of theirs (50 k), and ONE busy block does not fill the shared-memory pipe of its SM
// (56 % in the block time line against 77 % with two).  In raster order busy and idle blocks shared SMs all
// along and the kernel ended on the busy blocks of the last volume; sorted, the SMs hold two busy blocks each
// until the busy blocks of all volumes are done, and the cheap ones fill the end: 3.74 -> 3.61 ms per 16 BraTS
// volumes.  The four places of a unit share most of their planes, so those come from L2 (with single places
// interleaved over the volumes the kernel read 6 x the code image from DRAM; four shells of a volume in raster
// order kept the traffic but lost the gain: 3.74 ms).  Block 256 threads.  Dynamic shared memory:
//   ring [nchunk][P][4][64] u8 | hist [k+1][192] u32 | h2s [k+1] u32 | cols [ncols] u32 | rowmask [64] u32 |
//   mbarriers [48] u64 | gs, ge [17] int
// cols[c] = byte offset of column c inside a chunk relative to (plane = warp, x = lane):
// (wz+dz)*256 + 16 + dx.   R2 > 0: the ball is Ball<R2> (unit spacing), cols/gs/ge are not used.
template <int R2>
__global__ void __launch_bounds__(kXtThreads, kXtWarps <= 6 ? 2 : 1)
xcorr_tma_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t* __restrict__ codes,
                 float* __restrict__ out, const unsigned* __restrict__ cols_g, const unsigned* __restrict__ order,
                 const unsigned* __restrict__ h2, const XcStat* __restrict__ vstat, const __grid_constant__ XtGeo g) {
    extern __shared__ unsigned char xt_smem_raw[];
    const unsigned xt_unit = blockIdx.x / (4u * (unsigned)g.nb), xt_rem = blockIdx.x - xt_unit * 4u * (unsigned)g.nb;
    const int vol = (int)(xt_rem >> 2);
    const unsigned xt_place = __ldg(order + xt_unit * 4u + (xt_rem & 3u));
    if (xt_place == 0xffffffffu) return;   // a unit at the edge of the grid of places
    const int xt_bx = (int)(xt_place & 0xffffu), xt_by = (int)(xt_place >> 16);
    // TMA destinations must be 128-byte aligned: align the dynamic window by hand (128 spare bytes)
    unsigned char* xt_smem = xt_smem_raw + ((128u - (smem_u32(xt_smem_raw) & 127u)) & 127u);
    const int chunkB = g.P * kXtCH * kXtRowB;   // P is even: a multiple of 128
    const int nbin = g.k + 1;
    const int nchunk = g.nchunk;
    unsigned char* ring = xt_smem;
    unsigned* hist = reinterpret_cast<unsigned*>(ring + (size_t)nchunk * chunkB);
    unsigned* h2s = hist + (size_t)((nbin + 3) & ~3) * kXtCons;   // histogram rows padded to a multiple of 4 codes
    unsigned* cols = h2s + ((nbin + 3) & ~3);
    unsigned* rowmask = cols + ((g.ncols + 3) & ~3);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(rowmask + kXtMaxChunks * kXtCH);
    int* gs = reinterpret_cast<int*>(bars + 3 * kXtMaxChunks);
    int* ge = gs + 17;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    XT_T(tid == 0, 998);
    const int x0 = xt_bx * 32;
    const int zb = xt_by / g.nseg, sg = xt_by % g.nseg;
    const int z0 = zb * kXtWarps;
    const int y0 = sg * g.seg;
    const int T = min(g.ny, y0 + g.seg) - y0;        // rows of this segment (>= 1)
    const int r0 = y0 - g.wy;                        // image row of ring row 0
    const int total_chunks = (T + 2 * g.wy + kXtCH - 1) / kXtCH;
    const unsigned bar0 = smem_u32(bars);
    // barrier addresses: full[s] = bar0 + 8 s, ready[s] = + 8 (16 + s), empty[s] = + 8 (32 + s)

    for (int i = tid; i < ((nbin + 3) & ~3); i += kXtThreads) h2s[i] = (i >= 1 && i < nbin) ? h2[(size_t)vol * 256 + (i - 1)] : 0u;
    if (R2 == 0) {
        for (int i = tid; i < g.ncols; i += kXtThreads) cols[i] = cols_g[i];
        if (tid < 17) { gs[tid] = g.gstart[tid]; ge[tid] = g.gend[tid]; }
    }
    if (tid == 0) {
        for (int s = 0; s < nchunk; s++) {
            mbar_init(bar0 + 8 * s, 1);
            mbar_init(bar0 + 8 * (kXtMaxChunks + s), 1);
            mbar_init(bar0 + 8 * (2 * kXtMaxChunks + s), kXtWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
    XT_T(tid == 0, 0);

    if (warp == kXtWarps) {
        // ================================================================ copy warp
        // issues the chunk loads as ring slots come back from the consumers; nothing else, so
        // that up to (nchunk - live chunks) loads are in flight whatever the flag warp is doing
        if (lane == 0) {
            const unsigned ring_s = smem_u32(ring);
            int slot = 0;
            unsigned par = 0;   // parity of the `empty` phase that frees the slot for its next use
            for (int j = 0; j < total_chunks; j++) {
                if (j >= nchunk) mbar_wait_relaxed(bar0 + 8 * (2 * kXtMaxChunks + slot), par);
                mbar_arrive_expect_tx(bar0 + 8 * slot, (unsigned)chunkB);
                tma_load_4d(ring_s + slot * chunkB, &tmap, bar0 + 8 * slot, x0 - kXtHalo, r0 + j * kXtCH, z0 - g.wz, vol);
                XT_T(true, 1 + j);
                if (++slot == nchunk) { slot = 0; if (j >= nchunk) par ^= 1u; }
            }
        }
        return;
    }
    if (warp == kXtWarps + 1) {
        // ================================================================ flag warp
        // marks every (plane, row) segment of a landed chunk whose relevant bytes all equal the
        // block's reference code.  Lane = (segment of the round, 16-byte quarter): the 32 lanes
        // read 512 consecutive bytes, a ballot folds the four quarters of each segment.
        unsigned ref4;
        {
            const int zz = min(max(z0 - g.wz, 0), g.nz - 1), xx = min(max(x0 - g.wx, 0), g.nx - 1);
            const int yy = min(max(r0, 0), g.ny - 1);
            ref4 = (unsigned)__ldg(codes + (((size_t)vol * g.nz + zz) * g.ny + yy) * g.pitch + xx) * 0x01010101u;
        }
        // bytes of this lane's quarter that matter: inside the image and within the x reach of the
        // ball around the block's 32 voxels
        unsigned bm[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            bm[i] = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int wxo = 16 * (lane & 3) + 4 * i + b, xa = x0 - kXtHalo + wxo;
                if (xa >= 0 && xa < g.nx && wxo >= kXtHalo - g.wx && wxo < kXtHalo + 32 + g.wx) bm[i] |= 0xffu << (8 * b);
            }
        }
        const int nsegm = g.P * kXtCH;   // (plane, row) segments of a chunk, 8 per round
        int slot = 0;
        unsigned par = 0;
        for (int j = 0; j < total_chunks; j++) {
            mbar_wait(bar0 + 8 * slot, par);
            XT_T(lane == 0, 100 + j);
            const uint4* cbase = reinterpret_cast<const uint4*>(ring + (size_t)slot * chunkB);
            unsigned m = 0;   // lanes 0..3: plane mask of ring row 4*slot + lane
            for (int q0 = 0; q0 * 8 < nsegm; q0 += 4) {   // four rounds (8 planes) per trip: loads first
                uint4 v[4];
                bool need[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int sgm = (q0 + u) * 8 + (lane >> 2);   // plane 2q + (lane >> 4), row (lane >> 2) & 3
                    const int zz = z0 - g.wz + (sgm >> 2);
                    const int ry = r0 + j * kXtCH + (sgm & 3);   // rows above / below the image hold no voxels: flat
                    need[u] = sgm < nsegm && zz >= 0 && zz < g.nz && ry >= 0 && ry < g.ny;
                    v[u] = need[u] ? cbase[(q0 + u) * 32 + lane] : make_uint4(ref4, ref4, ref4, ref4);
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const bool flat = ((((v[u].x ^ ref4) & bm[0]) | ((v[u].y ^ ref4) & bm[1]) | ((v[u].z ^ ref4) & bm[2]) |
                                        ((v[u].w ^ ref4) & bm[3])) == 0);
                    const unsigned bal = __ballot_sync(0xffffffffu, flat);
                    // row `lane & 3` of planes 2q and 2q+1: quarters at bits 4*row .. 4*row+3 (+16)
                    const int sh = 4 * (lane & 3), q = q0 + u;
                    m |= (((bal >> sh) & 0xfu) == 0xfu ? 1u : 0u) << (2 * q);
                    m |= (((bal >> (16 + sh)) & 0xfu) == 0xfu ? 1u : 0u) << (2 * q + 1);
                }
            }
            if (lane < 4) rowmask[slot * kXtCH + lane] = m;
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8 * (kXtMaxChunks + slot));
            XT_T(lane == 0, 200 + j);
            if (++slot == nchunk) { slot = 0; par ^= 1u; }
        }
        return;
    }

    // ==================================================================== consumer warps
    const int x = x0 + lane, z = z0 + warp;
    const bool wactive = z < g.nz;            // warp-uniform
    const bool active = wactive && x < g.nx;
    const size_t nvox = (size_t)g.nx * g.ny * g.nz;
    float* ov = out + (size_t)vol * nvox + ((size_t)z * g.ny + y0) * g.nx + x;   // output of step 0
    const XcStat vs = vstat[vol];
    const long long n2 = vs.n2;
    const long long K = (long long)g.k;
    const unsigned hbase = smem_u32(hist) + tid * 4;
    const unsigned tb = smem_u32(ring) + warp * (kXtCH * kXtRowB) + lane;
    const unsigned pm = ((1u << (2 * g.wz + 1)) - 1u) << warp;
    const bool p32 = (unsigned long long)g.ball * (unsigned long long)n2 < 0xffffffffull;
    const int wy = g.wy, wy2 = 2 * g.wy;

    for (int c = 0; c < nbin; c++) hist[c * kXtCons + tid] = 0u;   // own column only

    int w_slot = 0;        // ring slot of the next chunk whose `ready` barrier this warp waits for
    unsigned w_par = 0;
    int rows_ready = 0;    // ring rows [0, rows_ready) are in shared memory, flat flags included
    auto ensure = [&](int rel) {
        while (rows_ready <= rel) {
            mbar_wait(bar0 + 8 * (kXtMaxChunks + w_slot), w_par);
            if (++w_slot == nchunk) { w_slot = 0; w_par ^= 1u; }
            rows_ready += kXtCH;
        }
    };
    const int x_lim = nchunk * kXtCH;
    int x_idx = 0;         // rowmask index of the next ring row to examine (slot * 4 + row)
    int run = 0;           // consecutive flat rows (for this warp's planes) ending at the last examined row
    auto examine_next = [&]() {
        run = ((rowmask[x_idx] & pm) == pm) ? run + 1 : 0;
        if (++x_idx == x_lim) x_idx = 0;
    };
    // byte offset of ring row `rel` (general form, used outside the loop)
    auto rowoff = [&](int rel) { return (unsigned)(((rel >> 2) % nchunk) * chunkB + (rel & 3) * kXtRowB); };

    XT_T(tid == 0, 300);
    ensure(wy2);
    XT_T(tid == 0, 301);
    for (int rel = 0; rel <= wy2; rel++) examine_next();
    unsigned S = 0;   // sum over ALL codes (the dummy code 0 included) of h^2
    // columns per y half-extent (compiled in for the known balls)
    auto ncols_of = [&](int e) -> int {
        if constexpr (R2 > 0) return ball_cnt<R2>(e);
        else return ge[e] - gs[e];
    };
    // every column of this warp's balls lies inside the image in x and z (warp-uniform)
    const bool xz_in = x0 - g.wx >= 0 && x0 + 31 + g.wx < g.nx && z - g.wz >= 0 && z + g.wz < g.nz;
    if (wactive) {
        const bool flat0 = run >= wy2 + 1;
        // flat: the whole neighbourhood of the first voxel is one value per column.  A column with half-extent
        // e has nin[e] of its 2e+1 voxels in rows of the image: they go to ONE counter (the reference code, or
        // code 0 where the column lies outside the image in x or z); the voxels in rows above / below the image
        // are not counted (code 0) whatever the column.
        unsigned nin[17];
        unsigned outside = 0;
        if (flat0) {
#pragma unroll
            for (int e = 0; e <= (R2 > 0 ? Ball<R2>::W : 16); e++) {
                if (e > wy) break;
                const int lo = max(0, y0 - e), hi = min(g.ny - 1, y0 + e);
                nin[e] = (unsigned)(hi - lo + 1);
                outside += (unsigned)ncols_of(e) * (unsigned)(2 * e + 1 - (hi - lo + 1));
            }
        }
        if constexpr (R2 == 0) {
            if (flat0) {
                const unsigned prow = tb + rowoff(wy);
                for (int e = 0; e <= wy; e++)
                    for (int c = gs[e]; c < ge[e]; c++)
                        atoms_add(hbase + lds_u8(prow + cols[c]) * kXtHS, nin[e]);
                hist[tid] += outside;
            } else {
                for (int e = 0; e <= wy; e++)
                    for (int c = gs[e]; c < ge[e]; c++) {
                        const unsigned co = cols[c];
                        for (int dy = -e; dy <= e; dy++)
                            atoms_add(hbase + lds_u8(tb + rowoff(wy + dy) + co) * kXtHS, 1u);
                    }
            }
        } else {
            // row by row: row dy of the ball holds the columns with half-extent >= |dy|.  A flat
            // neighbourhood needs one row only, every column weighted by its voxels inside the image.
            if (flat0) {
                xt_init_row<R2, 0>(hbase, tb + rowoff(wy), 0, nin);
                hist[tid] += outside;
            } else {
                for (int dy = -wy; dy <= wy; dy++)
                    xt_init_row<R2, 0>(hbase, tb + rowoff(wy + dy), dy < 0 ? -dy : dy, nullptr);
            }
        }
        for (int c = 0; c < nbin; c++) {
            const unsigned h = hist[c * kXtCons + tid];
            S += h * h;
        }
    }

    XT_T(tid == 0, 302);
    float res = 0.0f;
    bool dirty = true;    // warp-uniform: histograms changed since the last coefficient
    int slot_t = 0;       // ring slot of the chunk that holds row t
    for (int t = 0; t < T; t++) {
        if (wactive && dirty) {
            const unsigned hd = hist[tid];   // not-counted voxels of the ball (code 0)
            unsigned long long P = 0;
            // four codes per trip: one broadcast LDS.128 of the region histogram, four counter loads;
            // codes beyond chi (up to the padded end of h2s) meet zeros of h2s
            if (p32) {
                unsigned P32 = 0;
                for (int c = vs.clo & ~3; c <= vs.chi; c += 4) {
                    const uint4 w = *reinterpret_cast<const uint4*>(h2s + c);
                    const unsigned* hc = hist + c * kXtCons + tid;
                    P32 += hc[0] * w.x + hc[kXtCons] * w.y + hc[2 * kXtCons] * w.z + hc[3 * kXtCons] * w.w;
                }
                P = P32;
            } else {
                for (int c = vs.clo & ~3; c <= vs.chi; c += 4) {
                    const uint4 w = *reinterpret_cast<const uint4*>(h2s + c);
                    const unsigned* hc = hist + c * kXtCons + tid;
                    P += (unsigned long long)hc[0] * w.x + (unsigned long long)hc[kXtCons] * w.y +
                         (unsigned long long)hc[2 * kXtCons] * w.z + (unsigned long long)hc[3 * kXtCons] * w.w;
                }
            }
            const long long n1 = (long long)g.ball - (long long)hd;
            const long long s11 = (long long)(S - hd * hd);
            const long long num = K * (long long)P - n1 * n2;
            const long long d1 = K * s11 - n1 * n1;
            res = 0.0f;
            if (d1 > 0 && vs.d2pos)
                res = (float)__ddiv_rn((double)num, __dsqrt_rn(__dmul_rn((double)d1, vs.d2d)));
            dirty = false;
        }
        const int tq = t & 3;
        XT_T(tid == 0, 400 + t);
        // ---- four steps at once where nothing changes: the next four rows are flat as well, so
        // the slides t .. t+3 move nothing and rows t .. t+3 all get the current coefficient
        // (rows above / below the image count as flat, but a slide whose window reaches them does move
        // counters: those steps are taken one at a time below)
        if (tq == 0 && t + 4 < T && run >= wy2 + 1 && y0 + t >= wy && y0 + t + 4 + wy < g.ny) {
            ensure(t + wy2 + 4);
            XT_T(tid == 0, 600 + t);
            int i1 = x_idx + 1, i2 = x_idx + 2, i3 = x_idx + 3;
            if (i1 >= x_lim) i1 -= x_lim;
            if (i2 >= x_lim) i2 -= x_lim;
            if (i3 >= x_lim) i3 -= x_lim;
            const unsigned m4 = rowmask[x_idx] & rowmask[i1] & rowmask[i2] & rowmask[i3];
            if ((m4 & pm) == pm) {
                if (active) {
                    ov[0] = res; ov[g.nx] = res; ov[2 * (size_t)g.nx] = res; ov[3 * (size_t)g.nx] = res;
                }
                ov += 4 * (size_t)g.nx;
                run += 4;
                x_idx = x_idx + 4 >= x_lim ? x_idx + 4 - x_lim : x_idx + 4;
                __syncwarp();
                if (lane == 0) mbar_arrive(bar0 + 8 * (2 * kXtMaxChunks + slot_t));
                slot_t = slot_t + 1 == nchunk ? 0 : slot_t + 1;
                t += 3;
                continue;
            }
        }
        if (active) *ov = res;
        ov += g.nx;
        if (t + 1 >= T) break;
        // ---- slide the ball from row y0+t to y0+t+1: ring rows t .. t+2wy+1
        ensure(t + wy2 + 1);
        examine_next();
        // The slide moves nothing when its whole window (ring rows t .. t+2wy+1) is flat AND lies inside the
        // image in y.  Where the window is flat but reaches over the top or bottom of the image, a column moves
        // one voxel between code 0 (the row outside) and the reference code (the row inside) -- unless it lies
        // outside the image in x or z, where every row holds code 0: a known number of moves between two counters.
        const int yv = y0 + t;
        const bool ycross = yv < wy || yv + 1 + wy >= g.ny;
        const bool allflat = run >= wy2 + 2;
        if (wactive && allflat && ycross) {
            int sgn[17];   // per half-extent: +1 the column's leaving row is outside the image, -1 its entering row
            int dall = 0;  // voxels that move from code 0 to the reference code when every column is inside in x, z
#pragma unroll
            for (int e = 0; e <= (R2 > 0 ? Ball<R2>::W : 16); e++) {
                if (e > wy) break;
                const bool out_gone = yv - e < 0, in_gone = yv + 1 + e >= g.ny;   // leaving / entering row is outside
                sgn[e] = (out_gone && !in_gone ? 1 : 0) - (in_gone && !out_gone ? 1 : 0);
                dall += ncols_of(e) * sgn[e];
            }
            int d = dall;
            if (!xz_in) {   // columns outside the image in x or z hold code 0 in every row: they move nothing
                if constexpr (R2 > 0) {
                    d = ball_signed_cols<R2>(x, z, g.nx, g.nz, sgn);
                } else {
                    d = 0;
                    for (int e = 0; e <= wy; e++)
                        for (int c = gs[e]; c < ge[e]; c++) {
                            const int co = (int)cols[c];   // (wz + dz) * 256 + 16 + dx
                            const int dz = co / (kXtCH * kXtRowB) - g.wz, dx = co % (kXtCH * kXtRowB) - kXtHalo;
                            if ((unsigned)(x + dx) < (unsigned)g.nx && (unsigned)(z + dz) < (unsigned)g.nz) d += sgn[e];
                        }
                }
            }
            if (dall != 0 || !xz_in) {
                const unsigned refc = lds_u8(tb + rowoff(t + wy) + g.wz * (kXtCH * kXtRowB) + kXtHalo);   // own voxel
                if (refc != 0 && d != 0) {   // (a reference value that is not counted shares code 0 with the rows outside)
                    unsigned* hr = hist + refc * kXtCons + tid;
                    const unsigned h0 = hist[tid], h1 = *hr;
                    hist[tid] = h0 - (unsigned)d;
                    *hr = h1 + (unsigned)d;
                    S += 2u * (unsigned)d * (h1 - h0) + 2u * (unsigned)(d * d);
                }
                dirty = true;
            }
        } else if (wactive && !allflat) {
            unsigned A = 0;            // sum of counters before +1 minus sum of counters before -1
            if constexpr (R2 > 0) {
                constexpr int W = Ball<R2>::W;
                unsigned pin[W + 1], pout[W + 1];
#pragma unroll
                for (int e = 0; e <= W; e++) {
                    int sin = slot_t + ((tq + W + 1 + e) >> 2), sout = slot_t + ((tq + W - e) >> 2);
                    if (sin >= nchunk) sin -= nchunk;
                    if (sout >= nchunk) sout -= nchunk;
                    pin[e] = tb + sin * chunkB + ((tq + W + 1 + e) & 3) * kXtRowB;
                    pout[e] = tb + sout * chunkB + ((tq + W - e) & 3) * kXtRowB;
                }
                xt_slide<R2, 0>(hbase, pin, pout, A, nullptr, nullptr, 0);
            } else {
                unsigned pi0 = 0, pi1 = 0, pi2 = 0, pi3 = 0, po0 = 0, po1 = 0, po2 = 0, po3 = 0;   // previous batch
                for (int e = 0; e <= wy; e++) {
                    int c = gs[e];
                    const int cend = ge[e];
                    if (c >= cend) continue;
                    int sin = slot_t + ((tq + wy + 1 + e) >> 2), sout = slot_t + ((tq + wy - e) >> 2);
                    if (sin >= nchunk) sin -= nchunk;
                    if (sout >= nchunk) sout -= nchunk;
                    const unsigned pin = tb + sin * chunkB + ((tq + wy + 1 + e) & 3) * kXtRowB;
                    const unsigned pout = tb + sout * chunkB + ((tq + wy - e) & 3) * kXtRowB;
#pragma unroll 1
                    for (; c + 4 <= cend; c += 4) {
                        const uint4 co = *reinterpret_cast<const uint4*>(cols + c);   // gstart is 4-aligned
                        unsigned i0 = lds_u8(pin + co.x), o0 = lds_u8(pout + co.x);
                        unsigned i1 = lds_u8(pin + co.y), o1 = lds_u8(pout + co.y);
                        unsigned i2 = lds_u8(pin + co.z), o2 = lds_u8(pout + co.z);
                        unsigned i3 = lds_u8(pin + co.w), o3 = lds_u8(pout + co.w);
                        A += (pi0 - po0) + (pi1 - po1) + (pi2 - po2) + (pi3 - po3);   // returns of the previous batch
                        i0 = hbase + i0 * kXtHS; o0 = hbase + o0 * kXtHS;
                        i1 = hbase + i1 * kXtHS; o1 = hbase + o1 * kXtHS;
                        i2 = hbase + i2 * kXtHS; o2 = hbase + o2 * kXtHS;
                        i3 = hbase + i3 * kXtHS; o3 = hbase + o3 * kXtHS;
                        pi0 = atoms_add(i0, 1u); po0 = atoms_add(o0, 0xffffffffu);
                        pi1 = atoms_add(i1, 1u); po1 = atoms_add(o1, 0xffffffffu);
                        pi2 = atoms_add(i2, 1u); po2 = atoms_add(o2, 0xffffffffu);
                        pi3 = atoms_add(i3, 1u); po3 = atoms_add(o3, 0xffffffffu);
                    }
                    for (; c < cend; c++) {
                        const unsigned co = cols[c];
                        const unsigned ai = hbase + lds_u8(pin + co) * kXtHS, ao = hbase + lds_u8(pout + co) * kXtHS;
                        A += atoms_add(ai, 1u) - atoms_add(ao, 0xffffffffu);
                    }
                }
                A += (pi0 - po0) + (pi1 - po1) + (pi2 - po2) + (pi3 - po3);
            }
            S += 2u * A + 2u * (unsigned)g.nreal;
            dirty = true;
        }
        // ring row t is dead after this slide; its chunk goes back to the producer with its last row
        if (tq == 3) {
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8 * (2 * kXtMaxChunks + slot_t));
            slot_t = slot_t + 1 == nchunk ? 0 : slot_t + 1;
        }
    }
    XT_T(tid == 0, 999);
}

// ball columns (dx, dz, ey): offsets with (dx*sx)^2 + (dy*sy)^2 + (dz*sz)^2 <= rho^2, binary64
static bool build_columns(const float sp[3], float rho, std::vector<unsigned>* cols, long long* ball) {
    cols->clear();
    *ball = 0;
    if (!(rho >= 0.0f)) return true;
    const double r2 = (double)rho * (double)rho;
    auto in = [&](int dx, int dy, int dz) {
        double ax = dx * (double)sp[0], ay = dy * (double)sp[1], az = dz * (double)sp[2];
        double s = ax * ax;
        s = s + ay * ay;
        s = s + az * az;
        return s <= r2;
    };
    int wx = 0, wz = 0;
    while (in(wx + 1, 0, 0)) { if (++wx > 120) return false; }
    while (in(0, 0, wz + 1)) { if (++wz > 120) return false; }
    for (int dz = -wz; dz <= wz; dz++)
        for (int dx = -wx; dx <= wx; dx++) {
            if (!in(dx, 0, dz)) continue;
            int ey = 0;
            while (in(dx, ey + 1, dz)) { if (++ey > 120) return false; }
            cols->push_back((unsigned)(dx & 0xff) | ((unsigned)(dz & 0xff) << 8) | ((unsigned)ey << 16));
            *ball += 2 * ey + 1;
        }
    return true;
}

// ---- host: exact break points of xc_bin_of --------------------------------------------------
static unsigned fkey(float f) {
    unsigned u;
    memcpy(&u, &f, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static float fromkey(unsigned key) {
    unsigned u = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
// edge[i] (i = 0..k) = smallest binary32 v with  v >= m  and ( v >= M  or  xc_bin_of(v) >= i )
static void compute_edges(double m, double M, int k, float* e) {
    if (!(M > m)) {   // empty or NaN range: nothing is ever counted
        for (int i = 0; i <= k; i++) e[i] = INFINITY;
        return;
    }
    auto pred = [&](unsigned key, int i) {
        const float v = fromkey(key);
        const double dv = (double)v;
        if (!(dv >= m)) return false;
        if (!(dv < M)) return true;
        return xc_bin_of(v, m, M, k) >= i;
    };
    const unsigned kmin = fkey(-INFINITY), kmax = fkey(INFINITY);
    for (int i = 0; i <= k; i++) {
        // analytic guess, then gallop in units of ulps to bracket the break point, then bisect
        double guess = m + (M - m) * ((double)i / (double)k);
        float gf = (float)guess;
        if (!(gf == gf)) gf = 0.0f;
        unsigned lo, hi, c = fkey(gf);
        if (c < kmin) c = kmin;
        if (c > kmax) c = kmax;
        if (pred(c, i)) {   // answer <= c: walk down until pred fails
            hi = c;
            unsigned step = 1;
            lo = c;
            while (true) {
                if (lo - kmin <= step) { lo = kmin; if (pred(lo, i)) { hi = lo; } break; }
                lo -= step;
                if (!pred(lo, i)) break;
                hi = lo;
                step <<= 1;
            }
            if (hi == kmin) { e[i] = fromkey(hi); continue; }
        } else {            // answer > c
            lo = c;
            unsigned step = 1;
            hi = c;
            while (true) {
                if (kmax - hi <= step) { hi = kmax; break; }
                hi += step;
                if (pred(hi, i)) break;
                lo = hi;
                step <<= 1;
            }
        }
        // invariant: pred(lo) false, pred(hi) true (pred(+inf) is always true)
        while (hi - lo > 1) {
            const unsigned mid = lo + (hi - lo) / 2;
            if (pred(mid, i)) hi = mid; else lo = mid;
        }
        e[i] = fromkey(hi);
    }
}

static int upload_edges(OpGuard& og, Scratch& sc, int nb, const double* m, const double* M, int k, float** d_edges) {
    std::vector<float> e((size_t)nb * kEdgeStride, INFINITY);
    for (int v = 0; v < nb; v++) {
        if (v > 0 && m[v] == m[v - 1] && M[v] == M[v - 1])
            memcpy(&e[(size_t)v * kEdgeStride], &e[(size_t)(v - 1) * kEdgeStride], sizeof(float) * kEdgeStride);
        else
            compute_edges(m[v], M[v], k, &e[(size_t)v * kEdgeStride]);
    }
    VX_TRY(sc.alloc(sizeof(float) * e.size(), d_edges));
    // pageable source: staged by the runtime before the call returns
    VX_CUDA(cudaMemcpyAsync(*d_edges, e.data(), sizeof(float) * e.size(), cudaMemcpyHostToDevice, og.st));
    return VX_OK;
}

static dim3 stream_grid(size_t nvox, int nb) {
    int per_vol = (int)((nvox + 256 * 16 - 1) / (256 * 16));
    int cap = (kNumSMs * 8 + nb - 1) / nb;
    return dim3(per_vol < cap ? per_vol : (cap < 1 ? 1 : cap), nb);
}

static int code_pitch(int nx) { return (nx + 15) & ~15; }   // TMA: row stride a multiple of 16 bytes

// u8 codes of A (0 for values that are not counted), rows padded to code_pitch(nx)
static int xcorr_codes(OpGuard& og, Scratch& sc, Image* A, const float* d_edges, int32_t k, uint8_t** d_codes) {
    const int pitch = code_pitch(A->nx);
    VX_TRY(sc.alloc((size_t)pitch * A->ny * A->nz * A->nb + 64, d_codes));
    {
        VX_LAUNCH(og.c, og.s, "xc_bin_kernel");
        const dim3 grid = stream_grid(A->nvox(), A->nb);
        if (A->has_codes() && og.c->quantised_paths.load()) {   // bin the 16-bit codes: half the bytes, a table look-up per voxel
            if (A->q_dtype == VX_I16)
                xc_bin_q16_kernel<short><<<grid, 256, 0, og.st>>>((const short*)A->q16, *d_codes, A->nvox(), A->nx, pitch, d_edges, k,
                                                                  A->q_range(), A->q_slope, A->q_inter);
            else
                xc_bin_q16_kernel<unsigned short><<<grid, 256, 0, og.st>>>((const unsigned short*)A->q16, *d_codes, A->nvox(), A->nx,
                                                                           pitch, d_edges, k, A->q_range(), A->q_slope, A->q_inter);
        } else {
            xc_bin_kernel<<<grid, 256, 0, og.st>>>((const float*)A->d, *d_codes, A->nvox(), A->nx, pitch, d_edges, k);
        }
    }
    return check_launch("xc_bin_kernel");
}

// h2 of `src` (f32 values, or the u8 code image when coded) over region R into d_h2 ([nb][256] u32)
static int region_hist(OpGuard& og, const void* src, bool coded, Image* R, const unsigned* Rbits, const float* d_edges,
                       int32_t k, unsigned* d_h2) {
    const int nb = R->nb;
    const size_t nvox = R->nvox();
    const int pitch = code_pitch(R->nx);
    VX_CUDA(cudaMemsetAsync(d_h2, 0, sizeof(unsigned) * 256 * nb, og.st));
    {
        VX_LAUNCH(og.c, og.s, "xc_region_hist_kernel");
        if (coded)
            xc_region_hist_kernel<true><<<stream_grid(nvox, nb), 256, 0, og.st>>>(src, Rbits, nvox, R->nx, pitch, d_edges, k, d_h2);
        else
            xc_region_hist_kernel<false><<<stream_grid(nvox, nb), 256, 0, og.st>>>(src, Rbits, nvox, R->nx, pitch, d_edges, k, d_h2);
    }
    return check_launch("xc_region_hist_kernel");
}

// Shared back end: statistics of the region histogram d_h2 ([nb][256] u32 on the device, final)
// and the sliding-ball kernel over the codes.
static int xcorr_run(OpGuard& og, Scratch& sc, Image* A, const uint8_t* d_codes, const unsigned* d_h2, float rho,
                     int32_t k, vx_img* out) {
    std::vector<unsigned> cols;
    long long ball = 0;
    VX_REQUIRE(build_columns(A->sp, rho, &cols, &ball) && ball <= 65535 && cols.size() <= 8192,
               VX_E_UNSUPPORTED, "radius %g is too large for the sliding-histogram kernel", (double)rho);
    const int nb = A->nb;
    const int pitch = code_pitch(A->nx);
    unsigned* d_cols = nullptr;
    XcStat* d_vstat = nullptr;
    VX_TRY(sc.alloc(sizeof(XcStat) * nb, &d_vstat));
    const size_t max_places = 4 * (size_t)((A->nx + 31) / 32 / 2 + 1) * ((A->nz + kXtWarps - 1) / kXtWarps / 2 + 1) *
                              (size_t)std::max(1, A->ny / 40 + 1);
    VX_TRY(sc.alloc(sizeof(unsigned) * (cols.size() + 128 + max_places), &d_cols));
    {
        VX_LAUNCH(og.c, og.s, "xc_stat_kernel");
        xc_stat_kernel<<<(nb + 63) / 64, 64, 0, og.st>>>(d_h2, k, nb, d_vstat);
    }
    VX_TRY(check_launch("xc_stat_kernel"));
    // geometry of the ball
    int wx = 0, wy = 0, wz = 0;
    for (unsigned cw : cols) {
        int dx = (int)(signed char)(cw & 0xffu), dz = (int)(signed char)((cw >> 8) & 0xffu), ey = (int)(cw >> 16);
        wx = std::max(wx, std::abs(dx)); wz = std::max(wz, std::abs(dz)); wy = std::max(wy, ey);
    }
    // ---- the TMA kernel: window 16 + 32 + 16 voxels wide, 6 + 2 wz planes, ring of 4-row chunks
    const int P = kXtWarps + 2 * wz;
    const int nchunk = 2 + wy / 2 + kXtLook;   // live chunks (2 + wy/2) + kXtLook in flight
    std::vector<unsigned> tcols;
    XtGeo g;
    memset(&g, 0, sizeof(g));
    bool tma_ok = !cols.empty() && wx <= kXtHalo && wy <= 16 && P <= 32 && nchunk <= kXtMaxChunks;
    if (tma_ok) {
        for (int e = 0; e <= wy; e++) {
            while (tcols.size() & 3) tcols.push_back(0);   // alignment filler, never read
            g.gstart[e] = (int)tcols.size();
            int cnt = 0;
            for (unsigned cw : cols) {
                int dx = (int)(signed char)(cw & 0xffu), dz = (int)(signed char)((cw >> 8) & 0xffu), ey = (int)(cw >> 16);
                if (ey != e) continue;
                tcols.push_back((unsigned)((wz + dz) * (kXtCH * kXtRowB) + kXtHalo + dx));
                cnt++;
            }
            g.gend[e] = g.gstart[e] + cnt;
        }
    }
    const int nbin = k + 1;
    const size_t chunkB = (size_t)P * kXtCH * kXtRowB;
    const size_t smem_tma = 128 + nchunk * chunkB +
                            sizeof(unsigned) * ((size_t)((nbin + 3) & ~3) * kXtCons + ((nbin + 3) & ~3) + ((tcols.size() + 3) & ~(size_t)3) +
                                                kXtMaxChunks * kXtCH) +
                            sizeof(unsigned long long) * 3 * kXtMaxChunks + sizeof(int) * 34;
    tma_ok = tma_ok && smem_tma <= 227 * 1024;
    Image* O = nullptr;
    VX_TRY(new_image(og.c, VX_F32, A->nx, A->ny, A->nz, A->nb, A->sp, &O));
    int rc = VX_OK;
    if (tma_ok) {
        // y segments: whole columns when the batch already fills the machine many times over, shorter
        // ones for small batches.  The initial full-ball build is expensive (515 counter adds for rho = 5,
        // and in the block time line of scripts/ubench/xt_trace.py as long as 6-11 slide steps), so a
        // segment of 240 rows instead of 120 is worth more than the finer tail: 16 BraTS volumes
        // 4.47 -> 4.08 ms with 3328 blocks of 240 rows instead of 6656 of 120 (noise, where no block is
        // short: 9.21 -> 9.40).
        const long long ctas_per_seg = (long long)((A->nx + 31) / 32) * ((A->nz + kXtWarps - 1) / kXtWarps) * nb;
        const long long want = 10LL * kNumSMs * 2;
        int nseg = (int)std::min<long long>((want + ctas_per_seg - 1) / ctas_per_seg, std::max(1, A->ny / 40));
        nseg = std::max(1, nseg);
        const int seg = (A->ny + nseg - 1) / nseg;
        nseg = (A->ny + seg - 1) / seg;
        // units (2 x chunks by 2 plane blocks of one segment) from the centre of the volume outwards, four place
        // slots each (x chunk | (plane block * nseg + segment) << 16, or 0xffffffff), appended to the column
        // table so that one copy carries both
        const int gx = (A->nx + 31) / 32, nzb = (A->nz + kXtWarps - 1) / kXtWarps;
        const int ux = (gx + 1) / 2, uz = (nzb + 1) / 2;
        const size_t n_units = (size_t)ux * uz * nseg;
        const size_t places_at = (tcols.size() + 3) & ~(size_t)3;
        {
            std::vector<std::pair<float, unsigned>> keyed;   // (distance of the unit's centre from the volume's, unit)
            keyed.reserve(n_units);
            for (int sg = 0; sg < nseg; sg++)
                for (int kz = 0; kz < uz; kz++)
                    for (int kx = 0; kx < ux; kx++) {
                        const float cx = 0.5f * (2 * kx + std::min(2 * kx + 2, gx)), cz = 0.5f * (2 * kz + std::min(2 * kz + 2, nzb));
                        const float fx = cx / gx - 0.5f, fz = cz / nzb - 0.5f, fy = (sg + 0.5f) / nseg - 0.5f;
                        keyed.emplace_back(fx * fx + fz * fz + fy * fy, (unsigned)((sg * uz + kz) * ux + kx));
                    }
            std::stable_sort(keyed.begin(), keyed.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
            tcols.resize(places_at, 0u);
            for (auto& kv : keyed) {
                const int kx = (int)(kv.second % ux), kz = (int)(kv.second / ux % uz), sg = (int)(kv.second / ux / uz);
                for (int j = 0; j < 4; j++) {
                    const int bx = 2 * kx + (j & 1), zb = 2 * kz + (j >> 1);
                    tcols.push_back(bx < gx && zb < nzb ? (unsigned)bx | ((unsigned)(zb * nseg + sg) << 16) : 0xffffffffu);
                }
            }
        }
        g.nx = A->nx; g.ny = A->ny; g.nz = A->nz; g.nb = nb; g.k = k;
        g.seg = seg; g.nseg = nseg; g.ball = (int)ball;
        g.ncols = (int)places_at; g.nreal = (int)cols.size();   // (the places follow the columns in tcols)
        g.wx = wx; g.wy = wy; g.wz = wz; g.P = P; g.nchunk = nchunk; g.pitch = pitch;
        // unit spacing and an integer-valued ball: the kernels with the column list compiled in
        auto kern = xcorr_tma_kernel<0>;
        if (A->sp[0] == 1.0f && A->sp[1] == 1.0f && A->sp[2] == 1.0f) {
            const double r2 = (double)rho * (double)rho;
            const int ir2 = (int)floor(r2);   // dx^2+dy^2+dz^2 <= rho^2  <=>  <= floor(rho^2) for integers
            switch (ir2) {
                case 4: kern = xcorr_tma_kernel<4>; break;
                case 9: kern = xcorr_tma_kernel<9>; break;
                case 16: kern = xcorr_tma_kernel<16>; break;
                case 25: kern = xcorr_tma_kernel<25>; break;
                default: break;
            }
        }
        CUtensorMap tmap;
        const uint64_t dims[4] = {(uint64_t)A->nx, (uint64_t)A->ny, (uint64_t)A->nz, (uint64_t)nb};
        const uint64_t strides[3] = {(uint64_t)pitch, (uint64_t)pitch * A->ny, (uint64_t)pitch * A->ny * A->nz};
        const uint32_t box[4] = {(uint32_t)kXtRowB, (uint32_t)kXtCH, (uint32_t)P, 1u};
        rc = tma_encode_u8_4d(&tmap, d_codes, dims, strides, box);
        if (rc == VX_OK) {
            cudaError_t e = cudaMemcpyAsync(d_cols, tcols.data(), sizeof(unsigned) * tcols.size(), cudaMemcpyHostToDevice, og.st);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma);
            if (e == cudaSuccess && (gx > 0xffff || nzb * nseg > 0xffff || (unsigned long long)n_units * 4 * nb > 0x7fffffffull)) {
                rc = fail(VX_E_UNSUPPORTED, "crossCorrelation: %d x %d x %d blocks exceed the launch grid", gx, nzb * nseg, nb);
            }
            if (e != cudaSuccess) rc = fail(VX_E_CUDA, "crossCorrelation set-up: %s", cudaGetErrorString(e));
        }
        if (rc == VX_OK) {
            dim3 grid((unsigned)(n_units * 4 * nb));
            {
                VX_LAUNCH(og.c, og.s, "xcorr_kernel");
                kern<<<grid, kXtThreads, smem_tma, og.st>>>(tmap, d_codes, (float*)O->d, d_cols, d_cols + places_at, d_h2, d_vstat, g);
            }
            rc = check_launch("xcorr_kernel");
        }
    } else {
        XcGeo gd;
        gd.nx = A->nx; gd.ny = A->ny; gd.nz = A->nz; gd.k = k; gd.pitch = pitch;
        gd.ncols = (int)cols.size();
        int nseg = (A->ny + 63) / 64;
        const int seg = (A->ny + nseg - 1) / nseg;
        nseg = (A->ny + seg - 1) / seg;
        gd.nw = (k + 1) / 2; gd.seg = seg; gd.nseg = nseg;
        if (!cols.empty()) {
            cudaError_t e = cudaMemcpyAsync(d_cols, cols.data(), sizeof(unsigned) * cols.size(), cudaMemcpyHostToDevice, og.st);
            if (e != cudaSuccess) rc = fail(VX_E_CUDA, "crossCorrelation set-up: %s", cudaGetErrorString(e));
        }
        size_t smem = sizeof(unsigned) * ((size_t)gd.nw * kXcThreads + 2 * gd.nw + gd.ncols);
        if (rc == VX_OK && smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(xcorr_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) rc = fail(VX_E_CUDA, "smem attribute: %s", cudaGetErrorString(e));
        }
        if (rc == VX_OK) {
            dim3 grid((A->nx + 31) / 32, (A->nz * gd.nseg + kXcWarps - 1) / kXcWarps, nb);
            {
                VX_LAUNCH(og.c, og.s, "xcorr_direct_kernel");
                xcorr_direct_kernel<<<grid, kXcThreads, smem, og.st>>>(d_codes, (float*)O->d, d_cols, d_h2, d_vstat, gd);
            }
            rc = check_launch("xcorr_direct_kernel");
        }
    }
    if (rc != VX_OK) { drop(O); return rc; }
    return og.finish(O, out);
}

}  // namespace vx


using namespace vx;

extern "C" int32_t vx_cross_correlation(vx_ctx ctx, vx_img a, vx_img b, vx_img region, float rho,
                                        const double* m, const double* M, int32_t k, vx_img* out) {
    VX_REQUIRE(out != nullptr && m != nullptr && M != nullptr, VX_E_INVALID_ARG, "NULL argument");
    VX_REQUIRE(k >= 1 && k <= kXcMaxBins, VX_E_UNSUPPORTED, "k = %d bins; supported range is [1,%d]", k,
               kXcMaxBins);
    VX_REQUIRE(rho == rho, VX_E_INVALID_ARG, "rho is NaN");
    OpGuard og;
    VX_TRY(og.begin(ctx));
    Image *A = nullptr, *B = nullptr, *R = nullptr;
    VX_TRY(og.input(a, &A, VX_F32));
    VX_TRY(og.input(b, &B, VX_F32));
    const unsigned* Rbits = nullptr;
    VX_TRY(og.input_bits(region, &R, &Rbits));
    VX_TRY(same_shape(A, B));
    VX_TRY(same_shape(A, R));
    Scratch sc(og.c, og.s);
    float* d_edges = nullptr;
    unsigned* d_h2 = nullptr;
    uint8_t* d_codes = nullptr;
    VX_TRY(upload_edges(og, sc, A->nb, m, M, k, &d_edges));
    VX_TRY(sc.alloc(sizeof(unsigned) * 256 * A->nb, &d_h2));
    VX_TRY(xcorr_codes(og, sc, A, d_edges, k, &d_codes));
    // similarTo(a) correlates an image with itself: its codes are already there
    if (A == B) VX_TRY(region_hist(og, d_codes, true, R, Rbits, d_edges, k, d_h2));
    else VX_TRY(region_hist(og, B->d, false, R, Rbits, d_edges, k, d_h2));
    return xcorr_run(og, sc, A, d_codes, d_h2, rho, k, out);
}

extern "C" int32_t vx_region_histogram(vx_ctx ctx, vx_img b, vx_img region, const double* m,
                                       const double* M, int32_t k, uint64_t* h2) {
    VX_REQUIRE(h2 != nullptr && m != nullptr && M != nullptr, VX_E_INVALID_ARG, "NULL argument");
    VX_REQUIRE(k >= 1 && k <= kXcMaxBins, VX_E_UNSUPPORTED, "k = %d bins; supported range is [1,%d]", k,
               kXcMaxBins);
    OpGuard og;
    VX_TRY(og.begin(ctx));
    Image *B = nullptr, *R = nullptr;
    VX_TRY(og.input(b, &B, VX_F32));
    const unsigned* Rbits = nullptr;
    VX_TRY(og.input_bits(region, &R, &Rbits));
    VX_TRY(same_shape(B, R));
    const int nb = B->nb;
    Scratch sc(og.c, og.s);
    float* d_edges = nullptr;
    unsigned* d_h2 = nullptr;
    VX_TRY(upload_edges(og, sc, nb, m, M, k, &d_edges));
    VX_TRY(sc.alloc(sizeof(unsigned) * 256 * nb, &d_h2));
    VX_TRY(region_hist(og, B->d, false, R, Rbits, d_edges, k, d_h2));
    std::vector<unsigned> host((size_t)256 * nb);
    VX_CUDA(cudaMemcpyAsync(host.data(), d_h2, sizeof(unsigned) * 256 * nb, cudaMemcpyDeviceToHost, og.st));
    VX_CUDA(cudaStreamSynchronize(og.st));
    for (int v = 0; v < nb; v++)
        for (int i = 0; i < k; i++) h2[(size_t)v * k + i] = host[(size_t)v * 256 + i];
    return og.finish_scalar();
}

extern "C" int32_t vx_cross_correlation_h2(vx_ctx ctx, vx_img a, const uint64_t* h2, float rho,
                                           const double* m, const double* M, int32_t k, vx_img* out) {
    VX_REQUIRE(out != nullptr && h2 != nullptr && m != nullptr && M != nullptr, VX_E_INVALID_ARG, "NULL argument");
    VX_REQUIRE(k >= 1 && k <= kXcMaxBins, VX_E_UNSUPPORTED, "k = %d bins; supported range is [1,%d]", k,
               kXcMaxBins);
    VX_REQUIRE(rho == rho, VX_E_INVALID_ARG, "rho is NaN");
    OpGuard og;
    VX_TRY(og.begin(ctx));
    Image* A = nullptr;
    VX_TRY(og.input(a, &A, VX_F32));
    const int nb = A->nb;
    std::vector<unsigned> host((size_t)256 * nb, 0u);
    for (int v = 0; v < nb; v++)
        for (int i = 0; i < k; i++) {
            const uint64_t c = h2[(size_t)v * k + i];
            VX_REQUIRE(c <= 0xffffffffull, VX_E_UNSUPPORTED, "region histogram counter %llu exceeds 32 bits",
                       (unsigned long long)c);
            host[(size_t)v * 256 + i] = (unsigned)c;
        }
    Scratch sc(og.c, og.s);
    float* d_edges = nullptr;
    unsigned* d_h2 = nullptr;
    uint8_t* d_codes = nullptr;
    VX_TRY(upload_edges(og, sc, nb, m, M, k, &d_edges));
    VX_TRY(sc.alloc(sizeof(unsigned) * 256 * nb, &d_h2));
    // pageable source: cudaMemcpyAsync stages it before returning, `host` may die afterwards
    VX_CUDA(cudaMemcpyAsync(d_h2, host.data(), sizeof(unsigned) * 256 * nb, cudaMemcpyHostToDevice, og.st));
    VX_TRY(xcorr_codes(og, sc, A, d_edges, k, &d_codes));
    return xcorr_run(og, sc, A, d_codes, d_h2, rho, k, out);
}

#ifdef XT_TRACE
extern "C" __attribute__((visibility("default"))) int32_t vx_debug_xt_trace(long long* out) {
    cudaDeviceSynchronize();
    return (int32_t)cudaMemcpyFromSymbol(out, vx::xt_trace, sizeof(long long) * 1024);
}
#endif

// Host-only helper (no device needed): the k+1 break points of the crossCorrelation binning for
// one (m, M), as used by the kernels.  Exposed so the exactness of the table can be tested on
// a machine without a GPU.
extern "C" int32_t vx_xcorr_bin_edges(double m, double M, int32_t k, float* edges) {
    VX_REQUIRE(edges != nullptr, VX_E_INVALID_ARG, "edges is NULL");
    VX_REQUIRE(k >= 1 && k <= kXcMaxBins, VX_E_UNSUPPORTED, "k = %d bins; supported range is [1,%d]", k,
               kXcMaxBins);
    compute_edges(m, M, k, edges);
    return VX_OK;
}

// Host-only helper (no device needed): the region statistic the kernels use, (double)(k*S22 - n2^2)
// formed in 128 bits, for one k-bin histogram.  Exposed so the two-limb arithmetic and the
// hand-written rounding can be tested against the oracle's __int128 on a machine without a GPU.
extern "C" int32_t vx_xcorr_region_stat(const uint64_t* h2, int32_t k, double* d2, int32_t* positive) {
    VX_REQUIRE(h2 != nullptr && d2 != nullptr && positive != nullptr, VX_E_INVALID_ARG, "NULL argument");
    VX_REQUIRE(k >= 1 && k <= kXcMaxBins, VX_E_UNSUPPORTED, "k = %d bins; supported range is [1,%d]", k,
               kXcMaxBins);
    unsigned h[256] = {0};
    for (int i = 0; i < k; i++) {
        VX_REQUIRE(h2[i] <= 0xffffffffull, VX_E_UNSUPPORTED, "region histogram counter %llu exceeds 32 bits",
                   (unsigned long long)h2[i]);
        h[i] = (unsigned)h2[i];
    }
    const XcStat s = xc_region_stat(h, k);
    *d2 = s.d2d;
    *positive = s.d2pos;
    return VX_OK;
}
