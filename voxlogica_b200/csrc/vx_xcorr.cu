// A6 crossCorrelation(rho, a, b, region, m, M, k) -- texture similarity
// (Greta.pdf p.43 "Texture analysis", p.53, p.48 similarTo).
//
// For every voxel x:  h1 = k-bin histogram of a over {y : |y-x| <= rho} (physical
// units, clipped to the image),  h2 = k-bin histogram of b over `region`;  bin i is
// [m + i*(M-m)/k, m + (i+1)*(M-m)/k), values outside [m, M) are not counted; the result
// is the Pearson correlation of the two k-vectors.  Written with exact integers:
//     r = (k*P - n1*n2) / sqrt( (k*S11 - n1^2) * (k*S22 - n2^2) ),
//     P = sum h1_i h2_i,  S11 = sum h1_i^2,  S22 = sum h2_i^2,  n = sum h
// so the only roundings are two int64->binary64 conversions, one product, one sqrt and
// one division -- the same IEEE operations the oracle performs, hence bit-exact.
//
// This operator is NOT memory-bound: a radius-5 ball has 515 voxels.  The kernel
// slides the ball along y: lanes of a warp are 32 consecutive x (coalesced u8 bin
// loads), each thread owns a private k-bin histogram in shared memory (two u16 bins
// per 32-bit word, word index = pair*T + tid, so a warp never bank-conflicts) and per
// step only touches the voxels entering and leaving the ball (2 x #(dx,dz) columns,
// 162 instead of 515 for rho = 5), maintaining S11 and n1 incrementally
// (+-(2h+1)).  P is a k-term dot product against h2 (broadcast reads).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "vx_internal.h"

namespace vx {

constexpr int kXcThreads = 128;
constexpr int kXcWarps = kXcThreads / 32;
constexpr int kXcMaxBins = 254;  // bins are stored as u8; value k = not counted

__host__ __device__ __forceinline__ int xc_bin_of(float v, double m, double M, int k) {
    double dv = (double)v;
    if (!(dv >= m) || !(dv < M)) return k;  // the dummy bin: not counted
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

__device__ __forceinline__ unsigned bin_from_edges(float v, const float* __restrict__ e, int k, float inv_w) {
    if (!(v >= e[0]) || !(v < e[k])) return (unsigned)k;   // outside [m, M) or NaN: the dummy bin
    int g = (int)((v - e[0]) * inv_w);
    g = g < 0 ? 0 : (g > k - 1 ? k - 1 : g);
    while (g > 0 && v < e[g]) g--;
    while (g < k - 1 && v >= e[g + 1]) g++;
    return (unsigned)g;
}

// grid (blocks, nb): 16 voxels per thread (4 x float4 in, one uint4 out) when nvox % 16 == 0
__global__ void __launch_bounds__(256)
xc_bin_kernel(const float* __restrict__ a, uint8_t* __restrict__ bins, size_t nvox,
              const float* __restrict__ edges, int k) {
    __shared__ float e[kEdgeStride];
    const size_t vol = blockIdx.y;
    for (int i = threadIdx.x; i <= k; i += 256) e[i] = edges[vol * kEdgeStride + i];
    __syncthreads();
    const float inv_w = (e[k] > e[0]) ? (float)k / (e[k] - e[0]) : 0.0f;
    const float* src = a + vol * nvox;
    uint8_t* dst = bins + vol * nvox;
    if ((nvox & 15) == 0) {
        const size_t ng = nvox >> 4;
        for (size_t g = (size_t)blockIdx.x * 256 + threadIdx.x; g < ng; g += (size_t)gridDim.x * 256) {
            const float4* p = reinterpret_cast<const float4*>(src) + g * 4;
            unsigned w[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float4 f = __ldg(p + q);
                w[q] = bin_from_edges(f.x, e, k, inv_w) | (bin_from_edges(f.y, e, k, inv_w) << 8) |
                       (bin_from_edges(f.z, e, k, inv_w) << 16) | (bin_from_edges(f.w, e, k, inv_w) << 24);
            }
            reinterpret_cast<uint4*>(dst)[g] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvox; i += (size_t)gridDim.x * 256)
            dst[i] = (uint8_t)bin_from_edges(src[i], e, k, inv_w);
    }
}

// histogram of b over region: grid (blocks, nb), h2 is [nb][256] (zeroed by the caller).
// BINNED: `b` already holds u8 bins (b is the image that was binned for the sliding pass).
template <bool BINNED>
__global__ void __launch_bounds__(256)
xc_region_hist_kernel(const void* __restrict__ b_, const uint8_t* __restrict__ region, size_t nvox,
                      const float* __restrict__ edges, int k, unsigned* __restrict__ h2) {
    __shared__ unsigned sh[256];
    __shared__ float e[kEdgeStride];
    sh[threadIdx.x] = 0;
    const size_t vol = blockIdx.y;
    for (int i = threadIdx.x; i <= k; i += 256) e[i] = edges[vol * kEdgeStride + i];
    __syncthreads();
    const float inv_w = (e[k] > e[0]) ? (float)k / (e[k] - e[0]) : 0.0f;
    const uint8_t* reg = region + vol * nvox;
    if ((nvox & 15) == 0) {
        const size_t ng = nvox >> 4;
        for (size_t g = (size_t)blockIdx.x * 256 + threadIdx.x; g < ng; g += (size_t)gridDim.x * 256) {
            const uint4 r = __ldg(reinterpret_cast<const uint4*>(reg) + g);
            if ((r.x | r.y | r.z | r.w) == 0) continue;    // regions are small: most groups are empty
            const unsigned rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (rw[q] == 0) continue;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (((rw[q] >> (8 * j)) & 0xffu) == 0) continue;
                    const size_t i = (g << 4) + 4 * q + j;
                    unsigned bin;
                    if (BINNED) bin = reinterpret_cast<const uint8_t*>(b_)[vol * nvox + i];
                    else bin = bin_from_edges(reinterpret_cast<const float*>(b_)[vol * nvox + i], e, k, inv_w);
                    if (bin < (unsigned)k) atomicAdd(&sh[bin], 1u);
                }
            }
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvox; i += (size_t)gridDim.x * 256) {
            if (reg[i] == 0) continue;
            unsigned bin;
            if (BINNED) bin = reinterpret_cast<const uint8_t*>(b_)[vol * nvox + i];
            else bin = bin_from_edges(reinterpret_cast<const float*>(b_)[vol * nvox + i], e, k, inv_w);
            if (bin < (unsigned)k) atomicAdd(&sh[bin], 1u);
        }
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(h2 + vol * 256 + threadIdx.x, sh[threadIdx.x]);
}

// per volume: vstat[2*vol] = n2, vstat[2*vol+1] = k*S22 - n2^2
__global__ void xc_stat_kernel(const unsigned* __restrict__ h2, int k, long long* __restrict__ vstat) {
    const int vol = blockIdx.x * blockDim.x + threadIdx.x;
    if (vol >= (int)gridDim.x * (int)blockDim.x) return;
    long long n2 = 0;
    unsigned long long s22 = 0;
    for (int i = 0; i < k; i++) {
        unsigned long long h = h2[(size_t)vol * 256 + i];
        n2 += (long long)h;
        s22 += h * h;
    }
    vstat[2 * vol] = n2;
    vstat[2 * vol + 1] = (long long)((unsigned long long)k * s22 - (unsigned long long)n2 * (unsigned long long)n2);
}

struct XcGeo {
    int nx, ny, nz, k, ncols, seg, nseg, nw;  // nw = histogram words per thread = ceil(k/2)
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

// grid (ceil(nx/32), ceil(nz*nseg/4), nb), block 128.  Dynamic shared memory:
//   hist  [nw][128] u32   private histograms (two u16 bins per word)
//   h2s   [2*nw]    u32   region histogram of this volume (zero padded)
//   cols  [ncols]   u32   ball columns: dx | dz<<8 | ey<<16 (dx,dz as signed bytes)
__global__ void __launch_bounds__(kXcThreads)
xcorr_direct_kernel(const uint8_t* __restrict__ bins, float* __restrict__ out,
             const unsigned* __restrict__ cols_g, const unsigned* __restrict__ h2,
             const long long* __restrict__ vstat, XcGeo g) {
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
    const uint8_t* bv = bins + (size_t)vol * nvox;
    float* ov = out + (size_t)vol * nvox;
    const long long n2 = vstat[2 * vol], d2 = vstat[2 * vol + 1];
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
        const uint8_t* colp = bv + ((size_t)zz * g.ny) * g.nx + xx;
        for (int yy = ya; yy <= yb; yy++) {
            unsigned b = __ldg(colp + (size_t)yy * g.nx);
            if (b < K) hist_add(hb, b, S2, n1);
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
        if (d1 > 0 && d2 > 0)
            res = (float)__ddiv_rn((double)num, __dsqrt_rn(__dmul_rn((double)d1, (double)d2)));
        ov[((size_t)z * g.ny + y) * g.nx + x] = res;
        if (y + 1 >= y1) break;
        // ---- slide the ball to y+1
        for (int c = 0; c < g.ncols; c++) {
            const unsigned cw = cols[c];
            const int xx = x + (int)(signed char)(cw & 0xffu);
            const int zz = z + (int)(signed char)((cw >> 8) & 0xffu);
            const int ey = (int)(cw >> 16);
            if (xx < 0 || xx >= g.nx || zz < 0 || zz >= g.nz) continue;
            const uint8_t* colp = bv + ((size_t)zz * g.ny) * g.nx + xx;
            const int yin = y + 1 + ey, yout = y - ey;
            if (yin < g.ny) {
                unsigned b = __ldg(colp + (size_t)yin * g.nx);
                if (b < K) hist_add(hb, b, S2, n1);
            }
            if (yout >= 0) {
                unsigned b = __ldg(colp + (size_t)yout * g.nx);
                if (b < K) hist_sub(hb, b, S2, n1);
            }
        }
    }
}


// ---------------------------------------------------------------------------------------
// Ring-buffered variant (the default).  A block = 4 warps = 4 consecutive z planes of one
// 32-voxel x chunk, marching together along a y segment.  The u8 bins the block needs
// (x0-8..x0+39, 4+2*wz planes) are staged in a shared-memory ring of R >= 2*wy+3 rows:
// every step the block fetches ONE new row per plane with coalesced 32-bit loads (168
// words for rho = 5) instead of 162 scattered global byte loads per thread, and every
// bin read of the inner loop is a conflict-free LDS (32 lanes = 32 consecutive bytes).
//
// The inner loop is branch-free: voxels that are not counted (outside [m,M) or outside
// the image) carry the dummy bin k, which has its own histogram slot; its count is
// subtracted when the coefficient is emitted (n1 = |ball| - h[k], S11 excludes h[k]^2,
// h2[k] = 0).  Columns are sorted by their y half-extent so that the ring-row offsets of
// the entering / leaving voxels are loop invariants of each group, and all histogram
// traffic uses explicit shared-space addresses.
struct XrGeo {
    int nx, ny, nz, k, ncols, seg, nseg, nw, ball;
    int wy, wz, R, P;      // ring rows (2*wy+3) and planes (4 + 2*wz)
    int gstart[17], gend[17];  // columns with y half-extent e are [gstart[e], gend[e]), gstart 4-aligned
};
constexpr int kRingXW = 48;  // voxels per ring row: 8 + 32 + 8
constexpr int kRingXB = 96;  // bytes per ring row (one u16 per voxel)

// The ring does not hold bins but the byte offset of the bin's counter inside a thread's
// histogram column: word pair index = bin >> 1 (stride kXcThreads*4 = 512 bytes), half =
// bin & 1.  Converting once per loaded voxel saves five ALU instructions on each of the
// 162 counter accesses per step.
__device__ __forceinline__ unsigned bin_offset(unsigned bin) { return ((bin & 0xfeu) << 8) + ((bin & 1u) << 1); }
__device__ __forceinline__ unsigned offsets2(unsigned two_bins) {  // bins in bits 0-7 and 8-15
    return bin_offset(two_bins & 0xffu) | (bin_offset((two_bins >> 8) & 0xffu) << 16);
}
__device__ __forceinline__ unsigned lds_u16(unsigned addr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int ring_slot(int yr, int R) {
    int s = yr % R;
    return s < 0 ? s + R : s;
}
// slot of (row with slot s) + d, for -R < d < R, without a division
__device__ __forceinline__ int slot_add(int s, int d, int R) {
    int t = s + d;
    if (t >= R) t -= R;
    if (t < 0) t += R;
    return t;
}
// h[counter at offset] += delta on the calling thread's private histogram column (hbase =
// shared address of its word 0)
__device__ __forceinline__ void hist_bump(unsigned hbase, unsigned off, int delta) {
    const unsigned a = hbase + off;
    unsigned h;   // 32-bit registers: no PRMT juggling of 16-bit halves
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(h) : "r"(a));
    h += delta;
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(h));
}

// one voxel enters (counter offset oi), one leaves (oo): nothing to do when they are equal
// (the common case in flat regions), otherwise two predicated read-modify-writes
__device__ __forceinline__ void hist_move(unsigned hbase, unsigned oi, unsigned oo) {
    const unsigned ai = hbase + oi;
    const unsigned ao = hbase + oo;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .u32 h, g;\n\t"
        "setp.ne.u32 p, %0, %1;\n\t"
        "@p ld.shared.u16 h, [%0];\n\t"
        "@p ld.shared.u16 g, [%1];\n\t"
        "@p add.u32 h, h, 1;\n\t"
        "@p sub.u32 g, g, 1;\n\t"
        "@p st.shared.u16 [%0], h;\n\t"
        "@p st.shared.u16 [%1], g;\n\t"
        "}" ::"r"(ai), "r"(ao) : "memory");
}

// Per-thread view of the ring rows: thread t owns ring words t and t+128 of every row
// (a row is P planes x 12 words).  fetch() issues the global loads of one row into
// registers, commit() stores them into the ring two steps later (software prefetch, so
// the load latency hides behind the histogram work of the steps in between) and reports
// whether every in-image word equals ref_word (the row is "flat").
struct RingLoader {
    const uint8_t* src[2];  // address of x = xg in row y = 0 of the word's plane (nullptr: plane outside)
    int xg[2];              // x of the word's first byte (may be < 0 or >= nx)
    int ring_off[2];        // word index inside ring row slot 0, -1: thread owns no such word
    int nx, ny;
    unsigned dummy, ref_word, kbin;
    bool bytewise;

    __device__ __forceinline__ void init(const uint8_t* bv, const XrGeo& g, int x0, int z0, unsigned ref) {
        nx = g.nx; ny = g.ny;
        kbin = (unsigned)g.k;
        dummy = kbin * 0x01010101u;
        ref_word = ref;
        bytewise = (g.nx & 3) != 0;
        const int nwords = g.P * (kRingXW / 4);
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const int i = threadIdx.x + s * kXcThreads;
            src[s] = nullptr;
            ring_off[s] = -1;
            xg[s] = 0;
            if (i < nwords) {
                const int p = i / (kRingXW / 4), w = i % (kRingXW / 4);
                const int zz = z0 - g.wz + p;
                xg[s] = x0 - 8 + 4 * w;
                ring_off[s] = (p * g.R) * (kRingXW / 4) + w;
                if (zz >= 0 && zz < g.nz && xg[s] + 3 >= 0 && xg[s] < g.nx)
                    src[s] = bv + ((size_t)zz * g.ny) * g.nx;
            }
        }
    }
    // issue the global loads of row yr into v[] (nothing here waits for them)
    __device__ __forceinline__ void fetch(int yr, unsigned (&v)[2]) const {
        const bool row_ok = yr >= 0 && yr < ny;
#pragma unroll
        for (int s = 0; s < 2; s++) {
            v[s] = dummy;
            if (row_ok && src[s] != nullptr) {
                const uint8_t* a = src[s] + (size_t)yr * nx;
                if (!bytewise) {  // nx % 4 == 0: the word is entirely inside the row
                    v[s] = __ldg(reinterpret_cast<const unsigned*>(a + xg[s]));
                } else {
                    unsigned w = 0;
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const int xx = xg[s] + b;
                        w |= ((xx >= 0 && xx < nx) ? (unsigned)__ldg(a + xx) : kbin) << (8 * b);
                    }
                    v[s] = w;
                }
            }
        }
    }
    // store row yr into the ring; returns whether every in-image word of this thread equals
    // ref_word.  A row outside the image is all dummy and never flat (it changes the
    // histograms when it enters or leaves); columns/planes outside the image are dummy in
    // every row and are ignored.
    __device__ __forceinline__ bool commit(unsigned* ring32, int yr, int slot_of_yr, const unsigned (&v)[2]) const {
        const int slot = slot_of_yr * (kRingXW / 4);
        bool flat = yr >= 0 && yr < ny;
#pragma unroll
        for (int s = 0; s < 2; s++) {
            if (ring_off[s] < 0) continue;
            reinterpret_cast<uint2*>(ring32)[ring_off[s] + slot] = make_uint2(offsets2(v[s]), offsets2(v[s] >> 16));
            if (src[s] != nullptr) {
                const bool whole = xg[s] >= 0 && xg[s] + 3 < nx;
                flat = flat && whole && v[s] == ref_word;
            }
        }
        return flat;
    }
};

// grid (ceil(nx/32), ceil(nz/4)*nseg, nb), block 128.  Dynamic shared memory:
//   hist [nw][128] u32 | h2s [2*nw] u32 | cols [ncols] u32 | ring [P][R][48] u16
// cols[c] = byte offset of column c inside the ring relative to (plane = warp, row slot 0,
// x = lane): ((wz+dz)*R)*96 + 2*(8 + dx).
__global__ void __launch_bounds__(kXcThreads, 5)
xcorr_ring_kernel(const uint8_t* __restrict__ bins, float* __restrict__ out,
                  const unsigned* __restrict__ cols_g, const unsigned* __restrict__ h2,
                  const long long* __restrict__ vstat, const __grid_constant__ XrGeo g) {
    extern __shared__ __align__(16) unsigned smem[];
    unsigned* hist = smem;
    unsigned* h2s = smem + (size_t)g.nw * kXcThreads;
    unsigned* cols = h2s + ((2 * g.nw + 3) & ~3);
    unsigned* ring32 = cols + ((g.ncols + 3) & ~3);
    const int vol = blockIdx.z;
    for (int i = threadIdx.x; i < 2 * g.nw; i += kXcThreads) h2s[i] = i < g.k ? h2[(size_t)vol * 256 + i] : 0u;
    for (int i = threadIdx.x; i < g.ncols; i += kXcThreads) cols[i] = cols_g[i];
    // group bounds in shared memory: indexing the kernel parameter with a runtime e would
    // go through the (slow) indexed constant path
    __shared__ int gs[17], ge[17];
    if (threadIdx.x < 17) { gs[threadIdx.x] = g.gstart[threadIdx.x]; ge[threadIdx.x] = g.gend[threadIdx.x]; }

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = blockIdx.x * 32;
    const int zb = blockIdx.y / g.nseg, sg = blockIdx.y % g.nseg;
    const int z0 = zb * kXcWarps;
    const int x = x0 + lane, z = z0 + warp;
    const int y0 = sg * g.seg;
    const int y1 = min(g.ny, y0 + g.seg);
    const bool active = x < g.nx && z < g.nz;
    const size_t nvox = (size_t)g.nx * g.ny * g.nz;
    const uint8_t* bv = bins + (size_t)vol * nvox;
    float* ov = out + (size_t)vol * nvox;
    const long long n2 = vstat[2 * vol], d2 = vstat[2 * vol + 1];
    const unsigned K = (unsigned)g.k;
    const unsigned hbase = (unsigned)__cvta_generic_to_shared(hist) + threadIdx.x * 4;
    const unsigned tbase = (unsigned)__cvta_generic_to_shared(ring32) + warp * g.R * kRingXB + lane * 2;

    // Flat-region shortcut: when all 2*wy+2 ring rows a step touches hold one single value
    // (background, saturated tissue) the histogram cannot change, so the step is skipped and
    // the previous coefficient is stored again.  ref = bin of the first voxel of the block.
    unsigned ref_word;
    {
        const int zz = min(max(z0 - g.wz, 0), g.nz - 1), xx = min(max(x0 - 8, 0), g.nx - 1);
        const int yy = min(max(y0 - g.wy, 0), g.ny - 1);
        ref_word = (unsigned)__ldg(bv + ((size_t)zz * g.ny + yy) * g.nx + xx) * 0x01010101u;
    }
    RingLoader rl;
    rl.init(bv, g, x0, z0, ref_word);
    int flat_rows = 0;   // block-uniform: number of most recent consecutive flat rows
    const int R = g.R;
    int sy = ring_slot(y0, R);          // ring slot of row y (block-uniform, kept incrementally)
    // preload rows y0-wy .. y0+wy, four rows of loads in flight at a time
    {
        int sl = ring_slot(y0 - g.wy, R);
        for (int yr0 = y0 - g.wy; yr0 <= y0 + g.wy; yr0 += 4) {
            unsigned v[4][2];
#pragma unroll
            for (int q = 0; q < 4; q++) rl.fetch(yr0 + q <= y0 + g.wy ? yr0 + q : -1, v[q]);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (yr0 + q > y0 + g.wy) break;   // block-uniform
                const bool f = rl.commit(ring32, yr0 + q, sl, v[q]);
                sl = slot_add(sl, 1, R);
                flat_rows = __syncthreads_and(f) ? flat_rows + 1 : 0;
            }
        }
    }
    // software prefetch: rows y+1+wy and y+2+wy are already in registers when step y starts
    unsigned pfa[2], pfb[2];
    rl.fetch(y0 + 1 + g.wy, pfa);
    rl.fetch(y0 + 2 + g.wy, pfb);
    for (int w = 0; w < g.nw; w++) hist[w * kXcThreads + threadIdx.x] = 0u;
    __syncthreads();
    // words of h2 that hold a non-zero count (block-uniform; wlo > whi when the region is empty)
    int wlo = g.nw, whi = -1;
    for (int w = 0; w < g.nw; w++)
        if (h2s[2 * w] | h2s[2 * w + 1]) { wlo = min(wlo, w); whi = w; }
    const bool p32 = (unsigned long long)g.ball * (unsigned long long)n2 < 0xffffffffull;

    if (active) {
        if (flat_rows >= 2 * g.wy + 1) {
            // the whole neighbourhood of the first voxel is one value: every column adds its
            // 2e+1 voxels to one bin (ref, or the dummy bin for columns outside the image)
            for (int e = 0; e <= g.wy; e++)
                for (int c = gs[e]; c < ge[e]; c++)
                    hist_bump(hbase, lds_u16(tbase + sy * kRingXB + cols[c]), 2 * e + 1);
        } else {  // full ball around (x, y0, z)
            for (int e = 0; e <= g.wy; e++)
                for (int c = gs[e]; c < ge[e]; c++) {
                    const unsigned co = cols[c];
                    for (int dy = -e; dy <= e; dy++)
                        hist_bump(hbase, lds_u16(tbase + slot_add(sy, dy, R) * kRingXB + co), 1);
                }
        }
    }
    float res = 0.0f;
    bool changed = true;   // block-uniform: did the last slide touch the histograms?
    for (int y = y0; y < y1; y++) {
        if (active && changed) {
            // S11 over all bins; the dot product with h2 only over the words where h2 is not
            // zero (a region's intensities occupy a few of the k bins), in 32 bits when
            // ball * n2 cannot overflow them
            unsigned long long P = 0;
            unsigned S2 = 0;
            for (int w = 0; w < wlo; w++) {
                const unsigned v = hist[w * kXcThreads + threadIdx.x];
                const unsigned lo = v & 0xffffu, hi = v >> 16;
                S2 += lo * lo + hi * hi;
            }
            if (p32) {
                unsigned P32 = 0;
                for (int w = wlo; w <= whi; w++) {
                    const unsigned v = hist[w * kXcThreads + threadIdx.x];
                    const uint2 hh = *reinterpret_cast<const uint2*>(h2s + 2 * w);
                    const unsigned lo = v & 0xffffu, hi = v >> 16;
                    P32 += lo * hh.x + hi * hh.y;
                    S2 += lo * lo + hi * hi;
                }
                P = P32;
            } else {
                for (int w = wlo; w <= whi; w++) {
                    const unsigned v = hist[w * kXcThreads + threadIdx.x];
                    const uint2 hh = *reinterpret_cast<const uint2*>(h2s + 2 * w);
                    const unsigned lo = v & 0xffffu, hi = v >> 16;
                    P += (unsigned long long)lo * hh.x;
                    P += (unsigned long long)hi * hh.y;
                    S2 += lo * lo + hi * hi;
                }
            }
            for (int w = whi + 1; w < g.nw; w++) {
                const unsigned v = hist[w * kXcThreads + threadIdx.x];
                const unsigned lo = v & 0xffffu, hi = v >> 16;
                S2 += lo * lo + hi * hi;
            }
            // remove the dummy bin K (not-counted voxels) from the statistics
            const unsigned vd = hist[(K >> 1) * kXcThreads + threadIdx.x];
            const unsigned hd = (K & 1u) ? (vd >> 16) : (vd & 0xffffu);
            S2 -= hd * hd;
            const long long n1 = (long long)g.ball - (long long)hd;
            const long long num = (long long)K * (long long)P - n1 * n2;
            const long long d1 = (long long)K * (long long)S2 - n1 * n1;
            res = 0.0f;
            if (d1 > 0 && d2 > 0)
                res = (float)__ddiv_rn((double)num, __dsqrt_rn(__dmul_rn((double)d1, (double)d2)));
        }
        if (active) ov[((size_t)z * g.ny + y) * g.nx + x] = res;
        if (y + 1 >= y1) break;
        // the slot of row y+1+wy last held row y+1+wy-R < y-1-wy: nobody reads it any more
        const bool f = rl.commit(ring32, y + 1 + g.wy, slot_add(sy, 1 + g.wy, R), pfa);
        pfa[0] = pfb[0]; pfa[1] = pfb[1];
        rl.fetch(y + 3 + g.wy, pfb);
        flat_rows = __syncthreads_and(f) ? flat_rows + 1 : 0;
        changed = flat_rows < 2 * g.wy + 2;   // rows y-wy .. y+1+wy
        if (active && changed) {
            for (int e = 0; e <= g.wy; e++) {
                const unsigned pin = tbase + slot_add(sy, 1 + e, R) * kRingXB;
                const unsigned pout = tbase + slot_add(sy, -e, R) * kRingXB;
                int c = gs[e];
                const int cend = ge[e];
#pragma unroll 1
                for (; c + 4 <= cend; c += 4) {
                    const uint4 co = *reinterpret_cast<const uint4*>(cols + c);   // gstart is 4-aligned
                    const unsigned i0 = lds_u16(pin + co.x), o0 = lds_u16(pout + co.x);
                    const unsigned i1 = lds_u16(pin + co.y), o1 = lds_u16(pout + co.y);
                    const unsigned i2 = lds_u16(pin + co.z), o2 = lds_u16(pout + co.z);
                    const unsigned i3 = lds_u16(pin + co.w), o3 = lds_u16(pout + co.w);
                    hist_move(hbase, i0, o0);
                    hist_move(hbase, i1, o1);
                    hist_move(hbase, i2, o2);
                    hist_move(hbase, i3, o3);
                }
                for (; c < cend; c++) {
                    const unsigned co = cols[c];
                    const unsigned i0 = lds_u16(pin + co), o0 = lds_u16(pout + co);
                    hist_move(hbase, i0, o0);
                }
            }
        }
        sy = slot_add(sy, 1, R);
    }
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

static int upload_edges(OpGuard& og, int nb, const double* m, const double* M, int k, float** d_edges) {
    std::vector<float> e((size_t)nb * kEdgeStride, INFINITY);
    for (int v = 0; v < nb; v++) {
        if (v > 0 && m[v] == m[v - 1] && M[v] == M[v - 1])
            memcpy(&e[(size_t)v * kEdgeStride], &e[(size_t)(v - 1) * kEdgeStride], sizeof(float) * kEdgeStride);
        else
            compute_edges(m[v], M[v], k, &e[(size_t)v * kEdgeStride]);
    }
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(float) * e.size(), (void**)d_edges));
    // pageable source: staged by the runtime before the call returns
    VX_CUDA(cudaMemcpyAsync(*d_edges, e.data(), sizeof(float) * e.size(), cudaMemcpyHostToDevice, og.st));
    return VX_OK;
}

static dim3 stream_grid(size_t nvox, int nb) {
    int per_vol = (int)((nvox + 256 * 16 - 1) / (256 * 16));
    int cap = (kNumSMs * 8 + nb - 1) / nb;
    return dim3(per_vol < cap ? per_vol : (cap < 1 ? 1 : cap), nb);
}

// u8 bins of A (dummy bin k for values that are not counted)
static int xcorr_bins(OpGuard& og, Image* A, const float* d_edges, int32_t k, uint8_t** d_bins) {
    VX_TRY(scratch_alloc(og.c, og.s, A->count(), (void**)d_bins));
    {
        VX_LAUNCH(og.c, og.s, "xc_bin_kernel");
        xc_bin_kernel<<<stream_grid(A->nvox(), A->nb), 256, 0, og.st>>>((const float*)A->d, *d_bins, A->nvox(), d_edges, k);
    }
    return check_launch("xc_bin_kernel");
}

// h2 of `src` (f32 values, or u8 bins when binned) over region R into d_h2 ([nb][256] u32)
static int region_hist(OpGuard& og, const void* src, bool binned, Image* R, const float* d_edges, int32_t k,
                       unsigned* d_h2) {
    const int nb = R->nb;
    const size_t nvox = R->nvox();
    VX_CUDA(cudaMemsetAsync(d_h2, 0, sizeof(unsigned) * 256 * nb, og.st));
    {
        VX_LAUNCH(og.c, og.s, "xc_region_hist_kernel");
        if (binned)
            xc_region_hist_kernel<true><<<stream_grid(nvox, nb), 256, 0, og.st>>>(src, (const uint8_t*)R->d, nvox, d_edges, k, d_h2);
        else
            xc_region_hist_kernel<false><<<stream_grid(nvox, nb), 256, 0, og.st>>>(src, (const uint8_t*)R->d, nvox, d_edges, k, d_h2);
    }
    return check_launch("xc_region_hist_kernel");
}

// Shared back end: statistics of the region histogram d_h2 ([nb][256] u32 on the device, final)
// and the sliding-ball kernel over the bins.  Frees d_bins, d_edges and d_h2.
static int xcorr_run(OpGuard& og, Image* A, uint8_t* d_bins, float* d_edges, unsigned* d_h2, float rho, int32_t k,
                     vx_img* out) {
    std::vector<unsigned> cols;
    long long ball = 0;
    VX_REQUIRE(build_columns(A->sp, rho, &cols, &ball) && ball <= 65535 && cols.size() <= 8192,
               VX_E_UNSUPPORTED, "radius %g is too large for the sliding-histogram kernel", (double)rho);
    const int nb = A->nb;
    unsigned* d_cols = nullptr;
    long long* d_vstat = nullptr;
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(long long) * 2 * nb, (void**)&d_vstat));
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * (cols.size() + 128), (void**)&d_cols));
    {
        VX_LAUNCH(og.c, og.s, "xc_stat_kernel");
        xc_stat_kernel<<<nb, 1, 0, og.st>>>(d_h2, k, d_vstat);
    }
    VX_TRY(check_launch("xc_stat_kernel"));
    Image* O = nullptr;
    VX_TRY(new_image(og.c, VX_F32, A->nx, A->ny, A->nz, A->nb, A->sp, &O));
    // geometry of the ball
    int wx = 0, wy = 0, wz = 0;
    for (unsigned cw : cols) {
        int dx = (int)(signed char)(cw & 0xffu), dz = (int)(signed char)((cw >> 8) & 0xffu), ey = (int)(cw >> 16);
        wx = std::max(wx, std::abs(dx)); wz = std::max(wz, std::abs(dz)); wy = std::max(wy, ey);
    }
    const int nw = (k + 2) / 2;   // k bins + the dummy bin k
    // y segments: long enough to amortise the initial full-ball build, short enough to
    // give the batch at least a few waves of CTAs
    // (the build costs |ball| bumps, a step 2 * #columns moves: ~5% at 64 rows for rho = 5)
    int nseg = (A->ny + 63) / 64;
    {
        const long long ctas_per_seg = (long long)((A->nx + 31) / 32) * ((A->nz + kXcWarps - 1) / kXcWarps) * A->nb;
        const long long want = 4LL * kNumSMs * 5;   // four waves at five resident CTAs per SM
        while (nseg > 1 && ctas_per_seg * (nseg - 1) >= want) nseg--;
    }
    const int seg = (A->ny + nseg - 1) / nseg;
    nseg = (A->ny + seg - 1) / seg;
    const int ringR = 2 * wy + 3;   // rows y-wy .. y+1+wy of a step plus the one being written
    const int P = kXcWarps + 2 * wz;
    int rc = VX_OK;
    // ring offsets, sorted by y half-extent, every group padded to start 4-aligned
    std::vector<unsigned> ring_cols;
    XrGeo g;
    bool ring_ok = wx <= 8 && wy <= 16 && P * (kRingXW / 4) <= 2 * kXcThreads;
    if (ring_ok) {
        for (int e = 0; e <= wy; e++) {
            while (ring_cols.size() & 3) ring_cols.push_back(0);   // alignment filler, never read
            g.gstart[e] = (int)ring_cols.size();
            int cnt = 0;
            for (unsigned cw : cols) {
                int dx = (int)(signed char)(cw & 0xffu), dz = (int)(signed char)((cw >> 8) & 0xffu), ey = (int)(cw >> 16);
                if (ey != e) continue;
                ring_cols.push_back((unsigned)(((wz + dz) * ringR) * kRingXB + 2 * (8 + dx)));
                cnt++;
            }
            g.gend[e] = g.gstart[e] + cnt;
        }
    }
    const size_t smem_ring = sizeof(unsigned) * ((size_t)nw * kXcThreads + ((2 * nw + 3) & ~3) + ((ring_cols.size() + 3) & ~(size_t)3)) +
                             (size_t)P * ringR * kRingXB;
    if (ring_ok && smem_ring <= 160 * 1024) {
        g.nx = A->nx; g.ny = A->ny; g.nz = A->nz; g.k = k;
        g.ncols = (int)ring_cols.size();
        g.seg = seg; g.nseg = nseg; g.nw = nw; g.wy = wy; g.wz = wz; g.R = ringR; g.P = P;
        g.ball = (int)ball;
        VX_CUDA(cudaMemcpyAsync(d_cols, ring_cols.data(), sizeof(unsigned) * ring_cols.size(), cudaMemcpyHostToDevice, og.st));
        // 5 resident CTAs per SM (96 registers, ~44 KB): measured on B200 against 4 and 3 (extra
        // dynamic shared memory), alone and with other streams' kernels co-resident: 7.13 / 7.69 /
        // 8.98 ms per 16-volume launch, and the whole GTV step got slower with fewer CTAs as well
        const size_t smem_launch = smem_ring;
        if (smem_launch > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(xcorr_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_launch);
            if (e != cudaSuccess) rc = fail(VX_E_CUDA, "smem attribute: %s", cudaGetErrorString(e));
        }
        if (rc == VX_OK) {
            dim3 grid((A->nx + 31) / 32, ((A->nz + kXcWarps - 1) / kXcWarps) * nseg, nb);
            {
                VX_LAUNCH(og.c, og.s, "xcorr_kernel");
                xcorr_ring_kernel<<<grid, kXcThreads, smem_launch, og.st>>>(d_bins, (float*)O->d, d_cols, d_h2, d_vstat, g);
            }
            rc = check_launch("xcorr_kernel");
        }
    } else {
        XcGeo g;
        g.nx = A->nx; g.ny = A->ny; g.nz = A->nz; g.k = k;
        g.ncols = (int)cols.size();
        if (!cols.empty())
            VX_CUDA(cudaMemcpyAsync(d_cols, cols.data(), sizeof(unsigned) * cols.size(), cudaMemcpyHostToDevice, og.st));
        g.nw = nw; g.seg = seg; g.nseg = nseg;
        size_t smem = sizeof(unsigned) * ((size_t)g.nw * kXcThreads + 2 * g.nw + g.ncols);
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(xcorr_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) rc = fail(VX_E_CUDA, "smem attribute: %s", cudaGetErrorString(e));
        }
        if (rc == VX_OK) {
            dim3 grid((A->nx + 31) / 32, (A->nz * g.nseg + kXcWarps - 1) / kXcWarps, nb);
            {
                VX_LAUNCH(og.c, og.s, "xcorr_direct_kernel");
                xcorr_direct_kernel<<<grid, kXcThreads, smem, og.st>>>(d_bins, (float*)O->d, d_cols, d_h2, d_vstat, g);
            }
            rc = check_launch("xcorr_direct_kernel");
        }
    }
    if (rc == VX_OK) rc = scratch_free(og.c, og.s, d_edges);
    if (rc == VX_OK) rc = scratch_free(og.c, og.s, d_bins);
    if (rc == VX_OK) rc = scratch_free(og.c, og.s, d_h2);
    if (rc == VX_OK) rc = scratch_free(og.c, og.s, d_vstat);
    if (rc == VX_OK) rc = scratch_free(og.c, og.s, d_cols);
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
    VX_TRY(og.input(region, &R, VX_U8));
    VX_TRY(same_shape(A, B));
    VX_TRY(same_shape(A, R));
    float* d_edges = nullptr;
    unsigned* d_h2 = nullptr;
    uint8_t* d_bins = nullptr;
    VX_TRY(upload_edges(og, A->nb, m, M, k, &d_edges));
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * 256 * A->nb, (void**)&d_h2));
    VX_TRY(xcorr_bins(og, A, d_edges, k, &d_bins));
    // similarTo(a) correlates an image with itself: its bins are already there
    if (A == B) VX_TRY(region_hist(og, d_bins, true, R, d_edges, k, d_h2));
    else VX_TRY(region_hist(og, B->d, false, R, d_edges, k, d_h2));
    return xcorr_run(og, A, d_bins, d_edges, d_h2, rho, k, out);
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
    VX_TRY(og.input(region, &R, VX_U8));
    VX_TRY(same_shape(B, R));
    const int nb = B->nb;
    float* d_edges = nullptr;
    unsigned* d_h2 = nullptr;
    VX_TRY(upload_edges(og, nb, m, M, k, &d_edges));
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * 256 * nb, (void**)&d_h2));
    VX_TRY(region_hist(og, B->d, false, R, d_edges, k, d_h2));
    std::vector<unsigned> host((size_t)256 * nb);
    VX_CUDA(cudaMemcpyAsync(host.data(), d_h2, sizeof(unsigned) * 256 * nb, cudaMemcpyDeviceToHost, og.st));
    VX_CUDA(cudaStreamSynchronize(og.st));
    for (int v = 0; v < nb; v++)
        for (int i = 0; i < k; i++) h2[(size_t)v * k + i] = host[(size_t)v * 256 + i];
    VX_TRY(scratch_free(og.c, og.s, d_edges));
    VX_TRY(scratch_free(og.c, og.s, d_h2));
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
    float* d_edges = nullptr;
    unsigned* d_h2 = nullptr;
    uint8_t* d_bins = nullptr;
    VX_TRY(upload_edges(og, nb, m, M, k, &d_edges));
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * 256 * nb, (void**)&d_h2));
    // pageable source: cudaMemcpyAsync stages it before returning, `host` may die afterwards
    VX_CUDA(cudaMemcpyAsync(d_h2, host.data(), sizeof(unsigned) * 256 * nb, cudaMemcpyHostToDevice, og.st));
    VX_TRY(xcorr_bins(og, A, d_edges, k, &d_bins));
    return xcorr_run(og, A, d_bins, d_edges, d_h2, rho, k, out);
}

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
