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

#include <vector>

#include "vx_internal.h"

namespace vx {

constexpr int kXcThreads = 128;
constexpr int kXcWarps = kXcThreads / 32;
constexpr int kXcMaxBins = 254;  // bins are stored as u8, 255 = not counted

__host__ __device__ __forceinline__ int xc_bin_of(float v, double m, double M, int k) {
    double dv = (double)v;
    if (!(dv >= m) || !(dv < M)) return 255;
#ifdef __CUDA_ARCH__
    double t = __ddiv_rn(__dmul_rn(__dsub_rn(dv, m), (double)k), __dsub_rn(M, m));
#else
    double t = ((dv - m) * (double)k) / (M - m);
#endif
    int i = (int)t;
    if (i >= k) i = k - 1;
    return i;
}

// grid (blocks, nb)
__global__ void __launch_bounds__(256)
xc_bin_kernel(const float* __restrict__ a, uint8_t* __restrict__ bins, size_t nvox,
              const double* __restrict__ mM, int nb, int k) {
    const size_t vol = blockIdx.y;
    const double m = mM[vol], M = mM[nb + vol];
    const float* src = a + vol * nvox;
    uint8_t* dst = bins + vol * nvox;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvox; i += (size_t)gridDim.x * 256)
        dst[i] = (uint8_t)xc_bin_of(src[i], m, M, k);
}

// histogram of b over region: grid (blocks, nb), h2 is [nb][256] (zeroed by the caller)
__global__ void __launch_bounds__(256)
xc_region_hist_kernel(const float* __restrict__ b, const uint8_t* __restrict__ region, size_t nvox,
                      const double* __restrict__ mM, int nb, int k, unsigned* __restrict__ h2) {
    __shared__ unsigned sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const size_t vol = blockIdx.y;
    const double m = mM[vol], M = mM[nb + vol];
    const float* src = b + vol * nvox;
    const uint8_t* reg = region + vol * nvox;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvox; i += (size_t)gridDim.x * 256) {
        if (reg[i] == 0) continue;
        int bin = xc_bin_of(src[i], m, M, k);
        if (bin < k) atomicAdd(&sh[bin], 1u);
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
xcorr_kernel(const uint8_t* __restrict__ bins, float* __restrict__ out,
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
    std::vector<unsigned> cols;
    long long ball = 0;
    VX_REQUIRE(build_columns(A->sp, rho, &cols, &ball) && ball <= 65535 && cols.size() <= 8192,
               VX_E_UNSUPPORTED, "radius %g is too large for the sliding-histogram kernel", (double)rho);
    const int nb = A->nb;
    const size_t nvox = A->nvox(), total = A->count();
    // device scratch: m/M, bins, h2, vstat, columns
    double* d_mM = nullptr;
    uint8_t* d_bins = nullptr;
    unsigned *d_h2 = nullptr, *d_cols = nullptr;
    long long* d_vstat = nullptr;
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(double) * 2 * nb, (void**)&d_mM));
    VX_TRY(scratch_alloc(og.c, og.s, total, (void**)&d_bins));
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * 256 * nb, (void**)&d_h2));
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(long long) * 2 * nb, (void**)&d_vstat));
    VX_TRY(scratch_alloc(og.c, og.s, sizeof(unsigned) * (cols.size() + 1), (void**)&d_cols));
    VX_CUDA(cudaMemcpyAsync(d_mM, m, sizeof(double) * nb, cudaMemcpyHostToDevice, og.st));
    VX_CUDA(cudaMemcpyAsync(d_mM + nb, M, sizeof(double) * nb, cudaMemcpyHostToDevice, og.st));
    if (!cols.empty())
        VX_CUDA(cudaMemcpyAsync(d_cols, cols.data(), sizeof(unsigned) * cols.size(), cudaMemcpyHostToDevice, og.st));
    VX_CUDA(cudaMemsetAsync(d_h2, 0, sizeof(unsigned) * 256 * nb, og.st));
    int per_vol = (int)((nvox + 256 * 8 - 1) / (256 * 8));
    int cap = (kNumSMs * 8 + nb - 1) / nb;
    dim3 sgrid(per_vol < cap ? per_vol : cap, nb);
    {
        VX_LAUNCH(og.c, og.s, "xc_bin_kernel");
        xc_bin_kernel<<<sgrid, 256, 0, og.st>>>((const float*)A->d, d_bins, nvox, d_mM, nb, k);
    }
    VX_TRY(check_launch("xc_bin_kernel"));
    {
        VX_LAUNCH(og.c, og.s, "xc_region_hist_kernel");
        xc_region_hist_kernel<<<sgrid, 256, 0, og.st>>>((const float*)B->d, (const uint8_t*)R->d, nvox, d_mM, nb, k, d_h2);
    }
    VX_TRY(check_launch("xc_region_hist_kernel"));
    {
        VX_LAUNCH(og.c, og.s, "xc_stat_kernel");
        xc_stat_kernel<<<nb, 1, 0, og.st>>>(d_h2, k, d_vstat);
    }
    VX_TRY(check_launch("xc_stat_kernel"));
    Image* O = nullptr;
    VX_TRY(new_image(og.c, VX_F32, A->nx, A->ny, A->nz, A->nb, A->sp, &O));
    XcGeo g;
    g.nx = A->nx; g.ny = A->ny; g.nz = A->nz; g.k = k;
    g.ncols = (int)cols.size();
    g.nw = (k + 1) / 2;
    // y segments: long enough to amortise the initial full-ball build, short enough to
    // give the batch at least a few waves of CTAs
    g.seg = A->ny <= 64 ? A->ny : 64;
    g.nseg = (A->ny + g.seg - 1) / g.seg;
    size_t smem = sizeof(unsigned) * ((size_t)g.nw * kXcThreads + 2 * g.nw + g.ncols);
    int rc = VX_OK;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(xcorr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) rc = fail(VX_E_CUDA, "smem attribute: %s", cudaGetErrorString(e));
    }
    if (rc == VX_OK) {
        dim3 grid((A->nx + 31) / 32, (A->nz * g.nseg + kXcWarps - 1) / kXcWarps, nb);
        {
            VX_LAUNCH(og.c, og.s, "xcorr_kernel");
            xcorr_kernel<<<grid, kXcThreads, smem, og.st>>>(d_bins, (float*)O->d, d_cols, d_h2, d_vstat, g);
        }
        rc = check_launch("xcorr_kernel");
    }
    if (rc == VX_OK) rc = scratch_free(og.c, og.s, d_mM);
    if (rc == VX_OK) rc = scratch_free(og.c, og.s, d_bins);
    if (rc == VX_OK) rc = scratch_free(og.c, og.s, d_h2);
    if (rc == VX_OK) rc = scratch_free(og.c, og.s, d_vstat);
    if (rc == VX_OK) rc = scratch_free(og.c, og.s, d_cols);
    if (rc != VX_OK) { drop(O); return rc; }
    return og.finish(O, out);
}
