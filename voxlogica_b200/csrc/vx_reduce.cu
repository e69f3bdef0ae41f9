// A7 volume, A11 min/max (+ masked variants, sum): per-volume scalar reductions.
// Greta.pdf p.43 "Volume, max, min (global)".  Streaming, HBM-bound: every input
// byte is read once with 128-bit loads; partials are combined in a fixed order so the
// integer results (volume) and min/max are exact and run-to-run deterministic.
#include <math.h>

#include "vx_internal.h"

namespace vx {

constexpr int kRedThreads = 256;

enum RedKind { RED_VOLUME = 0, RED_MIN = 1, RED_MAX = 2, RED_SUM = 3, RED_MINMAX = 4 };

__device__ __forceinline__ double red_identity(int kind) {
    return kind == RED_MIN ? (double)INFINITY : kind == RED_MAX ? -(double)INFINITY : 0.0;
}
__device__ __forceinline__ double red_combine(int kind, double a, double b) {
    if (kind == RED_MIN) return fmin(a, b);
    if (kind == RED_MAX) return fmax(a, b);
    return a + b;
}

template <int KIND>
__device__ __forceinline__ double block_reduce(double v) {
    __shared__ double sh[kRedThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = red_combine(KIND, v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = threadIdx.x < kRedThreads / 32 ? sh[threadIdx.x] : red_identity(KIND);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) v = red_combine(KIND, v, __shfl_xor_sync(0xffffffffu, v, o));
    }
    return v;
}

// grid = (blocks_per_volume, nb).  u8 volume count.
__device__ __forceinline__ unsigned count16(uint4 w) {
    return __popc(__vcmpne4(w.x, 0) & 0x01010101u) + __popc(__vcmpne4(w.y, 0) & 0x01010101u) +
           __popc(__vcmpne4(w.z, 0) & 0x01010101u) + __popc(__vcmpne4(w.w, 0) & 0x01010101u);
}
__global__ void __launch_bounds__(kRedThreads)
reduce_volume_kernel(const uint8_t* __restrict__ a, size_t nvox, double* __restrict__ partial) {
    const size_t lo = (size_t)blockIdx.y * nvox, hi = lo + nvox;
    const size_t alo = (lo + 15) & ~(size_t)15, ahi = hi & ~(size_t)15;
    unsigned long long cnt = 0;
    const size_t tid = (size_t)blockIdx.x * kRedThreads + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * kRedThreads;
    if (alo < ahi) {
        const uint4* p = reinterpret_cast<const uint4*>(a + alo);
        const size_t n16 = (ahi - alo) >> 4;
        for (size_t i = tid; i < n16; i += nth) cnt += count16(__ldg(p + i));
        // ragged head and tail of this volume
        for (size_t j = lo + tid; j < alo; j += nth) cnt += a[j] != 0;
        for (size_t j = ahi + tid; j < hi; j += nth) cnt += a[j] != 0;
    } else {
        for (size_t j = lo + tid; j < hi; j += nth) cnt += a[j] != 0;
    }
    double v = block_reduce<RED_SUM>((double)cnt);  // exact: counts < 2^53
    if (threadIdx.x == 0) partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = v;
}

// the same count on the bit form: bits [vol * nvox, (vol + 1) * nvox) of the stream
__global__ void __launch_bounds__(kRedThreads)
reduce_volume_bits_kernel(const unsigned* __restrict__ bits, size_t nvox, double* __restrict__ partial) {
    const size_t lo = (size_t)blockIdx.y * nvox, hi = lo + nvox;
    const size_t w0 = lo >> 5, w1 = (hi - 1) >> 5;
    unsigned long long cnt = 0;
    for (size_t w = w0 + (size_t)blockIdx.x * kRedThreads + threadIdx.x; w <= w1; w += (size_t)gridDim.x * kRedThreads) {
        unsigned x = __ldg(bits + w);
        if (w == w0) x &= 0xffffffffu << (lo & 31);
        if (w == w1 && (hi & 31)) x &= (1u << (hi & 31)) - 1u;
        cnt += __popc(x);
    }
    double v = block_reduce<RED_SUM>((double)cnt);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = v;
}

// Accumulators of the f32 reductions.  min / max are exact in binary32 (fminf / fmaxf skip
// NaNs like their binary64 versions on the converted values would); the sum accumulates in
// binary64 in a fixed order.  RED_MINMAX produces both extrema from ONE read of the image.
template <int KIND>
struct Acc {
    float mn = INFINITY, mx = -INFINITY;
    double sum = 0.0;
    __device__ __forceinline__ void take(float x) {
        if (KIND == RED_MIN || KIND == RED_MINMAX) mn = fminf(mn, x);
        if (KIND == RED_MAX || KIND == RED_MINMAX) mx = fmaxf(mx, x);
        if (KIND == RED_SUM) sum += (double)x;
    }
    __device__ __forceinline__ void take4(float4 w, unsigned mk) {
        if (mk & 0xffu) take(w.x);
        if ((mk >> 8) & 0xffu) take(w.y);
        if ((mk >> 16) & 0xffu) take(w.z);
        if (mk >> 24) take(w.w);
    }
};

// partial layout: [nb][per_volume] (RED_MINMAX: the maxima follow at + nb * per_volume)
template <int KIND, bool MASKED>
__global__ void __launch_bounds__(kRedThreads)
reduce_f32_kernel(const float* __restrict__ a, const uint8_t* __restrict__ m, size_t nvox,
                  double* __restrict__ partial) {
    const size_t lo = (size_t)blockIdx.y * nvox, hi = lo + nvox;
    const size_t alo = (lo + 3) & ~(size_t)3, ahi = hi & ~(size_t)3;
    Acc<KIND> acc;
    const size_t tid = (size_t)blockIdx.x * kRedThreads + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * kRedThreads;
    auto mask4 = [&](size_t i) -> unsigned {
        return MASKED ? __ldg(reinterpret_cast<const unsigned*>(m + alo + 4 * i)) : 0x01010101u;
    };
    if (alo < ahi) {
        const float4* p = reinterpret_cast<const float4*>(a + alo);
        const size_t n4 = (ahi - alo) >> 2;
        size_t i = tid;
        for (; i + 3 * nth < n4; i += 4 * nth) {   // four independent loads in flight
            const float4 w0 = __ldg(p + i), w1 = __ldg(p + i + nth), w2 = __ldg(p + i + 2 * nth),
                         w3 = __ldg(p + i + 3 * nth);
            const unsigned k0 = mask4(i), k1 = mask4(i + nth), k2 = mask4(i + 2 * nth), k3 = mask4(i + 3 * nth);
            acc.take4(w0, k0); acc.take4(w1, k1); acc.take4(w2, k2); acc.take4(w3, k3);
        }
        for (; i < n4; i += nth) acc.take4(__ldg(p + i), mask4(i));
        for (size_t j = lo + tid; j < alo; j += nth) if (!MASKED || m[j]) acc.take(a[j]);
        for (size_t j = ahi + tid; j < hi; j += nth) if (!MASKED || m[j]) acc.take(a[j]);
    } else {
        for (size_t j = lo + tid; j < hi; j += nth) if (!MASKED || m[j]) acc.take(a[j]);
    }
    const size_t slot = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    if (KIND == RED_MIN || KIND == RED_MINMAX) {
        const double v = block_reduce<RED_MIN>((double)acc.mn);
        if (threadIdx.x == 0) partial[slot] = v;
    }
    if (KIND == RED_MINMAX) __syncthreads();   // block_reduce's staging array is reused
    if (KIND == RED_MAX || KIND == RED_MINMAX) {
        const double v = block_reduce<RED_MAX>((double)acc.mx);
        if (threadIdx.x == 0) partial[slot + (KIND == RED_MINMAX ? (size_t)gridDim.y * gridDim.x : 0)] = v;
    }
    if (KIND == RED_SUM) {
        const double v = block_reduce<RED_SUM>(acc.sum);
        if (threadIdx.x == 0) partial[slot] = v;
    }
}

// one block per (volume, result), fixed combination order.  grid (nb, nres): result r of kind
// kinds[r] reads partial + r * nb * per_volume and writes out + r * nb.
template <int KIND>
__device__ __forceinline__ double final_of(const double* __restrict__ part, int per_volume) {
    double acc = red_identity(KIND);
    for (int i = threadIdx.x; i < per_volume; i += kRedThreads) acc = red_combine(KIND, acc, part[i]);
    return block_reduce<KIND>(acc);
}
__global__ void __launch_bounds__(kRedThreads)
reduce_final_kernel(const double* __restrict__ partial, int per_volume, double* __restrict__ out, int kind0, int kind1) {
    const int kind = blockIdx.y == 0 ? kind0 : kind1;
    const double* part = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * per_volume;
    double v;
    if (kind == RED_MIN) v = final_of<RED_MIN>(part, per_volume);
    else if (kind == RED_MAX) v = final_of<RED_MAX>(part, per_volume);
    else v = final_of<RED_SUM>(part, per_volume);
    if (threadIdx.x == 0) out[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = v;
}

// min / max of a quantised image (the tag of vx_internal.h) without reading it: the scaling is
// monotone, so the extrema are the values of the smallest and the largest code of each volume, which
// the upload already recorded.  out: nb minima (want & 1), then nb maxima (want & 2), as reduce_device.
__global__ void quant_extrema_kernel(const int* __restrict__ qrange, int nb, int is_signed, int increasing, float slope,
                                     float inter, int want, double* __restrict__ out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nb) return;
    const int bias = is_signed ? 32768 : 0;
    const float flo = __fadd_rn(__fmul_rn((float)(qrange[2 * v] - bias), slope), inter);
    const float fhi = __fadd_rn(__fmul_rn((float)(qrange[2 * v + 1] - bias), slope), inter);
    const float mn = increasing ? flo : fhi, mx = increasing ? fhi : flo;
    int k = 0;
    if (want & 1) out[(k++) * nb + v] = (double)mn;
    if (want & 2) out[k * nb + v] = (double)mx;
}

// Runs the reduction and leaves nb doubles (RED_MINMAX: nb minima, then nb maxima) in device
// memory `d_out` (scratch owned by the caller)
int reduce_device(OpGuard& g, int kind, const Image* a, const Image* mask, double* d_out, const unsigned* a_bits) {
    const size_t nvox = a->nvox();
    if (a->has_codes() && mask == nullptr && g.c->quantised_paths.load() &&
        (kind == RED_MIN || kind == RED_MAX || kind == RED_MINMAX)) {
        const int want = kind == RED_MIN ? 1 : kind == RED_MAX ? 2 : 3;
        { VX_LAUNCH(g.c, g.s, "quant_extrema_kernel");
          quant_extrema_kernel<<<(a->nb + 127) / 128, 128, 0, g.st>>>(a->q_range(), a->nb, a->q_dtype == VX_I16, a->q_increasing ? 1 : 0,
                                                                   a->q_slope, a->q_inter, want, d_out); }
        return check_launch("quant_extrema_kernel");
    }
    int per_volume = (int)((nvox + (size_t)kRedThreads * 16 * 4 - 1) / ((size_t)kRedThreads * 16 * 4));
    int cap = (kNumSMs * 8 + a->nb - 1) / a->nb;
    if (per_volume > cap) per_volume = cap;
    if (per_volume < 1) per_volume = 1;
    const int nres = kind == RED_MINMAX ? 2 : 1;
    double* d_part = nullptr;
    VX_TRY(scratch_alloc(g.c, g.s, sizeof(double) * (size_t)per_volume * a->nb * nres, (void**)&d_part));
    dim3 grid(per_volume, a->nb);
    int k0 = RED_SUM, k1 = RED_SUM;
    if (kind == RED_VOLUME) {
        { VX_LAUNCH(g.c, g.s, "reduce_volume_kernel");
          if (a_bits) reduce_volume_bits_kernel<<<grid, kRedThreads, 0, g.st>>>(a_bits, nvox, d_part);
          else reduce_volume_kernel<<<grid, kRedThreads, 0, g.st>>>((const uint8_t*)a->d, nvox, d_part); }
        VX_TRY(check_launch("reduce_volume_kernel"));
    } else {
        const float* p = (const float*)a->d;
        const uint8_t* m = mask ? (const uint8_t*)mask->d : nullptr;
        { VX_LAUNCH(g.c, g.s, "reduce_f32_kernel");
          if (kind == RED_MIN && !m) reduce_f32_kernel<RED_MIN, false><<<grid, kRedThreads, 0, g.st>>>(p, m, nvox, d_part);
          else if (kind == RED_MIN) reduce_f32_kernel<RED_MIN, true><<<grid, kRedThreads, 0, g.st>>>(p, m, nvox, d_part);
          else if (kind == RED_MAX && !m) reduce_f32_kernel<RED_MAX, false><<<grid, kRedThreads, 0, g.st>>>(p, m, nvox, d_part);
          else if (kind == RED_MAX) reduce_f32_kernel<RED_MAX, true><<<grid, kRedThreads, 0, g.st>>>(p, m, nvox, d_part);
          else if (kind == RED_MINMAX && !m) reduce_f32_kernel<RED_MINMAX, false><<<grid, kRedThreads, 0, g.st>>>(p, m, nvox, d_part);
          else if (kind == RED_MINMAX) reduce_f32_kernel<RED_MINMAX, true><<<grid, kRedThreads, 0, g.st>>>(p, m, nvox, d_part);
          else if (!m) reduce_f32_kernel<RED_SUM, false><<<grid, kRedThreads, 0, g.st>>>(p, m, nvox, d_part);
          else reduce_f32_kernel<RED_SUM, true><<<grid, kRedThreads, 0, g.st>>>(p, m, nvox, d_part); }
        VX_TRY(check_launch("reduce_f32_kernel"));
        k0 = kind == RED_MINMAX ? RED_MIN : kind;
        k1 = RED_MAX;
    }
    { VX_LAUNCH(g.c, g.s, "reduce_final_kernel");
      reduce_final_kernel<<<dim3(a->nb, nres), kRedThreads, 0, g.st>>>(d_part, per_volume, d_out, k0, k1); }
    VX_TRY(check_launch("reduce_final_kernel"));
    return scratch_free(g.c, g.s, d_part);
}

static int reduce_to_host(vx_ctx ctx, int kind, vx_img a, vx_img mask, double* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    OpGuard g;
    VX_TRY(g.begin(ctx));
    Image *A = nullptr, *M = nullptr;
    const unsigned* Ab = nullptr;
    if (kind == RED_VOLUME) VX_TRY(g.input_bits(a, &A, &Ab));   // a popcount over the bit form
    else VX_TRY(g.input(a, &A, VX_F32));
    if (mask) {
        VX_TRY(g.input(mask, &M, VX_U8));
        VX_TRY(same_shape(A, M));
    }
    double* d_out = nullptr;
    VX_TRY(scratch_alloc(g.c, g.s, sizeof(double) * A->nb, (void**)&d_out));
    VX_TRY(reduce_device(g, kind, A, M, d_out, Ab));
    VX_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * A->nb, cudaMemcpyDeviceToHost, g.st));
    VX_TRY(scratch_free(g.c, g.s, d_out));
    VX_CUDA(cudaStreamSynchronize(g.st));
    return g.finish_scalar();
}

// min and max of every volume from one pass over the image (crossCorrelation's histogram range
// min(img) .. max(img) costs one read of the image instead of two)
static int minmax_to_host(vx_ctx ctx, vx_img a, vx_img mask, double* out_min, double* out_max) {
    VX_REQUIRE(out_min != nullptr && out_max != nullptr, VX_E_INVALID_ARG, "out is NULL");
    OpGuard g;
    VX_TRY(g.begin(ctx));
    Image *A = nullptr, *M = nullptr;
    VX_TRY(g.input(a, &A, VX_F32));
    if (mask) {
        VX_TRY(g.input(mask, &M, VX_U8));
        VX_TRY(same_shape(A, M));
    }
    double* d_out = nullptr;
    VX_TRY(scratch_alloc(g.c, g.s, sizeof(double) * A->nb * 2, (void**)&d_out));
    VX_TRY(reduce_device(g, RED_MINMAX, A, M, d_out, nullptr));
    VX_CUDA(cudaMemcpyAsync(out_min, d_out, sizeof(double) * A->nb, cudaMemcpyDeviceToHost, g.st));
    VX_CUDA(cudaMemcpyAsync(out_max, d_out + A->nb, sizeof(double) * A->nb, cudaMemcpyDeviceToHost, g.st));
    VX_TRY(scratch_free(g.c, g.s, d_out));
    VX_CUDA(cudaStreamSynchronize(g.st));
    return g.finish_scalar();
}

}  // namespace vx

using namespace vx;

extern "C" {
int32_t vx_volume(vx_ctx ctx, vx_img a, double* out) { return reduce_to_host(ctx, RED_VOLUME, a, 0, out); }
int32_t vx_min(vx_ctx ctx, vx_img a, double* out) { return reduce_to_host(ctx, RED_MIN, a, 0, out); }
int32_t vx_max(vx_ctx ctx, vx_img a, double* out) { return reduce_to_host(ctx, RED_MAX, a, 0, out); }
int32_t vx_sum(vx_ctx ctx, vx_img a, double* out) { return reduce_to_host(ctx, RED_SUM, a, 0, out); }
int32_t vx_minmax(vx_ctx ctx, vx_img a, vx_img mask, double* out_min, double* out_max) {
    return minmax_to_host(ctx, a, mask, out_min, out_max);
}
int32_t vx_min_masked(vx_ctx ctx, vx_img a, vx_img mask, double* out) {
    VX_REQUIRE(mask != 0, VX_E_INVALID_ARG, "mask handle is 0");
    return reduce_to_host(ctx, RED_MIN, a, mask, out);
}
int32_t vx_max_masked(vx_ctx ctx, vx_img a, vx_img mask, double* out) {
    VX_REQUIRE(mask != 0, VX_E_INVALID_ARG, "mask handle is 0");
    return reduce_to_host(ctx, RED_MAX, a, mask, out);
}
}
