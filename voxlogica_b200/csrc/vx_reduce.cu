// A7 volume, A11 min/max (+ masked variants, sum): per-volume scalar reductions.
// Greta.pdf p.43 "Volume, max, min (global)".  Streaming, HBM-bound: every input
// byte is read once with 128-bit loads; partials are combined in a fixed order so the
// integer results (volume) and min/max are exact and run-to-run deterministic.
#include <math.h>

#include "vx_internal.h"

namespace vx {

constexpr int kRedThreads = 256;

enum RedKind { RED_VOLUME = 0, RED_MIN = 1, RED_MAX = 2, RED_SUM = 3 };

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
        for (size_t i = tid; i < n16; i += nth) {
            uint4 w = __ldg(p + i);
            cnt += __popc(__vcmpne4(w.x, 0) & 0x01010101u) + __popc(__vcmpne4(w.y, 0) & 0x01010101u) +
                   __popc(__vcmpne4(w.z, 0) & 0x01010101u) + __popc(__vcmpne4(w.w, 0) & 0x01010101u);
        }
        // ragged head and tail of this volume
        for (size_t i = lo + tid; i < alo; i += nth) cnt += a[i] != 0;
        for (size_t i = ahi + tid; i < hi; i += nth) cnt += a[i] != 0;
    } else {
        for (size_t i = lo + tid; i < hi; i += nth) cnt += a[i] != 0;
    }
    double v = block_reduce<RED_SUM>((double)cnt);  // exact: counts < 2^53
    if (threadIdx.x == 0) partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = v;
}

template <int KIND, bool MASKED>
__global__ void __launch_bounds__(kRedThreads)
reduce_f32_kernel(const float* __restrict__ a, const uint8_t* __restrict__ m, size_t nvox,
                  double* __restrict__ partial) {
    const size_t lo = (size_t)blockIdx.y * nvox, hi = lo + nvox;
    const size_t alo = (lo + 3) & ~(size_t)3, ahi = hi & ~(size_t)3;
    double acc = red_identity(KIND);
    const size_t tid = (size_t)blockIdx.x * kRedThreads + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * kRedThreads;
    auto take = [&](float x, unsigned mk) {
        if (!MASKED || mk) acc = red_combine(KIND, acc, (double)x);
    };
    if (alo < ahi) {
        const float4* p = reinterpret_cast<const float4*>(a + alo);
        const size_t n4 = (ahi - alo) >> 2;
        for (size_t i = tid; i < n4; i += nth) {
            float4 w = __ldg(p + i);
            unsigned mk = 0x01010101u;
            if (MASKED) mk = *reinterpret_cast<const unsigned*>(m + alo + 4 * i);
            take(w.x, mk & 0xffu); take(w.y, (mk >> 8) & 0xffu);
            take(w.z, (mk >> 16) & 0xffu); take(w.w, mk >> 24);
        }
        for (size_t i = lo + tid; i < alo; i += nth) take(a[i], MASKED ? m[i] : 1);
        for (size_t i = ahi + tid; i < hi; i += nth) take(a[i], MASKED ? m[i] : 1);
    } else {
        for (size_t i = lo + tid; i < hi; i += nth) take(a[i], MASKED ? m[i] : 1);
    }
    double v = block_reduce<KIND>(acc);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = v;
}

// one block per volume, fixed combination order
template <int KIND>
__global__ void __launch_bounds__(kRedThreads)
reduce_final_kernel(const double* __restrict__ partial, int per_volume, double* __restrict__ out) {
    double acc = red_identity(KIND);
    for (int i = threadIdx.x; i < per_volume; i += kRedThreads)
        acc = red_combine(KIND, acc, partial[(size_t)blockIdx.x * per_volume + i]);
    double v = block_reduce<KIND>(acc);
    if (threadIdx.x == 0) out[blockIdx.x] = v;
}

// Runs the reduction and leaves nb doubles in device memory `d_out` (scratch owned by caller)
int reduce_device(OpGuard& g, int kind, const Image* a, const Image* mask, double* d_out) {
    const size_t nvox = a->nvox();
    int per_volume = (int)((nvox + (size_t)kRedThreads * 16 * 4 - 1) / ((size_t)kRedThreads * 16 * 4));
    int cap = (kNumSMs * 8 + a->nb - 1) / a->nb;
    if (per_volume > cap) per_volume = cap;
    if (per_volume < 1) per_volume = 1;
    double* d_part = nullptr;
    VX_TRY(scratch_alloc(g.c, g.s, sizeof(double) * (size_t)per_volume * a->nb, (void**)&d_part));
    dim3 grid(per_volume, a->nb);
    if (kind == RED_VOLUME) {
        { VX_LAUNCH(g.c, g.s, "reduce_volume_kernel");
          reduce_volume_kernel<<<grid, kRedThreads, 0, g.st>>>((const uint8_t*)a->d, nvox, d_part); }
        VX_TRY(check_launch("reduce_volume_kernel"));
        { VX_LAUNCH(g.c, g.s, "reduce_final_kernel");
          reduce_final_kernel<RED_SUM><<<a->nb, kRedThreads, 0, g.st>>>(d_part, per_volume, d_out); }
    } else {
        const float* p = (const float*)a->d;
        const uint8_t* m = mask ? (const uint8_t*)mask->d : nullptr;
        { VX_LAUNCH(g.c, g.s, "reduce_f32_kernel");
          if (kind == RED_MIN && !m) reduce_f32_kernel<RED_MIN, false><<<grid, kRedThreads, 0, g.st>>>(p, m, nvox, d_part);
          else if (kind == RED_MIN) reduce_f32_kernel<RED_MIN, true><<<grid, kRedThreads, 0, g.st>>>(p, m, nvox, d_part);
          else if (kind == RED_MAX && !m) reduce_f32_kernel<RED_MAX, false><<<grid, kRedThreads, 0, g.st>>>(p, m, nvox, d_part);
          else if (kind == RED_MAX) reduce_f32_kernel<RED_MAX, true><<<grid, kRedThreads, 0, g.st>>>(p, m, nvox, d_part);
          else if (!m) reduce_f32_kernel<RED_SUM, false><<<grid, kRedThreads, 0, g.st>>>(p, m, nvox, d_part);
          else reduce_f32_kernel<RED_SUM, true><<<grid, kRedThreads, 0, g.st>>>(p, m, nvox, d_part); }
        VX_TRY(check_launch("reduce_f32_kernel"));
        { VX_LAUNCH(g.c, g.s, "reduce_final_kernel");
          if (kind == RED_MIN) reduce_final_kernel<RED_MIN><<<a->nb, kRedThreads, 0, g.st>>>(d_part, per_volume, d_out);
          else if (kind == RED_MAX) reduce_final_kernel<RED_MAX><<<a->nb, kRedThreads, 0, g.st>>>(d_part, per_volume, d_out);
          else reduce_final_kernel<RED_SUM><<<a->nb, kRedThreads, 0, g.st>>>(d_part, per_volume, d_out); }
    }
    VX_TRY(check_launch("reduce_final_kernel"));
    return scratch_free(g.c, g.s, d_part);
}

static int reduce_to_host(vx_ctx ctx, int kind, vx_img a, vx_img mask, double* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    OpGuard g;
    VX_TRY(g.begin(ctx));
    Image *A = nullptr, *M = nullptr;
    VX_TRY(g.input(a, &A, kind == RED_VOLUME ? VX_U8 : VX_F32));
    if (mask) {
        VX_TRY(g.input(mask, &M, VX_U8));
        VX_TRY(same_shape(A, M));
    }
    double* d_out = nullptr;
    VX_TRY(scratch_alloc(g.c, g.s, sizeof(double) * A->nb, (void**)&d_out));
    VX_TRY(reduce_device(g, kind, A, M, d_out));
    VX_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * A->nb, cudaMemcpyDeviceToHost, g.st));
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
int32_t vx_min_masked(vx_ctx ctx, vx_img a, vx_img mask, double* out) {
    VX_REQUIRE(mask != 0, VX_E_INVALID_ARG, "mask handle is 0");
    return reduce_to_host(ctx, RED_MIN, a, mask, out);
}
int32_t vx_max_masked(vx_ctx ctx, vx_img a, vx_img mask, double* out) {
    VX_REQUIRE(mask != 0, VX_E_INVALID_ARG, "mask handle is 0");
    return reduce_to_host(ctx, RED_MAX, a, mask, out);
}
}
