// TMA (cp.async.bulk.tensor) and mbarrier helpers for sm_100a, shared by the kernels that
// stage image tiles through shared memory with the copy engine instead of per-thread loads.
// Host side: tensor maps are encoded with cuTensorMapEncodeTiled, fetched from the driver at
// run time (the library links the static runtime only, not libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vx {

// Tiled tensor map over a dense u8 tensor of rank 4 (x fastest).  strides[i] = byte stride of
// dimension i+1, every one a multiple of 16; box = tile extent per dimension (box[0] a multiple of
// 16 bytes).  Out-of-range elements of a tile are filled with zeros.
int tma_encode_u8_4d(CUtensorMap* map, const void* base, const uint64_t dims[4], const uint64_t strides[3],
                     const uint32_t box[4]);

#ifdef __CUDACC__
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// makes the initialised barriers visible to the async proxy (the copy engine) before first use
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// blocks until the phase with the given parity has completed (try_wait sleeps in hardware)
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// the same for a thread nobody is waiting for (a copy-issuing thread whose slots free up well ahead
// of need): back off between polls so that the polling does not take issue slots from the workers
__device__ __forceinline__ void mbar_wait_relaxed(unsigned bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(200);
}
// one tile of a rank-4 tensor map into shared memory; completes `bytes of the box` on the barrier
__device__ __forceinline__ void tma_load_4d(unsigned dst, const CUtensorMap* map, unsigned bar, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
#endif

}  // namespace vx
