// Latency and single-SM throughput of small rank-4 u8 TMA tile loads (the crossCorrelation ring
// chunk: box 64 x 4 x 16 x 1 = 4 KB in 64 rows of 64 bytes), measured with clock64 on B200.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>
#include "../../voxlogica_b200/csrc/vx_tma.h"
using namespace vx;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// depth loads in flight per block, n loads in total; out[block] = cycles
__global__ void lat(const __grid_constant__ CUtensorMap tmap, long long* out, int n, int depth, int bytes, int ny, int nz) {
    extern __shared__ unsigned char raw[];
    unsigned char* buf = raw + ((128u - (smem_u32(raw) & 127u)) & 127u);
    __shared__ __align__(8) unsigned long long bar[8];
    if (threadIdx.x == 0) { for (int i = 0; i < 8; i++) mbar_init(smem_u32(&bar[i]), 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    const int z0 = (blockIdx.x * 6) % (nz - 16);
    long long t0 = clock64();
    for (int j = 0; j < n + depth; j++) {
        if (j < n && lane == 0) {
            const int s = j % depth;
            mbar_arrive_expect_tx(smem_u32(&bar[s]), bytes);
            tma_load_4d(smem_u32(buf) + s * bytes, &tmap, smem_u32(&bar[s]), 32 * (blockIdx.x % 6), (4 * j) % (ny - 4), z0, blockIdx.x % 16);
        }
        const int jj = j - depth + 1;   // wait for the oldest once `depth` are in flight
        if (jj >= 0 && jj < n) mbar_wait(smem_u32(&bar[jj % depth]), (unsigned)(jj / depth) & 1u);
    }
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int nx = 240, ny = 240, nz = 155, nb = 16, pitch = 240;
    unsigned char* d; CK(cudaMalloc(&d, (size_t)pitch * ny * nz * nb)); CK(cudaMemset(d, 1, (size_t)pitch * ny * nz * nb));
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    long long* o; CK(cudaMalloc(&o, sizeof(long long) * 148 * 4));
    for (int promo = 0; promo < 2; promo++)
    for (int P = 16; P >= 4; P /= 4) {
        CUtensorMap tm;
        cuuint64_t gd[4] = {nx, ny, nz, nb}, gs[3] = {pitch, (cuuint64_t)pitch * ny, (cuuint64_t)pitch * ny * nz};
        cuuint32_t bx[4] = {64, 4, (cuuint32_t)P, 1}, es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         promo ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int bytes = 64 * 4 * P;
        for (int grid = 1; grid <= 296; grid *= 296)
            for (int depth = 1; depth <= 8; depth *= 2) {
                const int n = 200;
                CK(cudaFuncSetAttribute(lat, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4096 + 128));
                lat<<<grid, 64, depth * bytes + 128>>>(tm, o, n, depth, bytes, ny, nz);
                CK(cudaDeviceSynchronize());
                std::vector<long long> h(grid); CK(cudaMemcpy(h.data(), o, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
                double mean = 0; for (auto v : h) mean += (double)v; mean /= grid;
                printf("L2 promotion %s, box 64x4x%d (%d B, %d rows), %3d blocks, %d in flight: %.0f cycles per tile\n",
                       promo ? "none" : "128B", P, bytes, 4 * P, grid, depth, mean / n);
            }
    }
    return 0;
}
