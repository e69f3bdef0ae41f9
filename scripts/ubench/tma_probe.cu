// Probe: one rank-4 u8 TMA tile load (box 48 x 4 x P x 1, coordinates that may be negative) into
// shared memory, checked against the host copy.  Used to pin down the constraints of the copy
// engine before relying on it in vx_xcorr.cu.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>
#include "../../voxlogica_b200/csrc/vx_tma.h"
using namespace vx;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void probe(const __grid_constant__ CUtensorMap tmap, unsigned char* out, int bytes, int c0, int c1, int c2, int c3, int mode) {
    extern __shared__ unsigned char raw[];
    unsigned char* buf = raw + ((128u - (smem_u32(raw) & 127u)) & 127u);
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    __syncthreads();
    if (mode == 0) {
        if (threadIdx.x == 0) { mbar_arrive_expect_tx(smem_u32(&bar), bytes); tma_load_4d(smem_u32(buf), &tmap, smem_u32(&bar), c0, c1, c2, c3); }
    } else {
        if (threadIdx.x < 32) {
            unsigned pred = 0;
            asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
            if (pred) { mbar_arrive_expect_tx(smem_u32(&bar), bytes); tma_load_4d(smem_u32(buf), &tmap, smem_u32(&bar), c0, c1, c2, c3); }
        }
    }
    mbar_wait(smem_u32(&bar), 0);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = buf[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int nx = 48, ny = 20, nz = 9, nb = 2, pitch = 48;
    std::vector<unsigned char> h((size_t)pitch * ny * nz * nb);
    for (size_t i = 0; i < h.size(); i++) h[i] = (unsigned char)(1 + (i * 7 + i / 48) % 200);
    unsigned char* d; CK(cudaMalloc(&d, h.size())); CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const int v0 = argc > 1 ? atoi(argv[1]) : 0;
    for (int var = v0; var <= v0; var++) {
        const int inner = (var & 1) ? 64 : 48;
        const int mode = (var >> 1) & 1;
        const int c0 = (var & 4) ? 0 : -8;
        const int cz = (var & 8) ? 0 : -2;
        const int P = 10;
        CUtensorMap tm;
        cuuint64_t gd[4] = {nx, ny, nz, nb}, gs[3] = {pitch, (cuuint64_t)pitch * ny, (cuuint64_t)pitch * ny * nz};
        cuuint32_t bx[4] = {(cuuint32_t)inner, 4, (cuuint32_t)P, 1}, es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("variant %d: inner %d, %s, c0 %d: encode -> %d\n", var, inner, mode ? "elect.sync" : "lane0", c0, (int)r);
        if (r != CUDA_SUCCESS) continue;
        const int bytes = inner * 4 * P;
        unsigned char* o; CK(cudaMalloc(&o, bytes));
        probe<<<1, 64, bytes + 128>>>(tm, o, bytes, c0, cz, cz, 1, mode);
        cudaError_t e = cudaDeviceSynchronize();
        printf("  kernel: %s\n", cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<unsigned char> g(bytes);
        CK(cudaMemcpy(g.data(), o, bytes, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int p = 0; p < P; p++) for (int r4 = 0; r4 < 4; r4++) for (int x = 0; x < inner; x++) {
            const int xx = c0 + x, yy = cz + r4, zz = cz + p;
            unsigned char want = 0;
            if (xx >= 0 && xx < nx && yy >= 0 && yy < ny && zz >= 0 && zz < nz) want = h[(((size_t)1 * nz + zz) * ny + yy) * pitch + xx];
            if (g[(p * 4 + r4) * inner + x] != want) bad++;
        }
        printf("  mismatches: %d of %d\n", bad, bytes);
    }
    return 0;
}
