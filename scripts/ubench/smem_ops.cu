// Microbenchmark behind the crossCorrelation kernel design:
// SM-level cost of the shared-memory operations a private sliding histogram is made of,
// on B200, at the occupancies the kernel can run at.  Every thread owns a histogram column
// hist[bin][tid] (u32, conflict free); the "ring" is a byte array all lanes read at
// consecutive addresses.  Output: SM cycles per warp-level instruction (or per pair of
// updates) = the reciprocal throughput that bounds the kernel.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o smem_ops smem_ops.cu && ./smem_ops
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int T = 128;       // threads per block
// bins per histogram: runtime (101 = k + dummy; 51 lets more blocks fit per SM)
constexpr int RING = 8192;   // ring bytes

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned lds_u8(unsigned a) { unsigned v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ unsigned lds_u16(unsigned a) { unsigned v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ unsigned lds_u32(unsigned a) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u32(unsigned a, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(unsigned a, unsigned v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned atom_add(unsigned a, unsigned d) { unsigned v; asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(v) : "r"(a), "r"(d) : "memory"); return v; }
__device__ __forceinline__ void red_add(unsigned a, unsigned d) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(d) : "memory"); }

// mode: 0 LDS.U8 ring | 1 LDS.U32 private | 2 LDS+ADD+STS private (u32) | 3 ATOMS (return used)
//       4 RED | 5 proposed pair: 2 LDS.U8 + 2 LEA + 2 ATOMS + IADD3 | 6 same with RED (no S11)
//       7 round-1 style pair: 2 LDS.U16 ring, setp, 2 LDS.U16, add/sub, 2 STS.U16
//       8 pair with LDS/ADD/STS on u32 counters + S11 (no atomics)
__global__ void __launch_bounds__(T) k(int mode, int iters, int NBINS, unsigned long long* cyc, unsigned* sink) {
    extern __shared__ __align__(16) unsigned smem[];
    unsigned* hist = smem;                                   // [NBINS][T]
    unsigned char* ring = reinterpret_cast<unsigned char*>(smem + NBINS * T);
    for (int i = threadIdx.x; i < NBINS * T; i += T) hist[i] = 1000u;
    unsigned s = 12345u + blockIdx.x;
    for (int i = threadIdx.x; i < RING; i += T) {
        unsigned h = (unsigned)i * 2654435761u + s;
        h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
        ring[i] = (unsigned char)(h % NBINS);
    }
    __syncthreads();
    const unsigned hbase = (unsigned)__cvta_generic_to_shared(hist) + threadIdx.x * 4;
    const unsigned rbase = (unsigned)__cvta_generic_to_shared(ring) + (threadIdx.x & 31);
    unsigned acc = 0;
    unsigned a[8];
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = hbase + (unsigned)ring[(threadIdx.x * 8 + j) & (RING - 1)] * (T * 4);
    const long long t0 = clock64();
    if (mode == 0) {
        for (int it = 0; it < iters; it++) {
            const unsigned r = rbase + ((it * 352) & (RING - 512));
#pragma unroll
            for (int j = 0; j < 8; j++) acc += lds_u8(r + 48 * j);
        }
    } else if (mode == 1) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int j = 0; j < 8; j++) acc += lds_u32(a[j]);
        }
    } else if (mode == 2) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int j = 0; j < 8; j++) { unsigned h = lds_u32(a[j]); sts_u32(a[j], h + 1); acc += h; }
        }
    } else if (mode == 3) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int j = 0; j < 8; j++) acc += atom_add(a[j], (j & 1) ? 0xffffffffu : 1u);
        }
    } else if (mode == 4) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int j = 0; j < 8; j++) red_add(a[j], (j & 1) ? 0xffffffffu : 1u);
        }
    } else if (mode == 5 || mode == 6) {
        // software pipelined by hand: all ring reads of a batch first, then the address
        // arithmetic, then the atomics; the returned counts are only summed at the end
        for (int it = 0; it < iters; it++) {
            const unsigned rin = rbase + ((it * 352) & (RING - 1024));
            const unsigned rout = rin + 528;
            unsigned bi[8], bo[8];
#pragma unroll
            for (int j = 0; j < 8; j++) { bi[j] = lds_u8(rin + 48 * j + (j & 3)); bo[j] = lds_u8(rout + 48 * j + (j & 3)); }
#pragma unroll
            for (int j = 0; j < 8; j++) { bi[j] = hbase + bi[j] * (T * 4); bo[j] = hbase + bo[j] * (T * 4); }
            if (mode == 5) {
#pragma unroll
                for (int j = 0; j < 8; j++) { bi[j] = atom_add(bi[j], 1u); bo[j] = atom_add(bo[j], 0xffffffffu); }
#pragma unroll
                for (int j = 0; j < 8; j++) acc += bi[j] - bo[j];
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) { red_add(bi[j], 1u); red_add(bo[j], 0xffffffffu); }
            }
        }
    } else if (mode == 7) {
        for (int it = 0; it < iters; it++) {
            const unsigned rin = rbase * 1 + ((it * 352) & (RING - 1024));
            const unsigned rout = rin + 528;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                // ring holds u8 here; scale to the packed-u16 layout the round-1 kernel uses
                const unsigned bi = lds_u8(rin + 48 * j + (j & 3)), bo = lds_u8(rout + 48 * j + (j & 3));
                const unsigned ai = hbase + ((bi >> 1) * (T * 4)) + ((bi & 1) << 1), ao = hbase + ((bo >> 1) * (T * 4)) + ((bo & 1) << 1);
                if (ai != ao) {
                    unsigned h = lds_u16(ai), g = lds_u16(ao);
                    sts_u16(ai, h + 1); sts_u16(ao, g - 1);
                }
            }
        }
    } else if (mode == 8) {
        for (int it = 0; it < iters; it++) {
            const unsigned rin = rbase + ((it * 352) & (RING - 1024));
            const unsigned rout = rin + 528;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const unsigned bi = lds_u8(rin + 48 * j + (j & 3)), bo = lds_u8(rout + 48 * j + (j & 3));
                const unsigned ai = hbase + bi * (T * 4), ao = hbase + bo * (T * 4);
                if (ai != ao) {
                    unsigned h = lds_u32(ai), g = lds_u32(ao);
                    sts_u32(ai, h + 1); sts_u32(ao, g - 1);
                    acc += h - g;
                }
            }
        }
    }
    else if (mode == 9) {   // LDS.U16 ring, 32 lanes = 16 words
        const unsigned rb16 = (unsigned)__cvta_generic_to_shared(ring) + 2 * (threadIdx.x & 31);
        for (int it = 0; it < iters; it++) {
            const unsigned r = rb16 + ((it * 352) & (RING - 1024));
#pragma unroll
            for (int j = 0; j < 8; j++) acc += lds_u16(r + 96 * j);
        }
    } else if (mode == 10) {   // LDS.U32 ring, 32 lanes = 32 distinct words
        const unsigned rb32 = (unsigned)__cvta_generic_to_shared(ring) + 4 * (threadIdx.x & 31);
        for (int it = 0; it < iters; it++) {
            const unsigned r = rb32 + ((it * 352) & (RING - 2048));
#pragma unroll
            for (int j = 0; j < 8; j++) acc += lds_u32(r + 192 * j);
        }
    } else if (mode == 11) {   // LDS.U32 with the broadcast pattern of the u8 ring (lane>>2)
        const unsigned rb = (unsigned)__cvta_generic_to_shared(ring) + 4 * ((threadIdx.x & 31) >> 2);
        for (int it = 0; it < iters; it++) {
            const unsigned r = rb + ((it * 352) & (RING - 1024));
#pragma unroll
            for (int j = 0; j < 8; j++) acc += lds_u32(r + 48 * j);
        }
    } else if (mode == 12) {   // pair with a u32 ring of pre-scaled offsets: 2 LDS.32 + 2 IADD + 2 ATOMS + IADD3
        const unsigned rb32 = (unsigned)__cvta_generic_to_shared(ring) + 4 * (threadIdx.x & 31);
        for (int it = 0; it < iters; it++) {
            const unsigned rin = rb32 + ((it * 352) & (RING - 4096));
            const unsigned rout = rin + 2048;
            unsigned bi[8], bo[8];
#pragma unroll
            for (int j = 0; j < 8; j++) { bi[j] = lds_u32(rin + 192 * j + 4 * (j & 3)); bo[j] = lds_u32(rout + 192 * j + 4 * (j & 3)); }
#pragma unroll
            for (int j = 0; j < 8; j++) { bi[j] = hbase + ((bi[j] & 63u) * (T * 4)); bo[j] = hbase + ((bo[j] & 63u) * (T * 4)); }
#pragma unroll
            for (int j = 0; j < 8; j++) { bi[j] = atom_add(bi[j], 1u); bo[j] = atom_add(bo[j], 0xffffffffu); }
#pragma unroll
            for (int j = 0; j < 8; j++) acc += bi[j] - bo[j];
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0xdeadbeefu) sink[0] = acc + hist[threadIdx.x];
}

int main() {
    const char* names[] = {"LDS.U8 ring (per instr)", "LDS.U32 private (per instr)", "LDS+ADD+STS u32 (per RMW)",
                           "ATOMS add, return used (per instr)", "RED add (per instr)",
                           "pair: 2 LDS.U8 + 2 LEA + 2 ATOMS + IADD3 (per pair)", "pair: 2 LDS.U8 + 2 LEA + 2 RED (per pair)",
                           "pair round-1 style: u16 LDS/STS with equal-skip (per pair)", "pair: u32 LDS/ADD/STS + S11, equal-skip (per pair)", "LDS.U16 ring (per instr)", "LDS.U32 ring, distinct words (per instr)", "LDS.U32 broadcast lane>>2 (per instr)", "pair: 2 LDS.32 ring + 2 IMAD + 2 ATOMS + IADD3 (per pair)"};
    unsigned long long* d_cyc; unsigned* d_sink;
    CK(cudaMalloc(&d_cyc, sizeof(unsigned long long) * 148 * 8));
    CK(cudaMalloc(&d_sink, 64));
    const int iters = 2000;
    for (int nbins = 101; nbins >= 51; nbins -= 50) {
        const size_t smem = sizeof(unsigned) * nbins * T + RING;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int maxb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, k, T, smem));
        printf("nbins %d: smem per block %zu bytes, %d threads per block, %d blocks fit per SM\n", nbins, smem, T, maxb);
        for (int mode = 0; mode <= 12; mode++) {
            for (int bps = (nbins == 101 ? 1 : 4); bps <= maxb && bps <= 7; bps++) {
                const int grid = 148 * bps;
                k<<<grid, T, smem>>>(mode, 100, nbins, d_cyc, d_sink);   // warm-up
                CK(cudaDeviceSynchronize());
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                cudaEventRecord(e0);
                k<<<grid, T, smem>>>(mode, iters, nbins, d_cyc, d_sink);
                cudaEventRecord(e1);
                CK(cudaDeviceSynchronize());
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                static unsigned long long h[148 * 8];
                CK(cudaMemcpy(h, d_cyc, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost));
                unsigned long long mx = 0; double mean = 0;
                for (int i = 0; i < grid; i++) { if (h[i] > mx) mx = h[i]; mean += (double)h[i]; }
                mean /= grid;
                const double units_per_sm = (double)iters * 8 * (T / 32) * bps;   // warp-level units one SM retires
                printf("mode %2d %-62s warps/SM %2d: %.3f SM-cycles per unit (mean block %.0f cyc, max %llu, %.3f ms)\n", mode,
                       names[mode], bps * T / 32, mean / units_per_sm, mean, mx, ms);
            }
        }
    }
    return 0;
}
