// A1/A2: boolean and intensity pointwise operators, fused along
// the DAG.  One kernel executes a whole straight-line program per 16-voxel group:
// leaves are read once with 128-bit loads, intermediates live in a shared-memory
// register file sized by the program (zero bytes for a single operator), and only
// the root is written back.  HBM traffic of a chain = leaf bytes + root bytes.
//
// Semantics: Greta.pdf p.30 (not, and, true), p.35 (or, false), p.43 (thresholds);
// comparisons and arithmetic are IEEE-754 binary32, one rounding per operator
// (no FMA contraction can occur across the interpreter's dispatch).
#include "vx_internal.h"

namespace vx {

constexpr int kPwThreads = 256;
constexpr int kGroup = 16;  // voxels per thread per iteration

enum : uint8_t { K_REG = 0, K_LEAF = 1 };
enum : uint8_t { OPC_COPYB = 32, OPC_COPYF = 33 };

struct DevOp {
    uint8_t opcode, aux;
    int8_t dst;  // physical register, -1 = program output
    uint8_t a_kind, a_idx, b_kind, b_idx, pad;
    float imm;
};

struct DevProg {
    int n_ops;
    int n_breg, n_freg;
    int nb;
    unsigned long long nvox;     // voxels per volume
    unsigned long long count;    // voxels of the whole batch
    unsigned long long ngroups;  // 16-voxel groups over the whole batch (128-voxel groups in the bit kernel)
    const void* leaf[VXP_MAX_LEAVES];
    // a DEFERRED f32 leaf (Image::lazy): leaf = the 16-bit codes, leaf_tab = the per-volume tables
    // (nullptr for an ordinary leaf), leaf_mask = the mask's bit stream, leaf_flip = 0x8000 for int16 codes
    const float* leaf_tab[VXP_MAX_LEAVES];
    const void* leaf_mask[VXP_MAX_LEAVES];
    unsigned leaf_flip[VXP_MAX_LEAVES];
    // a deferred leaf that the program only compares with constants (leaf_rank = 1): the look-up is not
    // made at all.  The table is monotone, so `table[code] OP c` is `position OP' T` with the code's position
    // in value order inside the volume's code range (leaf_qrange, leaf_incr) and a break point T per
    // comparison and volume, found by pw_break_points_kernel; voxels outside the mask sit at -0.5.
    int leaf_rank[VXP_MAX_LEAVES];
    const int* leaf_qrange[VXP_MAX_LEAVES];
    int leaf_incr[VXP_MAX_LEAVES];
    // an f32 leaf read as the 16-bit codes of a quantised image (Image::q16): leaf = the codes,
    // v = fl(fl((float)q * slope) + inter) as vx_upload_scaled formed it; leaf_codes = 1 (int16) / 2 (uint16)
    int leaf_codes[VXP_MAX_LEAVES];
    float leaf_slope[VXP_MAX_LEAVES], leaf_inter[VXP_MAX_LEAVES];
    void* out;
    const float* bscal;  // [n_scalars][nb]
    DevOp ops[VXP_MAX_OPS];
};

struct F16 {
    float4 v[4];
};

// Boolean leaves and boolean results are BIT streams (Image::bits): the 16 voxels of group g are
// halfword g.  Inside the program a boolean register is 16 bytes of 0/1 (a uint4), which is what
// the comparisons produce and the select consumes.
__device__ __forceinline__ unsigned bytes_of_nib(unsigned nib) { return (nib * 0x00204081u) & 0x01010101u; }
__device__ __forceinline__ unsigned nib_of_bytes(unsigned w) {
    w = __vcmpne4(w, 0u) & 0x01010101u;
    return ((w * 0x01020408u) >> 24) & 0xfu;
}
__device__ __forceinline__ uint4 ld_b(const DevProg& P, uint8_t kind, uint8_t idx, size_t g,
                                      const uint4* sb) {
    if (kind == K_LEAF) {
        const unsigned h = __ldg(reinterpret_cast<const unsigned short*>(P.leaf[idx]) + g);
        return make_uint4(bytes_of_nib(h & 0xfu), bytes_of_nib((h >> 4) & 0xfu), bytes_of_nib((h >> 8) & 0xfu),
                          bytes_of_nib(h >> 12));
    }
    return sb[idx * kPwThreads + threadIdx.x];
}
// DEF: what the program's deferred leaves need -- 0: there are none, 1: table look-ups (and possibly rank
// mode), 2: rank mode only.  The kernel is instantiated per case so that the common programs do not carry
// the registers of the look-up path (80 -> 90 registers cost the f32 threshold a resident block per SM).
template <int DEF>
__device__ __forceinline__ F16 ld_f(const DevProg& P, uint8_t kind, uint8_t idx, size_t g,
                                    const float4* sf) {
    F16 r;
    if (DEF != 0 && kind == K_LEAF && P.leaf_tab[idx] != nullptr) {
        // deferred leaf: v = mask ? table[volume][u(code)] : 0 -- 2.125 bytes per voxel from HBM, the
        // look-ups hit L1 (a volume's codes span a few thousand table entries)
        const unsigned m = __ldg(reinterpret_cast<const unsigned short*>(P.leaf_mask[idx]) + g);
        float t[kGroup];
#pragma unroll
        for (int e = 0; e < kGroup; e++) t[e] = 0.f;
        // The codes are asked for together with the mask, not after it: skipping the 32 bytes of a group
        // outside the mask saved traffic nobody was short of and put two memory round trips in a row
        // (0.19 ms per 16 BraTS volumes for 90 MB).  (The code array ends with the batch's last 8-voxel group.)
        const uint4* qp = reinterpret_cast<const uint4*>(P.leaf[idx]) + g * 2;
        const size_t first = g * kGroup;
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        const uint4 a = __ldg(qp), b = first + 8 < P.count ? __ldg(qp + 1) : z4;
        if (m != 0u) {
            const unsigned qw[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            // volume of the group's first voxel (a 32-bit division whenever the batch allows it)
            const unsigned v0 = P.count <= 0xffffffffull ? (unsigned)first / (unsigned)P.nvox : (unsigned)(first / P.nvox);
            const bool one_volume = first - (size_t)v0 * P.nvox + (kGroup - 1) < P.nvox;
            const unsigned flip = P.leaf_flip[idx];
            const float* tab = P.leaf_tab[idx];
            if (DEF == 2 || P.leaf_rank[idx]) {
                const int* qr = P.leaf_qrange[idx];
                const bool incr = P.leaf_incr[idx] != 0;
                int lo = __ldg(qr + 2 * v0), hi = __ldg(qr + 2 * v0 + 1);
#pragma unroll
                for (int e = 0; e < kGroup; e++) {
                    if (!one_volume) {
                        const unsigned v = (unsigned)((first + e) / P.nvox);
                        lo = __ldg(qr + 2 * v); hi = __ldg(qr + 2 * v + 1);
                    }
                    const int u = (int)(((qw[e >> 1] >> (16 * (e & 1))) & 0xffffu) ^ flip);
                    t[e] = ((m >> e) & 1u) ? (float)(incr ? u - lo : hi - u) : -0.5f;
                }
            } else if (DEF == 1) {
#pragma unroll
                for (int e = 0; e < kGroup; e++) {
                    if (!((m >> e) & 1u)) continue;
                    const unsigned v = one_volume ? v0 : (unsigned)((first + e) / P.nvox);
                    const unsigned u = ((qw[e >> 1] >> (16 * (e & 1))) & 0xffffu) ^ flip;
                    t[e] = __ldg(tab + (size_t)v * 65536u + u);
                }
            }
        } else if (DEF == 2 || P.leaf_rank[idx]) {
#pragma unroll
            for (int e = 0; e < kGroup; e++) t[e] = -0.5f;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) r.v[q] = make_float4(t[4 * q], t[4 * q + 1], t[4 * q + 2], t[4 * q + 3]);
    } else if (kind == K_LEAF && P.leaf_codes[idx] != 0) {
        // quantised leaf: 2 bytes per voxel instead of 4, the same two roundings as the upload's conversion
        const uint4* qp = reinterpret_cast<const uint4*>(P.leaf[idx]) + g * 2;
        const size_t first = g * kGroup;
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        // (the code array ends with the batch's last 8-voxel group)
        const uint4 a = __ldg(qp), b = first + 8 < P.count ? __ldg(qp + 1) : z4;
        const unsigned qw[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        const bool sgn = P.leaf_codes[idx] == 1;
        const float slope = P.leaf_slope[idx], inter = P.leaf_inter[idx];
        float t[kGroup];
#pragma unroll
        for (int e = 0; e < kGroup; e++) {
            const unsigned h = (qw[e >> 1] >> (16 * (e & 1))) & 0xffffu;
            const float q = sgn ? (float)(short)h : (float)h;
            t[e] = __fadd_rn(__fmul_rn(q, slope), inter);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) r.v[q] = make_float4(t[4 * q], t[4 * q + 1], t[4 * q + 2], t[4 * q + 3]);
    } else if (kind == K_LEAF) {
        const float4* p = reinterpret_cast<const float4*>(P.leaf[idx]) + g * 4;
#pragma unroll
        for (int q = 0; q < 4; q++) r.v[q] = __ldg(p + q);
    } else {
#pragma unroll
        for (int q = 0; q < 4; q++) r.v[q] = sf[(idx * 4 + q) * kPwThreads + threadIdx.x];
    }
    return r;
}
__device__ __forceinline__ void st_b(const DevProg& P, int8_t dst, size_t g, uint4* sb, uint4 v) {
    if (dst < 0) {
        unsigned h = nib_of_bytes(v.x) | (nib_of_bytes(v.y) << 4) | (nib_of_bytes(v.z) << 8) | (nib_of_bytes(v.w) << 12);
        // bits past the last voxel of the batch stay 0 (the invariant of the bit form)
        if ((g + 1) * kGroup > P.count) h &= (1u << (unsigned)(P.count - g * kGroup)) - 1u;
        reinterpret_cast<unsigned short*>(P.out)[g] = (unsigned short)h;
    } else {
        sb[dst * kPwThreads + threadIdx.x] = v;
    }
}
__device__ __forceinline__ void st_f(const DevProg& P, int8_t dst, size_t g, float4* sf,
                                     const F16& v) {
    if (dst < 0) {
        float4* p = reinterpret_cast<float4*>(P.out) + g * 4;
#pragma unroll
        for (int q = 0; q < 4; q++) p[q] = v.v[q];
    } else {
#pragma unroll
        for (int q = 0; q < 4; q++) sf[(dst * 4 + q) * kPwThreads + threadIdx.x] = v.v[q];
    }
}

__device__ __forceinline__ unsigned cmp1(float x, float y, int op) {
    switch (op) {
        case VX_LT: return x < y;
        case VX_LE: return x <= y;
        case VX_GT: return x > y;
        case VX_GE: return x >= y;
        case VX_EQ: return x == y;
        default: return x != y;
    }
}
__device__ __forceinline__ float arith1(float x, float y, int op) {
    switch (op) {
        case VX_ADD: return __fadd_rn(x, y);
        case VX_SUB: return __fsub_rn(x, y);
        case VX_MUL: return __fmul_rn(x, y);
        case VX_DIV: return __fdiv_rn(x, y);
        case VX_RSUB: return __fsub_rn(y, x);
        case VX_RDIV: return __fdiv_rn(y, x);
        case VX_MIN: return fminf(x, y);
        default: return fmaxf(x, y);
    }
}
__device__ __forceinline__ unsigned pack4(unsigned a, unsigned b, unsigned c, unsigned d) {
    return a | (b << 8) | (c << 16) | (d << 24);
}
__device__ __forceinline__ uint4 cmp16(const F16& x, const F16& y, int op) {
    uint4 r;
    r.x = pack4(cmp1(x.v[0].x, y.v[0].x, op), cmp1(x.v[0].y, y.v[0].y, op),
                cmp1(x.v[0].z, y.v[0].z, op), cmp1(x.v[0].w, y.v[0].w, op));
    r.y = pack4(cmp1(x.v[1].x, y.v[1].x, op), cmp1(x.v[1].y, y.v[1].y, op),
                cmp1(x.v[1].z, y.v[1].z, op), cmp1(x.v[1].w, y.v[1].w, op));
    r.z = pack4(cmp1(x.v[2].x, y.v[2].x, op), cmp1(x.v[2].y, y.v[2].y, op),
                cmp1(x.v[2].z, y.v[2].z, op), cmp1(x.v[2].w, y.v[2].w, op));
    r.w = pack4(cmp1(x.v[3].x, y.v[3].x, op), cmp1(x.v[3].y, y.v[3].y, op),
                cmp1(x.v[3].z, y.v[3].z, op), cmp1(x.v[3].w, y.v[3].w, op));
    return r;
}
__device__ __forceinline__ F16 arith16(const F16& x, const F16& y, int op) {
    F16 r;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        r.v[q].x = arith1(x.v[q].x, y.v[q].x, op);
        r.v[q].y = arith1(x.v[q].y, y.v[q].y, op);
        r.v[q].z = arith1(x.v[q].z, y.v[q].z, op);
        r.v[q].w = arith1(x.v[q].w, y.v[q].w, op);
    }
    return r;
}
__device__ __forceinline__ F16 splat(float s) {
    F16 r;
#pragma unroll
    for (int q = 0; q < 4; q++) r.v[q] = make_float4(s, s, s, s);
    return r;
}
// per-volume scalar number `which` for the 16 voxels of group g
__device__ __forceinline__ F16 batch_scalar(const DevProg& P, int which, size_t g) {
    size_t first = g * kGroup;
    unsigned v0 = (unsigned)(first / P.nvox);
    unsigned v1 = (unsigned)((first + kGroup - 1) / P.nvox);
    const float* tab = P.bscal + (size_t)which * P.nb;
    if (v1 >= (unsigned)P.nb) v1 = P.nb - 1;
    if (v0 >= (unsigned)P.nb) v0 = P.nb - 1;
    if (v0 == v1) return splat(__ldg(tab + v0));
    F16 r;
    float tmp[kGroup];
#pragma unroll
    for (int e = 0; e < kGroup; e++) {
        unsigned v = (unsigned)((first + e) / P.nvox);
        if (v >= (unsigned)P.nb) v = P.nb - 1;
        tmp[e] = __ldg(tab + v);
    }
#pragma unroll
    for (int q = 0; q < 4; q++) r.v[q] = make_float4(tmp[4 * q], tmp[4 * q + 1], tmp[4 * q + 2], tmp[4 * q + 3]);
    return r;
}
__device__ __forceinline__ float sel1(unsigned m, int sh, float a, float fill) {
    return ((m >> sh) & 0xffu) ? a : fill;
}

// Four blocks per SM (64 registers, a few bytes of spill in the plain instantiation): these programs wait for
// memory with one group per thread in flight, so residency is what buys bandwidth -- measured on 16 BraTS
// volumes: threshold 0.119 -> 0.112 ms, deferred threshold 0.183 -> 0.156 ms; six blocks (40 registers)
// spill kilobytes and lose (0.148 / 0.189).
template <int DEF>
__global__ void __launch_bounds__(kPwThreads, 4)
pointwise_program_kernel(const __grid_constant__ DevProg P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4* sb = reinterpret_cast<uint4*>(smem_raw);
    float4* sf = reinterpret_cast<float4*>(smem_raw + (size_t)P.n_breg * kPwThreads * sizeof(uint4));
    const size_t stride = (size_t)gridDim.x * kPwThreads;
    for (size_t g = (size_t)blockIdx.x * kPwThreads + threadIdx.x; g < P.ngroups; g += stride) {
        for (int i = 0; i < P.n_ops; i++) {
            const DevOp op = P.ops[i];
            switch (op.opcode) {
                case VXP_NOT: {
                    uint4 x = ld_b(P, op.a_kind, op.a_idx, g, sb);
                    x.x ^= 0x01010101u; x.y ^= 0x01010101u; x.z ^= 0x01010101u; x.w ^= 0x01010101u;
                    st_b(P, op.dst, g, sb, x);
                } break;
                case VXP_AND: {
                    uint4 x = ld_b(P, op.a_kind, op.a_idx, g, sb), y = ld_b(P, op.b_kind, op.b_idx, g, sb);
                    x.x &= y.x; x.y &= y.y; x.z &= y.z; x.w &= y.w;
                    st_b(P, op.dst, g, sb, x);
                } break;
                case VXP_OR: {
                    uint4 x = ld_b(P, op.a_kind, op.a_idx, g, sb), y = ld_b(P, op.b_kind, op.b_idx, g, sb);
                    x.x |= y.x; x.y |= y.y; x.z |= y.z; x.w |= y.w;
                    st_b(P, op.dst, g, sb, x);
                } break;
                case VXP_XOR: {
                    uint4 x = ld_b(P, op.a_kind, op.a_idx, g, sb), y = ld_b(P, op.b_kind, op.b_idx, g, sb);
                    x.x ^= y.x; x.y ^= y.y; x.z ^= y.z; x.w ^= y.w;
                    st_b(P, op.dst, g, sb, x);
                } break;
                case OPC_COPYB: {
                    st_b(P, op.dst, g, sb, ld_b(P, op.a_kind, op.a_idx, g, sb));
                } break;
                case OPC_COPYF: {
                    st_f(P, op.dst, g, sf, ld_f<DEF>(P, op.a_kind, op.a_idx, g, sf));
                } break;
                case VXP_CMPS: {
                    F16 x = ld_f<DEF>(P, op.a_kind, op.a_idx, g, sf);
                    st_b(P, op.dst, g, sb, cmp16(x, splat(op.imm), op.aux));
                } break;
                case VXP_CMPSB: {
                    F16 x = ld_f<DEF>(P, op.a_kind, op.a_idx, g, sf);
                    st_b(P, op.dst, g, sb, cmp16(x, batch_scalar(P, op.b_idx, g), op.aux));
                } break;
                case VXP_CMP: {
                    F16 x = ld_f<DEF>(P, op.a_kind, op.a_idx, g, sf), y = ld_f<DEF>(P, op.b_kind, op.b_idx, g, sf);
                    st_b(P, op.dst, g, sb, cmp16(x, y, op.aux));
                } break;
                case VXP_ARITHS: {
                    F16 x = ld_f<DEF>(P, op.a_kind, op.a_idx, g, sf);
                    st_f(P, op.dst, g, sf, arith16(x, splat(op.imm), op.aux));
                } break;
                case VXP_ARITHSB: {
                    F16 x = ld_f<DEF>(P, op.a_kind, op.a_idx, g, sf);
                    st_f(P, op.dst, g, sf, arith16(x, batch_scalar(P, op.b_idx, g), op.aux));
                } break;
                case VXP_ARITH: {
                    F16 x = ld_f<DEF>(P, op.a_kind, op.a_idx, g, sf), y = ld_f<DEF>(P, op.b_kind, op.b_idx, g, sf);
                    st_f(P, op.dst, g, sf, arith16(x, y, op.aux));
                } break;
                case VXP_TOF32: {
                    uint4 x = ld_b(P, op.a_kind, op.a_idx, g, sb);
                    unsigned w[4] = {x.x, x.y, x.z, x.w};
                    F16 r;
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        r.v[q] = make_float4((float)(w[q] & 0xffu), (float)((w[q] >> 8) & 0xffu),
                                             (float)((w[q] >> 16) & 0xffu), (float)(w[q] >> 24));
                    st_f(P, op.dst, g, sf, r);
                } break;
                case VXP_TOBOOL: {
                    F16 x = ld_f<DEF>(P, op.a_kind, op.a_idx, g, sf);
                    st_b(P, op.dst, g, sb, cmp16(x, splat(0.f), VX_NE));
                } break;
                case VXP_CONSTB: {
                    unsigned w = op.aux ? 0x01010101u : 0u;
                    st_b(P, op.dst, g, sb, make_uint4(w, w, w, w));
                } break;
                case VXP_CONSTF: {
                    st_f(P, op.dst, g, sf, splat(op.imm));
                } break;
                case VXP_SELECT: {
                    uint4 m = ld_b(P, op.a_kind, op.a_idx, g, sb);
                    F16 x = ld_f<DEF>(P, op.b_kind, op.b_idx, g, sf);
                    unsigned w[4] = {m.x, m.y, m.z, m.w};
                    F16 r;
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        r.v[q] = make_float4(sel1(w[q], 0, x.v[q].x, op.imm), sel1(w[q], 8, x.v[q].y, op.imm),
                                             sel1(w[q], 16, x.v[q].z, op.imm), sel1(w[q], 24, x.v[q].w, op.imm));
                    st_f(P, op.dst, g, sf, r);
                } break;
                default: break;
            }
        }
    }
}

// Programs over boolean images only (not / and / or / xor chains) run on the bit form directly: a
// "group" is a uint4 = 128 voxels, the operators are the machine's own bitwise instructions, and
// the traffic is 1/8 of the byte form's.  Every thread runs the program on kWideU independent
// groups per iteration: all leaf loads of an instruction are issued before the first is used.
constexpr int kWideU = 4;
constexpr int kBitGroup = 128;   // voxels per uint4 of bits

__device__ __forceinline__ void ld_bw(const DevProg& P, uint8_t kind, uint8_t idx, const size_t (&g)[kWideU],
                                      const uint4* sb, uint4 (&x)[kWideU]) {
#pragma unroll
    for (int u = 0; u < kWideU; u++) {
        if (kind == K_LEAF)
            x[u] = g[u] < P.ngroups ? __ldg(reinterpret_cast<const uint4*>(P.leaf[idx]) + g[u]) : make_uint4(0, 0, 0, 0);
        else
            x[u] = sb[(idx * kWideU + u) * kPwThreads + threadIdx.x];
    }
}
__device__ __forceinline__ unsigned tail_word(unsigned w, long long valid) {   // keep the low `valid` bits
    return valid >= 32 ? w : valid <= 0 ? 0u : (w & ((1u << (unsigned)valid) - 1u));
}
__device__ __forceinline__ void st_bw(const DevProg& P, int8_t dst, const size_t (&g)[kWideU], uint4* sb,
                                      const uint4 (&x)[kWideU]) {
#pragma unroll
    for (int u = 0; u < kWideU; u++) {
        if (dst < 0) {
            if (g[u] < P.ngroups) {
                uint4 v = x[u];
                if ((g[u] + 1) * kBitGroup > P.count) {   // the last group: bits past the batch stay 0
                    const long long left = (long long)(P.count - g[u] * kBitGroup);
                    v.x = tail_word(v.x, left); v.y = tail_word(v.y, left - 32);
                    v.z = tail_word(v.z, left - 64); v.w = tail_word(v.w, left - 96);
                }
                reinterpret_cast<uint4*>(P.out)[g[u]] = v;
            }
        } else {
            sb[(dst * kWideU + u) * kPwThreads + threadIdx.x] = x[u];
        }
    }
}

__global__ void __launch_bounds__(kPwThreads)
boolean_program_kernel(const __grid_constant__ DevProg P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4* sb = reinterpret_cast<uint4*>(smem_raw);
    const size_t stride = (size_t)gridDim.x * kPwThreads;
    for (size_t g0 = (size_t)blockIdx.x * kPwThreads + threadIdx.x; g0 < P.ngroups; g0 += stride * kWideU) {
        size_t g[kWideU];
#pragma unroll
        for (int u = 0; u < kWideU; u++) g[u] = g0 + (size_t)u * stride;
        for (int i = 0; i < P.n_ops; i++) {
            const DevOp op = P.ops[i];
            uint4 x[kWideU], y[kWideU];
            if (op.opcode == VXP_CONSTB) {
                const unsigned w = op.aux ? 0xffffffffu : 0u;
#pragma unroll
                for (int u = 0; u < kWideU; u++) x[u] = make_uint4(w, w, w, w);
            } else {
                ld_bw(P, op.a_kind, op.a_idx, g, sb, x);
            }
            if (op.opcode == VXP_AND || op.opcode == VXP_OR || op.opcode == VXP_XOR) ld_bw(P, op.b_kind, op.b_idx, g, sb, y);
#pragma unroll
            for (int u = 0; u < kWideU; u++) {
                switch (op.opcode) {
                    case VXP_NOT: x[u].x = ~x[u].x; x[u].y = ~x[u].y; x[u].z = ~x[u].z; x[u].w = ~x[u].w; break;
                    case VXP_AND: x[u].x &= y[u].x; x[u].y &= y[u].y; x[u].z &= y[u].z; x[u].w &= y[u].w; break;
                    case VXP_OR: x[u].x |= y[u].x; x[u].y |= y[u].y; x[u].z |= y[u].z; x[u].w |= y[u].w; break;
                    case VXP_XOR: x[u].x ^= y[u].x; x[u].y ^= y[u].y; x[u].z ^= y[u].z; x[u].w ^= y[u].w; break;
                    default: break;   // OPC_COPYB, VXP_CONSTB
                }
            }
            st_bw(P, op.dst, g, sb, x);
        }
    }
}

// Break points of the comparisons of a rank-mode leaf (DevProg::leaf_rank).  One thread per (comparison,
// volume): T = first position p of the volume's code range, in value order, whose table entry is > c
// (GT, LE) or >= c (GE, LT); the comparison becomes  position >= T  (GT, GE)  or  position < T  (LT, LE).
// A voxel outside the mask has the value 0 and the position -0.5: table entries are >= 0, so it goes
// with the first position unless 0 itself fails / passes the comparison, in which case T = -1 puts
// the break below it.  NaN constants come out false everywhere, as they must.
struct BreakSpec {
    const float* table;
    const int* qrange;
    int incr, cmp;
    float c;
    int slot;   // row of the scalar array that receives T
};
struct BreakSpecs {
    int n;
    BreakSpec s[VXP_MAX_OPS];
};
__global__ void pw_break_points_kernel(const __grid_constant__ BreakSpecs B, int nb, float* __restrict__ bscal) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B.n * nb) return;
    const BreakSpec& b = B.s[t / nb];
    const int vol = t % nb;
    const int ulo = b.qrange[2 * vol], uhi = b.qrange[2 * vol + 1];
    const int span = uhi - ulo + 1;
    const float* tv = b.table + (size_t)vol * 65536u;
    const float c = b.c;
    const bool strict_gt = b.cmp == VX_GT || b.cmp == VX_LE;   // first entry > c; otherwise first entry >= c
    int lo = 0, hi = span > 0 ? span : 0;   // first p in [0, span] with pred(p), pred monotone
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const float v = tv[b.incr ? ulo + mid : uhi - mid];
        const bool pr = strict_gt ? v > c : v >= c;
        if (pr) hi = mid; else lo = mid + 1;
    }
    float T = (float)lo;
    bool zero_passes;
    switch (b.cmp) {
        case VX_GT: zero_passes = 0.f > c; break;
        case VX_GE: zero_passes = 0.f >= c; break;
        case VX_LT: zero_passes = 0.f < c; break;
        default: zero_passes = 0.f <= c; break;
    }
    if (b.cmp == VX_GT || b.cmp == VX_GE) { if (zero_passes) T = -1.f; }
    else if (!zero_passes) T = -1.f;
    bscal[(size_t)b.slot * nb + vol] = T;
}

// ---- one comparison of a 16-bit-coded image with a constant -----------------------------------------
// `flair <. 0.1` on a 16-bit upload and `percentiles(..) >. c` on a deferred percentile image are ONE
// comparison of a leaf that is read as 16-bit codes: 2 (+ 1/8) bytes in and 1/8 byte out per voxel.  In
// the interpreter a thread has one 16-voxel group (32 bytes) in flight and the block 8 KB: these two
// programs ran at 2.5 TB/s where the f32 threshold, with twice the bytes per thread, reaches 5.  Here a
// thread takes 32 voxels -- four 16-byte loads issued together, one word of the result's bit form out.
// RANK: the leaf is a deferred percentile image in rank mode (value = position of the code in the volume's
// code range, -0.5 outside the mask, against a per-volume break point); otherwise a quantised f32 leaf
// (value = fl(fl(q * slope) + inter), against the constant or a per-volume scalar).  Same arithmetic as
// ld_f / cmp16 of the interpreter, so the result is bit-identical.
struct CodeCmp {
    const uint4* codes;
    const unsigned* mask;      // RANK: bit form of the percentile mask
    const int* qrange;         // RANK: [nb][2] code range of every volume
    const float* thr;          // per-volume thresholds [nb], or nullptr: imm
    float imm, slope, inter;
    unsigned flip;             // RANK: 0x8000 for int16 codes
    int sgn, incr, nb;
    unsigned invert;           // 0xffffffff: the program negates the comparison (`!(a >. c)`, what filter asks for)
    unsigned long long nvox, count;
    unsigned* out;
};
template <bool RANK, int CMP>
__global__ void __launch_bounds__(kPwThreads)
code_compare_kernel(const __grid_constant__ CodeCmp a) {
    const size_t nw = (a.count + 31) / 32;
    for (size_t w = (size_t)blockIdx.x * kPwThreads + threadIdx.x; w < nw; w += (size_t)gridDim.x * kPwThreads) {
        const size_t first = w * 32;
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        uint4 q[4];
#pragma unroll
        for (int j = 0; j < 4; j++) q[j] = first + 8 * j < a.count ? __ldg(a.codes + w * 4 + j) : z4;
        const unsigned m = RANK ? __ldg(a.mask + w) : 0xffffffffu;
        unsigned v0 = a.count <= 0xffffffffull ? (unsigned)first / (unsigned)a.nvox : (unsigned)(first / a.nvox);
        const bool one_volume = first - (size_t)v0 * a.nvox + 31 < a.nvox;
        if (v0 >= (unsigned)a.nb) v0 = a.nb - 1;
        float thr = a.thr ? __ldg(a.thr + v0) : a.imm;
        int lo = 0, hi = 0;
        if (RANK) { lo = __ldg(a.qrange + 2 * v0); hi = __ldg(a.qrange + 2 * v0 + 1); }
        const unsigned qw[16] = {q[0].x, q[0].y, q[0].z, q[0].w, q[1].x, q[1].y, q[1].z, q[1].w,
                                 q[2].x, q[2].y, q[2].z, q[2].w, q[3].x, q[3].y, q[3].z, q[3].w};
        unsigned bits = 0;
        if (RANK && (CMP == VX_GE || CMP == VX_LT) && one_volume) {
            // position >= T on integers: T is -1 (everything passes, voxels outside the mask included) or a
            // position, and then `position >= T` is one comparison of the code itself with lo + T (hi - T
            // for a decreasing scaling) and voxels outside the mask fail.  (The float form below needs a
            // subtraction, a conversion, a select and a comparison per voxel: the kernel was issue-bound.)
            const int Ti = (int)thr;
            unsigned ge = 0;
            if (a.incr) {
                const int A = lo + Ti;
#pragma unroll
                for (int e = 0; e < 32; e++) {
                    const int u = (int)(((qw[e >> 1] >> (16 * (e & 1))) & 0xffffu) ^ a.flip);
                    ge |= (u >= A ? 1u : 0u) << e;
                }
            } else {
                const int A = hi - Ti;
#pragma unroll
                for (int e = 0; e < 32; e++) {
                    const int u = (int)(((qw[e >> 1] >> (16 * (e & 1))) & 0xffffu) ^ a.flip);
                    ge |= (u <= A ? 1u : 0u) << e;
                }
            }
            ge = Ti < 0 ? 0xffffffffu : (ge & m);
            bits = CMP == VX_GE ? ge : ~ge;
        } else
#pragma unroll
        for (int e = 0; e < 32; e++) {
            if (!one_volume) {
                unsigned v = (unsigned)((first + e) / a.nvox);
                if (v >= (unsigned)a.nb) v = a.nb - 1;
                if (a.thr) thr = __ldg(a.thr + v);
                if (RANK) { lo = __ldg(a.qrange + 2 * v); hi = __ldg(a.qrange + 2 * v + 1); }
            }
            const unsigned h = (qw[e >> 1] >> (16 * (e & 1))) & 0xffffu;
            float x;
            if (RANK) {
                const int u = (int)(h ^ a.flip);
                x = ((m >> e) & 1u) ? (float)(a.incr ? u - lo : hi - u) : -0.5f;
            } else {
                x = __fadd_rn(__fmul_rn(a.sgn ? (float)(short)h : (float)h, a.slope), a.inter);
            }
            bits |= cmp1(x, thr, CMP) << e;
        }
        bits ^= a.invert;
        // bits past the last voxel of the batch stay 0 (the invariant of the bit form)
        if (first + 32 > a.count) bits &= (1u << (unsigned)(a.count - first)) - 1u;
        a.out[w] = bits;
    }
}
template <bool RANK>
static void launch_code_compare(int cmp, int grid, cudaStream_t st, const CodeCmp& a) {
    switch (cmp) {
        case VX_LT: code_compare_kernel<RANK, VX_LT><<<grid, kPwThreads, 0, st>>>(a); break;
        case VX_LE: code_compare_kernel<RANK, VX_LE><<<grid, kPwThreads, 0, st>>>(a); break;
        case VX_GT: code_compare_kernel<RANK, VX_GT><<<grid, kPwThreads, 0, st>>>(a); break;
        case VX_GE: code_compare_kernel<RANK, VX_GE><<<grid, kPwThreads, 0, st>>>(a); break;
        case VX_EQ: code_compare_kernel<RANK, VX_EQ><<<grid, kPwThreads, 0, st>>>(a); break;
        default: code_compare_kernel<RANK, VX_NE><<<grid, kPwThreads, 0, st>>>(a); break;
    }
}

// ---- host side: validate, allocate registers, launch ---------------------------
enum { T_NONE = 0, T_B = 1, T_F = 2 };

struct PubReg {
    int type = T_NONE;
    int leaf = -1;  // alias of a leaf (written by VXP_LOAD)
    int phys = -1;  // physical register of its type
    int def = -1;   // defining instruction
};

static int result_type(int opcode) {
    switch (opcode) {
        case VXP_NOT: case VXP_AND: case VXP_OR: case VXP_XOR: case VXP_CMPS: case VXP_CMP:
        case VXP_TOBOOL: case VXP_CONSTB: case VXP_CMPSB:
            return T_B;
        case VXP_ARITHS: case VXP_ARITH: case VXP_TOF32: case VXP_CONSTF: case VXP_SELECT:
        case VXP_ARITHSB:
            return T_F;
    }
    return T_NONE;
}
// operand types: returns number of register operands and their required types
static int operand_types(int opcode, int* ta, int* tb) {
    *ta = *tb = T_NONE;
    switch (opcode) {
        case VXP_NOT: *ta = T_B; return 1;
        case VXP_AND: case VXP_OR: case VXP_XOR: *ta = *tb = T_B; return 2;
        case VXP_CMPS: case VXP_ARITHS: case VXP_TOBOOL: case VXP_CMPSB: case VXP_ARITHSB:
            *ta = T_F; return 1;
        case VXP_CMP: case VXP_ARITH: *ta = *tb = T_F; return 2;
        case VXP_TOF32: *ta = T_B; return 1;
        case VXP_SELECT: *ta = T_B; *tb = T_F; return 2;
        case VXP_CONSTB: case VXP_CONSTF: return 0;
    }
    return -1;
}

int run_pointwise(vx_ctx ctx, const vx_op* prog, int n_ops, const vx_img* leaves, int n_leaves,
                  const float* batch_scalars, int n_scalars, vx_img like, vx_img* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    VX_REQUIRE(prog != nullptr && n_ops > 0 && n_ops <= VXP_MAX_OPS, VX_E_INVALID_ARG,
               "program length %d outside [1,%d]", n_ops, VXP_MAX_OPS);
    VX_REQUIRE(n_leaves >= 0 && n_leaves <= VXP_MAX_LEAVES, VX_E_INVALID_ARG,
               "%d leaves, at most %d supported", n_leaves, VXP_MAX_LEAVES);
    OpGuard g;
    VX_TRY(g.begin(ctx));
    Image* L[VXP_MAX_LEAVES] = {};
    const void* leaf_ptr[VXP_MAX_LEAVES] = {};
    const float* leaf_tab[VXP_MAX_LEAVES] = {};
    const void* leaf_mask[VXP_MAX_LEAVES] = {};
    unsigned leaf_flip[VXP_MAX_LEAVES] = {};
    int leaf_codes[VXP_MAX_LEAVES] = {};
    const Image::Lazy* leaf_lazy[VXP_MAX_LEAVES] = {};
    for (int i = 0; i < n_leaves; i++) {
        int dt = 0;
        VX_REQUIRE(vx_info(leaves[i], &dt, nullptr, nullptr, nullptr, nullptr, nullptr) == VX_OK, VX_E_INVALID_ARG,
                   "invalid image handle %llu", (unsigned long long)leaves[i]);
        VX_REQUIRE(dt == VX_U8 || dt == VX_F32, VX_E_DTYPE,
                   "leaf %d has dtype %d; pointwise programs take u8 or f32", i, dt);
        if (dt == VX_U8) {   // boolean leaves are read as bits
            const unsigned* b = nullptr;
            VX_TRY(g.input_bits(leaves[i], &L[i], &b));
            leaf_ptr[i] = b;
        } else {
            VX_TRY(g.input_f32_deferred(leaves[i], &L[i], &leaf_ptr[i]));
            if (leaf_ptr[i] == nullptr) {   // no f32 array yet: read the deferred form
                const Image::Lazy* z = L[i]->lazy;
                leaf_ptr[i] = z->src->q16;
                leaf_tab[i] = z->table;
                leaf_mask[i] = z->mask->bits;
                leaf_flip[i] = z->src->q_dtype == VX_I16 ? 0x8000u : 0u;
                leaf_lazy[i] = z;
            } else if (L[i]->has_codes() && g.c->quantised_paths.load()) {
                leaf_ptr[i] = L[i]->q16;   // half the bytes of the f32 array, the same values
                leaf_codes[i] = L[i]->q_dtype == VX_I16 ? 1 : 2;
            }
        }
        if (i > 0) VX_TRY(same_shape(L[0], L[i]));
    }
    Image* shape = nullptr;
    if (n_leaves > 0) shape = L[0];
    else {
        VX_TRY(g.input_meta(like, &shape));   // only its shape is used
    }

    DevProg P;
    memset(&P, 0, sizeof(P));
    PubReg regs[VXP_MAX_REGS];
    // last use of the value defined at instruction i
    int last_use[VXP_MAX_OPS];
    int cur_def[VXP_MAX_REGS];
    for (int r = 0; r < VXP_MAX_REGS; r++) cur_def[r] = -1;
    for (int i = 0; i < n_ops; i++) last_use[i] = -1;
    for (int i = 0; i < n_ops; i++) {
        const vx_op& o = prog[i];
        VX_REQUIRE(o.dst >= 0 && o.dst < VXP_MAX_REGS, VX_E_INVALID_ARG,
                   "instruction %d: dst register %d out of range", i, o.dst);
        int ta, tb;
        int nopnd = (o.opcode == VXP_LOAD) ? 0 : operand_types(o.opcode, &ta, &tb);
        VX_REQUIRE(nopnd >= 0, VX_E_INVALID_ARG, "instruction %d: unknown opcode %d", i, o.opcode);
        if (nopnd >= 1) {
            VX_REQUIRE(o.a >= 0 && o.a < VXP_MAX_REGS && cur_def[o.a] >= 0, VX_E_INVALID_ARG,
                       "instruction %d: operand a reads undefined register %d", i, o.a);
            last_use[cur_def[o.a]] = i;
        }
        if (nopnd >= 2) {
            VX_REQUIRE(o.b >= 0 && o.b < VXP_MAX_REGS && cur_def[o.b] >= 0, VX_E_INVALID_ARG,
                       "instruction %d: operand b reads undefined register %d", i, o.b);
            last_use[cur_def[o.b]] = i;
        }
        cur_def[o.dst] = i;
    }

    std::vector<int> free_b, free_f;
    int n_b = 0, n_f = 0;
    int phys_of_def[VXP_MAX_OPS], type_of_def[VXP_MAX_OPS];
    for (int i = 0; i < n_ops; i++) { phys_of_def[i] = -1; type_of_def[i] = T_NONE; }
    int out_type = T_NONE;
    int nd = 0;
    for (int i = 0; i < n_ops; i++) {
        const vx_op& o = prog[i];
        const bool last = (i == n_ops - 1);
        if (o.opcode == VXP_LOAD) {
            VX_REQUIRE(o.a >= 0 && o.a < n_leaves, VX_E_INVALID_ARG,
                       "instruction %d: leaf %d out of range", i, o.a);
            PubReg& d = regs[o.dst];
            d.type = (L[o.a]->dtype == VX_U8) ? T_B : T_F;
            d.leaf = o.a; d.phys = -1; d.def = i;
            if (last) {
                DevOp& D = P.ops[nd++];
                D.opcode = d.type == T_B ? OPC_COPYB : OPC_COPYF;
                D.dst = -1; D.a_kind = K_LEAF; D.a_idx = (uint8_t)o.a;
                out_type = d.type;
            }
            continue;
        }
        int ta, tb;
        int nopnd = operand_types(o.opcode, &ta, &tb);
        DevOp& D = P.ops[nd++];
        D.opcode = (uint8_t)o.opcode;
        D.aux = (uint8_t)o.aux;
        D.imm = o.imm;
        if (o.opcode == VXP_CMPS || o.opcode == VXP_CMP || o.opcode == VXP_CMPSB)
            VX_REQUIRE(o.aux >= VX_LT && o.aux <= VX_NE, VX_E_INVALID_ARG,
                       "instruction %d: comparison %d unknown", i, o.aux);
        if (o.opcode == VXP_ARITHS || o.opcode == VXP_ARITH || o.opcode == VXP_ARITHSB)
            VX_REQUIRE(o.aux >= VX_ADD && o.aux <= VX_MAX, VX_E_INVALID_ARG,
                       "instruction %d: arithmetic operator %d unknown", i, o.aux);
        if (o.opcode == VXP_CMPSB || o.opcode == VXP_ARITHSB) {
            VX_REQUIRE(batch_scalars != nullptr && o.b >= 0 && o.b < n_scalars, VX_E_INVALID_ARG,
                       "instruction %d: per-volume scalar %d not supplied", i, o.b);
            D.b_idx = (uint8_t)o.b;
        }
        int deferred_free[2] = {-1, -1};
        for (int k = 0; k < nopnd; k++) {
            int r = k == 0 ? o.a : o.b;
            int want = k == 0 ? ta : tb;
            PubReg& pr = regs[r];
            VX_REQUIRE(pr.type == want, VX_E_DTYPE,
                       "instruction %d: operand %c is %s but the operator needs %s", i, 'a' + k,
                       pr.type == T_B ? "boolean" : "numeric", want == T_B ? "boolean" : "numeric");
            uint8_t kind = pr.leaf >= 0 ? K_LEAF : K_REG;
            uint8_t idx = (uint8_t)(pr.leaf >= 0 ? pr.leaf : pr.phys);
            if (k == 0) { D.a_kind = kind; D.a_idx = idx; }
            else { D.b_kind = kind; D.b_idx = idx; }
            if (pr.leaf < 0 && last_use[pr.def] == i) deferred_free[k] = pr.def;
        }
        // operands are read into registers before dst is written, so a register whose
        // last use is this instruction can be reused as its destination
        for (int k = 0; k < 2; k++) {
            int d = deferred_free[k];
            if (d < 0 || (k == 1 && d == deferred_free[0])) continue;
            if (type_of_def[d] == T_B) free_b.push_back(phys_of_def[d]);
            else free_f.push_back(phys_of_def[d]);
        }
        int rt = result_type(o.opcode);
        PubReg& d = regs[o.dst];
        // (a value that is never read keeps its physical register: dead code is the
        // caller's business, it only costs shared memory)
        d.type = rt; d.leaf = -1; d.def = i;
        if (last) {
            D.dst = -1;
            d.phys = -1;
            out_type = rt;
        } else {
            int p;
            if (rt == T_B) {
                if (!free_b.empty()) { p = free_b.back(); free_b.pop_back(); } else p = n_b++;
            } else {
                if (!free_f.empty()) { p = free_f.back(); free_f.pop_back(); } else p = n_f++;
            }
            d.phys = p;
            phys_of_def[i] = p; type_of_def[i] = rt;
            D.dst = (int8_t)p;
        }
    }
    size_t smem = (size_t)kPwThreads * (16 * (size_t)n_b + 64 * (size_t)n_f);
    VX_REQUIRE(smem <= 200 * 1024, VX_E_UNSUPPORTED,
               "program needs %d boolean + %d numeric live registers (%zu B shared memory); split it",
               n_b, n_f, smem);

    Image* o = nullptr;
    if (out_type == T_B) VX_TRY(new_bits_image(g.c, shape->nx, shape->ny, shape->nz, shape->nb, shape->sp, &o));
    else VX_TRY(new_image(g.c, VX_F32, shape->nx, shape->ny, shape->nz, shape->nb, shape->sp, &o));
    P.n_ops = nd;
    P.n_breg = n_b; P.n_freg = n_f;
    P.nb = shape->nb;
    P.nvox = shape->nvox();
    P.count = shape->count();
    P.ngroups = (shape->count() + kGroup - 1) / kGroup;
    for (int i = 0; i < n_leaves; i++) {
        P.leaf[i] = leaf_ptr[i];
        P.leaf_tab[i] = leaf_tab[i];
        P.leaf_mask[i] = leaf_mask[i];
        P.leaf_flip[i] = leaf_flip[i];
        P.leaf_codes[i] = leaf_codes[i];
        if (leaf_codes[i]) { P.leaf_slope[i] = L[i]->q_slope; P.leaf_inter[i] = L[i]->q_inter; }
    }
    P.out = out_type == T_B ? (void*)o->bits : o->d;
    // Deferred leaves that are only ever compared with constants are read in rank mode (DevProg::leaf_rank):
    // every such comparison becomes one against a per-volume break point, appended to the scalar array.
    BreakSpecs bs;
    bs.n = 0;
    if (n_scalars < 0 || !batch_scalars) n_scalars = 0;
    for (int i = 0; i < n_leaves; i++) {
        if (leaf_lazy[i] == nullptr || !leaf_lazy[i]->monotone) continue;
        bool ok = true;
        int uses = 0;
        for (int j = 0; j < nd && ok; j++) {
            const DevOp& D = P.ops[j];
            int ta, tb;
            const int nopnd = D.opcode == OPC_COPYF ? 1 : D.opcode == OPC_COPYB ? 0 : operand_types(D.opcode, &ta, &tb);
            const bool a_is = nopnd >= 1 && D.a_kind == K_LEAF && D.a_idx == i && (D.opcode == OPC_COPYF || ta == T_F);
            const bool b_is = nopnd >= 2 && D.b_kind == K_LEAF && D.b_idx == i && tb == T_F;
            if (b_is) ok = false;
            if (a_is) {
                if (D.opcode == VXP_CMPS && D.aux >= VX_LT && D.aux <= VX_GE) uses++;
                else ok = false;
            }
        }
        if (!ok || uses == 0 || n_scalars + bs.n + uses > 255) continue;
        for (int j = 0; j < nd; j++) {
            DevOp& D = P.ops[j];
            if (!(D.opcode == VXP_CMPS && D.a_kind == K_LEAF && D.a_idx == i)) continue;
            BreakSpec& b = bs.s[bs.n];
            b.table = leaf_lazy[i]->table;
            b.qrange = leaf_lazy[i]->src->q_range();
            b.incr = leaf_lazy[i]->src->q_increasing ? 1 : 0;
            b.cmp = D.aux;
            b.c = D.imm;
            b.slot = n_scalars + bs.n;
            D.opcode = VXP_CMPSB;
            D.b_idx = (uint8_t)b.slot;
            D.aux = (D.aux == VX_GT || D.aux == VX_GE) ? VX_GE : VX_LT;
            bs.n++;
        }
        P.leaf_rank[i] = 1;
        P.leaf_qrange[i] = leaf_lazy[i]->src->q_range();
        P.leaf_incr[i] = leaf_lazy[i]->src->q_increasing ? 1 : 0;
    }
    float* d_scal = nullptr;
    if (bs.n > 0 && n_scalars == 0) {
        int r = scratch_alloc(g.c, g.s, sizeof(float) * (size_t)bs.n * shape->nb, (void**)&d_scal);
        if (r != VX_OK) { drop(o); return r; }
        P.bscal = d_scal;
    }
    if (n_scalars > 0) {
        size_t bytes = sizeof(float) * (size_t)n_scalars * shape->nb;
        int r = scratch_alloc(g.c, g.s, sizeof(float) * (size_t)(n_scalars + bs.n) * shape->nb, (void**)&d_scal);
        if (r != VX_OK) { drop(o); return r; }
        // pageable source: the runtime stages it before returning
        cudaError_t e = cudaMemcpyAsync(d_scal, batch_scalars, bytes, cudaMemcpyHostToDevice, g.st);
        if (e != cudaSuccess) { drop(o); return fail(VX_E_CUDA, "scalar upload: %s", cudaGetErrorString(e)); }
        P.bscal = d_scal;
    }
    if (bs.n > 0) {
        {
            VX_LAUNCH(g.c, g.s, "pw_break_points_kernel");
            pw_break_points_kernel<<<(bs.n * shape->nb + 127) / 128, 128, 0, g.st>>>(bs, shape->nb, d_scal);
        }
        int r = check_launch("pw_break_points_kernel");
        if (r != VX_OK) { scratch_free(g.c, g.s, d_scal); drop(o); return r; }
    }
    int def_mode = 0;   // see ld_f
    for (int i = 0; i < n_leaves; i++)
        if (leaf_tab[i] != nullptr) def_mode = (P.leaf_rank[i] && def_mode != 1) ? 2 : 1;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(def_mode == 0 ? pointwise_program_kernel<0>
                                             : def_mode == 2 ? pointwise_program_kernel<2> : pointwise_program_kernel<1>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { drop(o); return fail(VX_E_CUDA, "smem attribute: %s", cudaGetErrorString(e)); }
    }
    // boolean-only programs take the wide kernel (kWideU groups per thread per iteration)
    bool boolean_only = n_f == 0 && (size_t)kPwThreads * 16 * kWideU * (size_t)n_b <= 48 * 1024;
    for (int i = 0; i < nd && boolean_only; i++) {
        const uint8_t oc = P.ops[i].opcode;
        boolean_only = oc == VXP_NOT || oc == VXP_AND || oc == VXP_OR || oc == VXP_XOR || oc == OPC_COPYB ||
                       oc == VXP_CONSTB;
    }
    if (boolean_only) {
        smem *= kWideU;
        P.ngroups = (shape->count() + kBitGroup - 1) / kBitGroup;
    }
    int ctas = smem == 0 ? 8 : (int)((220 * 1024) / smem);
    if (ctas < 1) ctas = 1;
    if (ctas > 8) ctas = 8;
    int grid = grid_for(boolean_only ? (P.ngroups + kWideU - 1) / kWideU : P.ngroups, kPwThreads, ctas);
    // one comparison of a coded leaf with a constant (or a per-volume scalar): the 32-voxels-per-thread kernel
    bool code_cmp = false;
    const bool negated = nd == 2 && P.ops[1].opcode == VXP_NOT && P.ops[1].a_kind == K_REG && P.ops[0].dst >= 0 &&
                         P.ops[1].a_idx == (uint8_t)P.ops[0].dst && P.ops[1].dst < 0;
    if ((nd == 1 || negated) && out_type == T_B && (P.ops[0].opcode == VXP_CMPS || P.ops[0].opcode == VXP_CMPSB) &&
        P.ops[0].a_kind == K_LEAF && g.c->quantised_paths.load()) {
        const int li = P.ops[0].a_idx;
        const bool rank = leaf_tab[li] != nullptr && P.leaf_rank[li] != 0;
        if (rank || (leaf_tab[li] == nullptr && leaf_codes[li] != 0)) {
            CodeCmp a;
            memset(&a, 0, sizeof(a));
            a.codes = reinterpret_cast<const uint4*>(leaf_ptr[li]);
            a.mask = reinterpret_cast<const unsigned*>(leaf_mask[li]);
            a.qrange = P.leaf_qrange[li];
            a.thr = P.ops[0].opcode == VXP_CMPSB ? P.bscal + (size_t)P.ops[0].b_idx * shape->nb : nullptr;
            a.imm = P.ops[0].imm;
            a.slope = P.leaf_slope[li]; a.inter = P.leaf_inter[li];
            a.flip = leaf_flip[li];
            a.sgn = leaf_codes[li] == 1; a.incr = P.leaf_incr[li]; a.nb = shape->nb;
            a.nvox = P.nvox; a.count = P.count;
            a.invert = negated ? 0xffffffffu : 0u;
            a.out = reinterpret_cast<unsigned*>(o->bits);
            const int cgrid = grid_for((P.count + 31) / 32, kPwThreads, 8);
            {
                VX_LAUNCH(g.c, g.s, "code_compare_kernel");
                if (rank) launch_code_compare<true>(P.ops[0].aux, cgrid, g.st, a);
                else launch_code_compare<false>(P.ops[0].aux, cgrid, g.st, a);
            }
            code_cmp = true;
        }
    }
    if (!code_cmp) {
        VX_LAUNCH(g.c, g.s, boolean_only ? "boolean_program_kernel" : "pointwise_program_kernel");
        if (boolean_only) boolean_program_kernel<<<grid, kPwThreads, smem, g.st>>>(P);
        else if (def_mode == 0) pointwise_program_kernel<0><<<grid, kPwThreads, smem, g.st>>>(P);
        else if (def_mode == 2) pointwise_program_kernel<2><<<grid, kPwThreads, smem, g.st>>>(P);
        else pointwise_program_kernel<1><<<grid, kPwThreads, smem, g.st>>>(P);
    }
    int r = check_launch(code_cmp ? "code_compare_kernel" : "pointwise_program_kernel");
    if (r == VX_OK && d_scal) r = scratch_free(g.c, g.s, d_scal);
    if (r != VX_OK) { drop(o); return r; }
    return g.finish(o, out);
}

}  // namespace vx

using namespace vx;

static int32_t unary(vx_ctx ctx, vx_img a, int opcode, int aux, float imm, vx_img* out) {
    vx_op p[2] = {{VXP_LOAD, 0, 0, 0, 0, 0.f}, {opcode, 1, 0, 0, aux, imm}};
    return run_pointwise(ctx, p, 2, &a, 1, nullptr, 0, 0, out);
}
static int32_t binary(vx_ctx ctx, vx_img a, vx_img b, int opcode, int aux, float imm, vx_img* out) {
    vx_op p[3] = {{VXP_LOAD, 0, 0, 0, 0, 0.f}, {VXP_LOAD, 1, 1, 0, 0, 0.f}, {opcode, 2, 0, 1, aux, imm}};
    vx_img l[2] = {a, b};
    return run_pointwise(ctx, p, 3, l, 2, nullptr, 0, 0, out);
}

extern "C" {

int32_t vx_fused_pointwise(vx_ctx ctx, const vx_op* program, int32_t n_ops, const vx_img* leaves,
                           int32_t n_leaves, const float* batch_scalars, int32_t n_scalars,
                           vx_img* out) {
    VX_REQUIRE(n_leaves > 0 && leaves != nullptr, VX_E_INVALID_ARG,
               "a fused program needs at least one leaf image");
    return run_pointwise(ctx, program, n_ops, leaves, n_leaves, batch_scalars, n_scalars, 0, out);
}

int32_t vx_const_bool(vx_ctx ctx, vx_img like, int32_t value, vx_img* out) {
    vx_op p = {VXP_CONSTB, 0, 0, 0, value ? 1 : 0, 0.f};
    return run_pointwise(ctx, &p, 1, nullptr, 0, nullptr, 0, like, out);
}
int32_t vx_const_f32(vx_ctx ctx, vx_img like, float value, vx_img* out) {
    vx_op p = {VXP_CONSTF, 0, 0, 0, 0, value};
    return run_pointwise(ctx, &p, 1, nullptr, 0, nullptr, 0, like, out);
}
int32_t vx_not(vx_ctx ctx, vx_img a, vx_img* out) { return unary(ctx, a, VXP_NOT, 0, 0.f, out); }
int32_t vx_and(vx_ctx ctx, vx_img a, vx_img b, vx_img* out) { return binary(ctx, a, b, VXP_AND, 0, 0.f, out); }
int32_t vx_or(vx_ctx ctx, vx_img a, vx_img b, vx_img* out) { return binary(ctx, a, b, VXP_OR, 0, 0.f, out); }
int32_t vx_xor(vx_ctx ctx, vx_img a, vx_img b, vx_img* out) { return binary(ctx, a, b, VXP_XOR, 0, 0.f, out); }
int32_t vx_cmp_scalar(vx_ctx ctx, vx_img a, int32_t op, float scalar, vx_img* out) {
    return unary(ctx, a, VXP_CMPS, op, scalar, out);
}
int32_t vx_cmp(vx_ctx ctx, vx_img a, vx_img b, int32_t op, vx_img* out) {
    return binary(ctx, a, b, VXP_CMP, op, 0.f, out);
}
int32_t vx_arith_scalar(vx_ctx ctx, vx_img a, int32_t op, float scalar, vx_img* out) {
    return unary(ctx, a, VXP_ARITHS, op, scalar, out);
}
int32_t vx_arith(vx_ctx ctx, vx_img a, vx_img b, int32_t op, vx_img* out) {
    return binary(ctx, a, b, VXP_ARITH, op, 0.f, out);
}
int32_t vx_cmp_scalar_batch(vx_ctx ctx, vx_img a, int32_t op, const float* scalars, vx_img* out) {
    VX_REQUIRE(scalars != nullptr, VX_E_INVALID_ARG, "scalars is NULL");
    vx_op p[2] = {{VXP_LOAD, 0, 0, 0, 0, 0.f}, {VXP_CMPSB, 1, 0, 0, op, 0.f}};
    return run_pointwise(ctx, p, 2, &a, 1, scalars, 1, 0, out);
}
int32_t vx_arith_scalar_batch(vx_ctx ctx, vx_img a, int32_t op, const float* scalars, vx_img* out) {
    VX_REQUIRE(scalars != nullptr, VX_E_INVALID_ARG, "scalars is NULL");
    vx_op p[2] = {{VXP_LOAD, 0, 0, 0, 0, 0.f}, {VXP_ARITHSB, 1, 0, 0, op, 0.f}};
    return run_pointwise(ctx, p, 2, &a, 1, scalars, 1, 0, out);
}
int32_t vx_to_f32(vx_ctx ctx, vx_img a, vx_img* out) { return unary(ctx, a, VXP_TOF32, 0, 0.f, out); }
int32_t vx_to_bool(vx_ctx ctx, vx_img a, vx_img* out) { return unary(ctx, a, VXP_TOBOOL, 0, 0.f, out); }
int32_t vx_mask(vx_ctx ctx, vx_img a, vx_img m, float fill, vx_img* out) {
    vx_op p[3] = {{VXP_LOAD, 0, 0, 0, 0, 0.f}, {VXP_LOAD, 1, 1, 0, 0, 0.f}, {VXP_SELECT, 2, 1, 0, 0, fill}};
    vx_img l[2] = {a, m};
    return run_pointwise(ctx, p, 3, l, 2, nullptr, 0, 0, out);
}

}  // extern "C"
