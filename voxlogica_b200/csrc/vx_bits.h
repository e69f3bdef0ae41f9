// Device helpers for the bit form of boolean images (Image::bits, vx_internal.h): a flat
// little-endian bit stream over the whole batch, bit (i & 31) of word (i >> 5) = voxel i.
// Row-structured kernels (connected components, distance balls, near/interior) work on the
// 32-voxel words of a row: word w of the row that starts at voxel `row * nx` is the 32 bits at
// stream position row * nx + 32 w, which need not be word aligned (nx = 240: every other row
// starts on a half word) -- one funnel shift on the way in, half-word stores on the way out.
#pragma once
#include <stdint.h>

namespace vx {

// bits [p, p + valid) of the stream in the low bits of the result, the rest 0  (1 <= valid <= 32)
__device__ __forceinline__ unsigned flat_load32(const unsigned* __restrict__ fb, unsigned long long p, int valid) {
    const unsigned long long w = p >> 5;
    const unsigned sh = (unsigned)p & 31u;
    unsigned v = __ldg(fb + w);
    if (sh) {
        const unsigned hi = (sh + (unsigned)valid > 32u) ? __ldg(fb + w + 1) : 0u;
        v = __funnelshift_r(v, hi, sh);
    }
    return valid >= 32 ? v : (v & ((1u << valid) - 1u));
}
// one bit of the stream
__device__ __forceinline__ unsigned flat_bit(const unsigned* __restrict__ fb, unsigned long long p) {
    return (__ldg(fb + (p >> 5)) >> ((unsigned)p & 31u)) & 1u;
}
// word w of row `row`: valid bits only, positions past the end of the row are 0
__device__ __forceinline__ unsigned flat_row_word(const unsigned* __restrict__ fb, unsigned long long row, int w, int nx) {
    const int valid = min(32, nx - w * 32);
    return flat_load32(fb, row * (unsigned long long)nx + (unsigned)w * 32u, valid);
}
// Store the low `valid` bits of v at stream position p.  align16: every row starts on a half word
// and has a whole number of half words (nx % 16 == 0), so a row word is one aligned word or two
// half words that nobody else writes.  Otherwise neighbouring row words share stream words: the
// destination must have been zeroed and the pieces are merged with atomicOr.
__device__ __forceinline__ void flat_store32(unsigned* __restrict__ fb, unsigned long long p, int valid, unsigned v,
                                             bool align16) {
    const unsigned sh = (unsigned)p & 31u;
    if (sh == 0 && valid == 32) {
        fb[p >> 5] = v;
    } else if (align16) {
        unsigned short* h = reinterpret_cast<unsigned short*>(fb) + (p >> 4);
        h[0] = (unsigned short)(v & 0xffffu);
        if (valid > 16) h[1] = (unsigned short)(v >> 16);
    } else if (v) {
        atomicOr(fb + (p >> 5), v << sh);
        if (sh + (unsigned)valid > 32u) atomicOr(fb + (p >> 5) + 1, v >> (32u - sh));
    }
}
__device__ __forceinline__ void flat_store_row_word(unsigned* __restrict__ fb, unsigned long long row, int w, int nx,
                                                    unsigned v, bool align16) {
    const int valid = min(32, nx - w * 32);
    flat_store32(fb, row * (unsigned long long)nx + (unsigned)w * 32u, valid, valid >= 32 ? v : (v & ((1u << valid) - 1u)),
                 align16);
}

}  // namespace vx
