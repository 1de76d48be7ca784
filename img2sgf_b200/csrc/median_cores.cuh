// median_cores.cuh -- bit-sliced saturated-window verdicts of the median kernel (preproc.cu).  Host+device so
// that tests/host/median_host.cpp can check the very same code against brute-force window counts on the CPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define I2S_HD __host__ __device__ __forceinline__
#else
#define I2S_HD inline
#endif

namespace i2s {

constexpr int MT_W = 64, MT_H = 32;                  // output tile of the median kernels

// bits [sh, sh + 32) of the 64-bit value hi:lo, 0 <= sh < 32
I2S_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, int sh)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> sh);
#endif
}

// Saturated-window verdicts for a whole tile, bit-sliced: with one bit per pixel (planes 8 and 9:
// pixel == 255 / pixel == 0) the number of saturated pixels in a B x B window is a sum of B*B bits,
// evaluated for 32 pixels at once with full adders on 32-bit words -- a vertical carry-save count of the
// B rows, then B horizontally shifted copies of that 3-bit number added into a 6-bit accumulator, then a
// comparison with the majority count.  One thread does one (window size, output row, 32-pixel word);
// ~10 logic operations per pixel for all three sizes and both planes, against ~90 for gathering and
// counting the windows per pixel.
I2S_HD void bs_full_add(uint32_t a, uint32_t b, uint32_t c, uint32_t &s, uint32_t &cy)
{
    s = a ^ b ^ c;
    cy = (a & b) | (c & (a ^ b));
}

template <int B, int RS, int HX, int SH, int GW>
I2S_HD uint32_t settle_word(const uint32_t (&plane)[SH][GW + 1], int ty, int j)
{
    constexpr int R = B / 2, RO = RS - R;
    constexpr int KM = (B * B) / 2 + 1;              // a value held by KM window pixels is the median
    constexpr int NB = B == 3 ? 4 : (B == 5 ? 5 : 6);   // bits of the count (max B*B)
    const int w0 = (HX + 32 * j - R) >> 5;           // the two staged words the shifted windows come from
    uint32_t v[3][2];                                // vertical count (3 bits) of the B rows, words w0 and w0 + 1
#pragma unroll
    for (int k = 0; k < 2; k++) {
        uint32_t r[B];
#pragma unroll
        for (int i = 0; i < B; i++) r[i] = plane[ty + RO + i][w0 + k];
        if (B == 3) { bs_full_add(r[0], r[1], r[2], v[0][k], v[1][k]); v[2][k] = 0; }
        else if (B == 5) {
            uint32_t s1, c1, c2;
            bs_full_add(r[0], r[1], r[2], s1, c1);
            bs_full_add(s1, r[3], r[4], v[0][k], c2);
            v[1][k] = c1 ^ c2; v[2][k] = c1 & c2;
        } else {
            uint32_t s1, c1, s2, c2, c3;
            bs_full_add(r[0], r[1], r[2], s1, c1);
            bs_full_add(r[3], r[4], r[5], s2, c2);
            bs_full_add(s1, s2, r[B - 1], v[0][k], c3);
            bs_full_add(c1, c2, c3, v[1][k], v[2][k]);
        }
    }
    uint32_t acc[NB];
#pragma unroll
    for (int i = 0; i < NB; i++) acc[i] = 0;
#pragma unroll
    for (int dx = -R; dx <= R; dx++) {
        const int sh = (HX + 32 * j + dx) - 32 * w0;   // 13 .. 19 (+ 32 never: w0 is the word of the leftmost window)
        uint32_t carry = 0;
#pragma unroll
        for (int i = 0; i < NB; i++) {
            const uint32_t nbit = i < 3 ? funnel_r(v[i][0], v[i][1], sh) : 0u;
            uint32_t sum;
            bs_full_add(acc[i], nbit, carry, sum, carry);
            acc[i] = sum;
        }
    }
    // count >= KM, most significant bit first
    uint32_t ge = 0, eq = 0xffffffffu;
#pragma unroll
    for (int i = NB - 1; i >= 0; i--) {
        if ((KM >> i) & 1) eq &= acc[i];
        else { ge |= eq & acc[i]; eq &= ~acc[i]; }
    }
    return ge | eq;
}


}  // namespace i2s
