// roll_cores.cuh -- per-lane arithmetic of the register-rolling kernels (Gaussian 3/5/7 and
// Sobel + L1 magnitude + non-maximum suppression).  One lane owns 4 adjacent pixels (one 32-bit
// word) of a row and walks down the image; two pixels travel in one register as unsigned 16-bit
// halves ("pair": low half = left pixel), so one packed add / multiply / min / max does two pixels.
// Everything here is exact integer arithmetic.
//
// The functions are host+device so that tests/host/roll_host.cpp can run the very same code
// lane by lane on the CPU (no GPU in the build container) and compare it with the CPU checker.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#include <cuda_fp16.h>
#define I2S_HD __host__ __device__ __forceinline__
#else
#define I2S_HD inline
#endif

namespace i2s {
namespace roll {

// ------------------------------------------------------------------ packed primitives
I2S_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
#endif
}
I2S_HD uint32_t max2(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __vmaxu2(a, b);
#else
    uint32_t al = a & 0xffff, bl = b & 0xffff, ah = a >> 16, bh = b >> 16;
    return (al > bl ? al : bl) | ((ah > bh ? ah : bh) << 16);
#endif
}
I2S_HD uint32_t min2(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __vminu2(a, b);
#else
    uint32_t al = a & 0xffff, bl = b & 0xffff, ah = a >> 16, bh = b >> 16;
    return (al < bl ? al : bl) | ((ah < bh ? ah : bh) << 16);
#endif
}
// d = c + a.lo16 * b.byte0 + a.hi16 * b.byte1
I2S_HD uint32_t dp2(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
    return __dp2a_lo(a, b, c);
#else
    return c + (a & 0xffff) * (b & 0xff) + (a >> 16) * ((b >> 8) & 0xff);
#endif
}
// pair helpers: a word holds pixels p0..p3 in bytes 0..3
I2S_HD uint32_t pair_lo(uint32_t w) { return prmt(w, 0, 0x4140); }      // (p0, p1)
I2S_HD uint32_t pair_hi(uint32_t w) { return prmt(w, 0, 0x4342); }      // (p2, p3)
// odd-aligned pair (hi half of a, lo half of b)
I2S_HD uint32_t pair_mid(uint32_t a, uint32_t b) { return prmt(a, b, 0x5432); }
constexpr uint32_t ONE2 = 0x00010001u;

// ------------------------------------------------------------------ Gaussian 3/5/7 (A.2)
// Vertical pass on raw bytes: rl/rh = pairs of the rows y-3 .. y+3.  Q8 sums (<= 65280) stay
// inside their 16-bit halves.  V[2k], V[2k+1] = pairs (p0,p1), (p2,p3) of kernel k (3,5,7).
I2S_HD void gauss_vertical(const uint32_t (&rl)[7], const uint32_t (&rh)[7], uint32_t (&V)[6])
{
    const uint32_t a0l = rl[3], a1l = rl[2] + rl[4], a2l = rl[1] + rl[5], a3l = rl[0] + rl[6];
    const uint32_t a0h = rh[3], a1h = rh[2] + rh[4], a2h = rh[1] + rh[5], a3h = rh[0] + rh[6];
    V[0] = 88u * a0l + 84u * a1l;
    V[1] = 88u * a0h + 84u * a1h;
    V[2] = 54u * a0l + 52u * a1l + 49u * a2l;
    V[3] = 54u * a0h + 52u * a1h + 49u * a2h;
    V[4] = 38u * (a0l + a1l) + 36u * a2l + 35u * a3l;
    V[5] = 38u * (a0h + a1h) + 36u * a2h + 35u * a3h;
}

I2S_HD uint32_t pack_q16(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3)
{
    // byte 2 of each Q16 result (rounding constant already added)
    return prmt(prmt(r0, r1, 0x0062), prmt(r2, r3, 0x0062), 0x5410);
}

// Horizontal passes on the 16-bit column sums.  E0,E1 = this lane's pairs (V0,V1), (V2,V3);
// Em1 = (V-2,V-1) and Em2 = (V-4,V-3) come from the left lane, Ep2 = (V4,V5), Ep3 = (V6,V7) from
// the right lane.  Result: 4 output bytes.
I2S_HD uint32_t gauss_h3(uint32_t Em1, uint32_t E0, uint32_t E1, uint32_t Ep2)
{
    constexpr uint32_t K10 = 84u | (88u << 8), K1_ = 84u, K_1 = 84u << 8, R = 32768u;
    const uint32_t Om1 = pair_mid(Em1, E0), O0 = pair_mid(E0, E1);
    const uint32_t r0 = dp2(Om1, K10, dp2(E0, K_1, R));
    const uint32_t r1 = dp2(E0, K10, dp2(E1, K1_, R));
    const uint32_t r2 = dp2(O0, K10, dp2(E1, K_1, R));
    const uint32_t r3 = dp2(E1, K10, dp2(Ep2, K1_, R));
    return pack_q16(r0, r1, r2, r3);
}

I2S_HD uint32_t gauss_h5(uint32_t Em1, uint32_t E0, uint32_t E1, uint32_t Ep2)
{
    constexpr uint32_t K21 = 49u | (52u << 8), K01 = 54u | (52u << 8), K2_ = 49u, K_2 = 49u << 8, R = 32768u;
    const uint32_t Om1 = pair_mid(Em1, E0), O0 = pair_mid(E0, E1), O1 = pair_mid(E1, Ep2);
    const uint32_t r0 = dp2(Em1, K21, dp2(E0, K01, dp2(E1, K2_, R)));
    const uint32_t r1 = dp2(Om1, K21, dp2(O0, K01, dp2(O1, K2_, R)));
    const uint32_t r2 = dp2(E0, K21, dp2(E1, K01, dp2(Ep2, K2_, R)));
    const uint32_t r3 = dp2(O0, K21, dp2(O1, K01, dp2(Ep2, K_2, R)));
    return pack_q16(r0, r1, r2, r3);
}

I2S_HD uint32_t gauss_h7(uint32_t Em2, uint32_t Em1, uint32_t E0, uint32_t E1, uint32_t Ep2, uint32_t Ep3)
{
    constexpr uint32_t K32 = 35u | (36u << 8), K10 = 38u | (38u << 8), K12 = 38u | (36u << 8), K3_ = 35u,
                       K_3 = 35u << 8, R = 32768u;
    const uint32_t Om2 = pair_mid(Em2, Em1), Om1 = pair_mid(Em1, E0), O0 = pair_mid(E0, E1), O1 = pair_mid(E1, Ep2);
    const uint32_t r0 = dp2(Om2, K32, dp2(Om1, K10, dp2(O0, K12, dp2(O1, K3_, R))));
    const uint32_t r1 = dp2(Em1, K32, dp2(E0, K10, dp2(E1, K12, dp2(Ep2, K3_, R))));
    const uint32_t r2 = dp2(Om1, K32, dp2(O0, K10, dp2(O1, K12, dp2(Ep2, K_3, R))));
    const uint32_t r3 = dp2(E0, K32, dp2(E1, K10, dp2(Ep2, K12, dp2(Ep3, K3_, R))));
    return pack_q16(r0, r1, r2, r3);
}

// ------------------------------------------------------------------ packed half-precision helpers
// Gradients (|d| <= 1020), magnitudes (<= 2040) and the sector quantities are small integers, exact in
// fp16: from the column sums on, the Sobel / NMS arithmetic runs as half2 operations (HADD2 / HFMA2 /
// HSET2 on the FMA pipe) instead of packed 16-bit integer min/max/logic on the half-rate ALU pipe.
// `h2` is the bit pattern of a half2 (low half = left pixel of the pair).  On the host the same
// functions are evaluated through float (exact for these ranges) so that tests/host/roll_host.cpp
// runs the identical dataflow.
typedef uint32_t h2;

#if !defined(__CUDA_ARCH__)
inline float half_bits_to_float(uint32_t hb)
{
    const uint32_t sign = (hb >> 15) & 1u, exp = (hb >> 10) & 31u, man = hb & 1023u;
    float v;
    if (exp == 0) v = (float)man * (1.0f / 16777216.0f);                    // subnormal: man * 2^-24
    else if (exp == 31) v = man ? __builtin_nanf("") : __builtin_inff();
    else {
        v = (float)(1024u + man);
        int e = (int)exp - 25;                                               // (1024 + man) * 2^(exp-25)
        while (e > 0) { v *= 2.0f; e--; }
        while (e < 0) { v *= 0.5f; e++; }
    }
    return sign ? -v : v;
}
inline uint32_t float_to_half_bits(float f)                                  // exact inputs only (integers, |f| <= 2048)
{
    _Float16 hv = (_Float16)f;
    uint16_t b;
    __builtin_memcpy(&b, &hv, 2);
    return b;
}
inline h2 h2_make(float lo, float hi) { return float_to_half_bits(lo) | (float_to_half_bits(hi) << 16); }
inline float h2_lo(h2 a) { return half_bits_to_float(a & 0xffffu); }
inline float h2_hi(h2 a) { return half_bits_to_float(a >> 16); }
#endif

I2S_HD h2 h2_const(float v)                            // both halves = v (an exactly representable value)
{
#if defined(__CUDA_ARCH__)
    const __half2 r = __float2half2_rn(v);
    return *reinterpret_cast<const uint32_t *>(&r);
#else
    return h2_make(v, v);
#endif
}
I2S_HD h2 h2_sub(h2 a, h2 b)
{
#if defined(__CUDA_ARCH__)
    const __half2 r = __hsub2(*reinterpret_cast<const __half2 *>(&a), *reinterpret_cast<const __half2 *>(&b));
    return *reinterpret_cast<const uint32_t *>(&r);
#else
    return h2_make(h2_lo(a) - h2_lo(b), h2_hi(a) - h2_hi(b));
#endif
}
I2S_HD h2 h2_add(h2 a, h2 b)
{
#if defined(__CUDA_ARCH__)
    const __half2 r = __hadd2(*reinterpret_cast<const __half2 *>(&a), *reinterpret_cast<const __half2 *>(&b));
    return *reinterpret_cast<const uint32_t *>(&r);
#else
    return h2_make(h2_lo(a) + h2_lo(b), h2_hi(a) + h2_hi(b));
#endif
}
I2S_HD h2 h2_mul(h2 a, h2 b)                           // only the SIGN of the result is used (overflow to inf is fine)
{
#if defined(__CUDA_ARCH__)
    const __half2 r = __hmul2(*reinterpret_cast<const __half2 *>(&a), *reinterpret_cast<const __half2 *>(&b));
    return *reinterpret_cast<const uint32_t *>(&r);
#else
    auto sgn = [](float x) { return x > 0 ? 1.0f : (x < 0 ? -1.0f : 0.0f); };
    return h2_make(sgn(h2_lo(a)) * sgn(h2_lo(b)), sgn(h2_hi(a)) * sgn(h2_hi(b)));
#endif
}
I2S_HD h2 h2_fma(h2 a, h2 b, h2 c)                     // a * b + c, exact for the operand ranges used here
{
#if defined(__CUDA_ARCH__)
    const __half2 r = __hfma2(*reinterpret_cast<const __half2 *>(&a), *reinterpret_cast<const __half2 *>(&b),
                              *reinterpret_cast<const __half2 *>(&c));
    return *reinterpret_cast<const uint32_t *>(&r);
#else
    return h2_make(h2_lo(a) * h2_lo(b) + h2_lo(c), h2_hi(a) * h2_hi(b) + h2_hi(c));
#endif
}
// 0xffff in every half where a > b / a >= b (HSET2.BM)
I2S_HD uint32_t h2_gt(h2 a, h2 b)
{
#if defined(__CUDA_ARCH__)
    return __hgt2_mask(*reinterpret_cast<const __half2 *>(&a), *reinterpret_cast<const __half2 *>(&b));
#else
    return (h2_lo(a) > h2_lo(b) ? 0xffffu : 0u) | (h2_hi(a) > h2_hi(b) ? 0xffff0000u : 0u);
#endif
}
I2S_HD uint32_t h2_ge(h2 a, h2 b)
{
#if defined(__CUDA_ARCH__)
    return __hge2_mask(*reinterpret_cast<const __half2 *>(&a), *reinterpret_cast<const __half2 *>(&b));
#else
    return (h2_lo(a) >= h2_lo(b) ? 0xffffu : 0u) | (h2_hi(a) >= h2_hi(b) ? 0xffff0000u : 0u);
#endif
}
// |a| + |b|, |a| * k + |b|, |a| > b: the absolute values are written with the half2 intrinsic inside the
// expression so that the compiler folds them into source modifiers of the consuming instruction
I2S_HD h2 h2_addabs(h2 a, h2 b)
{
#if defined(__CUDA_ARCH__)
    const __half2 r = __hadd2(__habs2(*reinterpret_cast<const __half2 *>(&a)), __habs2(*reinterpret_cast<const __half2 *>(&b)));
    return *reinterpret_cast<const uint32_t *>(&r);
#else
    return h2_add(a & 0x7fff7fffu, b & 0x7fff7fffu);
#endif
}
I2S_HD h2 h2_absadd(h2 a, h2 c)                        // |a| + c
{
#if defined(__CUDA_ARCH__)
    const __half2 r = __hadd2(__habs2(*reinterpret_cast<const __half2 *>(&a)), *reinterpret_cast<const __half2 *>(&c));
    return *reinterpret_cast<const uint32_t *>(&r);
#else
    return h2_add(a & 0x7fff7fffu, c);
#endif
}
I2S_HD h2 h2_fma_abs(h2 a, h2 k, h2 b)                 // |a| * k + |b|
{
#if defined(__CUDA_ARCH__)
    const __half2 r = __hfma2(__habs2(*reinterpret_cast<const __half2 *>(&a)), *reinterpret_cast<const __half2 *>(&k),
                              __habs2(*reinterpret_cast<const __half2 *>(&b)));
    return *reinterpret_cast<const uint32_t *>(&r);
#else
    return h2_fma(a & 0x7fff7fffu, k, b & 0x7fff7fffu);
#endif
}
I2S_HD uint32_t h2_absgt(h2 a, h2 b)                   // |a| > b
{
#if defined(__CUDA_ARCH__)
    return __hgt2_mask(__habs2(*reinterpret_cast<const __half2 *>(&a)), *reinterpret_cast<const __half2 *>(&b));
#else
    return h2_gt(a & 0x7fff7fffu, b);
#endif
}
constexpr uint32_t H2_BIAS = 0x64006400u;              // 1024.0 in both halves: (v | H2_BIAS) is the half2 1024 + v for 0 <= v < 1024

// ------------------------------------------------------------------ Sobel + L1 magnitude (A.4)
// One pixel row of one channel, seen by a lane: its own word plus the words of the two neighbour lanes.
// Kept per row: the three shifted pairs Nm = (p-1,p0), No = (p1,p2), Np = (p3,p4) as 16-bit integers and
// the horizontal smoothing s_j = p[j-1]+2p[j]+p[j+1] of pixels (0,1) / (2,3) as biased half2 (1024 + s).
struct SobelRow { uint32_t Nm, No, Np; h2 sA, sB; };

I2S_HD SobelRow sobel_row(uint32_t word, uint32_t left_word, uint32_t right_word)
{
    // left_word / right_word: the neighbour lanes' words of the same row (p-1 = byte 3 of the left
    // one, p4 = byte 0 of the right one); the zero bytes of the unpacked pairs serve as zero source
    const uint32_t lo = pair_lo(word), hi = pair_hi(word);
    SobelRow r;
    r.Nm = prmt(left_word, lo, 0x5453);        // (p-1, p0)
    r.No = pair_mid(lo, hi);                   // (p1, p2)
    r.Np = prmt(hi, right_word, 0x1412);       // (p3, p4)
    r.sA = (r.Nm + r.No + lo + lo) | H2_BIAS;  // sums <= 1020
    r.sB = (r.No + r.Np + hi + hi) | H2_BIAS;
    return r;
}

// Gradient of the middle row m from rows t (above), m, b (below): signed dx, dy per pixel as half2,
// pairs A = pixels 0,1 and B = pixels 2,3.  The biases of the operands cancel in the differences.
struct Grad { h2 dxA, dxB, dyA, dyB; };

I2S_HD Grad sobel_grad(const SobelRow &t, const SobelRow &m, const SobelRow &b)
{
    const h2 vNm = (t.Nm + b.Nm + m.Nm + m.Nm) | H2_BIAS;      // column sums at columns (-1,0)
    const h2 vNo = (t.No + b.No + m.No + m.No) | H2_BIAS;      // (1,2)
    const h2 vNp = (t.Np + b.Np + m.Np + m.Np) | H2_BIAS;      // (3,4)
    Grad g;
    g.dxA = h2_sub(vNo, vNm);                                   // dx(0,1) = v(1,2) - v(-1,0)
    g.dxB = h2_sub(vNp, vNo);                                   // dx(2,3) = v(3,4) - v(1,2)
    g.dyA = h2_sub(b.sA, t.sA);                                 // dy = s(below) - s(above)
    g.dyB = h2_sub(b.sB, t.sB);
    return g;
}

// L1 magnitude |dx| + |dy| (<= 2040, exact)
I2S_HD void grad_mag(const Grad &g, h2 &mA, h2 &mB)
{
    mA = h2_addabs(g.dxA, g.dyA);
    mB = h2_addabs(g.dxB, g.dyB);
}

I2S_HD uint32_t bitsel(uint32_t k, uint32_t a, uint32_t b) { return (a & k) | (b & ~k); }   // k ? a : b, one LOP3

// 3-channel Canny: per pixel keep the gradient of the channel with the largest |dx|+|dy|, the
// first channel winning ties (cv.Canny on a colour image, img2sgf.py:162).  `cur`/`mA,mB` hold the
// best so far and its magnitude; `g` is the next channel.
I2S_HD void grad_select(Grad &cur, h2 &mA, h2 &mB, const Grad &g)
{
    h2 gA, gB;
    grad_mag(g, gA, gB);
    const uint32_t kA = h2_gt(gA, mA), kB = h2_gt(gB, mB);     // 0xffff where g wins
    cur.dxA = bitsel(kA, g.dxA, cur.dxA); cur.dxB = bitsel(kB, g.dxB, cur.dxB);
    cur.dyA = bitsel(kA, g.dyA, cur.dyA); cur.dyB = bitsel(kB, g.dyB, cur.dyB);
    mA = bitsel(kA, gA, mA);
    mB = bitsel(kB, gB, mB);
}

// ------------------------------------------------------------------ non-maximum suppression (A.4)
// A magnitude row as seen by a lane: A = (m0,m1), B = (m2,m3) plus the shifted pairs
// Cm = (m-1,m0), Co = (m1,m2), Cp = (m3,m4) once the neighbour lanes' magnitudes are known (half2).
struct MagRow { h2 A, B, Cm, Co, Cp; };

I2S_HD MagRow mag_row(h2 A, h2 B, h2 leftB, h2 rightA)
{
    MagRow r;
    r.A = A; r.B = B;
    r.Cm = pair_mid(leftB, A);                 // (left lane's m3, m0)
    r.Co = pair_mid(A, B);
    r.Cp = pair_mid(B, rightA);                // (m3, right lane's m0)
    return r;
}

// Sector of the gradient direction, from ax = |dx| and ay = |dy| (both <= 1020):
//   horizontal  <=>  (ay << 15) < ax * 13573                <=>  ay <= hp,  hp = floor(ax*13573 / 2^15)
//   vertical    <=>  (ay << 15) > ax * 13573 + (ax << 16)   <=>  ay - 2 ax > hp
// (ax*13573 is a multiple of 2^15 only for ax = 0, where the two forms of the horizontal test differ at
// ay = 0 alone: magnitude 0, no candidate.)  hp comes from a 1021-entry table of half bit patterns
// (sector_table_entry), indexed with the integer ax recovered from the half by the 1024 bias.
I2S_HD uint16_t sector_table_entry(int ax)
{
    const int hp = (ax * 13573) >> 15;                     // <= 422
#if defined(__CUDA_ARCH__)
    return __half_as_ushort(__int2half_rn(hp));
#else
    return (uint16_t)float_to_half_bits((float)hp);
#endif
}
constexpr int SECTOR_TABLE = 1024;

// masks (0xffff per half): nh = NOT horizontal, v = vertical
I2S_HD void sector_masks(h2 dx, h2 dy, const uint16_t *tab, uint32_t &nh, uint32_t &v)
{
    const uint32_t axi = h2_absadd(dx, H2_BIAS) & 0x03ff03ffu;         // the two integers ax = |dx|
    const h2 hp = (uint32_t)tab[axi & 0xffffu] | ((uint32_t)tab[axi >> 16] << 16);
    nh = h2_absgt(dy, hp);                                               // ay > hp
    v = h2_gt(h2_fma_abs(dx, h2_const(-2.0f), dy), hp);                  // ay - 2 ax > hp
}

// NMS + thresholds for the 4 pixels of a lane.  up/c/dn = magnitude rows y-1, y, y+1; g = gradient
// of row y; low / high = the thresholds in both halves (clamped to 2047: no magnitude exceeds 2040).
// Returns the 4 state bytes: 0 none, 1 weak candidate, 3 strong candidate.
// nms_diag() is kept separate so a warp can skip it when no lane has a diagonal candidate.
struct NmsPartial { uint32_t PA, PB, dA, dB; };   // pass masks so far (axis sectors, above low); diagonal-sector candidates

// some pixel of this lane has a magnitude above `low`
I2S_HD bool any_above(const MagRow &c, h2 low) { return (h2_gt(c.A, low) | h2_gt(c.B, low)) != 0; }

I2S_HD NmsPartial nms_axis(const MagRow &up, const MagRow &c, const MagRow &dn, const Grad &g, h2 low, const uint16_t *tab)
{
    uint32_t nhA, vA, nhB, vB;
    sector_masks(g.dxA, g.dyA, tab, nhA, vA);
    sector_masks(g.dxB, g.dyB, tab, nhB, vB);
    const uint32_t lowA = h2_gt(c.A, low), lowB = h2_gt(c.B, low);
    // horizontal: m > left && m >= right ; vertical: m > up && m >= down
    const uint32_t hA = ~nhA & h2_gt(c.A, c.Cm) & h2_ge(c.A, c.Co), hB = ~nhB & h2_gt(c.B, c.Co) & h2_ge(c.B, c.Cp);
    const uint32_t wA = vA & h2_gt(c.A, up.A) & h2_ge(c.A, dn.A), wB = vB & h2_gt(c.B, up.B) & h2_ge(c.B, dn.B);
    NmsPartial p;
    p.PA = (hA | wA) & lowA;
    p.PB = (hB | wB) & lowB;
    p.dA = nhA & ~vA & lowA;               // diagonal sector and above low
    p.dB = nhB & ~vB & lowB;
    return p;
}

I2S_HD bool nms_needs_diag(const NmsPartial &p) { return (p.dA | p.dB) != 0; }

// diagonal sector: same sign of dx,dy -> m > UL && m > DR ; opposite sign -> m > UR && m > DL
I2S_HD void nms_diag(NmsPartial &p, const MagRow &up, const MagRow &c, const MagRow &dn, const Grad &g)
{
    const h2 zero = 0u;
    const uint32_t sdA = h2_gt(zero, h2_mul(g.dxA, g.dyA)), sdB = h2_gt(zero, h2_mul(g.dxB, g.dyB));   // signs differ
    const uint32_t sameA = h2_gt(c.A, up.Cm) & h2_gt(c.A, dn.Co), sameB = h2_gt(c.B, up.Co) & h2_gt(c.B, dn.Cp);
    const uint32_t oppA = h2_gt(c.A, up.Co) & h2_gt(c.A, dn.Cm), oppB = h2_gt(c.B, up.Cp) & h2_gt(c.B, dn.Co);
    p.PA |= p.dA & bitsel(sdA, oppA, sameA);
    p.PB |= p.dB & bitsel(sdB, oppB, sameB);
}

I2S_HD uint32_t nms_state(const NmsPartial &p, const MagRow &c, h2 high)
{
    // per half: 0 where rejected, else 1 + 2 * (m > high)
    const uint32_t stA = p.PA & ((h2_gt(c.A, high) & 0x00020002u) | 0x00010001u);
    const uint32_t stB = p.PB & ((h2_gt(c.B, high) & 0x00020002u) | 0x00010001u);
    return prmt(stA, stB, 0x6420);
}

}  // namespace roll
}  // namespace i2s
