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

// ------------------------------------------------------------------ Sobel + L1 magnitude (A.4)
// One pixel row of one channel, seen by a lane: its own word plus the words of the two neighbour lanes.  Kept per row: the three shifted pairs Nm = (p-1,p0), No = (p1,p2),
// Np = (p3,p4) and the horizontal smoothing sA = (s0,s1), sB = (s2,s3), s_j = p[j-1]+2p[j]+p[j+1].
struct SobelRow { uint32_t Nm, No, Np, sA, sB; };

I2S_HD SobelRow sobel_row(uint32_t word, uint32_t left_word, uint32_t right_word)
{
    // left_word / right_word: the neighbour lanes' words of the same row (p-1 = byte 3 of the left
    // one, p4 = byte 0 of the right one); the zero bytes of the unpacked pairs serve as zero source
    const uint32_t lo = pair_lo(word), hi = pair_hi(word);
    SobelRow r;
    r.Nm = prmt(left_word, lo, 0x5453);        // (p-1, p0)
    r.No = pair_mid(lo, hi);                   // (p1, p2)
    r.Np = prmt(hi, right_word, 0x1412);       // (p3, p4)
    r.sA = r.Nm + r.No + lo + lo;
    r.sB = r.No + r.Np + hi + hi;
    return r;
}

// Gradient of the middle row m from rows t (above), m, b (below): |dx|, |dy| per pixel as pairs
// A = pixels 0,1 and B = pixels 2,3, and two "sign carriers": fx != 0 <=> dx < 0, fy != 0 <=> dy < 0
// (per 16-bit half).
struct Grad { uint32_t axA, axB, ayA, ayB, fxA, fxB, fyA, fyB; };

I2S_HD Grad sobel_grad(const SobelRow &t, const SobelRow &m, const SobelRow &b)
{
    const uint32_t vNm = t.Nm + b.Nm + m.Nm + m.Nm;      // column sums at columns (-1,0)
    const uint32_t vNo = t.No + b.No + m.No + m.No;      // (1,2)
    const uint32_t vNp = t.Np + b.Np + m.Np + m.Np;      // (3,4)
    Grad g;
    uint32_t mx;
    mx = max2(vNo, vNm); g.axA = mx - min2(vNo, vNm); g.fxA = mx ^ vNo;      // dx(0,1) = v(1,2) - v(-1,0)
    mx = max2(vNp, vNo); g.axB = mx - min2(vNp, vNo); g.fxB = mx ^ vNp;      // dx(2,3) = v(3,4) - v(1,2)
    mx = max2(b.sA, t.sA); g.ayA = mx - min2(b.sA, t.sA); g.fyA = mx ^ b.sA;  // dy = s(below) - s(above)
    mx = max2(b.sB, t.sB); g.ayB = mx - min2(b.sB, t.sB); g.fyB = mx ^ b.sB;
    return g;
}

// 3-channel Canny: per pixel keep the gradient of the channel with the largest |dx|+|dy|, the
// first channel winning ties (cv.Canny on a colour image, img2sgf.py:162).  `cur`/`mcur` hold the
// best so far and its magnitude pairs; `g` is the next channel.
I2S_HD void grad_select(Grad &cur, uint32_t &mA, uint32_t &mB, const Grad &g)
{
    const uint32_t gA = g.axA + g.ayA, gB = g.axB + g.ayB;
    // take g where gA > mA  <=>  max(gA, mA + 1) == gA ... as a 0/0xffff mask per half
    const uint32_t tA = min2(max2(gA, mA + ONE2) ^ gA, ONE2), tB = min2(max2(gB, mB + ONE2) ^ gB, ONE2);
    const uint32_t kA = (tA ^ ONE2) * 0xffffu, kB = (tB ^ ONE2) * 0xffffu;    // 0xffff where g wins
    cur.axA = (g.axA & kA) | (cur.axA & ~kA); cur.axB = (g.axB & kB) | (cur.axB & ~kB);
    cur.ayA = (g.ayA & kA) | (cur.ayA & ~kA); cur.ayB = (g.ayB & kB) | (cur.ayB & ~kB);
    cur.fxA = (g.fxA & kA) | (cur.fxA & ~kA); cur.fxB = (g.fxB & kB) | (cur.fxB & ~kB);
    cur.fyA = (g.fyA & kA) | (cur.fyA & ~kA); cur.fyB = (g.fyB & kB) | (cur.fyB & ~kB);
    mA = (gA & kA) | (mA & ~kA);
    mB = (gB & kB) | (mB & ~kB);
}

// ------------------------------------------------------------------ non-maximum suppression (A.4)
// A magnitude row as seen by a lane: A = (m0,m1), B = (m2,m3) plus the shifted pairs
// Cm = (m-1,m0), Co = (m1,m2), Cp = (m3,m4) once the neighbour lanes' magnitudes are known.
struct MagRow { uint32_t A, B, Cm, Co, Cp; };

I2S_HD MagRow mag_row(uint32_t A, uint32_t B, uint32_t leftB, uint32_t rightA)
{
    MagRow r;
    r.A = A; r.B = B;
    r.Cm = pair_mid(leftB, A);                 // (left lane's m3, m0)
    r.Co = pair_mid(A, B);
    r.Cp = pair_mid(B, rightA);                // (m3, right lane's m0)
    return r;
}

// 0xffff in every half where v != 0
I2S_HD uint32_t nz_mask(uint32_t v) { return min2(v, ONE2) * 0xffffu; }

// Fail value of "m > a && m >= b" per half: zero where the test passes.
I2S_HD uint32_t fail_gt_ge(uint32_t m, uint32_t a, uint32_t b) { return max2(max2(a + ONE2, b), m) ^ m; }
// Fail value of "m > a && m > b"
I2S_HD uint32_t fail_gt_gt(uint32_t m, uint32_t a, uint32_t b) { return max2(max2(a, b) + ONE2, m) ^ m; }

// Sector of the gradient direction per half, from |dx| = ax and |dy| = ay (both <= 1020):
//   horizontal  <=>  (ay << 15) < ax * 13573                <=>  ay <= hp,  hp = floor(ax*13573 / 2^15)
//   vertical    <=>  (ay << 15) > ax * 13573 + (ax << 16)   <=>  ay > 2 ax + hp
// (13573 = 53*256 + 5, so hp = (ax*53 + ((ax*5) >> 8)) >> 7 stays inside 16 bits; the two forms of
// the horizontal test differ only when ax*13573 is a multiple of 2^15, i.e. ax = 0, where ay <= 0
// means magnitude 0 and the pixel is no candidate anyway.)
// Returns masks (0xffff per half): nh = NOT horizontal, v = vertical.
I2S_HD void sector_masks(uint32_t ax, uint32_t ay, uint32_t &nh, uint32_t &v)
{
    // horizontal  <=>  ay <= floor(wq / 128), wq = ax*53 + ((ax*5) >> 8)  <=>  128 * min(ay, 511) <= wq
    // (wq <= 54080, so an ay above 511 can never be horizontal and the product stays inside 16 bits)
    const uint32_t wq = ax * 53u + prmt(ax * 5u, 0, 0x4341);
    const uint32_t a128 = min2(ay, 0x01ff01ffu) * 128u;
    nh = nz_mask(a128 - min2(a128, wq));                       // 128 ay > wq
    // vertical  <=>  ay > 2 ax + floor(wq / 128)  <=>  128 (ay - 2 ax) > wq  (ay - 2ax clamped to 0..511)
    const uint32_t two = ax + ax;
    const uint32_t z128 = min2(ay - min2(ay, two), 0x01ff01ffu) * 128u;
    v = nz_mask(z128 - min2(z128, wq | 0x007f007fu));          // 128 z > wq  <=>  128 z > (wq | 127)
}

// NMS + thresholds for the 4 pixels of a lane.  up/c/dn = magnitude rows y-1, y, y+1; g = gradient
// of row y.  low1 = (low+1) in both halves, high1 likewise (saturated to 0xffff).
// Returns the 4 state bytes: 0 none, 1 weak candidate, 3 strong candidate.
// `need_diag` (out): some pixel of this lane sits in the diagonal sector and is above `low`;
// the caller then calls nms_diag() -- kept separate so a warp can skip it when no lane needs it.
struct NmsPartial { uint32_t FA, FB, dA, dB; };   // fail values so far; diagonal-sector masks

// some pixel of this lane has a magnitude above `low` (m >= low1)
I2S_HD bool any_above(const MagRow &c, uint32_t low1)
{
    const uint32_t mx = max2(c.A, c.B);
    return min2(max2(mx, low1) ^ mx, ONE2) != ONE2;
}

I2S_HD NmsPartial nms_axis(const MagRow &up, const MagRow &c, const MagRow &dn, const Grad &g, uint32_t low1)
{
    uint32_t nhA, vA, nhB, vB;
    sector_masks(g.axA, g.ayA, nhA, vA);
    sector_masks(g.axB, g.ayB, nhB, vB);
    const uint32_t ThA = fail_gt_ge(c.A, c.Cm, c.Co), ThB = fail_gt_ge(c.B, c.Co, c.Cp);
    const uint32_t TvA = fail_gt_ge(c.A, up.A, dn.A), TvB = fail_gt_ge(c.B, up.B, dn.B);
    const uint32_t lowA = max2(c.A, low1) ^ c.A, lowB = max2(c.B, low1) ^ c.B;      // zero where m > low
    NmsPartial p;
    p.FA = (ThA & ~nhA) | (TvA & vA) | lowA;
    p.FB = (ThB & ~nhB) | (TvB & vB) | lowB;
    p.dA = nhA & ~vA;                      // diagonal sector
    p.dB = nhB & ~vB;
    return p;
}

I2S_HD bool nms_needs_diag(const NmsPartial &p, const MagRow &c, uint32_t low1)
{
    const uint32_t lowA = max2(c.A, low1) ^ c.A, lowB = max2(c.B, low1) ^ c.B;
    return ((p.dA & ~nz_mask(lowA)) | (p.dB & ~nz_mask(lowB))) != 0;
}

// diagonal sector: same sign of dx,dy -> m > UL && m > DR ; opposite sign -> m > UR && m > DL
I2S_HD void nms_diag(NmsPartial &p, const MagRow &up, const MagRow &c, const MagRow &dn, const Grad &g)
{
    const uint32_t sdA = nz_mask(g.fxA) ^ nz_mask(g.fyA), sdB = nz_mask(g.fxB) ^ nz_mask(g.fyB);   // signs differ
    const uint32_t sameA = fail_gt_gt(c.A, up.Cm, dn.Co), sameB = fail_gt_gt(c.B, up.Co, dn.Cp);
    const uint32_t oppA = fail_gt_gt(c.A, up.Co, dn.Cm), oppB = fail_gt_gt(c.B, up.Cp, dn.Co);
    p.FA |= p.dA & ((oppA & sdA) | (sameA & ~sdA));
    p.FB |= p.dB & ((oppB & sdB) | (sameB & ~sdB));
}

I2S_HD uint32_t nms_state(const NmsPartial &p, const MagRow &c, uint32_t high1)
{
    // per half: 0 where rejected, else 3 - 2 * (m <= high)
    const uint32_t nsA = min2(max2(c.A, high1) ^ c.A, ONE2), nsB = min2(max2(c.B, high1) ^ c.B, ONE2);
    const uint32_t stA = (0x00030003u - nsA - nsA) & ~nz_mask(p.FA);
    const uint32_t stB = (0x00030003u - nsB - nsB) & ~nz_mask(p.FB);
    return prmt(stA, stB, 0x6420);
}

}  // namespace roll
}  // namespace i2s
