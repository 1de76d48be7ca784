// circles.cu -- gradient-voting Hough circle detector, the 10-call stack, and circle masking.
// Reference call sites: cv.HoughCircles(.., HOUGH_GRADIENT, 1, 10, [], 100, 30, 1, 30)
// img2sgf.py:179-186; masking loop :191-198.  Arithmetic: SURVEY.md Appendix A.5, A.6
// (integer votes; float32 step vectors / distances / radius scoring with IEEE rn ops).
#include "canny.cuh"
#include "circles.cuh"
#include "preproc.cuh"
#include "sort.cuh"
#include "profile.cuh"
#include <cuda_fp16.h>

namespace i2s {

constexpr int MAX_R = 30, ACC_THR = 30, NBINS = 290;      // minRadius 1: radii 1..30 are |t| of the signed radius t != 0
constexpr int CANNY_LOW = 50, CANNY_HIGH = 100;

// ------------------------------------------------------------------ K5a: edge lists
// The edge pixels of every map are compacted once into a global list of (position, Q10 gradient
// step), bucketed by 32x32 pixel tile.  One WARP owns one bucket: two 128-bit loads per lane fetch
// its state bytes, the edge bits are squeezed into a 32-bit mask per lane, one shuffle scan gives
// every lane its slot, lane 0 reserves the bucket's slice of the map's list with a single atomic
// and publishes the directory entry.  The positions go through a per-warp shared-memory list so
// that the expensive part -- recomputing the Sobel gradient of each edge pixel from the source image
// and sx = cvRound(dx*1024/mag), sy likewise -- runs with all lanes busy.
// No block-wide barrier.  Order inside a bucket is arbitrary; votes commute.  (Staging the bucket's
// 34x34 image patch in shared memory with aligned word loads instead of the 8 byte loads per edge pixel
// was measured: 7.4 -> 10.4 ms per 1024 images, most buckets hold too few edge pixels to pay for it.)
constexpr int EB = 32;
#ifndef I2S_EL_WARPS
#define I2S_EL_WARPS 4
#endif
constexpr int EL_WARPS = I2S_EL_WARPS;

__device__ __forceinline__ uint32_t ldg_u32(const uint8_t *p) { return __ldg(reinterpret_cast<const uint32_t *>(p)); }
// sum of (unsigned bytes of a) x (signed bytes of b) + c
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ uint32_t edge_nibble(uint32_t v) { return (((v >> 1) & 0x01010101u) * 0x01020408u) >> 24; }

__global__ void __launch_bounds__(EL_WARPS * 32, 2048 / (EL_WARPS * 32)) k_edge_list(const MapSet ms, const Dims dims, const uint8_t *__restrict__ state,
                                                            int spitch, size_t sstride, uint2 *__restrict__ edges, size_t estride,
                                                            int32_t *ecount, int2 *dir, int nbx, int nby)
{
    __shared__ uint16_t s_pos[EL_WARPS][EB * EB];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int map = blockIdx.z, by = blockIdx.y, bx = blockIdx.x * EL_WARPS + warp;
    if (bx >= nbx) return;                                       // warp-uniform from here on
    const int2 wh = dims.of(map % ms.n);
    const int w = wh.x, h = wh.y;
    if (bx * EB >= w || by * EB >= h) return;                    // bucket outside this image: never read by anyone
    const uint8_t *stm = state + map * sstride;
    // state bytes of the bucket: lane -> 16 pixels of row (lane >> 1) [+16 in the second pass]
    const int x = bx * EB + 16 * (lane & 1);
    uint32_t M = 0;                                              // bit 16*pass + j: pixel (x + j, row)
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
        const int y = by * EB + 16 * pass + (lane >> 1);
        if (y < h && x < w) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(stm + (size_t)y * spitch + x));
            uint32_t m16 = edge_nibble(v.x) | (edge_nibble(v.y) << 4) | (edge_nibble(v.z) << 8) | (edge_nibble(v.w) << 12);
            if (w - x < 16) m16 &= (1u << (w - x)) - 1u;        // columns beyond the image
            M |= m16 << (16 * pass);
        }
    }
    const int c = __popc(M);
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    // One elected lane reserves the bucket's slice.  The reply is not needed before the first store, so the
    // atomic's round trip overlaps the gradient loads below instead of preceding them.  (elect.sync: behind
    // a plain `lane == 0` test ptxas aggregates the atomic across the "active" lanes and reads the reply
    // back at once.)
    int off = 0;
    uint32_t leader = 0;
    if (total == 0) {
        if (lane == 0) dir[((size_t)map * nby + by) * nbx + bx] = make_int2(0, 0);
        return;
    }
    asm volatile("{\n\t.reg .pred p;\n\telect.sync %1|p, 0xffffffff;\n\t@p atom.global.add.u32 %0, [%2], %3;\n\t}"
                 : "+r"(off), "=r"(leader) : "l"(ecount + map), "r"(total) : "memory");
    uint16_t *lst = s_pos[warp];
    {
        int p = incl - c;
        uint32_t m = M;
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            lst[p++] = (uint16_t)((((b & 16) + (lane >> 1)) << 5) | (((lane & 1) << 4) + (b & 15)));
        }
    }
    __syncwarp();
    int ipitch;
    const uint8_t *img = ms.plane(map, ipitch);
    const bool al4 = (((uintptr_t)img | (uint32_t)ipitch) & 3u) == 0;                 // warp-uniform
    auto entry = [&](int i) -> uint2 {
        const int pos = lst[i];
        const int px = bx * EB + (pos & 31), py = by * EB + (pos >> 5);
        // Sobel 3x3, replicate border (A.4)
        const uint32_t xm = (uint32_t)max(px - 1, 0), xp = (uint32_t)min(px + 1, w - 1);
        const uint32_t o0 = (uint32_t)max(py - 1, 0) * (uint32_t)ipitch, o1 = (uint32_t)py * (uint32_t)ipitch,
                       o2 = (uint32_t)min(py + 1, h - 1) * (uint32_t)ipitch;
        int dx, dy;
        if (al4 && px >= 1 && px + 1 < w) {
            // interior pixel of a 4-byte aligned plane: the three columns of a row come out of the (at most
            // two) aligned words that hold them with one funnel shift, and the Sobel sums are byte dot
            // products with signed weights -- three to six word loads and five IDP.4A instead of eight byte loads
            const uint32_t a = (uint32_t)(px - 1) & ~3u, b = (uint32_t)(px + 1) & ~3u, sh = 8u * ((uint32_t)(px - 1) & 3u);
            const bool two = sh >= 16u;                        // the three columns straddle two words
            const uint32_t r0 = __funnelshift_r(ldg_u32(img + o0 + a), two ? ldg_u32(img + o0 + b) : 0u, sh);
            const uint32_t r1 = __funnelshift_r(ldg_u32(img + o1 + a), two ? ldg_u32(img + o1 + b) : 0u, sh);
            const uint32_t r2 = __funnelshift_r(ldg_u32(img + o2 + a), two ? ldg_u32(img + o2 + b) : 0u, sh);
            dx = dp4a_us(r0, 0x000100FFu, dp4a_us(r1, 0x000200FEu, dp4a_us(r2, 0x000100FFu, 0)));   // (-1, 0, 1), (-2, 0, 2)
            dy = dp4a_us(r2, 0x00010201u, dp4a_us(r0, 0x00FFFEFFu, 0));                             // (1, 2, 1), -(1, 2, 1)
        } else {
            const int p00 = __ldg(img + o0 + xm), p01 = __ldg(img + o0 + px), p02 = __ldg(img + o0 + xp);
            const int p10 = __ldg(img + o1 + xm), p12 = __ldg(img + o1 + xp);
            const int p20 = __ldg(img + o2 + xm), p21 = __ldg(img + o2 + px), p22 = __ldg(img + o2 + xp);
            dx = (p02 + 2 * p12 + p22) - (p00 + 2 * p10 + p20);
            dy = (p20 + 2 * p21 + p22) - (p00 + 2 * p01 + p02);
        }
        int sx = 0, sy = 0;
        if (dx != 0 || dy != 0) {
            const float vx = (float)dx, vy = (float)dy;
            const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)));
            if (!(mag < 1.0f)) {
                sx = __float2int_rn(__fdiv_rn(__fmul_rn(vx, 1024.0f), mag));
                sy = __float2int_rn(__fdiv_rn(__fmul_rn(vy, 1024.0f), mag));
            }
        }
        return make_uint2(((uint32_t)py << 16) | (uint32_t)px, (uint32_t)(sx & 0xffff) | ((uint32_t)sy << 16));   // (0,0) step = no vote
    };
    // the first two rounds (a bucket holds 56 edge pixels on average) before the offset is looked at
    uint2 v0 = make_uint2(0, 0), v1 = make_uint2(0, 0);
    if (lane < total) v0 = entry(lane);
    if (lane + 32 < total) v1 = entry(lane + 32);
    off = __shfl_sync(0xffffffffu, off, leader);
    if (lane == 0) dir[((size_t)map * nby + by) * nbx + bx] = make_int2(off, total);
    uint2 *out = edges + map * estride + off;
    if (lane < total) out[lane] = v0;
    if (lane + 32 < total) out[lane + 32] = v1;
    for (int i = lane + 64; i < total; i += 32) out[i] = entry(i);
}

// ------------------------------------------------------------------ K5+K6: voting fused with peak finding
// One block owns a 128x128 tile of accumulator cells, plus the 1-cell ring the 4-neighbour test
// reads and a 2-cell guard band, as int32 in shared memory; the accumulator never exists in global
// memory.  Every edge pixel within 30 px of the ring can vote into the tile: the block walks the
// edge-list buckets that overlap that region.  Each (pixel, direction) ray is clipped in float to
// the range of radii whose cells fall inside the tile -- conservatively, at most one extra step per
// side, which the guard band absorbs -- so the inner loop is a bare shared-memory atomic per vote
// with no bounds test.  Rays are monotone in x and y, hence "cells inside the tile" is one interval
// of radii and dropping out-of-image cells equals the reference's break.  Votes are integers, so
// the result does not depend on the order of the atomics.
//
// Every warp owns one contiguous slice of the tile's item sequence (lanes interleaved), so the bucket
// pointer moves by a step or two per iteration after one binary search per warp.  (Dealing rounds of 32
// items to the warps in turn, to even out short rim rays and long interior rays, measured 3 % slower.)
//
// The vote loop.  The reference's cell for signed radius t is ((x<<10) + t*sx) >> 10 = x + floor(t*sx/1024)
// (x is an integer), so the OFFSET of the cell from the pixel depends on (t, sx, sy) only.  Both offsets
// travel in one register: U(t) = (t*sy + 2^15) << 16 | (t*sx + 2^15) = 0x80008000 + t*S with S = sy*65536 + sx;
// |t*s| <= 30*1024 keeps each half inside its 16 bits, so a single 32-bit add steps both coordinates.
// Bits 10..15 of each half are floor(t*s/1024) + 32; (U >> 8) & 0x00FC00FC holds four times those two
// numbers in its 16-bit halves, and one 16-bit x 8-bit dot product (IDP.2A) with the byte pair
// (1, pitch) turns that into the byte address of the cell.  Per vote: add, shift, and, dot, atomic -- and
// in the unrolled loop the shift and the AND are shared by two votes (vote_at2).
constexpr int AT = 128;                      // tile edge in accumulator cells
constexpr int AG = 2;                        // guard cells around the ring
constexpr int AS = AT + 2 + 2 * AG;          // shared rows / used columns
constexpr int AP = AS + 1;                   // shared pitch (odd: column walks are conflict free)
#ifndef I2S_VOTE_UNROLL
#define I2S_VOTE_UNROLL 8               // a power of two; 8 measured 1.5 % faster than 4, 16 20 % slower
#endif
constexpr int VOTE_THREADS = 512;            // 384 alike, 640 / 672 slower
constexpr int VOTE_SMEM = AS * AP * 4;
constexpr int VB = 7;                        // buckets per axis that can overlap a tile's region
static_assert(AP < 256, "the pitch is a byte operand of the address dot product");
static_assert((I2S_VOTE_UNROLL & (I2S_VOTE_UNROLL - 1)) == 0 && I2S_VOTE_UNROLL >= 2, "the remainder is peeled in powers of two");

// 1/v for |v| >= 1 (a non-zero Q10 step): the bare MUFU.RCP, 1 ulp -- the clip interval has 0.25 of slack
__device__ __forceinline__ float rcp_approx(float v)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ void vote_at(uint32_t a0, uint32_t U)
{
    const uint32_t m = (U >> 8) & 0x00FC00FCu;
    const uint32_t addr = __dp2a_lo(m, 1u | ((uint32_t)AP << 8), a0);
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}

// Two votes at once: one byte permute gathers the high bytes of the four halves of (U, V) -- (x, y) of the
// first vote, (x, y) of the second --, one AND clears the two fraction bits under each, and a 4-way byte dot
// product with (1, pitch, 0, 0) resp. (0, 0, 1, pitch) gives each address: 4 instructions per vote with the add.
__device__ __forceinline__ void vote_at2(uint32_t a0, uint32_t U, uint32_t V, uint32_t W0, uint32_t W1)
{
    const uint32_t m = __byte_perm(U, V, 0x7531) & 0xFCFCFCFCu;
    const uint32_t addr0 = __dp4a(m, W0, a0), addr1 = __dp4a(m, W1, a0);
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr0) : "memory");
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr1) : "memory");
}

__device__ __forceinline__ void vote_item(int *s_acc, uint32_t s_base, uint2 e, int cx0, int cy0, int X0, int X1, int Y0, int Y1,
                                          uint32_t W0, uint32_t W1)
{
    const int x = e.x & 0xffff, y = e.x >> 16;
    const int sx = (int)(short)(e.y & 0xffff), sy = (int)e.y >> 16;
    float lo = -(float)MAX_R, hi = (float)MAX_R;
    // (small integers: the float differences are exact, one conversion per axis instead of two)
    if (sx != 0) {
        const float inv = 1024.0f * rcp_approx((float)sx), fx = (float)x;
        const float ta = ((float)X0 - fx) * inv, tb = ((float)(X1 + 1) - fx) * inv;
        lo = fmaxf(lo, fminf(ta, tb)); hi = fminf(hi, fmaxf(ta, tb));
    } else if (x < X0 || x > X1) return;
    if (sy != 0) {
        const float inv = 1024.0f * rcp_approx((float)sy), fy = (float)y;
        const float ta = ((float)Y0 - fy) * inv, tb = ((float)(Y1 + 1) - fy) * inv;
        lo = fmaxf(lo, fminf(ta, tb)); hi = fminf(hi, fmaxf(ta, tb));
    } else if (y < Y0 || y > Y1) return;
    // 0.25 of slack covers the error of the approximate divide; at most one extra step per side
    const int t_lo = max(-MAX_R, (int)floorf(lo - 0.25f)), t_hi = min(MAX_R, (int)ceilf(hi + 0.25f));
    if (t_lo > t_hi) return;
    // Both rays are one arithmetic sequence in the signed radius t (forward t > 0, backward t < 0, r = |t|).
    // One loop over t_lo..t_hi votes them all -- a warp then runs max(len) instead of max(forward) +
    // max(backward) -- and the vote the loop casts at t = 0 (the pixel's own cell, which the reference
    // never votes) is taken back afterwards.
    const int S = sy * 65536 + sx;
    const uint32_t a0 = s_base + 4u * (uint32_t)((y - cy0 - 32) * AP + (x - cx0 - 32));
    uint32_t bias = 0x80008000u;
    asm volatile("" : "+r"(bias));          // opaque: keeps the bias inside the running value instead of one add per vote
    uint32_t U = bias + (uint32_t)t_lo * (uint32_t)S;
    constexpr int UN = I2S_VOTE_UNROLL;
    const int nvotes = t_hi - t_lo + 1;
    for (int i = nvotes / UN; i > 0; i--, U += (uint32_t)UN * (uint32_t)S) {
#pragma unroll
        for (int k = 0; k + 1 < UN; k += 2) vote_at2(a0, U + (uint32_t)k * (uint32_t)S, U + (uint32_t)(k + 1) * (uint32_t)S, W0, W1);
        if (UN & 1) vote_at(a0, U + (uint32_t)(UN - 1) * (uint32_t)S);
    }
    int rem = nvotes % UN;                  // the remainder without a loop: pairs, then a single vote
#pragma unroll
    for (int k = UN / 2; k >= 2; k >>= 1)
        if (rem & k) {
#pragma unroll
            for (int j = 0; j < k; j += 2) vote_at2(a0, U + (uint32_t)j * (uint32_t)S, U + (uint32_t)(j + 1) * (uint32_t)S, W0, W1);
            U += (uint32_t)k * (uint32_t)S;
        }
    if (rem & 1) vote_at(a0, U);
    if (t_lo <= 0 && t_hi >= 0)             // the pixel's own cell: offset (32, 32) from a0
        asm volatile("red.shared.add.u32 [%0], 0xffffffff;" ::"r"(a0 + 4u * (uint32_t)(32 * AP + 32)) : "memory");
}

#ifndef I2S_VOTE_MINB
#define I2S_VOTE_MINB 3               // three blocks per SM is what the 72 KB accumulator allows: up to 40 registers
#endif
__global__ void __launch_bounds__(VOTE_THREADS, I2S_VOTE_MINB) k_vote_peaks(const uint2 *__restrict__ edges, size_t estride,
                                                             const int2 *__restrict__ dir, int nbx, int nby, const Dims dims,
                                                             int n_images, int32_t *cand, int32_t *ncand, int cand_cap, const uint2 vw)
{
    // vw: the byte weights (1, pitch) of the address dot products for the first / second vote of a pair.  A kernel
    // parameter so that they are constant-bank operands of the instruction, not immediates rebuilt inside the loop.
    extern __shared__ __align__(16) unsigned char s_raw[];
    int *s_acc = reinterpret_cast<int *>(s_raw);                       // AS x AP
    __shared__ int s_boff[VB * VB], s_bend[VB * VB + 1];               // bucket slice start / running item end
    constexpr int NW = VOTE_THREADS / 32;
    const int map = blockIdx.z;
    const int2 wh = dims.of(map % n_images);
    const int w = wh.x, h = wh.y;
    const int tx0 = blockIdx.x * AT, ty0 = blockIdx.y * AT;
    if (tx0 >= w || ty0 >= h) return;                                  // tile outside this image (ragged batch)
    const uint2 *elist = edges + map * estride;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cx0 = tx0 - 1 - AG, cy0 = ty0 - 1 - AG;
    const int X0 = max(tx0 - 1, 0), X1 = min(tx0 + AT, w - 1);
    const int Y0 = max(ty0 - 1, 0), Y1 = min(ty0 + AT, h - 1);
    const int rx0 = max(tx0 - 1 - MAX_R, 0), rx1 = min(tx0 + AT + MAX_R, w - 1);
    const int ry0 = max(ty0 - 1 - MAX_R, 0), ry1 = min(ty0 + AT + MAX_R, h - 1);
    const int bx0 = rx0 / EB, bx1 = rx1 / EB, by0 = ry0 / EB, by1 = ry1 / EB;
    const int nbw = bx1 - bx0 + 1, nb = nbw * (by1 - by0 + 1);      // <= VB*VB
    {
        uint4 *z = reinterpret_cast<uint4 *>(s_acc);                   // 16-byte stores; the tail word by word
        for (int i = threadIdx.x; i < AS * AP / 4; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
        for (int i = (AS * AP / 4) * 4 + threadIdx.x; i < AS * AP; i += blockDim.x) s_acc[i] = 0;
    }
    if (threadIdx.x < nb) {
        const int b = threadIdx.x;
        const int2 d = dir[((size_t)map * nby + by0 + b / nbw) * nbx + bx0 + b % nbw];
        s_boff[b] = d.x;
        s_bend[b + 1] = d.y;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        s_bend[0] = 0;
        for (int b = 0; b < nb; b++) { run += s_bend[b + 1]; s_bend[b + 1] = run; }
    }
    __syncthreads();
    const int items = s_bend[nb];
    {
        const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(s_acc);
        const int per = ((items + NW - 1) / NW + 31) & ~31;          // items per warp, whole rounds of 32
        const int i0 = warp * per, i1 = min(i0 + per, items), istep = 32;
        int b = 0;
        if (i0 < i1) {                                               // bucket holding item i0: s_bend[b] <= i0 < s_bend[b+1]
            int lo = 0, hi = nb - 1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s_bend[mid + 1] <= i0) lo = mid + 1; else hi = mid;
            }
            b = lo;
        }
        for (int it = i0 + lane; it < i1; it += istep) {
            while (it >= s_bend[b + 1]) b++;                         // `it` only grows: b is monotone
            const uint2 e = __ldg(elist + s_boff[b] + (it - s_bend[b]));
            if (e.y == 0) continue;                                  // no gradient direction: no votes
            // (pixels of the walked buckets that lie beyond the reach of the tile need no test of their own:
            // their clipped range of radii comes out empty)
            vote_item(s_acc, s_base, e, cx0, cy0, X0, X1, Y0, Y1, vw.x, vw.y);
        }
    }
    __syncthreads();
    // Cells outside the image never receive votes in the reference.  The conservative extra step can
    // leave a vote at most one cell outside the clip box; of those cells the peak test below only ever
    // reads column w and row h (right / lower neighbours of the last column / row): clear them.
    if (w - cx0 < AS || h - cy0 < AS) {
        if (w - cx0 < AS)
            for (int ly = threadIdx.x; ly < AS; ly += blockDim.x) s_acc[ly * AP + (w - cx0)] = 0;
        if (h - cy0 < AS)
            for (int lx = threadIdx.x; lx < AS; lx += blockDim.x) s_acc[(h - cy0) * AP + lx] = 0;
        __syncthreads();
    }
    // K6: 4-neighbour peaks above the accumulator threshold, interior cells only (x,y >= 1)
    const int aw = w + 2;
    for (int ty = warp; ty < AT; ty += NW) {
        const int cy = ty0 + ty;
        if (cy < 1 || cy >= h) continue;                             // warp-uniform
        const int *row = s_acc + (ty + 1 + AG) * AP + 1 + AG;
#pragma unroll
        for (int j = 0; j < AT / 32; j++) {
            const int tx = lane + 32 * j, cx = tx0 + tx;
            const int v = row[tx];
            if (v > ACC_THR && cx >= 1 && cx < w) {
                const int *c = row + tx;
                if (v > c[-1] && v >= c[1] && v > c[-AP] && v >= c[AP]) {
                    int slot = atomicAdd(ncand + map, 1);
                    if (slot < cand_cap) cand[(size_t)map * cand_cap + slot] = cy * aw + cx;
                }
            }
        }
    }
}

// ------------------------------------------------------------------ K7a: radius estimation
// One warp per candidate centre.  Only pixels within 30 px can contribute, so the warp scans
// the edge-list buckets around the centre instead of the whole non-zero list.
__device__ __forceinline__ float radius_of_q(int q)
{
    // (upbin + j)/2.f / nBinsPerDr * dr + minRadius, every step rounded to float32
    return __fadd_rn(__fdiv_rn(__fdiv_rn((float)q, 2.0f), 10.0f), 1.0f);
}

#ifndef I2S_RADIUS_GRIDX
#define I2S_RADIUS_GRIDX 16
#endif
constexpr int RW = 8;          // warps per block
#ifndef I2S_RADIUS_BATCH
#define I2S_RADIUS_BATCH 4
#endif
constexpr int RB = I2S_RADIUS_BATCH;   // candidates per warp and round
constexpr int RBINS = 320;     // NBINS padded to a multiple of 32 (pad stays zero)
constexpr int RQ = 576;        // radius table size: q = upbin + j <= 289 + 279

struct RadiusTables { float rtab[RQ]; uint16_t binlut[900]; };

// Centres sit on half-integers and edge pixels on integers, so the float32 squared distance
// (cx+.5-px)^2 + (cy+.5-py)^2 is exactly q + 0.5 with q = dx(dx+1) + dy(dy+1) an integer: the
// sqrt / rint chain of SURVEY A.5 step 4 is tabulated over the 899 admissible q, once per call
// (it used to be recomputed by every block: a fifth of the radius kernel's instructions).
__global__ void __launch_bounds__(256) k_radius_tables(RadiusTables *t)
{
    for (int q = threadIdx.x; q < RQ; q += blockDim.x) t->rtab[q] = radius_of_q(q);
    for (int q = threadIdx.x; q < 900; q += blockDim.x) {
        const float dd = __fsqrt_rn((float)q + 0.5f);
        const int bin = __float2int_rn(__fmul_rn(__fsub_rn(dd, 1.0f), 10.0f));
        t->binlut[q] = (uint16_t)min(max(bin, 0), NBINS - 1);
    }
}

__global__ void __launch_bounds__(RW * 32, 5) k_radius(const RadiusTables *__restrict__ tables, const uint2 *__restrict__ edges, size_t estride,
                                                   const int2 *__restrict__ dir, int nbx, int nby, const Dims dims,
                                                   int n_images, const int32_t *__restrict__ cand,
                                                   const int32_t *__restrict__ ncand, int cand_cap, unsigned long long *est,
                                                   int32_t *nest, int32_t *status)
{
    __shared__ __align__(16) int s_bins[RW][RB][RBINS];   // histograms, then (prefix sum << 16) | (1 + highest non-empty bin at or below)
    __shared__ float s_rtab[RQ];
    __shared__ __align__(4) uint16_t s_binlut[900];   // histogram bin of squared distance q + 0.5, q = 1..899
    const int map = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int n = ncand[map];
    if (n > cand_cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(status + map % n_images, I2S_ST_CAND_OVERFLOW);
        n = cand_cap;
    }
    if (blockIdx.x * RW * RB >= n) return;
    const int2 wh = dims.of(map % n_images);
    const int w = wh.x, h = wh.y;
    const uint2 *elist = edges + map * estride;
    const int2 *mdir = dir + (size_t)map * nbx * nby;
    const int aw = w + 2;
    // the radius and distance-bin tables (k_radius_tables) from the workspace into shared memory
    for (int q = threadIdx.x; q < RQ; q += blockDim.x) s_rtab[q] = __ldg(tables->rtab + q);
    for (int q = threadIdx.x; q < 900 / 2; q += blockDim.x)
        reinterpret_cast<uint32_t *>(s_binlut)[q] = __ldg(reinterpret_cast<const uint32_t *>(tables->binlut) + q);
    __syncthreads();
    // A warp takes RB consecutive candidates per round: their histograms are filled one after the other
    // by all lanes, then the window scans -- a sequential walk per candidate -- run side by side, one
    // candidate per lane, instead of one after the other with 32 lanes computing the same thing.
    for (int c0 = (blockIdx.x * RW + warp) * RB; c0 < n; c0 += gridDim.x * RW * RB) {
        const int nc = min(RB, n - c0);
        const int mybase = lane < nc ? cand[(size_t)map * cand_cap + c0 + lane] : 0;
        {
            uint4 *z = reinterpret_cast<uint4 *>(&s_bins[warp][0][0]);
            for (int i = lane; i < RB * RBINS / 4; i += 32) z[i] = make_uint4(0, 0, 0, 0);
        }
        __syncwarp();
        for (int k = 0; k < nc; k++) {
            int *bins = s_bins[warp][k];
            const int base = __shfl_sync(0xffffffffu, mybase, k);
            const int cy = base / aw, cx = base - cy * aw;
            // histogram of the distances to the edge pixels within 30 px: walk the edge-list buckets
            // that overlap the 60x60 window (at most 3x3 of them).  (A finer 16x16 directory was tried:
            // fewer entries to reject, but cells of ~18 entries leave half a warp idle -- slower.)
            // Both differences travel in one register: (centre + 0x8000) - entry keeps the low half from
            // borrowing, the xor turns it back into a signed 16-bit dx next to dy (|d| <= 92 inside 3x3
            // buckets), and q = dx^2 + dy^2 + dx + dy is two 16-bit x 8-bit dot products.
            const int xlo = max(cx - 29, 0), xhi = min(cx + 30, w - 1);
            const int ylo = max(cy - 29, 0), yhi = min(cy + 30, h - 1);
            const uint32_t cpk = (((uint32_t)cy << 16) | (uint32_t)cx) + 0x8000u;
            auto tally = [&](uint32_t e) {
                const int a = (int)((cpk - e) ^ 0x8000u);                        // (dy << 16) | (dx & 0xffff)
                const int b = (int)__byte_perm((uint32_t)a, 0u, 0x4420);         // bytes (dx, dy)
                const int q = __dp2a_lo(a, 0x0101, __dp2a_lo(a, b, 0));          // 1 <= r2 <= 900  <=>  1 <= q <= 899
                if ((unsigned)(q - 1) < 899u) atomicAdd(bins + s_binlut[q], 1);
            };
            // The walk is bound by the latency of its loads, so they are issued in bulk: the (at most nine)
            // directory entries by nine lanes at once, then per bucket row the first 64 entries of its three
            // buckets -- six loads per lane in flight -- before any of them is tallied.  A lane without an
            // entry tallies the centre itself (q = 0: rejected), so there is no branch around the tally.
            const int bx0 = xlo / EB, by0 = ylo / EB, nbc = xhi / EB - bx0 + 1, nbr = yhi / EB - by0 + 1;   // <= 3 each
            int2 dl = make_int2(0, 0);
            if (lane < 9) {
                const int r = lane / 3, c = lane - 3 * r;
                if (r < nbr && c < nbc) dl = __ldg(mdir + (by0 + r) * nbx + bx0 + c);
            }
            const uint32_t none = cpk - 0x8000u;
            for (int r = 0; r < nbr; r++) {
                int off[3], cnt[3];
                uint32_t e[6];
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    off[c] = __shfl_sync(0xffffffffu, dl.x, 3 * r + c);
                    cnt[c] = __shfl_sync(0xffffffffu, dl.y, 3 * r + c);
                }
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    e[2 * c] = lane < cnt[c] ? __ldg(&elist[off[c] + lane].x) : none;
                    e[2 * c + 1] = lane + 32 < cnt[c] ? __ldg(&elist[off[c] + lane + 32].x) : none;
                }
#pragma unroll
                for (int i = 0; i < 6; i++) tally(e[i]);
#pragma unroll
                for (int c = 0; c < 3; c++)
                    for (int i = lane + 64; i < cnt[c]; i += 32) tally(__ldg(&elist[off[c] + i].x));
            }
        }
        __syncwarp();
        // in place: inclusive prefix sums (10 bins per lane; at most 2827 pixels lie within 30 px) and, per
        // bin, one plus the highest non-empty bin at or below it (0: none)
        for (int k = 0; k < nc; k++) {
            int *bins = s_bins[warp][k];
            int loc[10], sum = 0, top = 0;
#pragma unroll
            for (int i = 0; i < 10; i++) { loc[i] = bins[lane * 10 + i]; sum += loc[i]; if (loc[i]) top = lane * 10 + i + 1; }
            int incl = sum, below = top;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o), u = __shfl_up_sync(0xffffffffu, below, o);
                if (lane >= o) { incl += t; below = max(below, u); }
            }
            below = __shfl_up_sync(0xffffffffu, below, 1);                   // over the lanes before this one
            if (lane == 0) below = 0;
            int run = incl - sum;
#pragma unroll
            for (int i = 0; i < 10; i++) {
                run += loc[i];
                if (loc[i]) below = lane * 10 + i + 1;
                bins[lane * 10 + i] = (run << 16) | below;
            }
        }
        __syncwarp();
        // OpenCV's scan from the top bin: every non-zero bin opens a 10-bin window, the bin just below
        // the window is skipped.  Lane k walks candidate k.
        if (lane < nc) {
            const uint32_t *P = reinterpret_cast<const uint32_t *>(s_bins[warp][lane]);
            int maxCount = 0, bestq = 0;
            float rBest = 0.0f;
            int j = NBINS - 1;
            while (j > 0) {
                const int up = (int)(P[j] & 0xffffu) - 1;
                if (up <= 0) break;
                int jn = up - 10, cur = (int)(P[up] >> 16);
                if (jn >= 0) cur -= (int)(P[jn] >> 16);
                else jn = -1;
                float rCur = s_rtab[up + jn];
                if ((__fmul_rn((float)cur, rBest) >= __fmul_rn((float)maxCount, rCur)) ||
                    (rBest < 1.1920929e-07f && cur >= maxCount)) {
                    rBest = rCur; maxCount = cur; bestq = up + jn;
                }
                j = jn - 1;
            }
            if (maxCount > ACC_THR) {
                const int cy = mybase / aw, cx = mybase - cy * aw;
                int slot = atomicAdd(nest + map, 1);
                if (slot < cand_cap)
                    est[(size_t)map * cand_cap + slot] = ((unsigned long long)(4095 - maxCount) << 38) |
                                                         ((unsigned long long)(1023 - bestq) << 28) |
                                                         ((unsigned long long)cx << 14) | (unsigned long long)cy;
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ K7b: total-order sort + greedy minDist
// One block per map: bitonic sort of the packed keys (support desc, radius desc, x asc, y asc), then
// OpenCV's sequential suppression -- a circle is kept iff it is >= 10 px from every circle kept
// before it -- evaluated in parallel.  All candidates are hashed into a grid of cells at least 16 px
// wide; every round each undecided candidate looks at the EARLIER candidates within 10 px (chains of
// its 3x3 cell neighbourhood): one of them kept => rejected; all of them rejected => kept; otherwise
// it waits.  The earliest undecided candidate is always decided, so the loop ends, and every decision
// is the sequential algorithm's by induction over the sorted order -- a handful of rounds instead of
// one dependent step per candidate.  The kept circles are compacted in sorted order by a block scan.
//
// Working arrays (16 bytes per candidate) live in shared memory up to 8192 candidates per map and in
// the caller's workspace above that, so any cand_cap the limits accept runs.
struct FinishBufs {
    unsigned long long *keys;        // np2cap
    short2 *xy;                      // np2cap
    uint16_t *next;                  // np2cap
    uint8_t *state;                  // 2 x np2cap   0 undecided, 1 kept, 2 rejected (two copies, see the rounds)
};
constexpr int FINISH_SMEM_CAP = 8192;
constexpr int FINISH_THREADS = 1024;         // a map with thousands of candidates (the edge-map input) sets the launch's critical path
constexpr int FINISH_MAX_CELLS = 4096;
__host__ __device__ static inline size_t finish_bytes_per_map(int np2cap) { return (size_t)np2cap * 16; }

__global__ void __launch_bounds__(FINISH_THREADS) k_circles_finish(const unsigned long long *__restrict__ est,
                                                        const int32_t *__restrict__ nest, int cand_cap, int np2cap,
                                                        int cshift, int cells_x, int cells_y, float *circ,
                                                        int32_t *ncirc, int circle_cap, int32_t *status, int n_images,
                                                        unsigned char *gbuf)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ int s_warp_tot[FINISH_THREADS / 32];
    const int map = blockIdx.x;
    int *head = reinterpret_cast<int *>(s_raw);                                            // cells_x * cells_y
    unsigned char *arr = gbuf ? gbuf + (size_t)map * finish_bytes_per_map(np2cap)
                              : s_raw + (((size_t)cells_x * cells_y * 4 + 15) & ~(size_t)15);
    FinishBufs B;
    B.keys = reinterpret_cast<unsigned long long *>(arr);
    B.xy = reinterpret_cast<short2 *>(B.keys + np2cap);
    B.next = reinterpret_cast<uint16_t *>(B.xy + np2cap);
    B.state = reinterpret_cast<uint8_t *>(B.next + np2cap);
    const int n = min(nest[map], cand_cap);
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = threadIdx.x; i < np2; i += blockDim.x) B.keys[i] = i < n ? est[(size_t)map * cand_cap + i] : ~0ull;
    for (int i = threadIdx.x; i < cells_x * cells_y; i += blockDim.x) head[i] = -1;
    __syncthreads();
    bitonic_sort_block(B.keys, np2);
    // hash every candidate into its cell (chain order is irrelevant: the index decides priority)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long k = B.keys[i];
        const int x = (int)((k >> 14) & 0x3fff), y = (int)(k & 0x3fff);
        B.xy[i] = make_short2((short)x, (short)y);
        B.state[i] = 0;
        const int prev = atomicExch(&head[(y >> cshift) * cells_x + (x >> cshift)], i);
        B.next[i] = (uint16_t)(prev < 0 ? 0xffff : prev);
    }
    // Rounds on two copies of the state: decisions are taken from the previous round's copy and written to the
    // other one, so no thread ever reads a byte another thread may be writing.
    uint8_t *cur = B.state, *nxt = B.state + np2cap;
    __syncthreads();
    while (true) {
        bool waiting = false;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            uint8_t s = cur[i];
            if (!s) {
                const short2 p = B.xy[i];
                const int ccx = p.x >> cshift, ccy = p.y >> cshift;
                bool any_kept = false, all_rejected = true;
                for (int dy = -1; dy <= 1; dy++) {
                    const int ny = ccy + dy;
                    if (ny < 0 || ny >= cells_y) continue;
                    for (int dx = -1; dx <= 1; dx++) {
                        const int nx = ccx + dx;
                        if (nx < 0 || nx >= cells_x) continue;
                        for (int j = head[ny * cells_x + nx]; j >= 0; j = (B.next[j] == 0xffff) ? -1 : (int)B.next[j]) {
                            if (j >= i) continue;
                            const short2 q = B.xy[j];
                            const int ex = q.x - p.x, ey = q.y - p.y;
                            if (ex * ex + ey * ey >= 100) continue;
                            const int sj = cur[j];
                            any_kept |= sj == 1;
                            all_rejected &= sj == 2;
                        }
                    }
                }
                s = any_kept ? 2 : (all_rejected ? 1 : 0);
                waiting |= s == 0;
            }
            nxt[i] = s;
        }
        uint8_t *t = cur; cur = nxt; nxt = t;
        if (!__syncthreads_or(waiting)) break;
    }
    B.state = cur;
    // kept circles in sorted order: contiguous slice per thread, block scan of the kept counts
    const int per = (n + (int)blockDim.x - 1) / (int)blockDim.x;
    const int i0 = min((int)threadIdx.x * per, n), i1 = min(i0 + per, n);
    int mine = 0;
    for (int i = i0; i < i1; i++) mine += B.state[i] == 1;
    int incl = mine;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    int before = incl - mine, nk = 0;
    for (int k = 0; k < FINISH_THREADS / 32; k++) {
        if (k < warp) before += s_warp_tot[k];
        nk += s_warp_tot[k];
    }
    float *out = circ + (size_t)map * circle_cap * 3;
    for (int i = i0; i < i1; i++) {
        if (B.state[i] != 1) continue;
        if (before < circle_cap) {
            const unsigned long long k = B.keys[i];
            out[3 * before] = (float)B.xy[i].x + 0.5f;
            out[3 * before + 1] = (float)B.xy[i].y + 0.5f;
            out[3 * before + 2] = radius_of_q(1023 - (int)((k >> 28) & 0x3ff));
        }
        before++;
    }
    if (threadIdx.x == 0) {
        ncirc[map] = nk;
        if (nk > circle_cap) atomicOr(status + map % n_images, I2S_ST_CIRCLE_OVERFLOW);
    }
}

// ------------------------------------------------------------------ stacking in `blurs` order
// blurs = [grey, edges, median1 (=grey), gauss1 (=grey), median3, gauss3, median5, gauss5,
// median7, gauss7] (img2sgf.py:171-175); internal map order is [grey, edges, med3, gau3, ...].
__constant__ int c_call_map[I2S_N_CALLS] = {0, 1, 0, 0, 2, 3, 4, 5, 6, 7};

__global__ void __launch_bounds__(256) k_stack(const float *__restrict__ mcirc, const int32_t *__restrict__ mcount,
                                               int n, int circle_cap, float *out, int32_t *counts, int32_t *status,
                                               int2 *dup)
{
    const int img = blockIdx.x;
    int off = 0;
    float *o = out + (size_t)img * circle_cap * 3;
    for (int c = 0; c < I2S_N_CALLS; c++) {
        int map = c_call_map[c] * n + img;
        int cnt = min(mcount[map], circle_cap);
        const float *src = mcirc + (size_t)map * circle_cap * 3;
        for (int i = threadIdx.x; i < cnt * 3; i += blockDim.x)
            if (off * 3 + i < circle_cap * 3) o[off * 3 + i] = src[i];
        off += cnt;
    }
    if (threadIdx.x == 0) {
        counts[img] = off;
        if (off > circle_cap) atomicOr(status + img, I2S_ST_CIRCLE_OVERFLOW);
        // calls 0, 2 and 3 are the same list (grey, median 1, Gaussian 1): k_mask may ignore the first
        // two copies, the third one covers the same pixels later in the sequence
        const int n0 = min(mcount[img], circle_cap), n1 = min(mcount[n + img], circle_cap);
        dup[img] = (off <= circle_cap) ? make_int2(n0, n1) : make_int2(0, 0);
    }
}

// ------------------------------------------------------------------ K8: masking
// Sequential semantics of the reference loop reduce to: a pixel covered by any rectangle takes
// the value decided by the LAST covering circle i* (255 on that circle's 5-px plus, else 0),
// because every circle's plus lies inside its own rectangle.
constexpr int KT = 64, KCHUNK = 1024;

__device__ __forceinline__ void circle_rect(const float *c, int &x0, int &y0, int &x1, int &y1)
{
    float r = __fadd_rn(c[2], 2.0f);
    x0 = __float2int_rn(__fsub_rn(c[0], r)); y0 = __float2int_rn(__fsub_rn(c[1], r));
    x1 = __float2int_rn(__fadd_rn(c[0], r)); y1 = __float2int_rn(__fadd_rn(c[1], r));
}

// a tile-relative coordinate, clamped to the ring around the tile, as a half2 pair of equal halves
__device__ __forceinline__ uint32_t tile_half2(int v)
{
    const __half2 p = __half2half2(__int2half_rn(min(max(v, -1), KT)));
    return *reinterpret_cast<const uint32_t *>(&p);
}
__device__ __forceinline__ __half2 as_half2(uint32_t v) { return *reinterpret_cast<const __half2 *>(&v); }

__global__ void __launch_bounds__(256) k_mask(const uint8_t *__restrict__ edges, uint8_t *__restrict__ masked,
                                              const Dims dims, int pitch, size_t stride, const float *__restrict__ circles,
                                              const int32_t *__restrict__ counts, int circle_cap,
                                              const int2 *__restrict__ dup)
{
    __shared__ uint4 s_rect[KCHUNK];                               // (x0, x1, y0, y1), each a half2 pair of one number
    __shared__ uint32_t s_idx[KCHUNK];                             // (i + 1) in both halves
    __shared__ int s_n;
    const int img = blockIdx.z;
    const int2 wh = dims.of(img);
    const int w = wh.x, h = wh.y;
    const int tx0 = blockIdx.x * KT, ty0 = blockIdx.y * KT;
    if (tx0 >= w || ty0 >= h) return;                              // tile outside this image (ragged batch)
    const float *circ = circles + (size_t)img * circle_cap * 3;
    const int n = min(counts[img], circle_cap);
    // circles [0, n0) and [n0 + n1, 2 n0 + n1) are repeated verbatim at [2 n0 + n1, 3 n0 + n1) (see k_stack)
    const int2 dd = dup ? dup[img] : make_int2(0, 0);
    const int skip_a = dd.x, skip_b0 = dd.x + dd.y, skip_b1 = 2 * dd.x + dd.y;
    const int tx1 = min(tx0 + KT, w) - 1, ty1 = min(ty0 + KT, h) - 1;
    // this thread's pixels: rows (threadIdx.x/16) + 16*q, q<4 ; cols (threadIdx.x%16)*4 + k, k<4.
    // last[2q + k/2]: 1 + index of the last covering circle of pixel (q, k) in 16 bits (0 = none).  Tile
    // coordinates (-1..64) are exact in half precision: the inside tests of two pixels are one HSET2 each.
    const int lx = (threadIdx.x & 15) * 4, ly = threadIdx.x >> 4;
    const __half2 cx01 = __floats2half2_rn((float)lx, (float)(lx + 1)), cx23 = __floats2half2_rn((float)(lx + 2), (float)(lx + 3));
    const __half2 cy01 = __floats2half2_rn((float)ly, (float)(ly + 16)), cy23 = __floats2half2_rn((float)(ly + 32), (float)(ly + 48));
    uint32_t last[8];
#pragma unroll
    for (int k = 0; k < 8; k++) last[k] = 0;
    for (int c0 = 0; c0 < n; c0 += KCHUNK) {
        __syncthreads();
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        for (int i = c0 + threadIdx.x; i < min(n, c0 + KCHUNK); i += blockDim.x) {
            if (i < skip_a || (i >= skip_b0 && i < skip_b1)) continue;
            int x0, y0, x1, y1;
            circle_rect(circ + 3 * i, x0, y0, x1, y1);
            if (x1 >= tx0 && x0 <= tx1 && y1 >= ty0 && y0 <= ty1) {
                int s = atomicAdd(&s_n, 1);
                s_rect[s] = make_uint4(tile_half2(x0 - tx0), tile_half2(x1 - tx0), tile_half2(y0 - ty0), tile_half2(y1 - ty0));
                s_idx[s] = (uint32_t)(i + 1) * 0x10001u;
            }
        }
        __syncthreads();
        const int m = s_n;
        for (int j = 0; j < m; j++) {
            const uint4 r = s_rect[j];
            const uint32_t V = s_idx[j];
            const __half2 X0 = as_half2(r.x), X1 = as_half2(r.y), Y0 = as_half2(r.z), Y1 = as_half2(r.w);
            const uint32_t v01 = __hge2_mask(cx01, X0) & __hle2_mask(cx01, X1) & V;
            const uint32_t v23 = __hge2_mask(cx23, X0) & __hle2_mask(cx23, X1) & V;
            const uint32_t y01 = __hge2_mask(cy01, Y0) & __hle2_mask(cy01, Y1);
            const uint32_t y23 = __hge2_mask(cy23, Y0) & __hle2_mask(cy23, Y1);
            const uint32_t r0 = __byte_perm(y01, 0, 0x1010), r1 = __byte_perm(y01, 0, 0x3232);
            const uint32_t r2 = __byte_perm(y23, 0, 0x1010), r3 = __byte_perm(y23, 0, 0x3232);
            last[0] = __vmaxu2(last[0], v01 & r0); last[1] = __vmaxu2(last[1], v23 & r0);
            last[2] = __vmaxu2(last[2], v01 & r1); last[3] = __vmaxu2(last[3], v23 & r1);
            last[4] = __vmaxu2(last[4], v01 & r2); last[5] = __vmaxu2(last[5], v23 & r2);
            last[6] = __vmaxu2(last[6], v01 & r3); last[7] = __vmaxu2(last[7], v23 & r3);
        }
    }
    const bool al = ((reinterpret_cast<uintptr_t>(edges) | reinterpret_cast<uintptr_t>(masked) | (uintptr_t)pitch | (uintptr_t)stride) & 3) == 0;
    const int wlim = write_limit(w, pitch, 4);
    const int x = tx0 + lx;
    if (x >= w) return;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        int y = ty0 + ly + 16 * q;
        if (y >= h) continue;
        const size_t o = img * stride + (size_t)y * pitch;
        uint32_t src = 0;
        if (al && x + 3 < wlim) src = *reinterpret_cast<const uint32_t *>(edges + o + x);
        else
            for (int k = 0; k < 4 && x + k < w; k++) src |= (uint32_t)edges[o + x + k] << (8 * k);
        uint32_t res = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int li = (int)((last[q * 2 + (k >> 1)] >> (16 * (k & 1))) & 0xffffu) - 1;
            uint32_t v;
            if (li < 0) v = (src >> (8 * k)) & 0xffu;
            else {
                int mx = __float2int_rn(circ[3 * li]), my = __float2int_rn(circ[3 * li + 1]);
                v = (abs(x + k - mx) + abs(y - my) <= 1) ? 255u : 0u;
            }
            res |= v << (8 * k);
        }
        store4(masked + o, x, w, wlim, al, res);
    }
}

// ------------------------------------------------------------------ host orchestration
static int p2_of(int v)
{
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

size_t circles_scratch_bytes(int maps, int h, int w, const i2s_limits_t &lim)
{
    const size_t plane = (size_t)h * canvas_pitch(w);
    size_t b = 0;
    b += align_up(maps * plane, 256);                               // state maps
    b += align_up(maps * plane * 8, 256);                           // edge lists (position, Q10 step), worst case
    b += align_up((size_t)maps * cdiv(w, EB) * cdiv(h, EB) * 8, 256);   // bucket directory
    b += align_up((size_t)maps * lim.cand_cap * 4, 256);            // candidate centres
    b += align_up((size_t)maps * lim.cand_cap * 8, 256);            // estimated circle keys
    b += align_up((size_t)maps * 4 * 4, 256);                       // counters
    b += align_up(sizeof(RadiusTables), 256);
    if (lim.cand_cap > FINISH_SMEM_CAP) b += align_up((size_t)maps * finish_bytes_per_map(p2_of(lim.cand_cap)), 256);
    b += canny_scratch_bytes(maps, h, w);
    return b + 4096;
}

// HoughCircles on every map of `ms`; per-map circles [maps][circle_cap][3] + counts [maps]
int hough_circles_maps(const MapSet &ms, const Dims &dims, float *mcirc, int32_t *mcount, int32_t *status,
                       const i2s_limits_t &lim, Arena &ar, cudaStream_t st)
{
    const int maps = ms.count * ms.n;
    const int h = dims.h, w = dims.w;
    const int spitch = canvas_pitch(w);
    const size_t plane = (size_t)h * spitch;
    uint8_t *state = ar.take<uint8_t>(maps * plane);
    uint2 *edges = ar.take<uint2>(maps * plane);
    const int nbx = cdiv(w, EB), nby = cdiv(h, EB);
    int2 *dir = ar.take<int2>((size_t)maps * nbx * nby);
    int32_t *cand = ar.take<int32_t>((size_t)maps * lim.cand_cap);
    unsigned long long *est = ar.take<unsigned long long>((size_t)maps * lim.cand_cap);
    int32_t *ctr = ar.take<int32_t>((size_t)maps * 3);
    RadiusTables *rtables = ar.take<RadiusTables>(1);
    const int np2 = p2_of(lim.cand_cap);
    unsigned char *gbuf = lim.cand_cap > FINISH_SMEM_CAP ? ar.take<unsigned char>((size_t)maps * finish_bytes_per_map(np2)) : nullptr;
    void *cscratch = ar.take<uint8_t>(canny_scratch_bytes(maps, h, w));
    if (!ar.ok()) { set_error("hough_circles: workspace too small"); return I2S_E_WORKSPACE; }
    I2S_ARG(maps < 65536 && nby < 65536);
    int32_t *ncand = ctr, *nest = ctr + maps, *ecount = ctr + 2 * maps;

    int rc = canny_states(ms, dims, 1, state, spitch, plane, CANNY_LOW, CANNY_HIGH, lim.hyst_passes, status, cscratch, st,
                          nullptr, 0, 0);
    if (rc) return rc;
    I2S_CUDA(cudaMemsetAsync(ctr, 0, sizeof(int32_t) * maps * 3, st));
    {
        ScopedSection sec(SEC_EDGE_LIST, st);
        k_edge_list<<<dim3(cdiv(nbx, EL_WARPS), nby, maps), EL_WARPS * 32, 0, st>>>(ms, dims, state, spitch, plane, edges, plane,
                                                                                   ecount, dir, nbx, nby);
        I2S_CHECK_LAUNCH("k_edge_list");
    }
    {
        ScopedSection sec(SEC_VOTE, st);
        I2S_CUDA(cudaFuncSetAttribute(k_vote_peaks, cudaFuncAttributeMaxDynamicSharedMemorySize, VOTE_SMEM));
        k_vote_peaks<<<dim3(cdiv(w, AT), cdiv(h, AT), maps), VOTE_THREADS, VOTE_SMEM, st>>>(edges, plane, dir, nbx, nby, dims, ms.n,
                                                                                           cand, ncand, lim.cand_cap,
                                                                                           make_uint2(1u | ((uint32_t)AP << 8), (1u | ((uint32_t)AP << 8)) << 16));
        I2S_CHECK_LAUNCH("k_vote_peaks");
    }
    {
        ScopedSection sec(SEC_RADIUS, st);
        // (a variant with four centres per warp -- 8 lanes each, 16-bit bins -- cut the instruction count by
        // 18 % but not the time; row-sorted bucket lists with per-row starts, so that a centre walks only the
        // rows inside its window, scanned 37 % fewer entries but ran slower: shorter per-bucket loops leave more
        // lanes idle, and the row-per-lane loads made the list kernel 10 % slower)
        k_radius_tables<<<1, 256, 0, st>>>(rtables);
        I2S_CHECK_LAUNCH("k_radius_tables");
        k_radius<<<dim3(I2S_RADIUS_GRIDX, maps), RW * 32, 0, st>>>(rtables, edges, plane, dir, nbx, nby, dims, ms.n, cand, ncand, lim.cand_cap, est, nest,
                                                     status);
        I2S_CHECK_LAUNCH("k_radius");
    }
    ScopedSection sec(SEC_CIRCLES_FINISH, st);
    int cshift = 4;                                            // cells of >= 16 px (> minDist)
    while ((size_t)((w >> cshift) + 1) * ((h >> cshift) + 1) > FINISH_MAX_CELLS) cshift++;
    const int cells_x = (w >> cshift) + 1, cells_y = (h >> cshift) + 1;
    size_t smem = (((size_t)cells_x * cells_y * 4 + 15) & ~(size_t)15) + (gbuf ? 0 : finish_bytes_per_map(np2));
    I2S_CUDA(cudaFuncSetAttribute(k_circles_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_circles_finish<<<maps, FINISH_THREADS, smem, st>>>(est, nest, lim.cand_cap, np2, cshift, cells_x, cells_y, mcirc, mcount,
                                              lim.circle_cap, status, ms.n, gbuf);
    I2S_CHECK_LAUNCH("k_circles_finish");
    return I2S_OK;
}

int mask_circles(const uint8_t *edges, uint8_t *masked, const Dims &dims, int n, int pitch, size_t stride,
                 const float *circles, const int32_t *counts, int circle_cap, cudaStream_t st, const int2 *dup)
{
    ScopedSection sec(SEC_MASK, st);
    I2S_ARG(n < 65536);
    k_mask<<<dim3(cdiv(dims.w, KT), cdiv(dims.h, KT), n), 256, 0, st>>>(edges, masked, dims, pitch, stride, circles, counts,
                                                                       circle_cap, dup);
    I2S_CHECK_LAUNCH("k_mask");
    return I2S_OK;
}

size_t find_circles_scratch_bytes(int n, int h, int w, const i2s_limits_t &lim)
{
    const size_t plane = (size_t)h * canvas_pitch(w);
    size_t b = 6 * align_up((size_t)n * plane, 256);                       // the six blurred copies
    b += align_up((size_t)n * I2S_N_UNIQUE * lim.circle_cap * 12, 256);    // per-map circles
    b += align_up((size_t)n * I2S_N_UNIQUE * 4, 256);
    b += align_up((size_t)n * sizeof(int2), 256);
    return b + circles_scratch_bytes(n * I2S_N_UNIQUE, h, w, lim) + 4096;
}

// grey / edges: [n] planes of `pitch` bytes per row, dims.h rows apart (the caller's planes); the blurred
// copies live in the workspace on the library's own canvas pitch
int find_circles(const uint8_t *grey, const uint8_t *edges, const Dims &dims, int n, int pitch, float *circles,
                 int32_t *counts, uint8_t *masked, int32_t *status, const i2s_limits_t &lim, Arena &ar, cudaStream_t st)
{
    const int h = dims.h, w = dims.w;
    const int bpitch = canvas_pitch(w);
    const size_t bplane = (size_t)h * bpitch;
    uint8_t *blur[6];
    for (int k = 0; k < 6; k++) blur[k] = ar.take<uint8_t>((size_t)n * bplane);
    const int maps = n * I2S_N_UNIQUE;
    float *mcirc = ar.take<float>((size_t)maps * lim.circle_cap * 3);
    int32_t *mcount = ar.take<int32_t>(maps);
    int2 *dup = ar.take<int2>(n);
    if (!ar.ok()) { set_error("find_circles: workspace too small"); return I2S_E_WORKSPACE; }
    // internal order: grey, edges, med3, gau3, med5, gau5, med7, gau7
    int rc;
    const size_t gstride = (size_t)h * pitch;
    if ((rc = gauss357(grey, pitch, gstride, blur[1], blur[3], blur[5], bpitch, bplane, dims, n, st))) return rc;
    if ((rc = median357(grey, pitch, gstride, blur[0], blur[2], blur[4], bpitch, bplane, dims, n, st))) return rc;
    MapSet ms{};
    ms.n = n;
    ms.add(grey, pitch, h);
    ms.add(edges, pitch, h);
    for (int k = 0; k < 6; k++) ms.add(blur[k], bpitch, h);
    if ((rc = hough_circles_maps(ms, dims, mcirc, mcount, status, lim, ar, st))) return rc;
    {
        ScopedSection sec(SEC_STACK, st);
        k_stack<<<n, 256, 0, st>>>(mcirc, mcount, n, lim.circle_cap, circles, counts, status, dup);
        I2S_CHECK_LAUNCH("k_stack");
    }
    return mask_circles(edges, masked, dims, n, pitch, (size_t)h * pitch, circles, counts, lim.circle_cap, st, dup);
}

int check_limits(const i2s_limits_t *lim)
{
    I2S_ARG(lim && lim->cand_cap >= 32 && lim->cand_cap <= 16384 && lim->circle_cap >= 1 && lim->circle_cap <= I2S_MAX_CIRCLE_CAP && lim->line_cap >= 2 &&
            lim->line_cap <= 4096 && lim->hyst_passes >= 1);
    return I2S_OK;
}

}  // namespace i2s

using namespace i2s;

extern "C" size_t i2s_hough_circles_workspace_bytes(int n, int h, int w, const i2s_limits_t *lim)
{
    if (n <= 0 || h <= 0 || w <= 0 || !lim) return 4096;
    return circles_scratch_bytes(n, h, w, *lim);
}

extern "C" int i2s_hough_circles(const uint8_t *img, int pitch, int n, int h, int w, float *circles, int32_t *counts,
                                 int32_t *status, const i2s_limits_t *lim, void *ws, size_t ws_bytes, void *stream)
{
    I2S_ARG(img && circles && counts && status && ws && n >= 0 && h > 0 && w > 0 && h < 16384 && w < 16384);
    if (pitch == 0) pitch = w;
    I2S_ARG(pitch >= w);
    int rc = check_limits(lim);
    if (rc) return rc;
    if (n == 0) return I2S_OK;
    Arena ar(ws, ws_bytes);
    MapSet ms = MapSet::single(img, pitch, h, n);
    return hough_circles_maps(ms, Dims::uniform(h, w), circles, counts, status, *lim, ar, (cudaStream_t)stream);
}

extern "C" int i2s_mask_circles(const uint8_t *edges, uint8_t *masked, int pitch, int n, int h, int w, const float *circles,
                                const int32_t *counts, int circle_cap, void *stream)
{
    I2S_ARG(edges && masked && circles && counts && n >= 0 && h > 0 && w > 0 && circle_cap > 0 && circle_cap <= I2S_MAX_CIRCLE_CAP);
    if (pitch == 0) pitch = w;
    I2S_ARG(pitch >= w);
    if (n == 0) return I2S_OK;
    return mask_circles(edges, masked, Dims::uniform(h, w), n, pitch, (size_t)h * pitch, circles, counts, circle_cap,
                        (cudaStream_t)stream, nullptr);
}

extern "C" size_t i2s_find_circles_workspace_bytes(int n, int h, int w, const i2s_limits_t *lim)
{
    if (n <= 0 || h <= 0 || w <= 0 || !lim) return 4096;
    return find_circles_scratch_bytes(n, h, w, *lim);
}

extern "C" int i2s_find_circles(const uint8_t *grey, const uint8_t *edges, int pitch, int n, int h, int w, float *circles,
                                int32_t *counts, uint8_t *masked, int32_t *status, const i2s_limits_t *lim, void *ws,
                                size_t ws_bytes, void *stream)
{
    I2S_ARG(grey && edges && circles && counts && masked && status && ws && n >= 0 && h > 0 && w > 0 && h < 16384 &&
            w < 16384);
    if (pitch == 0) pitch = w;
    I2S_ARG(pitch >= w);
    int rc = check_limits(lim);
    if (rc) return rc;
    if (n == 0) return I2S_OK;
    Arena ar(ws, ws_bytes);
    return find_circles(grey, edges, Dims::uniform(h, w), n, pitch, circles, counts, masked, status, *lim, ar,
                        (cudaStream_t)stream);
}
