// preproc.cu -- greyscale, PIL contrast/brightness, fused Gaussian 3/5/7, median 3/5/7.
// Reference call sites: img2sgf.py:142-149 (contrast, brightness), :153 (grey), :174 (median),
// :175 (Gaussian).  Arithmetic: SURVEY.md Appendix A.1, A.2, A.3, A.9 (all integer / fixed point
// or separately rounded float32, bit-exact).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "preproc.cuh"
#include "profile.cuh"
#include "tma.cuh"
#include "roll_cores.cuh"
#include "median_cores.cuh"

namespace i2s {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count()
{
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        n <= 0)
        return 148;
    return n;
}

// ------------------------------------------------------------------ grey (A.1)
__device__ __forceinline__ uint32_t luma_q15(uint32_t c0, uint32_t c1, uint32_t c2)
{
    return (3735u * c0 + 19235u * c1 + 9798u * c2 + 16384u) >> 15;
}

// 4 pixels (12 bytes in, 4 bytes out) per thread; grid (words of a row / 256, rows, images)
__global__ void __launch_bounds__(256) k_grey(const uint8_t *__restrict__ rgb, int rgb_pitch, uint8_t *__restrict__ grey,
                                              int pitch, int h, int w)
{
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x >= w) return;
    const uint8_t *src = rgb + ((size_t)blockIdx.z * h + y) * rgb_pitch + (size_t)x * 3;
    uint8_t *dst = grey + ((size_t)blockIdx.z * h + y) * pitch;
    const bool al_in = ((reinterpret_cast<uintptr_t>(rgb) | (uintptr_t)rgb_pitch) & 3) == 0;
    const bool al_out = ((reinterpret_cast<uintptr_t>(grey) | (uintptr_t)pitch) & 3) == 0;
    uint32_t out = 0;
    if (al_in && x + 3 < w) {
        const uint32_t *p = reinterpret_cast<const uint32_t *>(src);
        const uint32_t a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        out = luma_q15(a & 255, (a >> 8) & 255, (a >> 16) & 255) | (luma_q15(a >> 24, b & 255, (b >> 8) & 255) << 8) |
              (luma_q15((b >> 16) & 255, b >> 24, c & 255) << 16) | (luma_q15((c >> 8) & 255, (c >> 16) & 255, c >> 24) << 24);
    } else {
        for (int k = 0; k < 4 && x + k < w; k++) out |= luma_q15(src[3 * k], src[3 * k + 1], src[3 * k + 2]) << (8 * k);
    }
    store4(dst, x, w, write_limit(w, pitch, 4), al_out, out);
}

// ------------------------------------------------------------------ contrast + brightness (A.9)
// PIL: ImageEnhance.Contrast(img).enhance(fc) = blend(solid(m), img, fc), m = int(mean(L) + 0.5) with
// L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16; ImageEnhance.Brightness(img).enhance(fb) =
// blend(black, img, fb).  ImagingBlend computes, per channel value, t = a + f * (b - a) in float32 and
// stores 0 if t <= 0, 255 if t >= 255, else trunc(t) -- so both steps are 256-entry tables per image.
__global__ void __launch_bounds__(256) k_luma_sum(const MapSet ms, const Dims dims, int ch, unsigned long long *sums)
{
    const int img_i = blockIdx.y;
    const int2 wh = dims.of(img_i);
    int pitch;
    const uint8_t *img = ms.plane(img_i, pitch);
    unsigned long long s = 0;
    for (int y = blockIdx.x; y < wh.y; y += gridDim.x) {
        const uint8_t *row = img + (size_t)y * pitch;
        if (ch == 1)         // a mode-"L" source converted to RGB has R = G = B = v, whose L is v
            for (int x = threadIdx.x; x < wh.x; x += blockDim.x) s += row[x];
        else
            for (int x = threadIdx.x; x < wh.x; x += blockDim.x)
                s += (19595u * row[3 * x] + 38470u * row[3 * x + 1] + 7471u * row[3 * x + 2] + 0x8000u) >> 16;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(sums + img_i, s);
}

__device__ __forceinline__ uint32_t blend_u8(float a, float f, uint32_t b)
{
    const float t = __fadd_rn(a, __fmul_rn(f, __fsub_rn((float)b, a)));
    return t <= 0.f ? 0u : (t >= 255.f ? 255u : (uint32_t)t);
}

__global__ void __launch_bounds__(256) k_enhance(const MapSet ms, const Dims dims, int ch, uint8_t *__restrict__ out, int opitch,
                                                 size_t ostride, const unsigned long long *sums, float fc, float fb)
{
    __shared__ uint8_t s_lut[256];
    const int img_i = blockIdx.y;
    const int2 wh = dims.of(img_i);
    int pitch;
    const uint8_t *img = ms.plane(img_i, pitch);
    uint8_t *o = out + (size_t)img_i * ostride;
    {
        const int m = (int)((double)sums[img_i] / (double)((long long)wh.x * wh.y) + 0.5);
        const uint32_t c = fc == 1.0f ? threadIdx.x : blend_u8((float)m, fc, threadIdx.x);
        s_lut[threadIdx.x] = (uint8_t)(fb == 1.0f ? c : blend_u8(0.0f, fb, c));
    }
    __syncthreads();
    for (int y = blockIdx.x; y < wh.y; y += gridDim.x) {
        const uint8_t *row = img + (size_t)y * pitch;
        uint8_t *orow = o + (size_t)y * opitch;
        for (int x = threadIdx.x; x < ch * wh.x; x += blockDim.x) orow[x] = s_lut[row[x]];
    }
}

int enhance(const MapSet &ms, const Dims &dims, int ch, uint8_t *out, int opitch, size_t ostride, void *scratch8n, float fc,
            float fb, cudaStream_t st)
{
    ScopedSection sec(SEC_ENHANCE, st);
    unsigned long long *sums = (unsigned long long *)scratch8n;
    I2S_CUDA(cudaMemsetAsync(sums, 0, sizeof(unsigned long long) * ms.n, st));
    dim3 grid(min(dims.h, max(1, 8 * sm_count() / max(ms.n, 1))), ms.n);
    if (fc != 1.0f) {
        k_luma_sum<<<grid, 256, 0, st>>>(ms, dims, ch, sums);
        I2S_CHECK_LAUNCH("k_luma_sum");
    }
    k_enhance<<<grid, 256, 0, st>>>(ms, dims, ch, out, opitch, ostride, sums, fc, fb);
    I2S_CHECK_LAUNCH("k_enhance");
    return I2S_OK;
}

// A single-channel source batch (any pitch / offsets) onto the library's canvas
__global__ void __launch_bounds__(256) k_to_canvas(const MapSet ms, const Dims dims, uint8_t *__restrict__ out, int opitch,
                                                   size_t ostride)
{
    const int img_i = blockIdx.y;
    const int2 wh = dims.of(img_i);
    int pitch;
    const uint8_t *img = ms.plane(img_i, pitch);
    uint8_t *o = out + (size_t)img_i * ostride;
    for (int y = blockIdx.x; y < wh.y; y += gridDim.x)
        for (int x = threadIdx.x; x < wh.x; x += blockDim.x) o[(size_t)y * opitch + x] = img[(size_t)y * pitch + x];
}

int to_canvas(const MapSet &ms, const Dims &dims, uint8_t *out, int opitch, size_t ostride, cudaStream_t st)
{
    ScopedSection sec(SEC_GREY, st);
    dim3 grid(min(dims.h, max(1, 8 * sm_count() / max(ms.n, 1))), ms.n);
    k_to_canvas<<<grid, 256, 0, st>>>(ms, dims, out, opitch, ostride);
    I2S_CHECK_LAUNCH("k_to_canvas");
    return I2S_OK;
}

// ------------------------------------------------------------------ Gaussian 3/5/7, register rolling (A.2)
// One warp walks down a strip of 128 loaded / 120 stored columns: each lane owns one 32-bit word
// (4 pixels) per row and keeps the last 7 rows as unpacked pairs in registers.  Per row: one
// coalesced 128-byte load per warp, the three vertical passes on packed 16-bit halves (Q8 sums
// fit), an exchange of the edge pairs with the neighbour lanes (8 shuffles) and the three
// horizontal passes as 16-bit x 8-bit dot products (IDP.2A) into Q16, one rounding.  Lanes 0 and
// 31 only supply the halo.  No shared memory, no barriers; arithmetic in roll_cores.cuh.
constexpr int GR_TH = 64;          // output rows per warp strip (128 rows / 4 warps measured slower)
constexpr int GR_OW = 120;         // output columns per warp strip
constexpr int GR_WARPS = 8;

__device__ __forceinline__ uint32_t load_word_border(const uint8_t *__restrict__ row, int x, int w, bool al, int mode)
{
    if (al && x >= 0 && x + 3 < w) return __ldg(reinterpret_cast<const uint32_t *>(row + x));
    if (x >= w + 8 || x < -8) return 0;                       // beyond any halo: value never used
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) v |= (uint32_t)__ldg(row + border_index(x + k, w, mode)) << (8 * k);
    return v;
}

__global__ void __launch_bounds__(GR_WARPS * 32, 4) k_gauss357_roll(const uint8_t *__restrict__ src, uint8_t *__restrict__ d3,
                                                                 uint8_t *__restrict__ d5, uint8_t *__restrict__ d7,
                                                                 const Dims dims, int spitch, size_t sstride, int pitch,
                                                                 size_t stride, int strips_x, int strips_y, int total)
{
    const int lane = threadIdx.x & 31;
    const int strip = blockIdx.x * GR_WARPS + (threadIdx.x >> 5);
    if (strip >= total) return;                                // warp-uniform
    const int sx = strip % strips_x, t = strip / strips_x, sy = t % strips_y, img = t / strips_y;
    const int2 wh = dims.of(img);
    const int w = wh.x, h = wh.y;
    if (sx * GR_OW >= w || sy * GR_TH >= h) return;            // strip outside this image (ragged batch)
    const uint8_t *im = src + img * sstride;
    const bool al = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)spitch | (uintptr_t)sstride) & 3) == 0;
    const bool al_out = ((reinterpret_cast<uintptr_t>(d3) | reinterpret_cast<uintptr_t>(d5) | reinterpret_cast<uintptr_t>(d7) |
                          (uintptr_t)pitch | (uintptr_t)stride) & 3) == 0;
    const int wlim = write_limit(w, pitch, 4);
    const int x = sx * GR_OW - 4 + 4 * lane;
    const int y0 = sy * GR_TH, y1 = min(y0 + GR_TH, h);
    const bool store_lane = lane >= 1 && lane <= 30 && x < w;
    uint32_t wl[7], wh7[7];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const uint32_t v = load_word_border(im + (size_t)border_index(y0 - 3 + k, h, BORDER_REFLECT101) * spitch, x, w, al,
                                            BORDER_REFLECT101);
        wl[k] = roll::pair_lo(v); wh7[k] = roll::pair_hi(v);
    }
    // the row entering the window is loaded one iteration ahead of its use
    uint32_t nxt = load_word_border(im + (size_t)border_index(y0 + 3, h, BORDER_REFLECT101) * spitch, x, w, al, BORDER_REFLECT101);
#pragma unroll 1
    for (int yb = y0; yb < y1; yb += 7) {
#pragma unroll
        for (int u = 0; u < 7; u++) {
            const int y = yb + u;
            if (y < y1) {                                      // warp-uniform
                const uint32_t v = nxt;
                nxt = load_word_border(im + (size_t)border_index(y + 4, h, BORDER_REFLECT101) * spitch, x, w, al,
                                       BORDER_REFLECT101);
                wl[(u + 6) % 7] = roll::pair_lo(v); wh7[(u + 6) % 7] = roll::pair_hi(v);
                const uint32_t rl[7] = {wl[u % 7], wl[(u + 1) % 7], wl[(u + 2) % 7], wl[(u + 3) % 7], wl[(u + 4) % 7],
                                        wl[(u + 5) % 7], wl[(u + 6) % 7]};
                const uint32_t rh[7] = {wh7[u % 7], wh7[(u + 1) % 7], wh7[(u + 2) % 7], wh7[(u + 3) % 7], wh7[(u + 4) % 7],
                                        wh7[(u + 5) % 7], wh7[(u + 6) % 7]};
                uint32_t V[6];
                roll::gauss_vertical(rl, rh, V);
                // neighbour pairs: left lane's (V2,V3) [and (V0,V1) for the 7-tap], right lane's (V0,V1) [and (V2,V3)]
                const uint32_t l3 = __shfl_up_sync(0xffffffffu, V[1], 1), r3 = __shfl_down_sync(0xffffffffu, V[0], 1);
                const uint32_t l5 = __shfl_up_sync(0xffffffffu, V[3], 1), r5 = __shfl_down_sync(0xffffffffu, V[2], 1);
                const uint32_t l7 = __shfl_up_sync(0xffffffffu, V[5], 1), r7 = __shfl_down_sync(0xffffffffu, V[4], 1);
                const uint32_t ll7 = __shfl_up_sync(0xffffffffu, V[4], 1), rr7 = __shfl_down_sync(0xffffffffu, V[5], 1);
                if (store_lane) {
                    const size_t o = img * stride + (size_t)y * pitch;
                    if (d3) store4(d3 + o, x, w, wlim, al_out, roll::gauss_h3(l3, V[0], V[1], r3));
                    if (d5) store4(d5 + o, x, w, wlim, al_out, roll::gauss_h5(l5, V[2], V[3], r5));
                    if (d7) store4(d7 + o, x, w, wlim, al_out, roll::gauss_h7(ll7, l7, V[4], V[5], r7, rr7));
                }
            }
        }
    }
}

int gauss357(const uint8_t *src, int spitch, size_t sstride, uint8_t *d3, uint8_t *d5, uint8_t *d7, int pitch, size_t stride,
             const Dims &dims, int n, cudaStream_t st)
{
    ScopedSection sec(SEC_GAUSS, st);
    const int strips_x = cdiv(dims.w, GR_OW), strips_y = cdiv(dims.h, GR_TH);
    const long long total = (long long)n * strips_x * strips_y;
    I2S_ARG(total < (1ll << 31));
    k_gauss357_roll<<<(unsigned)((total + GR_WARPS - 1) / GR_WARPS), GR_WARPS * 32, 0, st>>>(src, d3, d5, d7, dims, spitch, sstride,
                                                                                            pitch, stride, strips_x, strips_y, (int)total);
    I2S_CHECK_LAUNCH("k_gauss357_roll");
    return I2S_OK;
}

// ------------------------------------------------------------------ median (A.3)
// Exact b x b medians with BORDER_REPLICATE, bit-sliced.  The staged tile is transposed once per block
// into eight 1-bit planes (plus "== 255" and "== 0").
//   * Saturated-window verdicts (settle_word, median_cores.cuh): more than half of the window pixels
//     equal to 255 (0) => the median is 255 (0), evaluated for 32 pixels per logic instruction.  Printed
//     diagrams are mostly saturated paper and ink: the verdicts settle 95-98 % of the pixels.
//   * The 4-pixel groups that still hold an unsettled pixel (5-9 % on diagrams -- thin bands along the stone
//     rims --, 40-65 % on noisy scans) are queued in shared memory and every lane finishes one group by
//     MSB-first rank selection on the window bits (select_group), so the warps doing the expensive part
//     are dense whatever the content.  (Also tried: rank selection for 32 pixels per thread with carry-save
//     adder trees on whole words -- no POPC, 3.5x fewer instructions on noisy content, but 130 registers
//     per thread: slower in practice.)
constexpr int MED_THREADS = 256;

template <int MASK, int RS, int HX, int SH, int GW>
__device__ __forceinline__ void median_settle(const uint32_t (&s_bits)[10][SH][GW + 1], uint32_t (&s_set)[3][2][MT_H][MT_W / 32])
{
    constexpr int WORDS = MT_W / 32, UNITS = 3 * MT_H * WORDS;
    for (int u = threadIdx.x; u < UNITS; u += blockDim.x) {
        const int bi = u / (MT_H * WORDS), rest = u - bi * (MT_H * WORDS), ty = rest / WORDS, j = rest - ty * WORDS;
        if (!((MASK >> bi) & 1)) continue;
        uint32_t a, b;
        if (bi == 0) { a = settle_word<3, RS, HX, SH, GW>(s_bits[8], ty, j); b = settle_word<3, RS, HX, SH, GW>(s_bits[9], ty, j); }
        else if (bi == 1) { a = settle_word<5, RS, HX, SH, GW>(s_bits[8], ty, j); b = settle_word<5, RS, HX, SH, GW>(s_bits[9], ty, j); }
        else { a = settle_word<7, RS, HX, SH, GW>(s_bits[8], ty, j); b = settle_word<7, RS, HX, SH, GW>(s_bits[9], ty, j); }
        s_set[bi][0][ty][j] = a;
        s_set[bi][1][ty][j] = b;
    }
}

// Rank selection for ONE group of 4 adjacent pixels (row ty, columns gx..gx+3 of the tile): for each plane
// the window bits of the four pixels (B rows x (B+3) columns, rows packed at a stride of B+3 bits) are
// gathered once; per pixel the median is selected MSB-first with C = the set of still-possible window
// elements and k = the rank wanted inside C: the next result bit is 0 iff k < popc(C & ~plane).  The groups
// are taken from a queue (see k_median), one per lane.
template <int B, int RS, int HX, int SH, int GW>
__device__ __forceinline__ uint32_t select_group(const uint32_t (&s_bits)[10][SH][GW + 1], int ty, int gx)
{
    constexpr int F = B + 3;                         // window columns of 4 adjacent pixels
    constexpr int RPW = 32 / F;                      // window rows packed per 32-bit word
    constexpr int NW = (B + RPW - 1) / RPW;          // words per window
    constexpr int RO = RS - B / 2;                   // first window row inside the staged halo
    uint32_t fm[NW];                                 // per-word masks of the B low bits of every packed row field
#pragma unroll
    for (int wd = 0; wd < NW; wd++) {
        fm[wd] = 0;
#pragma unroll
        for (int r = wd * RPW; r < B && r < (wd + 1) * RPW; r++) fm[wd] |= ((1u << B) - 1u) << ((r - wd * RPW) * F);
    }
    const int start = gx + HX - B / 2, wi = start >> 5, sh = start & 31;
    uint32_t P[8][NW];
#pragma unroll
    for (int bit = 0; bit < 8; bit++) {
#pragma unroll
        for (int wd = 0; wd < NW; wd++) P[bit][wd] = 0;
#pragma unroll
        for (int r = 0; r < B; r++) {
            const uint32_t *rw = &s_bits[bit][ty + RO + r][wi];
            const uint32_t bits = __funnelshift_r(rw[0], rw[1], sh) & ((1u << F) - 1u);
            P[bit][r / RPW] |= bits << ((r % RPW) * F);
        }
    }
    uint32_t packed = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint32_t C[NW];                              // candidate set, kept shifted to pixel j's window columns
#pragma unroll
        for (int wd = 0; wd < NW; wd++) C[wd] = fm[wd] << j;
        int k = (B * B) / 2;
        uint32_t val = 0;
#pragma unroll
        for (int bit = 7; bit >= 0; bit--) {
            uint32_t Z[NW], O[NW];
            int nz = 0;
#pragma unroll
            for (int wd = 0; wd < NW; wd++) {
                const uint32_t pj = P[bit][wd];
                Z[wd] = C[wd] & ~pj;
                O[wd] = C[wd] & pj;
                nz += __popc(Z[wd]);
            }
            const bool zero = k < nz;                // the median has a 0 in this bit
#pragma unroll
            for (int wd = 0; wd < NW; wd++) C[wd] = zero ? Z[wd] : O[wd];
            if (!zero) { k -= nz; val |= 1u << bit; }
        }
        packed |= val << (8 * j);
    }
    return packed;
}

// MASK selects the window sizes computed from ONE staged tile and ONE set of bit planes:
// bit 0 -> 3x3 into dst3, bit 1 -> 5x5 into dst5, bit 2 -> 7x7 into dst7.
__device__ __forceinline__ void store_unit(uint8_t *row, int x, int w, int wlim, bool al, const uint32_t (&out)[8])
{
#pragma unroll
    for (int g = 0; g < 8; g++)
        if (x + 4 * g < w) store4(row, x + 4 * g, w, wlim, al, out[g]);
}

template <int MASK> __global__ void __launch_bounds__(MED_THREADS, 4) k_median(const uint8_t *__restrict__ src,
                                                                         uint8_t *__restrict__ dst3, uint8_t *__restrict__ dst5,
                                                                         uint8_t *__restrict__ dst7, const Dims dims, int spitch,
                                                                         size_t sstride, int pitch, size_t stride)
{
    constexpr int RS = (MASK & 4) ? 3 : (MASK & 2) ? 2 : 1, HX = 16;   // x halo 16: bulk-copy rows are 16-byte aligned
    constexpr int SW = MT_W + 2 * HX, SH = MT_H + 2 * RS;
    constexpr int GW = (SW + 31) / 32;               // 32-pixel groups per tile row
    constexpr int WORDS = MT_W / 32, UNITS = MT_H * WORDS;
    __shared__ __align__(128) uint8_t s_in[SH * SW];
    __shared__ uint32_t s_bits[10][SH][GW + 1];      // planes 0..7: bits of the pixel; 8: pixel == 255; 9: pixel == 0
    __shared__ uint32_t s_set[3][2][MT_H][MT_W / 32]; // per window size: median settled at 255 / at 0, one bit per pixel
    __shared__ uint16_t s_glist[3][UNITS * 8];       // per window size: the 4-pixel groups queued for select_group (unit << 3 | group)
    __shared__ int s_ngrp[3];
    __shared__ uint64_t s_bar;
    const int2 wh = dims.of(blockIdx.z);
    const int w = wh.x, h = wh.y;
    const int x0 = blockIdx.x * MT_W, y0 = blockIdx.y * MT_H;
    if (x0 >= w || y0 >= h) return;                  // tile outside this image (ragged batch)
    const uint8_t *img = src + blockIdx.z * sstride;
    const uintptr_t src_al = reinterpret_cast<uintptr_t>(src) | (uintptr_t)spitch | (uintptr_t)sstride;
    const bool al_in = (src_al & 3) == 0, bulk = (src_al & 15) == 0;
    const bool al = ((reinterpret_cast<uintptr_t>(dst3) | reinterpret_cast<uintptr_t>(dst5) | reinterpret_cast<uintptr_t>(dst7) |
                      (uintptr_t)pitch | (uintptr_t)stride) & 3) == 0;
    const int wlim = write_limit(w, pitch, 4);
    if (threadIdx.x < 3) s_ngrp[threadIdx.x] = 0;
    stage_tile_bulk(s_in, img, h, w, spitch, x0 - HX, y0 - RS, SW, SH, BORDER_REPLICATE, bulk, al_in, &s_bar);
    {   // bit planes: s_bits[b][row][g] bit i = bit b of tile pixel (row, 32 g + i).  A thread turns
        // 8 adjacent pixels into one byte of every plane: bit b of the 4 bytes of a word is gathered
        // into a nibble by one multiply ((x & 0x01010101) * 0x01020408 puts byte j's bit at 24 + j).
        static_assert(SW % 32 == 0, "whole words of a plane row");
        uint8_t *planes = reinterpret_cast<uint8_t *>(&s_bits[0][0][0]);
        constexpr int ROWB = (GW + 1) * 4, PLB = SH * ROWB;           // bytes per plane row / per plane
        auto nib = [](uint32_t v) { return ((v & 0x01010101u) * 0x01020408u) >> 24; };
        auto all8 = [](uint32_t v) { v &= v >> 1; v &= v >> 2; v &= v >> 4; return v; };   // bit 0 of each byte: byte == 0xff
        for (int i = threadIdx.x; i < SH * (SW / 8); i += blockDim.x) {
            const int row = i / (SW / 8), k = i - row * (SW / 8);
            const uint2 v = *reinterpret_cast<const uint2 *>(s_in + row * SW + 8 * k);
            uint8_t *dstb = planes + row * ROWB + k;
#pragma unroll
            for (int bit = 0; bit < 8; bit++) dstb[bit * PLB] = (uint8_t)(nib(v.x >> bit) | (nib(v.y >> bit) << 4));
            dstb[8 * PLB] = (uint8_t)(nib(all8(v.x)) | (nib(all8(v.y)) << 4));
            dstb[9 * PLB] = (uint8_t)(nib(all8(~v.x)) | (nib(all8(~v.y)) << 4));
        }
        for (int i = threadIdx.x; i < 10 * SH; i += blockDim.x) s_bits[i / SH][i % SH][GW] = 0;   // pad word
    }
    __syncthreads();
    median_settle<MASK, RS, HX, SH, GW>(s_bits, s_set);
    __syncthreads();
    // One thread per (window size, row, 32-pixel word): write the verdict bytes of the 32 pixels (0 for the
    // unsettled ones, for now) and queue the 4-pixel groups that hold an unsettled pixel.
    for (int t = threadIdx.x; t < 3 * UNITS; t += blockDim.x) {
        const int bi = 2 - t / UNITS, u = t % UNITS, ty = u / WORDS, j = u % WORDS;     // 7x7 first
        if (!((MASK >> bi) & 1)) continue;
        const int y = y0 + ty, x = x0 + 32 * j;
        if (y >= h || x >= w) continue;
        uint8_t *dst = bi == 2 ? dst7 : (bi == 1 ? dst5 : dst3);
        const uint32_t m255 = s_set[bi][0][ty][j];
        uint32_t un = ~(m255 | s_set[bi][1][ty][j]);                          // unsettled pixels
        if (w - x < 32) un &= (1u << (w - x)) - 1u;                            // columns beyond the image do not count
        un |= un >> 1; un |= un >> 2;                                          // bit 4g: group g has an unsettled pixel
        const uint32_t gm = ((((un & 0x1111u) * 0x1248u) >> 12) & 0x0fu) | ((((un >> 16) & 0x1111u) * 0x1248u) >> 8 & 0xf0u);
        uint32_t out[8];
#pragma unroll
        for (int g = 0; g < 8; g++) out[g] = ((((m255 >> (4 * g)) & 0xfu) * 0x00204081u) & 0x01010101u) * 0xffu;
        store_unit(dst + blockIdx.z * stride + (size_t)y * pitch, x, w, wlim, al, out);
        if (gm) {
            int at = atomicAdd(&s_ngrp[bi], __popc(gm));
            for (uint32_t m = gm; m; m &= m - 1) s_glist[bi][at++] = (uint16_t)((u << 3) | (__ffs(m) - 1));
        }
    }
    __syncthreads();
    // the queued groups of all sizes as one sequence, largest windows first: one group per lane
    const int n7 = s_ngrp[2], n5 = s_ngrp[1], n3 = s_ngrp[0];
    for (int i = threadIdx.x; i < n7 + n5 + n3; i += blockDim.x) {
        const int bi = i < n7 ? 2 : (i < n7 + n5 ? 1 : 0);
        const int e = s_glist[bi][i - (bi == 2 ? 0 : (bi == 1 ? n7 : n7 + n5))];
        const int u = e >> 3, ty = u / WORDS, gx = 32 * (u % WORDS) + 4 * (e & 7);
        uint8_t *dst = (bi == 2 ? dst7 : (bi == 1 ? dst5 : dst3)) + blockIdx.z * stride + (size_t)(y0 + ty) * pitch;
        uint32_t v;
        if (bi == 2) v = select_group<7, RS, HX, SH, GW>(s_bits, ty, gx);
        else if (bi == 1) v = select_group<5, RS, HX, SH, GW>(s_bits, ty, gx);
        else v = select_group<3, RS, HX, SH, GW>(s_bits, ty, gx);
        store4(dst, x0 + gx, w, wlim, al, v);
    }
}

// ---- a lone 3x3: 19-exchange selection network on two pixels per register (N. Devillard's opt_med9
// ordering, checked exhaustively with the 0-1 principle).  An exchange is one VIMNMX.U16x2 min + one max.
__device__ __forceinline__ void cex(uint32_t &a, uint32_t &b)
{
    uint32_t lo = __vminu2(a, b);
    b = __vmaxu2(a, b);
    a = lo;
}

__device__ __forceinline__ uint32_t net_median9(uint32_t (&p)[9])
{
    cex(p[1], p[2]); cex(p[4], p[5]); cex(p[7], p[8]); cex(p[0], p[1]); cex(p[3], p[4]); cex(p[6], p[7]);
    cex(p[1], p[2]); cex(p[4], p[5]); cex(p[7], p[8]); cex(p[0], p[3]); cex(p[5], p[8]); cex(p[4], p[7]);
    cex(p[3], p[6]); cex(p[1], p[4]); cex(p[2], p[5]); cex(p[4], p[7]); cex(p[4], p[2]); cex(p[6], p[4]);
    cex(p[4], p[2]);
    return p[4];
}

__global__ void __launch_bounds__(256) k_median3_net(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                     const Dims dims, int spitch, size_t sstride, int pitch, size_t stride)
{
    constexpr int B = 3, R = 1, HX = 16;
    constexpr int SW = MT_W + 2 * HX, SH = MT_H + 2 * R;
    __shared__ __align__(128) uint8_t s_in[SH * SW];
    __shared__ uint64_t s_bar;
    const int2 wh = dims.of(blockIdx.z);
    const int w = wh.x, h = wh.y;
    const int x0 = blockIdx.x * MT_W, y0 = blockIdx.y * MT_H;
    if (x0 >= w || y0 >= h) return;
    const uint8_t *img = src + blockIdx.z * sstride;
    uint8_t *out = dst + blockIdx.z * stride;
    const uintptr_t src_al = reinterpret_cast<uintptr_t>(src) | (uintptr_t)spitch | (uintptr_t)sstride;
    const bool al_in = (src_al & 3) == 0, bulk = (src_al & 15) == 0;
    const bool al = ((reinterpret_cast<uintptr_t>(dst) | (uintptr_t)pitch | (uintptr_t)stride) & 3) == 0;
    const int wlim = write_limit(w, pitch, 4);
    stage_tile_bulk(s_in, img, h, w, spitch, x0 - HX, y0 - R, SW, SH, BORDER_REPLICATE, bulk, al_in, &s_bar);
    for (int idx = threadIdx.x; idx < MT_H * (MT_W / 4); idx += blockDim.x) {
        int ty = idx / (MT_W / 4), gx = (idx - ty * (MT_W / 4)) * 4;
        int y = y0 + ty, x = x0 + gx;
        if (y >= h || x >= w) continue;
        // e[dy][k] = (pixel x-R+k | pixel x-R+k+1 << 16): pair (x,x+1) uses k = 0..B-1, pair (x+2,x+3) k = 2..B+1
        uint32_t e[B][B + 2];
#pragma unroll
        for (int dy = 0; dy < B; dy++) {
            const uint32_t *rw = reinterpret_cast<const uint32_t *>(s_in + (ty + dy) * SW + gx + HX - 4);   // bytes x-4 .. x+7
            uint32_t X[6];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                uint32_t v = rw[j];
                X[2 * j] = __byte_perm(v, 0, 0x4140);
                X[2 * j + 1] = __byte_perm(v, 0, 0x4342);
            }
#pragma unroll
            for (int k = 0; k < B + 2; k++) {
                const int off = 4 - R + k;                        // byte offset of the first pixel from x-4
                e[dy][k] = (off & 1) ? __funnelshift_r(X[off >> 1], X[(off >> 1) + 1], 16) : X[off >> 1];
            }
        }
        uint32_t p[9], q[9];
#pragma unroll
        for (int dy = 0; dy < B; dy++)
#pragma unroll
            for (int dx = 0; dx < B; dx++) { p[dy * B + dx] = e[dy][dx]; q[dy * B + dx] = e[dy][dx + 2]; }
        const uint32_t pa = net_median9(p), pb = net_median9(q);
        const uint32_t packed = (pa & 0xffu) | ((pa >> 8) & 0xff00u) | ((pb & 0xffu) << 16) | ((pb << 8) & 0xff000000u);
        store4(out + (size_t)y * pitch, x, w, wlim, al, packed);
    }
}

// medianBlur 3, 5 and 7 of the same images from one staged tile and one set of bit planes (the blur
// pyramid of img2sgf.py:171-175 needs all three).
int median357(const uint8_t *src, int spitch, size_t sstride, uint8_t *d3, uint8_t *d5, uint8_t *d7, int pitch, size_t stride,
              const Dims &dims, int n, cudaStream_t st)
{
    if (n == 0) return I2S_OK;
    dim3 grid(cdiv(dims.w, MT_W), cdiv(dims.h, MT_H), n);
    ScopedSection sec(SEC_MEDIAN, st);
    k_median<7><<<grid, MED_THREADS, 0, st>>>(src, d3, d5, d7, dims, spitch, sstride, pitch, stride);
    I2S_CHECK_LAUNCH("k_median");
    return I2S_OK;
}

}  // namespace i2s

using namespace i2s;

extern "C" const char *i2s_last_error(void) { return i2s::g_err; }
extern "C" int i2s_version(void) { return 200; }
extern "C" int i2s_canvas_pitch(int w) { return w > 0 ? i2s::canvas_pitch(w) : 0; }
extern "C" void i2s_default_limits(i2s_limits_t *lim)
{
    lim->cand_cap = 4096;
    lim->circle_cap = 4096;
    lim->line_cap = 1024;
    lim->hyst_passes = 5;
}
extern "C" void i2s_default_params(i2s_params_t *p)
{
    p->line_threshold = 0;          // choose_threshold() per image, img2sgf.py:606-613,638
    p->black_threshold = 128;       // :45
    p->canny_low = 50;              // :47
    p->canny_high = 200;            // :48
    p->contrast_factor = 1.0f;      // input already enhanced
    p->brightness_factor = 1.0f;
}

extern "C" int i2s_grey(const uint8_t *rgb, int rgb_pitch, uint8_t *grey, int pitch, int n, int h, int w, void *stream)
{
    I2S_ARG(rgb && grey && n >= 0 && h > 0 && w > 0 && h < 65536 && n < 65536);
    if (rgb_pitch == 0) rgb_pitch = 3 * w;
    if (pitch == 0) pitch = w;
    I2S_ARG(rgb_pitch >= 3 * w && pitch >= w);
    if (n == 0) return I2S_OK;
    cudaStream_t st = (cudaStream_t)stream;
    ScopedSection sec(SEC_GREY, st);
    k_grey<<<dim3(cdiv(cdiv(w, 4), 256), h, n), 256, 0, st>>>(rgb, rgb_pitch, grey, pitch, h, w);
    I2S_CHECK_LAUNCH("k_grey");
    return I2S_OK;
}

extern "C" int i2s_enhance(const uint8_t *rgb, int rgb_pitch, uint8_t *out, int out_pitch, void *scratch8n, int n, int h,
                           int w, double contrast_factor, double brightness_factor, void *stream)
{
    I2S_ARG(rgb && out && scratch8n && n >= 0 && h > 0 && w > 0 && n < 65536);
    if (rgb_pitch == 0) rgb_pitch = 3 * w;
    if (out_pitch == 0) out_pitch = 3 * w;
    I2S_ARG(rgb_pitch >= 3 * w && out_pitch >= 3 * w);
    if (n == 0) return I2S_OK;
    MapSet ms = MapSet::single(rgb, rgb_pitch, h, n);
    return enhance(ms, Dims::uniform(h, w), 3, out, out_pitch, (size_t)h * out_pitch, scratch8n, (float)contrast_factor,
                   (float)brightness_factor, (cudaStream_t)stream);
}

extern "C" int i2s_gauss357(const uint8_t *src, uint8_t *dst3, uint8_t *dst5, uint8_t *dst7, int n, int h, int w,
                            int pitch, void *stream)
{
    I2S_ARG(src && n >= 0 && h > 0 && w > 0);
    if (pitch == 0) pitch = w;
    I2S_ARG(pitch >= w);
    if (n == 0) return I2S_OK;
    return gauss357(src, pitch, (size_t)h * pitch, dst3, dst5, dst7, pitch, (size_t)h * pitch, Dims::uniform(h, w), n,
                    (cudaStream_t)stream);
}

extern "C" int i2s_median(const uint8_t *src, uint8_t *dst, int n, int h, int w, int pitch, int b, void *stream)
{
    I2S_ARG(src && dst && n >= 0 && n < 65536 && h > 0 && w > 0 && (b == 1 || b == 3 || b == 5 || b == 7));
    if (pitch == 0) pitch = w;
    I2S_ARG(pitch >= w);
    if (n == 0) return I2S_OK;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(cdiv(w, MT_W), cdiv(h, MT_H), n);
    const Dims dims = Dims::uniform(h, w);
    const size_t stride = (size_t)h * pitch;
    ScopedSection sec(SEC_MEDIAN, st);
    if (b == 1) I2S_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * stride, cudaMemcpyDeviceToDevice, st));
    else if (b == 3) k_median3_net<<<grid, 256, 0, st>>>(src, dst, dims, pitch, stride, pitch, stride);
    else if (b == 5) k_median<2><<<grid, MED_THREADS, 0, st>>>(src, nullptr, dst, nullptr, dims, pitch, stride, pitch, stride);
    else k_median<4><<<grid, MED_THREADS, 0, st>>>(src, nullptr, nullptr, dst, dims, pitch, stride, pitch, stride);
    I2S_CHECK_LAUNCH("k_median");
    return I2S_OK;
}
