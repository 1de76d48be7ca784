// circles.cu -- gradient-voting Hough circle detector, the 10-call stack, and circle masking.
// Reference call sites: cv.HoughCircles(.., HOUGH_GRADIENT, 1, 10, [], 100, 30, 1, 30)
// img2sgf.py:179-186; masking loop :191-198.  Arithmetic: SURVEY.md Appendix A.5, A.6
// (integer votes; float32 step vectors / distances / radius scoring with IEEE rn ops).
#include "canny.cuh"
#include "circles.cuh"
#include "sort.cuh"
#include "profile.cuh"

namespace i2s {

constexpr int MIN_R = 1, MAX_R = 30, ACC_THR = 30, NBINS = 290;
constexpr int CANNY_LOW = 50, CANNY_HIGH = 100;

// ------------------------------------------------------------------ K5a: edge lists
// The edge pixels of every map are compacted once into a global list of (position, Q10 gradient
// step), bucketed by 32x32 pixel tile: one block scans its tile of the state map (ballot
// compaction, no per-pixel atomics), reserves a contiguous slice of the map's list with a single
// atomic, recomputes the Sobel gradient of each edge pixel from the source image and stores
// sx = cvRound(dx*1024/mag), sy likewise.  Order inside a bucket is arbitrary;
// votes commute.
constexpr int EB = 32;

__device__ __forceinline__ void sobel_at(const uint8_t *__restrict__ img, int h, int w, int x, int y, int &dx,
                                         int &dy)
{
    int xm = x > 0 ? x - 1 : 0, xp = x < w - 1 ? x + 1 : w - 1;
    const uint8_t *r0 = img + (size_t)(y > 0 ? y - 1 : 0) * w;
    const uint8_t *r1 = img + (size_t)y * w;
    const uint8_t *r2 = img + (size_t)(y < h - 1 ? y + 1 : h - 1) * w;
    int p00 = __ldg(r0 + xm), p01 = __ldg(r0 + x), p02 = __ldg(r0 + xp);
    int p10 = __ldg(r1 + xm), p12 = __ldg(r1 + xp);
    int p20 = __ldg(r2 + xm), p21 = __ldg(r2 + x), p22 = __ldg(r2 + xp);
    dx = (p02 + 2 * p12 + p22) - (p00 + 2 * p10 + p20);
    dy = (p20 + 2 * p21 + p22) - (p00 + 2 * p01 + p02);
}

// Second phase shared by both compaction kernels: reserve a slice of the map's list per bucket,
// publish the directory entry, then recompute the Sobel gradient of every compacted edge pixel
// and store (position, Q10 step).
__device__ __forceinline__ void edge_emit(const uint8_t *__restrict__ img, int h, int w, int map, size_t plane,
                                          uint32_t (*s_pos)[EB * EB], int *s_n, int *s_off, int *s_end,
                                          uint2 *__restrict__ edges, int32_t *ecount, int2 *dir, int nbx, int nby)
{
    __syncthreads();
    if (threadIdx.x < 4) {
        const int sub = threadIdx.x;
        const int bxx = blockIdx.x * 2 + (sub & 1), byy = blockIdx.y * 2 + (sub >> 1);
        const int n = s_n[sub];
        int off = 0;
        if (bxx < nbx && byy < nby) {
            off = n ? atomicAdd(ecount + map, n) : 0;
            dir[((size_t)map * nby + byy) * nbx + bxx] = make_int2(off, n);
        }
        s_off[sub] = off;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        s_end[0] = 0;
        for (int k = 0; k < 4; k++) s_end[k + 1] = s_end[k] + s_n[k];
    }
    __syncthreads();
    uint2 *out = edges + (size_t)map * plane;
    const int total = s_end[4];
    int sub = 0;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        while (i >= s_end[sub + 1]) sub++;
        const int li = i - s_end[sub];
        const uint32_t e = s_pos[sub][li];
        const int x = e & 0xffff, y = e >> 16;
        int dx, dy;
        sobel_at(img, h, w, x, y, dx, dy);
        int sx = 0, sy = 0;
        if (dx != 0 || dy != 0) {
            float vx = (float)dx, vy = (float)dy;
            float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)));
            if (!(mag < 1.0f)) {
                sx = __float2int_rn(__fdiv_rn(__fmul_rn(vx, 1024.0f), mag));
                sy = __float2int_rn(__fdiv_rn(__fmul_rn(vy, 1024.0f), mag));
            }
        }
        out[s_off[sub] + li] = make_uint2(e, (uint32_t)(sx & 0xffff) | ((uint32_t)sy << 16));   // (0,0) step = no vote
    }
}

__global__ void __launch_bounds__(256) k_edge_buckets(const MapSet ms, const uint8_t *__restrict__ state, int h, int w,
                                                      bool al, uint2 *__restrict__ edges, int32_t *ecount, int2 *dir,
                                                      int nbx, int nby)
{
    // one block compacts a 2x2 group of buckets (64x64 pixels)
    __shared__ uint32_t s_pos[4][EB * EB];
    __shared__ int s_n[4], s_off[4], s_end[5];
    static_assert(EB * (EB / 4) == 256, "one 32-bit word of a bucket's state tile per thread");
    const size_t plane = (size_t)h * w;
    const int map = blockIdx.z;
    const uint8_t *img = ms.plane(map, plane);
    const uint8_t *stm = state + map * plane;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 4) s_n[threadIdx.x] = 0;
    __syncthreads();
    // all four state words of this thread first (independent loads in flight), then the compaction
    uint32_t vv[4];
#pragma unroll
    for (int sub = 0; sub < 4; sub++) {
        const int bxx = blockIdx.x * 2 + (sub & 1), byy = blockIdx.y * 2 + (sub >> 1);
        const int ty = threadIdx.x / (EB / 4), gx = (threadIdx.x % (EB / 4)) * 4;
        const int y = byy * EB + ty, x = bxx * EB + gx;
        uint32_t v = 0;
        if (bxx < nbx && byy < nby && y < h && x < w) {
            const uint8_t *p = stm + (size_t)y * w + x;
            if (al && x + 3 < w) v = __ldg(reinterpret_cast<const uint32_t *>(p));
            else
                for (int k = 0; k < 4 && x + k < w; k++) v |= (uint32_t)__ldg(p + k) << (8 * k);
        }
        vv[sub] = v;
    }
#pragma unroll
    for (int sub = 0; sub < 4; sub++) {
        const int bxx = blockIdx.x * 2 + (sub & 1), byy = blockIdx.y * 2 + (sub >> 1);
        if (bxx >= nbx || byy >= nby) continue;                      // block-uniform
        const int ty = threadIdx.x / (EB / 4), gx = (threadIdx.x % (EB / 4)) * 4;
        const int y = byy * EB + ty, x = bxx * EB + gx;
        uint32_t v = vv[sub];
        v &= 0x02020202u;
        const int nb = __popc(v);                                    // 0..4 edge pixels in this word
        const uint32_t b0 = __ballot_sync(0xffffffffu, nb & 1), b1 = __ballot_sync(0xffffffffu, nb & 2),
                       b2 = __ballot_sync(0xffffffffu, nb & 4);
        const uint32_t lt = (1u << lane) - 1u;
        const int total = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
        int base = 0;
        if (lane == 0 && total) base = atomicAdd(&s_n[sub], total);
        base = __shfl_sync(0xffffffffu, base, 0);
        int pos = base + __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt);
        while (v) {
            int k = (__ffs(v) - 1) >> 3;
            v &= ~(0xffu << (8 * k));
            s_pos[sub][pos++] = ((uint32_t)y << 16) | (uint32_t)(x + k);
        }
    }
    edge_emit(img, h, w, map, plane, s_pos, s_n, s_off, s_end, edges, ecount, dir, nbx, nby);
}

// Same result with 16 pixels (one 128-bit load) per thread: thread t owns row t/4 and the 16-pixel
// column group t%4 of the block's 64x64 pixels, so a warp covers 8 rows of two buckets (lanes with
// the same bit 1 share a bucket) and needs one 5-bit ballot prefix per 16 pixels instead of one
// 3-bit prefix per 4.  Requires w % 16 == 0 and a 16-byte aligned state map.
__global__ void __launch_bounds__(256) k_edge_buckets16(const MapSet ms, const uint8_t *__restrict__ state, int h, int w,
                                                        uint2 *__restrict__ edges, int32_t *ecount, int2 *dir,
                                                        int nbx, int nby)
{
    __shared__ uint32_t s_pos[4][EB * EB];
    __shared__ int s_n[4], s_off[4], s_end[5];
    const size_t plane = (size_t)h * w;
    const int map = blockIdx.z;
    const uint8_t *img = ms.plane(map, plane);
    const uint8_t *stm = state + map * plane;
    const int lane = threadIdx.x & 31;
    const int row = threadIdx.x >> 2, cg = threadIdx.x & 3;
    const int y = blockIdx.y * (2 * EB) + row, x = blockIdx.x * (2 * EB) + cg * 16;
    const int sub = (row >> 5) * 2 + (cg >> 1);
    if (threadIdx.x < 4) s_n[threadIdx.x] = 0;
    __syncthreads();
    uint4 v = make_uint4(0, 0, 0, 0);
    if (y < h && x < w) v = __ldg(reinterpret_cast<const uint4 *>(stm + (size_t)y * w + x));
    uint32_t wd[4] = {v.x & 0x02020202u, v.y & 0x02020202u, v.z & 0x02020202u, v.w & 0x02020202u};
    const int nbits = __popc(wd[0]) + __popc(wd[1]) + __popc(wd[2]) + __popc(wd[3]);      // 0..16
    const uint32_t bm = (lane & 2) ? 0xccccccccu : 0x33333333u;                            // lanes of my bucket
    const uint32_t lt = ((1u << lane) - 1u) & bm;
    int pre = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const uint32_t b = __ballot_sync(0xffffffffu, (nbits >> k) & 1);
        pre += __popc(b & lt) << k;
        tot += __popc(b & bm) << k;
    }
    int base = 0;
    if ((lane == 0 || lane == 2) && tot) base = atomicAdd(&s_n[sub], tot);
    base = __shfl_sync(0xffffffffu, base, lane & 2);
    int pos = base + pre;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t m = wd[k];
        while (m) {
            const int j = (__ffs(m) - 1) >> 3;
            m &= m - 1;
            s_pos[sub][pos++] = ((uint32_t)y << 16) | (uint32_t)(x + 4 * k + j);
        }
    }
    edge_emit(img, h, w, map, plane, s_pos, s_n, s_off, s_end, edges, ecount, dir, nbx, nby);
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
constexpr int AT = 128;                      // tile edge in accumulator cells
constexpr int AG = 2;                        // guard cells around the ring
constexpr int AS = AT + 2 + 2 * AG;          // shared rows / used columns
constexpr int AP = AS + 1;                   // shared pitch (odd: column walks are conflict free)
constexpr int VOTE_THREADS = 512;
constexpr int VOTE_SMEM = AS * AP * 4;
constexpr int VB = 7;                        // buckets per axis that can overlap a tile's region

__global__ void __launch_bounds__(VOTE_THREADS) k_vote_peaks(const uint2 *__restrict__ edges,
                                                             const int2 *__restrict__ dir, int nbx, int nby, int h,
                                                             int w, int32_t *cand, int32_t *ncand, int cand_cap)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    int *s_acc = reinterpret_cast<int *>(s_raw);                       // AS x AP
    __shared__ int s_boff[VB * VB], s_bend[VB * VB + 1];               // bucket slice start / running item end
    const size_t plane = (size_t)h * w;
    const int map = blockIdx.z;
    const uint2 *elist = edges + map * plane;
    const int tx0 = blockIdx.x * AT, ty0 = blockIdx.y * AT;      // first cell of the tile proper
    const int cx0 = tx0 - 1 - AG, cy0 = ty0 - 1 - AG;            // cell coordinates of shared (0,0)
    // cells of the ring that exist in the image: the clip box of the rays
    const int X0 = max(tx0 - 1, 0), X1 = min(tx0 + AT, w - 1);
    const int Y0 = max(ty0 - 1, 0), Y1 = min(ty0 + AT, h - 1);
    // pixels that can reach the ring, and the buckets holding them
    const int rx0 = max(tx0 - 1 - MAX_R, 0), rx1 = min(tx0 + AT + MAX_R, w - 1);
    const int ry0 = max(ty0 - 1 - MAX_R, 0), ry1 = min(ty0 + AT + MAX_R, h - 1);
    const int bx0 = rx0 / EB, bx1 = rx1 / EB, by0 = ry0 / EB, by1 = ry1 / EB;
    const int nbw = bx1 - bx0 + 1, nb = nbw * (by1 - by0 + 1);      // <= VB*VB
    for (int i = threadIdx.x; i < AS * AP; i += blockDim.x) s_acc[i] = 0;
    if (threadIdx.x < nb) {
        int b = threadIdx.x;
        int2 d = dir[((size_t)map * nby + by0 + b / nbw) * nbx + bx0 + b % nbw];
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
    const int items = s_bend[nb];                                   // one item = one edge pixel, both rays
    int b = 0;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        while (it >= s_bend[b + 1]) b++;                             // `it` only grows: b is monotone
        const uint2 e = __ldg(elist + s_boff[b] + (it - s_bend[b]));
        const int x = e.x & 0xffff, y = e.x >> 16;
        if (x < rx0 || x > rx1 || y < ry0 || y > ry1) continue;
        const int sx = (int)(short)(e.y & 0xffff), sy = (int)e.y >> 16;
        if (sx == 0 && sy == 0) continue;
        // Signed radii t (cell = pixel + t * step, t in [-30,-1] U [1,30]) whose cell lies in
        // [X0,X1] x [Y0,Y1]: one interval [lo,hi] in float, conservative by < 1 step each side.
        float lo = -(float)MAX_R, hi = (float)MAX_R;
        if (sx != 0) {
            float inv = __fdividef(1024.0f, (float)sx);
            float ta = (float)(X0 - x) * inv, tb = (float)(X1 + 1 - x) * inv;
            lo = fmaxf(lo, fminf(ta, tb)); hi = fminf(hi, fmaxf(ta, tb));
        } else if (x < X0 || x > X1) continue;
        if (sy != 0) {
            float inv = __fdividef(1024.0f, (float)sy);
            float ta = (float)(Y0 - y) * inv, tb = (float)(Y1 + 1 - y) * inv;
            lo = fmaxf(lo, fminf(ta, tb)); hi = fminf(hi, fmaxf(ta, tb));
        } else if (y < Y0 || y > Y1) continue;
        // 0.25 of slack covers the error of the approximate divide; at most one extra step per side
        const int t_lo = max(-MAX_R, (int)floorf(lo - 0.25f)), t_hi = min(MAX_R, (int)ceilf(hi + 0.25f));
        const int xb = (x - cx0) * 1024, yb = (y - cy0) * 1024;
        {   // forward ray: t = max(t_lo,1) .. t_hi
            const int r0 = max(t_lo, MIN_R);
            int x1 = xb + r0 * sx, y1 = yb + r0 * sy;
#pragma unroll 4
            for (int r = r0; r <= t_hi; r++, x1 += sx, y1 += sy)
                atomicAdd(s_acc + (y1 >> 10) * AP + (x1 >> 10), 1);
        }
        {   // backward ray: t = -1 down to t_lo, i.e. radius r = -t
            const int r0 = max(-t_hi, MIN_R), r1 = -t_lo;
            int x1 = xb - r0 * sx, y1 = yb - r0 * sy;
#pragma unroll 4
            for (int r = r0; r <= r1; r++, x1 -= sx, y1 -= sy)
                atomicAdd(s_acc + (y1 >> 10) * AP + (x1 >> 10), 1);
        }
    }
    __syncthreads();
    // cells outside the image never receive votes in the reference: clear what the conservative
    // extra steps may have left there (border tiles only)
    if (cx0 < 0 || cy0 < 0 || cx0 + AS > w || cy0 + AS > h) {
        for (int i = threadIdx.x; i < AS * AS; i += blockDim.x) {
            int ly = i / AS, lx = i - ly * AS;
            int cx = cx0 + lx, cy = cy0 + ly;
            if (cx < 0 || cy < 0 || cx >= w || cy >= h) s_acc[ly * AP + lx] = 0;
        }
        __syncthreads();
    }
    // K6: 4-neighbour peaks above the accumulator threshold, interior cells only (x,y >= 1)
    const int aw = w + 2;
    for (int idx = threadIdx.x; idx < AT * AT; idx += blockDim.x) {
        int ty = idx / AT, tx = idx - ty * AT;
        int cx = tx0 + tx, cy = ty0 + ty;
        if (cx < 1 || cy < 1 || cx >= w || cy >= h) continue;    // cells x==w / y==h never receive votes
        const int *c = s_acc + (ty + 1 + AG) * AP + tx + 1 + AG;
        int v = c[0];
        if (v > ACC_THR && v > c[-1] && v >= c[1] && v > c[-AP] && v >= c[AP]) {
            int slot = atomicAdd(ncand + map, 1);
            if (slot < cand_cap) cand[(size_t)map * cand_cap + slot] = cy * aw + cx;
        }
    }
}

// ---- second generation of the same kernel -------------------------------------------------------
// Same tile, same clipping, same atomics.  What changed: every warp owns one contiguous slice of the
// tile's item sequence (lanes interleaved), so the bucket pointer moves by a step or two per
// iteration after one binary search per warp, consecutive lanes hold consecutive contour pixels
// (similar ray lengths, neighbouring banks), and the peak scan reads rows with lane-consecutive
// columns, testing the threshold before anything else.
template <int UNROLL>
__device__ __forceinline__ void vote_item(int *s_acc, uint2 e, int cx0, int cy0, int X0, int X1, int Y0, int Y1)
{
    const int x = e.x & 0xffff, y = e.x >> 16;
    const int sx = (int)(short)(e.y & 0xffff), sy = (int)e.y >> 16;
    float lo = -(float)MAX_R, hi = (float)MAX_R;
    if (sx != 0) {
        float inv = __fdividef(1024.0f, (float)sx);
        float ta = (float)(X0 - x) * inv, tb = (float)(X1 + 1 - x) * inv;
        lo = fmaxf(lo, fminf(ta, tb)); hi = fminf(hi, fmaxf(ta, tb));
    } else if (x < X0 || x > X1) return;
    if (sy != 0) {
        float inv = __fdividef(1024.0f, (float)sy);
        float ta = (float)(Y0 - y) * inv, tb = (float)(Y1 + 1 - y) * inv;
        lo = fmaxf(lo, fminf(ta, tb)); hi = fminf(hi, fmaxf(ta, tb));
    } else if (y < Y0 || y > Y1) return;
    const int t_lo = max(-MAX_R, (int)floorf(lo - 0.25f)), t_hi = min(MAX_R, (int)ceilf(hi + 0.25f));
    if (t_lo > t_hi) return;
    // Both rays are one arithmetic sequence in the signed radius t: cell(t) = floor((p*1024 + t*step) / 1024)
    // for t > 0 (forward) and t < 0 (backward, r = -t).  One loop over t_lo..t_hi votes them all -- a warp
    // then runs max(len) instead of max(forward) + max(backward) -- and the vote the loop casts at t = 0
    // (the pixel's own cell, which the reference never votes) is taken back afterwards.
    const int xb = (x - cx0) * 1024, yb = (y - cy0) * 1024;
    int x1 = xb + t_lo * sx, y1 = yb + t_lo * sy;
#pragma unroll UNROLL
    for (int t = t_lo; t <= t_hi; t++, x1 += sx, y1 += sy) atomicAdd(s_acc + (y1 >> 10) * AP + (x1 >> 10), 1);
    if (t_lo <= 0 && t_hi >= 0) atomicAdd(s_acc + (y - cy0) * AP + (x - cx0), -1);
}

template <int UNROLL>
__global__ void __launch_bounds__(VOTE_THREADS) k_vote_peaks2(const uint2 *__restrict__ edges,
                                                              const int2 *__restrict__ dir, int nbx, int nby, int h,
                                                              int w, int32_t *cand, int32_t *ncand, int cand_cap)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    int *s_acc = reinterpret_cast<int *>(s_raw);                       // AS x AP
    __shared__ int s_boff[VB * VB], s_bend[VB * VB + 1];               // bucket slice start / running item end
    constexpr int NW = VOTE_THREADS / 32;
    const size_t plane = (size_t)h * w;
    const int map = blockIdx.z;
    const uint2 *elist = edges + map * plane;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tx0 = blockIdx.x * AT, ty0 = blockIdx.y * AT;
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
        const int per = ((items + NW - 1) / NW + 31) & ~31;          // items per warp, whole rounds of 32
        const int i0 = warp * per, i1 = min(i0 + per, items);
        int b = 0;
        if (i0 < i1) {                                               // bucket holding item i0: s_bend[b] <= i0 < s_bend[b+1]
            int lo = 0, hi = nb - 1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s_bend[mid + 1] <= i0) lo = mid + 1; else hi = mid;
            }
            b = lo;
        }
        for (int it = i0 + lane; it < i1; it += 32) {
            while (it >= s_bend[b + 1]) b++;                         // `it` only grows: b is monotone
            const uint2 e = __ldg(elist + s_boff[b] + (it - s_bend[b]));
            const int x = e.x & 0xffff, y = e.x >> 16;
            if (e.y == 0 || x < rx0 || x > rx1 || y < ry0 || y > ry1) continue;
            vote_item<UNROLL>(s_acc, e, cx0, cy0, X0, X1, Y0, Y1);
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
// the 60x60 window of the edge map around the centre instead of the whole non-zero list.
__device__ __forceinline__ float radius_of_q(int q)
{
    // (upbin + j)/2.f / nBinsPerDr * dr + minRadius, every step rounded to float32
    return __fadd_rn(__fdiv_rn(__fdiv_rn((float)q, 2.0f), 10.0f), 1.0f);
}

constexpr int RW = 8;          // warps per block
constexpr int RBINS = 320;     // NBINS padded to a multiple of 32 (pad stays zero)
constexpr int RQ = 576;        // radius table size: q = upbin + j <= 289 + 279

__global__ void __launch_bounds__(RW * 32) k_radius(const uint2 *__restrict__ edges, const int2 *__restrict__ dir,
                                                   int nbx, int nby, int h, int w,
                                                   const int32_t *__restrict__ cand, const int32_t *__restrict__ ncand,
                                                   int cand_cap, unsigned long long *est, int32_t *nest, int32_t *status,
                                                   int n_images)
{
    __shared__ int s_bins[RW][RBINS];
    __shared__ int s_pref[RW][RBINS];
    __shared__ uint32_t s_mask[RW][RBINS / 32];
    __shared__ float s_rtab[RQ];
    __shared__ uint16_t s_binlut[900];             // histogram bin of squared distance q + 0.5, q = 1..899
    const int map = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint2 *elist = edges + (size_t)map * h * w;
    const int2 *mdir = dir + (size_t)map * nbx * nby;
    const int aw = w + 2;
    int n = ncand[map];
    if (n > cand_cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(status + map % n_images, I2S_ST_CAND_OVERFLOW);
        n = cand_cap;
    }
    if (blockIdx.x * RW >= n) return;
    for (int q = threadIdx.x; q < RQ; q += blockDim.x) s_rtab[q] = radius_of_q(q);
    // Centres sit on half-integers and edge pixels on integers, so the float32 squared distance
    // (cx+.5-px)^2 + (cy+.5-py)^2 is exactly q + 0.5 with q = dx(dx+1) + dy(dy+1) an integer: the
    // sqrt / rint chain of SURVEY A.5 step 4 is tabulated once per block over the 899 admissible q.
    for (int q = threadIdx.x; q < 900; q += blockDim.x) {
        const float dd = __fsqrt_rn((float)q + 0.5f);
        const int bin = __float2int_rn(__fmul_rn(__fsub_rn(dd, 1.0f), 10.0f));
        s_binlut[q] = (uint16_t)min(max(bin, 0), NBINS - 1);
    }
    __syncthreads();
    int *bins = s_bins[warp], *pref = s_pref[warp];
    uint32_t *mask = s_mask[warp];
    for (int c = blockIdx.x * RW + warp; c < n; c += gridDim.x * RW) {
        int base = cand[(size_t)map * cand_cap + c];
        int cy = base / aw, cx = base - cy * aw;
        for (int b = lane; b < RBINS; b += 32) bins[b] = 0;
        __syncwarp();
        // histogram of the distances to the edge pixels within 30 px: walk the edge-list buckets
        // that overlap the 60x60 window (at most 3x3 of them).  (A finer 16x16 directory was tried:
        // fewer entries to reject, but cells of ~18 entries leave half a warp idle -- slower.)
        const int xlo = max(cx - 29, 0), xhi = min(cx + 30, w - 1);
        const int ylo = max(cy - 29, 0), yhi = min(cy + 30, h - 1);
        for (int by = ylo / EB; by <= yhi / EB; by++)
            for (int bx = xlo / EB; bx <= xhi / EB; bx++) {
                const int2 d = __ldg(mdir + by * nbx + bx);
                // two list entries per lane and round: the loop is bound by the latency of its loads
                for (int i = lane; i < d.y; i += 64) {
                    const uint32_t e0 = __ldg(&elist[d.x + i].x);
                    const bool two = i + 32 < d.y;
                    const uint32_t e1 = two ? __ldg(&elist[d.x + i + 32].x) : 0u;
                    {
                        const int dxi = cx - (int)(e0 & 0xffff), dyi = cy - (int)(e0 >> 16);
                        const int q = dxi * dxi + dxi + dyi * dyi + dyi;     // 1 <= r2 <= 900  <=>  1 <= q <= 899
                        if ((unsigned)(q - 1) < 899u) atomicAdd(bins + s_binlut[q], 1);
                    }
                    if (two) {
                        const int dxi = cx - (int)(e1 & 0xffff), dyi = cy - (int)(e1 >> 16);
                        const int q = dxi * dxi + dxi + dyi * dyi + dyi;
                        if ((unsigned)(q - 1) < 899u) atomicAdd(bins + s_binlut[q], 1);
                    }
                }
            }
        __syncwarp();
        // inclusive prefix sums (10 bins per lane) and the non-zero bitmap
        {
            int loc[10], sum = 0;
#pragma unroll
            for (int k = 0; k < 10; k++) { loc[k] = bins[lane * 10 + k]; sum += loc[k]; }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            int run = incl - sum;
#pragma unroll
            for (int k = 0; k < 10; k++) { run += loc[k]; pref[lane * 10 + k] = run; }
#pragma unroll
            for (int wd = 0; wd < RBINS / 32; wd++) {
                uint32_t m = __ballot_sync(0xffffffffu, bins[wd * 32 + lane] != 0);
                if (lane == 0) mask[wd] = m;
            }
        }
        __syncwarp();
        // OpenCV's scan from the top bin: every non-zero bin opens a 10-bin window, the bin just below
        // the window is skipped.  Uniform across lanes.
        int maxCount = 0, bestq = 0;
        float rBest = 0.0f;
        int j = NBINS - 1;
        while (j > 0) {
            int wd = j >> 5;
            uint32_t m = mask[wd] & (0xffffffffu >> (31 - (j & 31)));
            while (m == 0 && --wd >= 0) m = mask[wd];
            if (m == 0) break;
            int up = wd * 32 + 31 - __clz(m);
            if (up <= 0) break;
            int jn = up - 10, cur;
            if (jn >= 0) cur = pref[up] - pref[jn];
            else { cur = pref[up]; jn = -1; }
            float rCur = s_rtab[up + jn];
            if ((__fmul_rn((float)cur, rBest) >= __fmul_rn((float)maxCount, rCur)) ||
                (rBest < 1.1920929e-07f && cur >= maxCount)) {
                rBest = rCur; maxCount = cur; bestq = up + jn;
            }
            j = jn - 1;
        }
        if (lane == 0 && maxCount > ACC_THR) {
            int slot = atomicAdd(nest + map, 1);
            if (slot < cand_cap)
                est[(size_t)map * cand_cap + slot] = ((unsigned long long)(4095 - maxCount) << 38) |
                                                     ((unsigned long long)(1023 - bestq) << 28) |
                                                     ((unsigned long long)cx << 14) | (unsigned long long)cy;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ K7b: total-order sort + greedy minDist
// One block per map: bitonic sort of the packed keys (support desc, radius desc, x asc, y asc),
// then warp 0 runs the sequential suppression (kept iff >= 10 px from every kept circle).  Kept
// circles are hashed into a grid of cells at least 16 px wide, so a candidate only has to be
// compared with the chains of its 3x3 cell neighbourhood (lanes 0..8, one cell each).
__global__ void __launch_bounds__(256) k_circles_finish(const unsigned long long *__restrict__ est,
                                                        const int32_t *__restrict__ nest, int cand_cap, int np2cap,
                                                        int cshift, int cells_x, int cells_y, float *circ,
                                                        int32_t *ncirc, int circle_cap, int32_t *status, int n_images)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(s_raw);          // np2cap
    short2 *kept = reinterpret_cast<short2 *>(keys + np2cap);                          // np2cap
    uint16_t *next = reinterpret_cast<uint16_t *>(kept + np2cap);                      // np2cap
    uint16_t *head = next + np2cap;                                                    // cells_x * cells_y
    const int map = blockIdx.x;
    int n = min(nest[map], cand_cap);
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = threadIdx.x; i < np2; i += blockDim.x)
        keys[i] = i < n ? est[(size_t)map * cand_cap + i] : ~0ull;
    for (int i = threadIdx.x; i < cells_x * cells_y; i += blockDim.x) head[i] = 0xffff;
    __syncthreads();
    bitonic_sort_block(keys, np2);
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    const int ddx = lane % 3 - 1, ddy = lane / 3 - 1;      // lanes 0..8: the 3x3 cell neighbourhood
    int nk = 0;
    float *out = circ + (size_t)map * circle_cap * 3;
    for (int i = 0; i < n; i++) {
        unsigned long long k = keys[i];
        int x = (int)((k >> 14) & 0x3fff), y = (int)(k & 0x3fff);
        const int cx = x >> cshift, cy = y >> cshift;
        bool clash = false;
        if (lane < 9) {
            int ncx = cx + ddx, ncy = cy + ddy;
            if (ncx >= 0 && ncx < cells_x && ncy >= 0 && ncy < cells_y) {
                for (int j = head[ncy * cells_x + ncx]; j != 0xffff; j = next[j]) {
                    int dx = kept[j].x - x, dy = kept[j].y - y;
                    clash |= dx * dx + dy * dy < 100;
                }
            }
        }
        if (__any_sync(0xffffffffu, clash)) continue;
        if (lane == 0) {
            kept[nk] = make_short2((short)x, (short)y);
            next[nk] = head[cy * cells_x + cx];
            head[cy * cells_x + cx] = (uint16_t)nk;
            if (nk < circle_cap) {
                out[3 * nk] = (float)x + 0.5f;
                out[3 * nk + 1] = (float)y + 0.5f;
                out[3 * nk + 2] = radius_of_q(1023 - (int)((k >> 28) & 0x3ff));
            }
        }
        nk++;
        __syncwarp();
    }
    if (lane == 0) {
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

__global__ void __launch_bounds__(256) k_mask(const uint8_t *__restrict__ edges, uint8_t *__restrict__ masked, int h,
                                              int w, const float *__restrict__ circles,
                                              const int32_t *__restrict__ counts, int circle_cap,
                                              const int2 *__restrict__ dup)
{
    __shared__ short4 s_rect[KCHUNK];
    __shared__ int s_idx[KCHUNK];
    __shared__ int s_n;
    const int img = blockIdx.z;
    const size_t plane = (size_t)h * w;
    const float *circ = circles + (size_t)img * circle_cap * 3;
    const int n = min(counts[img], circle_cap);
    // circles [0, n0) and [n0 + n1, 2 n0 + n1) are repeated verbatim at [2 n0 + n1, 3 n0 + n1) (see k_stack)
    const int2 dd = dup ? dup[img] : make_int2(0, 0);
    const int skip_a = dd.x, skip_b0 = dd.x + dd.y, skip_b1 = 2 * dd.x + dd.y;
    const int tx0 = blockIdx.x * KT, ty0 = blockIdx.y * KT;
    const int tx1 = min(tx0 + KT, w) - 1, ty1 = min(ty0 + KT, h) - 1;
    int last[16];
#pragma unroll
    for (int k = 0; k < 16; k++) last[k] = -1;
    // this thread's pixels: rows (threadIdx.x/16) + 16*q, q<4 ; cols (threadIdx.x%16)*4 + k, k<4
    const int lx = (threadIdx.x & 15) * 4, ly = threadIdx.x >> 4;
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
                s_rect[s] = make_short4((short)max(x0, -32768), (short)max(y0, -32768), (short)min(x1, 32767),
                                        (short)min(y1, 32767));
                s_idx[s] = i;
            }
        }
        __syncthreads();
        const int m = s_n;
        for (int j = 0; j < m; j++) {
            short4 r = s_rect[j];
            int i = s_idx[j];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                int y = ty0 + ly + 16 * q;
                if (y < r.y || y > r.w) continue;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    int x = tx0 + lx + k;
                    if (x >= r.x && x <= r.z) last[q * 4 + k] = max(last[q * 4 + k], i);
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        int y = ty0 + ly + 16 * q;
        if (y >= h) continue;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int x = tx0 + lx + k;
            if (x >= w) continue;
            size_t o = img * plane + (size_t)y * w + x;
            int li = last[q * 4 + k];
            uint8_t v;
            if (li < 0) v = edges[o];
            else {
                int mx = __float2int_rn(circ[3 * li]), my = __float2int_rn(circ[3 * li + 1]);
                v = (abs(x - mx) + abs(y - my) <= 1) ? 255 : 0;
            }
            masked[o] = v;
        }
    }
}

// ------------------------------------------------------------------ host orchestration
size_t circles_scratch_bytes(int maps, int h, int w, const i2s_limits_t &lim)
{
    size_t plane = (size_t)h * w;
    size_t b = 0;
    b += align_up(maps * plane, 256);                               // state maps
    b += align_up(maps * plane * 8, 256);                           // edge lists (position, Q10 step), worst case
    b += align_up((size_t)maps * cdiv(w, EB) * cdiv(h, EB) * 8, 256);   // bucket directory
    b += align_up((size_t)maps * lim.cand_cap * 4, 256);            // candidate centres
    b += align_up((size_t)maps * lim.cand_cap * 8, 256);            // estimated circle keys
    b += align_up((size_t)maps * 4 * 4, 256);                       // counters
    b += align_up((size_t)maps * lim.circle_cap * 12, 256);         // per-map circles
    b += canny_scratch_bytes(maps, h, w);
    return b + 4096;
}

// HoughCircles on every map of `ms`; per-map circles [maps][circle_cap][3] + counts [maps]
int hough_circles_maps(const MapSet &ms, int h, int w, float *mcirc, int32_t *mcount, int32_t *status,
                       const i2s_limits_t &lim, Arena &ar, cudaStream_t st)
{
    const int maps = ms.count * ms.n;
    const size_t plane = (size_t)h * w;
    uint8_t *state = ar.take<uint8_t>(maps * plane);
    uint2 *edges = ar.take<uint2>(maps * plane);
    const int nbx = cdiv(w, EB), nby = cdiv(h, EB);
    int2 *dir = ar.take<int2>((size_t)maps * nbx * nby);
    int32_t *cand = ar.take<int32_t>((size_t)maps * lim.cand_cap);
    unsigned long long *est = ar.take<unsigned long long>((size_t)maps * lim.cand_cap);
    int32_t *ctr = ar.take<int32_t>((size_t)maps * 3);
    void *cscratch = ar.take<uint8_t>(canny_scratch_bytes(maps, h, w));
    if (!ar.ok()) { set_error("hough_circles: workspace too small"); return I2S_E_WORKSPACE; }
    int32_t *ncand = ctr, *nest = ctr + maps, *ecount = ctr + 2 * maps;

    int rc = canny_states(ms, 1, state, h, w, CANNY_LOW, CANNY_HIGH, lim.hyst_passes, status, cscratch, st);
    if (rc) return rc;
    I2S_CUDA(cudaMemsetAsync(ctr, 0, sizeof(int32_t) * maps * 3, st));
    bool al = (w & 3) == 0 && ((uintptr_t)state & 3) == 0;
    const bool fine = (w & 15) == 0 && ((uintptr_t)state & 15) == 0 && !legacy_enabled("edges");   // list has 16x16 sub-buckets
    {
        ScopedSection sec(SEC_EDGE_LIST, st);
        const dim3 eg(cdiv(nbx, 2), cdiv(nby, 2), maps);
        if (fine)
            k_edge_buckets16<<<eg, 256, 0, st>>>(ms, state, h, w, edges, ecount, dir, nbx, nby);
        else
            k_edge_buckets<<<eg, 256, 0, st>>>(ms, state, h, w, al, edges, ecount, dir, nbx, nby);
        I2S_CHECK_LAUNCH("k_edge_buckets");
    }
    {
        ScopedSection sec(SEC_VOTE, st);
        auto kern = legacy_enabled("vote") ? k_vote_peaks : legacy_enabled("vote4") ? k_vote_peaks2<4> : k_vote_peaks2<8>;
        I2S_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, VOTE_SMEM));
        kern<<<dim3(cdiv(w, AT), cdiv(h, AT), maps), VOTE_THREADS, VOTE_SMEM, st>>>(edges, dir, nbx, nby, h, w, cand, ncand,
                                                                                   lim.cand_cap);
        I2S_CHECK_LAUNCH("k_vote_peaks");
    }
    {
        ScopedSection sec(SEC_RADIUS, st);
        // (a variant with four centres per warp -- 8 lanes each, 16-bit bins -- cut the instruction count by
        // 18 % but not the time: the loop over the bucket entries dominates, not the per-centre scan)
        k_radius<<<dim3(16, maps), RW * 32, 0, st>>>(edges, dir, nbx, nby, h, w, cand, ncand, lim.cand_cap, est, nest, status, ms.n);
        I2S_CHECK_LAUNCH("k_radius");
    }
    ScopedSection sec(SEC_CIRCLES_FINISH, st);
    int np2 = 1;
    while (np2 < lim.cand_cap) np2 <<= 1;
    int cshift = 4;                                            // cells of >= 16 px (> minDist), at most 16384 of them
    while ((size_t)((w >> cshift) + 1) * ((h >> cshift) + 1) > 16384) cshift++;
    const int cells_x = (w >> cshift) + 1, cells_y = (h >> cshift) + 1;
    size_t smem = (size_t)np2 * (8 + 4 + 2) + (size_t)cells_x * cells_y * 2;
    I2S_CUDA(cudaFuncSetAttribute(k_circles_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_circles_finish<<<maps, 256, smem, st>>>(est, nest, lim.cand_cap, np2, cshift, cells_x, cells_y, mcirc, mcount,
                                              lim.circle_cap, status, ms.n);
    I2S_CHECK_LAUNCH("k_circles_finish");
    return I2S_OK;
}

int mask_circles(const uint8_t *edges, uint8_t *masked, int n, int h, int w, const float *circles,
                 const int32_t *counts, int circle_cap, cudaStream_t st, const int2 *dup)
{
    ScopedSection sec(SEC_MASK, st);
    k_mask<<<dim3(cdiv(w, KT), cdiv(h, KT), n), 256, 0, st>>>(edges, masked, h, w, circles, counts, circle_cap, dup);
    I2S_CHECK_LAUNCH("k_mask");
    return I2S_OK;
}

size_t find_circles_scratch_bytes(int n, int h, int w, const i2s_limits_t &lim)
{
    size_t plane = (size_t)h * w;
    size_t b = 6 * align_up((size_t)n * plane, 256);                       // the six blurred copies
    b += align_up((size_t)n * I2S_N_UNIQUE * lim.circle_cap * 12, 256);    // per-map circles
    b += align_up((size_t)n * I2S_N_UNIQUE * 4, 256);
    b += align_up((size_t)n * sizeof(int2), 256);
    return b + circles_scratch_bytes(n * I2S_N_UNIQUE, h, w, lim) + 4096;
}

int find_circles(const uint8_t *grey, const uint8_t *edges, int n, int h, int w, float *circles, int32_t *counts,
                 uint8_t *masked, int32_t *status, const i2s_limits_t &lim, Arena &ar, cudaStream_t st)
{
    const size_t plane = (size_t)h * w;
    uint8_t *blur[6];
    for (int k = 0; k < 6; k++) blur[k] = ar.take<uint8_t>((size_t)n * plane);
    const int maps = n * I2S_N_UNIQUE;
    float *mcirc = ar.take<float>((size_t)maps * lim.circle_cap * 3);
    int32_t *mcount = ar.take<int32_t>(maps);
    int2 *dup = ar.take<int2>(n);
    if (!ar.ok()) { set_error("find_circles: workspace too small"); return I2S_E_WORKSPACE; }
    // internal order: grey, edges, med3, gau3, med5, gau5, med7, gau7
    int rc;
    if ((rc = i2s_gauss357(grey, blur[1], blur[3], blur[5], n, h, w, st))) return rc;
    if ((rc = median357(grey, blur[0], blur[2], blur[4], n, h, w, st))) return rc;
    MapSet ms{};
    ms.src[0] = grey; ms.src[1] = edges;
    for (int k = 0; k < 6; k++) ms.src[2 + k] = blur[k];
    ms.count = I2S_N_UNIQUE; ms.n = n;
    if ((rc = hough_circles_maps(ms, h, w, mcirc, mcount, status, lim, ar, st))) return rc;
    {
        ScopedSection sec(SEC_STACK, st);
        k_stack<<<n, 256, 0, st>>>(mcirc, mcount, n, lim.circle_cap, circles, counts, status, dup);
        I2S_CHECK_LAUNCH("k_stack");
    }
    return mask_circles(edges, masked, n, h, w, circles, counts, lim.circle_cap, st, legacy_enabled("mask") ? nullptr : dup);
}

}  // namespace i2s

using namespace i2s;

static int check_limits(const i2s_limits_t *lim)
{
    I2S_ARG(lim && lim->cand_cap >= 32 && lim->cand_cap <= 16384 && lim->circle_cap >= 1 && lim->line_cap >= 2 &&
            lim->line_cap <= 4096 && lim->hyst_passes >= 1);
    return I2S_OK;
}

extern "C" size_t i2s_hough_circles_workspace_bytes(int n, int h, int w, const i2s_limits_t *lim)
{
    if (n <= 0 || h <= 0 || w <= 0 || !lim) return 4096;
    return circles_scratch_bytes(n, h, w, *lim);
}

extern "C" int i2s_hough_circles(const uint8_t *img, int n, int h, int w, float *circles, int32_t *counts,
                                 int32_t *status, const i2s_limits_t *lim, void *ws, size_t ws_bytes, void *stream)
{
    I2S_ARG(img && circles && counts && status && ws && n >= 0 && h > 0 && w > 0 && h < 16384 && w < 16384);
    int rc = check_limits(lim);
    if (rc) return rc;
    if (n == 0) return I2S_OK;
    Arena ar(ws, ws_bytes);
    MapSet ms = MapSet::single(img, n);
    return hough_circles_maps(ms, h, w, circles, counts, status, *lim, ar, (cudaStream_t)stream);
}

extern "C" int i2s_mask_circles(const uint8_t *edges, uint8_t *masked, int n, int h, int w, const float *circles,
                                const int32_t *counts, int circle_cap, void *stream)
{
    I2S_ARG(edges && masked && circles && counts && n >= 0 && h > 0 && w > 0 && circle_cap > 0);
    if (n == 0) return I2S_OK;
    return mask_circles(edges, masked, n, h, w, circles, counts, circle_cap, (cudaStream_t)stream, nullptr);
}

extern "C" size_t i2s_find_circles_workspace_bytes(int n, int h, int w, const i2s_limits_t *lim)
{
    if (n <= 0 || h <= 0 || w <= 0 || !lim) return 4096;
    return find_circles_scratch_bytes(n, h, w, *lim);
}

extern "C" int i2s_find_circles(const uint8_t *grey, const uint8_t *edges, int n, int h, int w, float *circles,
                                int32_t *counts, uint8_t *masked, int32_t *status, const i2s_limits_t *lim, void *ws,
                                size_t ws_bytes, void *stream)
{
    I2S_ARG(grey && edges && circles && counts && masked && status && ws && n >= 0 && h > 0 && w > 0 && h < 16384 &&
            w < 16384);
    int rc = check_limits(lim);
    if (rc) return rc;
    if (n == 0) return I2S_OK;
    Arena ar(ws, ws_bytes);
    return find_circles(grey, edges, n, h, w, circles, counts, masked, status, *lim, ar, (cudaStream_t)stream);
}
