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

// ------------------------------------------------------------------ K5: voting
// One block per 64x64 pixel tile.  Edge pixels are compacted into shared memory, then every
// (edge pixel, direction) pair is one work item walking up to 30 accumulator cells.
constexpr int VT = 64;

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

__global__ void __launch_bounds__(256) k_vote(const MapSet ms, const uint8_t *__restrict__ state,
                                              int32_t *__restrict__ acc, int h, int w, bool al)
{
    __shared__ uint32_t s_edge[VT * VT];
    __shared__ int s_n;
    const size_t plane = (size_t)h * w;
    const int map = blockIdx.z;
    const uint8_t *img = ms.plane(map, plane);
    const uint8_t *stm = state + map * plane;
    int32_t *accm = acc + (size_t)map * (h + 2) * (w + 2);
    const int aw = w + 2;
    const int x0 = blockIdx.x * VT, y0 = blockIdx.y * VT;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int idx = threadIdx.x; idx < VT * (VT / 4); idx += blockDim.x) {
        int ty = idx / (VT / 4), gx = (idx - ty * (VT / 4)) * 4;
        int y = y0 + ty, x = x0 + gx;
        if (y >= h || x >= w) continue;
        uint32_t v = 0;
        const uint8_t *p = stm + (size_t)y * w + x;
        if (al && x + 3 < w) v = __ldg(reinterpret_cast<const uint32_t *>(p));
        else
            for (int k = 0; k < 4 && x + k < w; k++) v |= (uint32_t)__ldg(p + k) << (8 * k);
        v &= 0x02020202u;
        while (v) {
            int k = (__ffs(v) - 1) >> 3;
            v &= ~(0xffu << (8 * k));
            s_edge[atomicAdd(&s_n, 1)] = ((uint32_t)y << 16) | (uint32_t)(x + k);
        }
    }
    __syncthreads();
    const int items = 2 * s_n;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        uint32_t e = s_edge[it >> 1];
        int x = e & 0xffff, y = e >> 16;
        int dx, dy;
        sobel_at(img, h, w, x, y, dx, dy);
        if (dx == 0 && dy == 0) continue;
        float vx = (float)dx, vy = (float)dy;
        float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)));
        if (mag < 1.0f) continue;
        int sx = __float2int_rn(__fdiv_rn(__fmul_rn(vx, 1024.0f), mag));
        int sy = __float2int_rn(__fdiv_rn(__fmul_rn(vy, 1024.0f), mag));
        if (it & 1) { sx = -sx; sy = -sy; }
        int x1 = x * 1024 + MIN_R * sx, y1 = y * 1024 + MIN_R * sy;
#pragma unroll 2
        for (int r = MIN_R; r <= MAX_R; r++, x1 += sx, y1 += sy) {
            int x2 = x1 >> 10, y2 = y1 >> 10;
            if ((unsigned)x2 >= (unsigned)w || (unsigned)y2 >= (unsigned)h) break;
            atomicAdd(accm + (size_t)y2 * aw + x2, 1);
        }
    }
}

// ------------------------------------------------------------------ K6: accumulator peaks
__global__ void __launch_bounds__(256) k_peaks(const int32_t *__restrict__ acc, int h, int w, int32_t *cand,
                                               int32_t *ncand, int cand_cap)
{
    const int aw = w + 2;
    const int map = blockIdx.z;
    const int32_t *a = acc + (size_t)map * (h + 2) * aw;
    int x = 1 + blockIdx.x * 64 + (threadIdx.x & 63);
    int yb = 1 + blockIdx.y * 16 + (threadIdx.x >> 6) * 4;
    if (x > w) return;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int y = yb + k;
        if (y > h) break;
        size_t base = (size_t)y * aw + x;
        int v = __ldg(a + base);
        if (v > ACC_THR && v > __ldg(a + base - 1) && v >= __ldg(a + base + 1) && v > __ldg(a + base - aw) &&
            v >= __ldg(a + base + aw)) {
            int slot = atomicAdd(ncand + map, 1);
            if (slot < cand_cap) cand[(size_t)map * cand_cap + slot] = (int32_t)base;
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

constexpr int RW = 8;   // warps per block

__global__ void __launch_bounds__(RW * 32) k_radius(const uint8_t *__restrict__ state, int h, int w,
                                                   const int32_t *__restrict__ cand, const int32_t *__restrict__ ncand,
                                                   int cand_cap, unsigned long long *est, int32_t *nest, int32_t *status,
                                                   int n_images)
{
    __shared__ int s_bins[RW][NBINS + 6];
    const int map = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t plane = (size_t)h * w;
    const uint8_t *stm = state + map * plane;
    const int aw = w + 2;
    int n = ncand[map];
    if (n > cand_cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(status + map % n_images, I2S_ST_CAND_OVERFLOW);
        n = cand_cap;
    }
    int *bins = s_bins[warp];
    for (int c = blockIdx.x * RW + warp; c < n; c += gridDim.x * RW) {
        int base = cand[(size_t)map * cand_cap + c];
        int cy = base / aw, cx = base - cy * aw;
        for (int b = lane; b < NBINS; b += 32) bins[b] = 0;
        __syncwarp();
        const float fcx = (float)cx + 0.5f, fcy = (float)cy + 0.5f;
        for (int i = lane; i < 60 * 60; i += 32) {
            int wy = i / 60, wx = i - wy * 60;
            int py = cy - 29 + wy, px = cx - 29 + wx;
            if (px < 0 || px >= w || py < 0 || py >= h) continue;
            if (!(__ldg(stm + (size_t)py * w + px) & 2)) continue;
            float ddx = fcx - (float)px, ddy = fcy - (float)py;
            float r2 = __fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy));
            if (r2 >= 1.0f && r2 <= 900.0f) {
                float d = __fsqrt_rn(r2);
                int bin = __float2int_rn(__fmul_rn(__fsub_rn(d, 1.0f), 10.0f));
                bin = min(max(bin, 0), NBINS - 1);
                atomicAdd(bins + bin, 1);
            }
        }
        __syncwarp();
        // the scan (all lanes redundantly, uniform control flow)
        int maxCount = 0, bestq = 0;
        float rBest = 0.0f;
        for (int j = NBINS - 1; j > 0; j--) {
            if (bins[j]) {
                int up = j, cur = 0;
                for (; j > up - 10 && j >= 0; j--) cur += bins[j];
                float rCur = radius_of_q(up + j);
                if ((__fmul_rn((float)cur, rBest) >= __fmul_rn((float)maxCount, rCur)) ||
                    (rBest < 1.1920929e-07f && cur >= maxCount)) {
                    rBest = rCur; maxCount = cur; bestq = up + j;
                }
            }
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
// then warp 0 runs the sequential suppression (kept iff >= 10 px from every kept circle).
__global__ void __launch_bounds__(256) k_circles_finish(const unsigned long long *__restrict__ est,
                                                        const int32_t *__restrict__ nest, int cand_cap, float *circ,
                                                        int32_t *ncirc, int circle_cap, int32_t *status, int n_images)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(s_raw);
    const int map = blockIdx.x;
    int n = min(nest[map], cand_cap);
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    short2 *kept = reinterpret_cast<short2 *>(keys + np2);
    for (int i = threadIdx.x; i < np2; i += blockDim.x)
        keys[i] = i < n ? est[(size_t)map * cand_cap + i] : ~0ull;
    __syncthreads();
    bitonic_sort_block(keys, np2);
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    int nk = 0;
    float *out = circ + (size_t)map * circle_cap * 3;
    for (int i = 0; i < n; i++) {
        unsigned long long k = keys[i];
        int x = (int)((k >> 14) & 0x3fff), y = (int)(k & 0x3fff);
        bool clash = false;
        for (int j = lane; j < nk; j += 32) {
            int dx = kept[j].x - x, dy = kept[j].y - y;
            clash |= dx * dx + dy * dy < 100;
        }
        if (__any_sync(0xffffffffu, clash)) continue;
        if (lane == 0) {
            kept[nk] = make_short2((short)x, (short)y);
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
                                               int n, int circle_cap, float *out, int32_t *counts, int32_t *status)
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
                                              const int32_t *__restrict__ counts, int circle_cap)
{
    __shared__ short4 s_rect[KCHUNK];
    __shared__ int s_idx[KCHUNK];
    __shared__ int s_n;
    const int img = blockIdx.z;
    const size_t plane = (size_t)h * w;
    const float *circ = circles + (size_t)img * circle_cap * 3;
    const int n = min(counts[img], circle_cap);
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
    size_t plane = (size_t)h * w, aplane = (size_t)(h + 2) * (w + 2);
    size_t b = 0;
    b += align_up(maps * plane, 256);                               // state maps
    b += align_up(maps * aplane * 4, 256);                          // accumulators
    b += align_up((size_t)maps * lim.cand_cap * 4, 256);            // candidate centres
    b += align_up((size_t)maps * lim.cand_cap * 8, 256);            // estimated circle keys
    b += align_up((size_t)maps * 4 * 3, 256);                       // counters
    b += align_up((size_t)maps * lim.circle_cap * 12, 256);         // per-map circles
    b += canny_scratch_bytes(maps, h, w);
    return b + 4096;
}

// HoughCircles on every map of `ms`; per-map circles [maps][circle_cap][3] + counts [maps]
int hough_circles_maps(const MapSet &ms, int h, int w, float *mcirc, int32_t *mcount, int32_t *status,
                       const i2s_limits_t &lim, Arena &ar, cudaStream_t st)
{
    const int maps = ms.count * ms.n;
    const size_t plane = (size_t)h * w, aplane = (size_t)(h + 2) * (w + 2);
    uint8_t *state = ar.take<uint8_t>(maps * plane);
    int32_t *acc = ar.take<int32_t>(maps * aplane);
    int32_t *cand = ar.take<int32_t>((size_t)maps * lim.cand_cap);
    unsigned long long *est = ar.take<unsigned long long>((size_t)maps * lim.cand_cap);
    int32_t *ctr = ar.take<int32_t>((size_t)maps * 2);
    void *cscratch = ar.take<uint8_t>(canny_scratch_bytes(maps, h, w));
    if (!ar.ok()) { set_error("hough_circles: workspace too small"); return I2S_E_WORKSPACE; }
    int32_t *ncand = ctr, *nest = ctr + maps;

    int rc = canny_states(ms, 1, state, h, w, CANNY_LOW, CANNY_HIGH, lim.hyst_passes, status, cscratch, st);
    if (rc) return rc;
    {
        ScopedSection sec(SEC_ACC_CLEAR, st);
        I2S_CUDA(cudaMemsetAsync(acc, 0, maps * aplane * sizeof(int32_t), st));
        I2S_CUDA(cudaMemsetAsync(ctr, 0, sizeof(int32_t) * maps * 2, st));
    }
    bool al = (w & 3) == 0 && ((uintptr_t)state & 3) == 0;
    {
        ScopedSection sec(SEC_VOTE, st);
        k_vote<<<dim3(cdiv(w, VT), cdiv(h, VT), maps), 256, 0, st>>>(ms, state, acc, h, w, al);
        I2S_CHECK_LAUNCH("k_vote");
    }
    {
        ScopedSection sec(SEC_PEAKS, st);
        k_peaks<<<dim3(cdiv(w, 64), cdiv(h, 16), maps), 256, 0, st>>>(acc, h, w, cand, ncand, lim.cand_cap);
        I2S_CHECK_LAUNCH("k_peaks");
    }
    {
        ScopedSection sec(SEC_RADIUS, st);
        k_radius<<<dim3(16, maps), RW * 32, 0, st>>>(state, h, w, cand, ncand, lim.cand_cap, est, nest, status, ms.n);
        I2S_CHECK_LAUNCH("k_radius");
    }
    ScopedSection sec(SEC_CIRCLES_FINISH, st);
    int np2 = 1;
    while (np2 < lim.cand_cap) np2 <<= 1;
    size_t smem = (size_t)np2 * 8 + (size_t)np2 * 4;
    I2S_CUDA(cudaFuncSetAttribute(k_circles_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_circles_finish<<<maps, 256, smem, st>>>(est, nest, lim.cand_cap, mcirc, mcount, lim.circle_cap, status, ms.n);
    I2S_CHECK_LAUNCH("k_circles_finish");
    return I2S_OK;
}

int mask_circles(const uint8_t *edges, uint8_t *masked, int n, int h, int w, const float *circles,
                 const int32_t *counts, int circle_cap, cudaStream_t st)
{
    ScopedSection sec(SEC_MASK, st);
    k_mask<<<dim3(cdiv(w, KT), cdiv(h, KT), n), 256, 0, st>>>(edges, masked, h, w, circles, counts, circle_cap);
    I2S_CHECK_LAUNCH("k_mask");
    return I2S_OK;
}

size_t find_circles_scratch_bytes(int n, int h, int w, const i2s_limits_t &lim)
{
    size_t plane = (size_t)h * w;
    size_t b = 6 * align_up((size_t)n * plane, 256);                       // the six blurred copies
    b += align_up((size_t)n * I2S_N_UNIQUE * lim.circle_cap * 12, 256);    // per-map circles
    b += align_up((size_t)n * I2S_N_UNIQUE * 4, 256);
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
    if (!ar.ok()) { set_error("find_circles: workspace too small"); return I2S_E_WORKSPACE; }
    // internal order: grey, edges, med3, gau3, med5, gau5, med7, gau7
    int rc;
    if ((rc = i2s_gauss357(grey, blur[1], blur[3], blur[5], n, h, w, st))) return rc;
    if ((rc = i2s_median(grey, blur[0], n, h, w, 3, st))) return rc;
    if ((rc = i2s_median(grey, blur[2], n, h, w, 5, st))) return rc;
    if ((rc = i2s_median(grey, blur[4], n, h, w, 7, st))) return rc;
    MapSet ms{};
    ms.src[0] = grey; ms.src[1] = edges;
    for (int k = 0; k < 6; k++) ms.src[2 + k] = blur[k];
    ms.count = I2S_N_UNIQUE; ms.n = n;
    if ((rc = hough_circles_maps(ms, h, w, mcirc, mcount, status, lim, ar, st))) return rc;
    {
        ScopedSection sec(SEC_STACK, st);
        k_stack<<<n, 256, 0, st>>>(mcirc, mcount, n, lim.circle_cap, circles, counts, status);
        I2S_CHECK_LAUNCH("k_stack");
    }
    return mask_circles(edges, masked, n, h, w, circles, counts, lim.circle_cap, st);
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
    return mask_circles(edges, masked, n, h, w, circles, counts, circle_cap, (cudaStream_t)stream);
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
