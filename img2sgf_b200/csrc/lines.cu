// lines.cu -- near-axis Hough line detection (7 angles from one read) and 1-D clustering.
// Reference call sites: find_lines img2sgf.py:230-255 (three cv.HoughLines calls: H with 3
// angles around pi/2, V1 with 2 angles at 0, V2 with 2 angles at pi; V2 rho negated, V1 rows
// before V2 rows), find_clusters_fixed_threshold + get_cluster_centres :268-292.
// Arithmetic: SURVEY.md Appendix A.7, A.8.
#include <math.h>
#include "lines.cuh"
#include "sort.cuh"
#include "profile.cuh"

namespace i2s {

constexpr int NANG = 7;                       // H: 0..2, V1: 3..4, V2: 5..6
constexpr int ACC_ROWS = 13;                  // (3+2) + (2+2) + (2+2) accumulator rows incl. zero borders
__constant__ int c_acc_row[NANG] = {1, 2, 3, 6, 7, 10, 11};

struct Trig { float c[NANG], s[NANG]; };

// cv::createTrigTable: ang starts at float(min_theta) and is incremented in float32 by float(theta)
static Trig make_trig()
{
    const double delta = M_PI / 180 * 1.0;                // angle_delta, img2sgf.py:52-53
    const float theta_f = (float)(M_PI / 180.0);
    const double mins[3] = {M_PI / 2 - delta, 0.0, M_PI - delta};
    const int cnt[3] = {3, 2, 2};
    Trig t;
    int a = 0;
    for (int call = 0; call < 3; call++) {
        float ang = (float)mins[call];
        for (int k = 0; k < cnt[call]; k++, ang += theta_f, a++) {
            t.s[a] = (float)sin((double)ang);
            t.c[a] = (float)cos((double)ang);
        }
    }
    return t;
}

__device__ __forceinline__ int rho_of(const Trig &t, int a, int j, int i)
{
    return __float2int_rn(__fadd_rn(__fmul_rn((float)j, t.c[a]), __fmul_rn((float)i, t.s[a])));
}

// ------------------------------------------------------------------ K9a: voting
// One block per 256x32 pixel tile; per-angle rho windows of the tile live in shared memory
// and are flushed with one global atomic per touched bin.
constexpr int LT_W = 256, LT_H = 32, LBINS = 320;

__global__ void __launch_bounds__(256) k_line_vote(const uint8_t *__restrict__ masked, int pitch, size_t stride,
                                                   int32_t *__restrict__ acc, size_t acc_stride, const Dims dims,
                                                   const Trig trig)
{
    __shared__ int s_acc[NANG][LBINS];
    __shared__ int s_base[NANG];
    const int img = blockIdx.z;
    const int2 wh = dims.of(img);
    const int w = wh.x, h = wh.y;
    const int x0 = blockIdx.x * LT_W, y0 = blockIdx.y * LT_H;
    if (x0 >= w || y0 >= h) return;                               // tile outside this image (ragged batch)
    const uint8_t *src = masked + img * stride;
    const bool al = ((reinterpret_cast<uintptr_t>(masked) | (uintptr_t)pitch | (uintptr_t)stride) & 3) == 0;
    const int numrho = 2 * (w + h) + 1, aw = numrho + 2, half = (numrho - 1) / 2;
    int32_t *accm = acc + (size_t)img * acc_stride;
    const int x1 = min(x0 + LT_W, w) - 1, y1 = min(y0 + LT_H, h) - 1;
    for (int i = threadIdx.x; i < NANG * LBINS; i += blockDim.x) (&s_acc[0][0])[i] = 0;
    if (threadIdx.x < NANG) {
        int a = threadIdx.x;
        int r = min(min(rho_of(trig, a, x0, y0), rho_of(trig, a, x1, y0)),
                    min(rho_of(trig, a, x0, y1), rho_of(trig, a, x1, y1)));
        s_base[a] = r;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < LT_H * (LT_W / 4); idx += blockDim.x) {
        int ty = idx / (LT_W / 4), gx = (idx - ty * (LT_W / 4)) * 4;
        int y = y0 + ty, x = x0 + gx;
        if (y >= h || x >= w) continue;
        uint32_t v = 0;
        const uint8_t *p = src + (size_t)y * pitch + x;
        if (al && x + 3 < w) v = __ldg(reinterpret_cast<const uint32_t *>(p));
        else
            for (int k = 0; k < 4 && x + k < w; k++) v |= (uint32_t)__ldg(p + k) << (8 * k);
        if (!v) continue;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (!((v >> (8 * k)) & 0xff)) continue;
#pragma unroll
            for (int a = 0; a < NANG; a++) {
                int r = rho_of(trig, a, x + k, y);
                int b = r - s_base[a];
                if ((unsigned)b < (unsigned)LBINS) atomicAdd(&s_acc[a][b], 1);
                else atomicAdd(accm + (size_t)c_acc_row[a] * aw + r + half + 1, 1);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NANG * LBINS; i += blockDim.x) {
        int a = i / LBINS, b = i - a * LBINS;
        int v = s_acc[a][b];
        if (v) atomicAdd(accm + (size_t)c_acc_row[a] * aw + s_base[a] + b + half + 1, v);
    }
}

// ------------------------------------------------------------------ K9b: peaks + sort
// grid (2, n): direction 0 = the H call, direction 1 = V1 then V2 (rho negated), as find_lines
// stacks them.  Keys: (votes desc, OpenCV linear index asc).
__device__ void line_call(const int32_t *__restrict__ acc_call, int na, int numrho, int thr, unsigned long long *keys,
                          int cap_p2, int line_cap, int *s_cnt, float sign, float *out, int &out_n, bool &overflow)
{
    const int aw = numrho + 2;
    if (threadIdx.x == 0) *s_cnt = 0;
    __syncthreads();
    for (int idx = threadIdx.x; idx < numrho * na; idx += blockDim.x) {
        int r = idx / na, a = idx - r * na;
        int base = (a + 1) * aw + r + 1;
        int v = __ldg(acc_call + base);
        if (v > thr && v > __ldg(acc_call + base - 1) && v >= __ldg(acc_call + base + 1) &&
            v > __ldg(acc_call + base - aw) && v >= __ldg(acc_call + base + aw)) {
            int s = atomicAdd(s_cnt, 1);
            if (s < cap_p2) keys[s] = ((unsigned long long)(0xffffffffu - (unsigned)v) << 32) | (unsigned)base;
        }
    }
    __syncthreads();
    int cnt = *s_cnt;
    if (cnt > line_cap) { overflow = true; cnt = line_cap; }
    int np2 = 1;
    while (np2 < cnt) np2 <<= 1;
    for (int i = cnt + threadIdx.x; i < np2; i += blockDim.x) keys[i] = ~0ull;
    __syncthreads();
    bitonic_sort_block(keys, np2);
    const float half = (float)(numrho - 1) * 0.5f;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        int base = (int)(keys[i] & 0xffffffffu);
        int r = base % aw - 1;
        if (out_n + i < line_cap) out[out_n + i] = sign * ((float)r - half);
    }
    out_n += cnt;
    __syncthreads();
}

// threshold.get() of image i: its own value, else the batch's, else choose_threshold() (img2sgf.py:606-613)
__device__ __forceinline__ int line_threshold_of(const Dims &dims, int i, int thr, int w, int h)
{
    if (dims.images && dims.images[i].line_threshold > 0) return dims.images[i].line_threshold;
    if (thr > 0) return thr;
    const int t = (int)((double)min(w, h) / 12.8 + 16.0);
    return min(max(t, 20), 200);
}

__global__ void __launch_bounds__(256) k_line_peaks(const int32_t *__restrict__ acc, size_t acc_stride, const Dims dims,
                                                    int thr_all, float *rho, int32_t *counts, int line_cap, int cap_p2,
                                                    int32_t *status)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(s_raw);
    __shared__ int s_cnt;
    const int dir = blockIdx.x, img = blockIdx.y;
    const int2 wh = dims.of(img);
    const int w = wh.x, h = wh.y;
    const int thr = line_threshold_of(dims, img, thr_all, w, h);
    const int numrho = 2 * (w + h) + 1, aw = numrho + 2;
    const int32_t *accm = acc + (size_t)img * acc_stride;
    float *out = rho + ((size_t)img * 2 + dir) * line_cap;
    int out_n = 0;
    bool overflow = false;
    if (dir == 0) {
        line_call(accm, 3, numrho, thr, keys, cap_p2, line_cap, &s_cnt, 1.0f, out, out_n, overflow);
    } else {
        line_call(accm + (size_t)5 * aw, 2, numrho, thr, keys, cap_p2, line_cap, &s_cnt, 1.0f, out, out_n, overflow);
        line_call(accm + (size_t)9 * aw, 2, numrho, thr, keys, cap_p2, line_cap, &s_cnt, -1.0f, out, out_n, overflow);
    }
    if (threadIdx.x == 0) {
        if (out_n > line_cap) overflow = true;
        counts[img * 2 + dir] = min(out_n, line_cap);
        if (overflow) atomicOr(status + img, I2S_ST_LINE_OVERFLOW);
    }
}

// ------------------------------------------------------------------ K10: 1-D clustering
// Single linkage at distance 10 on sorted 1-D data == split where the gap is >= 10.
__global__ void __launch_bounds__(256) k_cluster(const float *__restrict__ rho, const int32_t *__restrict__ counts,
                                                 int line_cap, int cap_p2, double *centres, int32_t *ncentres)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    float *v = reinterpret_cast<float *>(s_raw);
    const int dir = blockIdx.x, img = blockIdx.y;
    const int slot = img * 2 + dir;
    const int n = min(counts[slot], line_cap);
    double *out = centres + (size_t)slot * line_cap;
    if (n < 2) {               // AgglomerativeClustering.fit raises -> [] (img2sgf.py:273-278)
        if (threadIdx.x == 0) ncentres[slot] = 0;
        return;
    }
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    const float *src = rho + (size_t)slot * line_cap;
    for (int i = threadIdx.x; i < np2; i += blockDim.x) v[i] = i < n ? src[i] : INFINITY;
    __syncthreads();
    bitonic_sort_block(v, np2);
    if (threadIdx.x == 0) {
        int k = 0, start = 0;
        float sum = v[0];
        for (int i = 1; i <= n; i++) {
            if (i == n || __fsub_rn(v[i], v[i - 1]) >= 10.0f) {
                out[k++] = (double)__fdiv_rn(sum, (float)(i - start));
                start = i;
                sum = 0.0f;
            }
            if (i < n) sum = __fadd_rn(sum, v[i]);
        }
        ncentres[slot] = k;
    }
}

size_t lines_scratch_bytes(int n, int h, int w)
{
    size_t aw = 2 * (size_t)(w + h) + 3;
    return align_up((size_t)n * ACC_ROWS * aw * 4, 256) + 1024;
}

static int p2_of(int v)
{
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// masked: [n] planes of `pitch` bytes per row, `stride` bytes apart; threshold 0 = per-image value
// (i2s_image_t.line_threshold, else choose_threshold())
int find_lines(const uint8_t *masked, const Dims &dims, int n, int pitch, size_t stride, int threshold, float *rho,
               int32_t *counts, int line_cap, int32_t *status, Arena &ar, cudaStream_t st)
{
    const int h = dims.h, w = dims.w;
    const size_t acc_stride = ACC_ROWS * (2 * (size_t)(w + h) + 3);      // sized for the canvas; every image uses its own width
    int32_t *acc = ar.take<int32_t>((size_t)n * acc_stride);
    if (!ar.ok()) { set_error("find_lines: workspace too small"); return I2S_E_WORKSPACE; }
    I2S_ARG(n < 65536);
    I2S_CUDA(cudaMemsetAsync(acc, 0, (size_t)n * acc_stride * 4, st));
    const Trig trig = make_trig();
    {
        ScopedSection sec(SEC_LINE_VOTE, st);
        k_line_vote<<<dim3(cdiv(w, LT_W), cdiv(h, LT_H), n), 256, 0, st>>>(masked, pitch, stride, acc, acc_stride, dims, trig);
        I2S_CHECK_LAUNCH("k_line_vote");
    }
    ScopedSection sec(SEC_LINE_PEAKS, st);
    int cap_p2 = p2_of(line_cap);
    size_t smem = (size_t)cap_p2 * 8;
    I2S_CUDA(cudaFuncSetAttribute(k_line_peaks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_line_peaks<<<dim3(2, n), 256, smem, st>>>(acc, acc_stride, dims, threshold, rho, counts, line_cap, cap_p2, status);
    I2S_CHECK_LAUNCH("k_line_peaks");
    return I2S_OK;
}

int cluster(const float *rho, const int32_t *counts, int n, int line_cap, double *centres, int32_t *ncentres,
            cudaStream_t st)
{
    int cap_p2 = p2_of(line_cap);
    ScopedSection sec(SEC_CLUSTER, st);
    k_cluster<<<dim3(2, n), 256, (size_t)cap_p2 * 4, st>>>(rho, counts, line_cap, cap_p2, centres, ncentres);
    I2S_CHECK_LAUNCH("k_cluster");
    return I2S_OK;
}

}  // namespace i2s

using namespace i2s;

extern "C" size_t i2s_find_lines_workspace_bytes(int n, int h, int w)
{
    if (n <= 0 || h <= 0 || w <= 0) return 1024;
    return lines_scratch_bytes(n, h, w);
}

extern "C" int i2s_find_lines(const uint8_t *masked, int pitch, int n, int h, int w, int threshold, float *rho,
                              int32_t *counts, int line_cap, int32_t *status, void *ws, size_t ws_bytes, void *stream)
{
    I2S_ARG(masked && rho && counts && status && ws && n >= 0 && h > 0 && w > 0 && line_cap >= 2 && line_cap <= 4096 &&
            threshold >= 0);
    if (pitch == 0) pitch = w;
    I2S_ARG(pitch >= w);
    if (n == 0) return I2S_OK;
    Arena ar(ws, ws_bytes);
    return find_lines(masked, Dims::uniform(h, w), n, pitch, (size_t)h * pitch, threshold, rho, counts, line_cap, status, ar,
                      (cudaStream_t)stream);
}

extern "C" int i2s_cluster(const float *rho, const int32_t *counts, int n, int line_cap, double *centres,
                           int32_t *ncentres, void *stream)
{
    I2S_ARG(rho && counts && centres && ncentres && n >= 0 && line_cap >= 2 && line_cap <= 4096);
    if (n == 0) return I2S_OK;
    return cluster(rho, counts, n, line_cap, centres, ncentres, (cudaStream_t)stream);
}
