// board.cu -- grid validation/completion and per-intersection stone classification.
// Reference call sites: complete_grid img2sgf.py:335-397, truncate_grid :400-417,
// validate_grid :420-445, closest_index :448-459, average_intensity :468-481,
// identify_board :497-515,537-543.  All float64 arithmetic with Python round-half-even;
// compiled with -fmad=false so every operation is a separately rounded IEEE op.
#include "board.cuh"
#include "profile.cuh"

namespace i2s {

constexpr int BS = I2S_BOARD_SIZE;
constexpr double MIN_GRID_SPACING = 10.0;   // img2sgf.py:54
constexpr double BIG_SPACE_RATIO = 1.6;     // img2sgf.py:55

struct View { const double *p; int n; };    // n < 0 <=> None

__device__ View truncate_view(View v)
{
    if (v.n < 0) return v;
    if (v.n == BS + 2) return View{v.p + 1, v.n - 2};
    if (v.n == BS + 1) return View{v.p, v.n - 1};
    return v;
}

// complete_grid: returns the input view when no gap needs filling, a view onto `buf`
// (<= BS+3 entries) when gaps were filled, n = -1 for None.
__device__ View complete_grid(View x, double *buf)
{
    if (x.n <= 1) return View{nullptr, -1};
    double min_space = INFINITY;
    for (int i = 0; i + 1 < x.n; i++) min_space = fmin(min_space, x.p[i + 1] - x.p[i]);
    if (min_space < MIN_GRID_SPACING) return View{nullptr, -1};
    const double bound = min_space * BIG_SPACE_RATIO;
    int nbig = 0, nsmall = 0;
    double max_space = -INFINITY;
    for (int i = 0; i + 1 < x.n; i++) {
        double s = x.p[i + 1] - x.p[i];
        if (s > bound) nbig++;
        else { nsmall++; max_space = fmax(max_space, s); }
    }
    if (nbig == 0) return x;
    const double average_space = (min_space + max_space) / 2;
    int total = nsmall;
    for (int i = 0; i + 1 < x.n; i++) {
        double s = x.p[i + 1] - x.p[i];
        if (s > bound) {
            total += (int)rint(s / average_space);
            if (total > BS + 2) return View{nullptr, -1};
        }
    }
    if (total > BS + 2) return View{nullptr, -1};
    total += 1;
    if (x.n >= total) return x;
    int i = 1, j = 1;
    buf[0] = x.p[0];
    for (int q = 0; q + 1 < x.n; q++) {
        double s = x.p[q + 1] - x.p[q];
        if (s <= max_space) { buf[i++] = x.p[j++]; }
        else {
            int m = (int)rint(s / average_space);
            for (int k = 0; k < m; k++) buf[i++] = x.p[j - 1] + (double)(k + 1) * s / (double)m;
            j++;
        }
    }
    return View{buf, total};
}

__global__ void k_validate(const double *__restrict__ centres, const int32_t *__restrict__ ncentres, int n,
                           int line_cap, i2s_grid_t *grids, int32_t *status)
{
    int img = blockIdx.x * blockDim.x + threadIdx.x;
    if (img >= n) return;
    i2s_grid_t *g = grids + img;
    g->valid = 0; g->hsize = 0; g->vsize = 0; g->pad_ = 0; g->hspace = 0; g->vspace = 0;
    for (int k = 0; k < I2S_MAX_GRID; k++) { g->hcentres[k] = 0; g->vcentres[k] = 0; }
    double bh[BS + 4], bv[BS + 4];
    View hv = truncate_view(complete_grid(truncate_view(View{centres + (size_t)(img * 2) * line_cap, ncentres[img * 2]}), bh));
    if (hv.n < 0) return;
    View vv = truncate_view(complete_grid(truncate_view(View{centres + (size_t)(img * 2 + 1) * line_cap, ncentres[img * 2 + 1]}), bv));
    if (vv.n < 0) return;
    g->valid = 1;
    g->vsize = hv.n;                       // number of horizontal lines (img2sgf.py:435)
    g->hsize = vv.n;
    g->hspace = (hv.p[hv.n - 1] - hv.p[0]) / (double)hv.n;
    g->vspace = (vv.p[vv.n - 1] - vv.p[0]) / (double)vv.n;
    for (int k = 0; k < hv.n && k < I2S_MAX_GRID; k++) g->hcentres[k] = hv.p[k];
    for (int k = 0; k < vv.n && k < I2S_MAX_GRID; k++) g->vcentres[k] = vv.p[k];
    if (hv.n > I2S_MAX_GRID || vv.n > I2S_MAX_GRID) atomicOr(status + img, I2S_ST_GRID_OVERFLOW);
}

__device__ __forceinline__ int closest_index(double a, const double *x, int n)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (x[mid] < a) lo = mid + 1; else hi = mid;
    }
    if (lo == 0) return 0;
    if (lo == n) return n - 1;
    return (a - x[lo - 1] <= x[lo] - a) ? lo - 1 : lo;
}

// numpy slice semantics for a[lo:hi] with lo >= 0 already clipped
__device__ __forceinline__ void py_slice(int len, int &lo, int &hi)
{
    if (hi < 0) { hi += len; if (hi < 0) hi = 0; }
    if (hi > len) hi = len;
    if (lo > len) lo = len;
}

// identify_board in three small kernels so that the window means -- the only real work -- spread over
// the whole machine instead of one block per image:
//   k_classify_mark : per image, record header + STONE marks from the radius-filtered circles (:441-443, :504-505)
//   k_classify_mean : one warp per (image, intersection), mean of the clipped half-open window (:468-481)
//   k_classify_count: per image, stone counts and the brightness list in (i, j) scan order (:506-515)
__global__ void __launch_bounds__(256) k_classify_mark(const float *__restrict__ circles, const int32_t *__restrict__ counts,
                                                       int circle_cap, const i2s_grid_t *__restrict__ grids,
                                                       i2s_record_t *records, const int32_t *status)
{
    __shared__ double s_hc[BS], s_vc[BS];
    const int img = blockIdx.x;
    const i2s_grid_t *g = grids + img;
    i2s_record_t *rec = records + img;
    const int ncirc = min(counts[img], circle_cap);
    const bool ready = g->valid && g->hsize <= BS && g->vsize <= BS;
    uint32_t *recw = reinterpret_cast<uint32_t *>(rec);
    for (int i = threadIdx.x; i < (int)sizeof(i2s_record_t) / 4; i += blockDim.x) recw[i] = 0;
    if (ready && threadIdx.x < BS) { s_hc[threadIdx.x] = g->hcentres[threadIdx.x]; s_vc[threadIdx.x] = g->vcentres[threadIdx.x]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        rec->valid = (uint8_t)(g->valid != 0);
        rec->board_ready = (uint8_t)ready;
        rec->hsize = (uint8_t)min(g->hsize, 255);
        rec->vsize = (uint8_t)min(g->vsize, 255);
        rec->n_circles = counts[img];
        rec->status = status ? status[img] : 0;
    }
    if (!ready) return;
    const int hs = g->hsize, vs = g->vsize;
    const double hspace = g->hspace, vspace = g->vspace;
    // validate_grid's radius filter (:441-443) then nearest-intersection snap (:504-505)
    const double lo = fmin(hspace, vspace) * 0.3, hi = fmax(hspace, vspace) * 0.65;
    const float *circ = circles + (size_t)img * circle_cap * 3;
    for (int c = threadIdx.x; c < ncirc; c += blockDim.x) {
        double r = (double)circ[3 * c + 2];
        if (!(lo < r && r < hi)) continue;
        int i = closest_index((double)circ[3 * c], s_vc, hs);
        int j = closest_index((double)circ[3 * c + 1], s_hc, vs);
        rec->board[i * BS + j] = 3;
    }
}

constexpr int CM_WARPS = 8;

__global__ void __launch_bounds__(CM_WARPS * 32) k_classify_mean(const uint8_t *__restrict__ grey, int pitch, size_t stride,
                                                                const Dims dims, int n, const i2s_grid_t *__restrict__ grids,
                                                                int black_thr, i2s_record_t *records, double *brightness)
{
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * CM_WARPS + (threadIdx.x >> 5);          // (image, intersection)
    if (item >= n * BS * BS) return;
    const int img = item / (BS * BS), cell = item - img * (BS * BS);
    i2s_record_t *rec = records + img;
    if (rec->board[cell] != 3) return;                                    // warp-uniform
    const int i = cell / BS, j = cell - i * BS;
    const i2s_grid_t *g = grids + img;
    const int2 wh = dims.of(img);
    const int w = wh.x, h = wh.y;
    const double hspace = g->hspace, vspace = g->vspace;
    const double x = g->vcentres[i], y = g->hcentres[j];
    int xmin = (int)rint(x - hspace / 2), xmax = (int)rint(x + hspace / 2);
    int ymin = (int)rint(y - vspace / 2), ymax = (int)rint(y + vspace / 2);
    xmin = max(0, xmin); ymin = max(0, ymin);
    xmax = min(w, xmax); ymax = min(h, ymax);
    py_slice(w, xmin, xmax);
    py_slice(h, ymin, ymax);
    const int ww = xmax - xmin, wh_ = ymax - ymin;
    unsigned long long sum64 = 0;
    const uint8_t *plane = grey + img * stride;
    if (ww > 0 && wh_ > 0 && (((uintptr_t)plane | (uint32_t)pitch) & 3u) == 0) {
        // aligned 32-bit words, four pixels per dot product; the bytes of the first / last word outside
        // [xmin, xmax) are masked off.  The warp is cut into rows of `lpr` lanes (the power of two that
        // covers a window row) so that several rows are in flight per iteration.
        const int xa = xmin & ~3, nwords = (xmax - xa + 3) >> 2;
        int lsh = 5;
        while (lsh > 0 && (16 >> (5 - lsh)) >= nwords) lsh--;              // lanes per row = 1 << lsh >= min(nwords, 32)
        const int lpr = 1 << lsh, k0 = lane & (lpr - 1), rstep = 32 >> lsh;
        const uint8_t *base = plane + (size_t)ymin * pitch + xa;
#pragma unroll 4
        for (int py = lane >> lsh; py < wh_; py += rstep) {
            const uint32_t *row = reinterpret_cast<const uint32_t *>(base + (size_t)py * pitch);
            unsigned int sum = 0;
            for (int k = k0; k < nwords; k += lpr) {
                uint32_t v = __ldg(row + k);
                if (k == 0) v &= 0xffffffffu << (8 * (xmin - xa));
                const int nb = xmax - (xa + 4 * k);
                if (nb < 4) v &= (1u << (8 * nb)) - 1u;
                sum = __dp4a(v, 0x01010101u, sum);
            }
            sum64 += sum;
        }
    } else if (ww > 0 && wh_ > 0) {
        const uint8_t *base = plane + (size_t)ymin * pitch + xmin;
        for (int py = 0; py < wh_; py++) {
            const uint8_t *row = base + (size_t)py * pitch;
            unsigned int sum = 0;
            for (int px = lane; px < ww; px += 32) sum += __ldg(row + px);
            sum64 += sum;
        }
    }
    for (int o = 16; o > 0; o >>= 1) sum64 += __shfl_down_sync(0xffffffffu, sum64, o);
    if (lane == 0) {
        const double mean = (ww > 0 && wh_ > 0) ? (double)sum64 / (double)((long long)ww * wh_) : nan("");
        if (brightness) brightness[(size_t)img * BS * BS + cell] = mean;
        rec->board[cell] = (mean <= (double)black_thr) ? 1 : 2;          // NaN -> WHITE, like the reference
    }
}

__global__ void __launch_bounds__(32) k_classify_count(i2s_record_t *records, double *brightness)
{
    const int img = blockIdx.x, lane = threadIdx.x;
    i2s_record_t *rec = records + img;
    if (!rec->board_ready) {
        if (brightness)
            for (int c = lane; c < BS * BS; c += 32) brightness[(size_t)img * BS * BS + c] = 0.0;
        return;
    }
    // cells in (i, j) scan order = ascending board index; ballot keeps the order
    int nb = 0, nw = 0, k = 0;
    double *br = brightness ? brightness + (size_t)img * BS * BS : nullptr;
    for (int c0 = 0; c0 < BS * BS; c0 += 32) {
        const int c = c0 + lane;
        const int v = c < BS * BS ? rec->board[c] : 0;
        const double m = (br && v) ? br[c] : 0.0;
        const uint32_t stones = __ballot_sync(0xffffffffu, v != 0);
        nb += __popc(__ballot_sync(0xffffffffu, v == 1));
        nw += __popc(__ballot_sync(0xffffffffu, v == 2));
        __syncwarp();
        // forward compaction in place: the destination index never exceeds the source index
        if (br && v) br[k + __popc(stones & ((1u << lane) - 1u))] = m;
        k += __popc(stones);
        __syncwarp();
    }
    if (br)
        for (int c = k + lane; c < BS * BS; c += 32) br[c] = 0.0;
    if (lane == 0) { rec->n_black = nb; rec->n_white = nw; }
}

int validate_grid(const double *centres, const int32_t *ncentres, int n, int line_cap, i2s_grid_t *grids,
                  int32_t *status, cudaStream_t st)
{
    ScopedSection sec(SEC_VALIDATE, st);
    k_validate<<<cdiv(n, 64), 64, 0, st>>>(centres, ncentres, n, line_cap, grids, status);
    I2S_CHECK_LAUNCH("k_validate");
    return I2S_OK;
}

int classify_stones(const uint8_t *grey, const Dims &dims, int n, int pitch, size_t stride, const float *circles,
                    const int32_t *counts, int circle_cap, const i2s_grid_t *grids, int black_threshold,
                    i2s_record_t *records, double *brightness, const int32_t *status, cudaStream_t st)
{
    ScopedSection sec(SEC_CLASSIFY, st);
    k_classify_mark<<<n, 256, 0, st>>>(circles, counts, circle_cap, grids, records, status);
    I2S_CHECK_LAUNCH("k_classify_mark");
    k_classify_mean<<<cdiv(n * BS * BS, CM_WARPS), CM_WARPS * 32, 0, st>>>(grey, pitch, stride, dims, n, grids, black_threshold,
                                                                         records, brightness);
    I2S_CHECK_LAUNCH("k_classify_mean");
    k_classify_count<<<n, 32, 0, st>>>(records, brightness);
    I2S_CHECK_LAUNCH("k_classify_count");
    return I2S_OK;
}

}  // namespace i2s

using namespace i2s;

extern "C" int i2s_validate_grid(const double *centres, const int32_t *ncentres, int n, int line_cap, i2s_grid_t *grids,
                                 int32_t *status, void *stream)
{
    I2S_ARG(centres && ncentres && grids && status && n >= 0 && line_cap >= 2);
    if (n == 0) return I2S_OK;
    return validate_grid(centres, ncentres, n, line_cap, grids, status, (cudaStream_t)stream);
}

extern "C" int i2s_classify_stones(const uint8_t *grey, int pitch, int n, int h, int w, const float *circles,
                                   const int32_t *counts, int circle_cap, const i2s_grid_t *grids, int black_threshold,
                                   i2s_record_t *records, double *brightness, void *stream)
{
    I2S_ARG(grey && circles && counts && grids && records && n >= 0 && h > 0 && w > 0 && circle_cap > 0);
    if (pitch == 0) pitch = w;
    I2S_ARG(pitch >= w);
    if (n == 0) return I2S_OK;
    return classify_stones(grey, Dims::uniform(h, w), n, pitch, (size_t)h * pitch, circles, counts, circle_cap, grids,
                           black_threshold, records, brightness, nullptr, (cudaStream_t)stream);
}
