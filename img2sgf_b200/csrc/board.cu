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

__global__ void __launch_bounds__(512) k_classify(const uint8_t *__restrict__ grey, int h, int w,
                                                  const float *__restrict__ circles, const int32_t *__restrict__ counts,
                                                  int circle_cap, const i2s_grid_t *__restrict__ grids, int black_thr,
                                                  i2s_record_t *records, double *brightness, const int32_t *status)
{
    __shared__ uint8_t s_board[BS * BS];
    __shared__ double s_mean[BS * BS];
    __shared__ double s_hc[BS], s_vc[BS];
    const int img = blockIdx.x;
    const i2s_grid_t *g = grids + img;
    i2s_record_t *rec = records + img;
    const int ncirc = min(counts[img], circle_cap);
    const bool ready = g->valid && g->hsize <= BS && g->vsize <= BS;
    uint8_t *recb = reinterpret_cast<uint8_t *>(rec);
    for (int i = threadIdx.x; i < (int)sizeof(i2s_record_t); i += blockDim.x) recb[i] = 0;
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
    for (int i = threadIdx.x; i < BS * BS; i += blockDim.x) { s_board[i] = 0; s_mean[i] = 0.0; }
    if (threadIdx.x < BS) { s_hc[threadIdx.x] = g->hcentres[threadIdx.x]; s_vc[threadIdx.x] = g->vcentres[threadIdx.x]; }
    __syncthreads();
    // validate_grid's radius filter (:441-443) then nearest-intersection snap (:504-505)
    const double lo = fmin(hspace, vspace) * 0.3, hi = fmax(hspace, vspace) * 0.65;
    const float *circ = circles + (size_t)img * circle_cap * 3;
    for (int c = threadIdx.x; c < ncirc; c += blockDim.x) {
        double r = (double)circ[3 * c + 2];
        if (!(lo < r && r < hi)) continue;
        int i = closest_index((double)circ[3 * c], s_vc, hs);
        int j = closest_index((double)circ[3 * c + 1], s_hc, vs);
        s_board[i * vs + j] = 3;
    }
    __syncthreads();
    // one warp per stone: mean of the clipped half-open window (:468-481)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint8_t *gimg = grey + (size_t)img * h * w;
    for (int cell = warp; cell < hs * vs; cell += blockDim.x >> 5) {
        if (s_board[cell] != 3) continue;
        int i = cell / vs, j = cell - i * vs;
        double x = s_vc[i], y = s_hc[j];
        int xmin = (int)rint(x - hspace / 2), xmax = (int)rint(x + hspace / 2);
        int ymin = (int)rint(y - vspace / 2), ymax = (int)rint(y + vspace / 2);
        xmin = max(0, xmin); ymin = max(0, ymin);
        xmax = min(w, xmax); ymax = min(h, ymax);
        py_slice(w, xmin, xmax);
        py_slice(h, ymin, ymax);
        const int ww = xmax - xmin, wh = ymax - ymin;
        unsigned long long sum = 0;
        if (ww > 0 && wh > 0) {
            const int tot = ww * wh;
            for (int p = lane; p < tot; p += 32) {
                int py = p / ww, px = p - py * ww;
                sum += __ldg(gimg + (size_t)(ymin + py) * w + xmin + px);
            }
        }
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
        if (lane == 0) {
            double mean = (ww > 0 && wh > 0) ? (double)sum / (double)((long long)ww * wh) : nan("");
            s_mean[cell] = mean;
            s_board[cell] = (mean <= (double)black_thr) ? 1 : 2;     // NaN -> WHITE, like the reference
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < hs * vs; c += blockDim.x) {
        int i = c / vs, j = c - i * vs;
        rec->board[i * BS + j] = s_board[c];
    }
    if (threadIdx.x == 0) {
        int k = 0, nb = 0, nw = 0;
        double *br = brightness ? brightness + (size_t)img * BS * BS : nullptr;
        for (int c = 0; c < hs * vs; c++) {
            if (!s_board[c]) continue;
            if (br) br[k] = s_mean[c];
            k++;
            if (s_board[c] == 1) nb++; else nw++;
        }
        if (br) for (; k < BS * BS; k++) br[k] = 0.0;
        rec->n_black = nb; rec->n_white = nw;
    }
}

int validate_grid(const double *centres, const int32_t *ncentres, int n, int line_cap, i2s_grid_t *grids,
                  int32_t *status, cudaStream_t st)
{
    ScopedSection sec(SEC_VALIDATE, st);
    k_validate<<<cdiv(n, 64), 64, 0, st>>>(centres, ncentres, n, line_cap, grids, status);
    I2S_CHECK_LAUNCH("k_validate");
    return I2S_OK;
}

int classify_stones(const uint8_t *grey, int n, int h, int w, const float *circles, const int32_t *counts,
                    int circle_cap, const i2s_grid_t *grids, int black_threshold, i2s_record_t *records,
                    double *brightness, const int32_t *status, cudaStream_t st)
{
    ScopedSection sec(SEC_CLASSIFY, st);
    k_classify<<<n, 512, 0, st>>>(grey, h, w, circles, counts, circle_cap, grids, black_threshold, records, brightness,
                                  status);
    I2S_CHECK_LAUNCH("k_classify");
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

extern "C" int i2s_classify_stones(const uint8_t *grey, int n, int h, int w, const float *circles, const int32_t *counts,
                                   int circle_cap, const i2s_grid_t *grids, int black_threshold, i2s_record_t *records,
                                   double *brightness, void *stream)
{
    I2S_ARG(grey && circles && counts && grids && records && n >= 0 && h > 0 && w > 0 && circle_cap > 0);
    if (n == 0) return I2S_OK;
    return classify_stones(grey, n, h, w, circles, counts, circle_cap, grids, black_threshold, records, brightness,
                           nullptr, (cudaStream_t)stream);
}
