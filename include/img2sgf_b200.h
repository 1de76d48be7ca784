/*
 * img2sgf_b200.h -- C ABI of the B200-native diagram-recognition hot path.
 *
 * The reference (hanysz/img2sgf, /root/reference/img2sgf.py) has no FFI: its hot path is a
 * set of module-level Python functions that call cv2/sklearn/numpy and exchange state
 * through globals.  Each entry point below replaces one of those call sites;
 * the citation on each is the reference line range whose result it reproduces bit-exactly.
 * The Python host layer (img2sgf_b200/api.py) binds these with ctypes and mirrors the
 * reference's function names/return conventions; INTEGRATION.md shows the binding a
 * maintainer would add to img2sgf.py.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - images are batches of n tightly packed planes, u8, [n][h][w] (or [n][h][w][3] rgb);
 *   - outputs are caller-allocated with explicit capacities; counts are written on device;
 *   - `ws` is a caller-allocated device workspace of at least the size reported by the
 *     matching *_workspace_bytes() call; nothing is allocated inside;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no host
 *     synchronisation happens inside unless stated;
 *   - return value: 0 = enqueued OK, negative = I2S_E_* (i2s_last_error() gives the text);
 *   - data-dependent failures (a capacity exceeded, hysteresis not converged within the
 *     pass budget) cannot be known at enqueue time: they are reported in a device-side
 *     status word (one int32 per image, I2S_ST_* bits) that the caller reads back with
 *     its results.  A non-zero status means "outputs invalid, retry with larger limits".
 */
#ifndef IMG2SGF_B200_H
#define IMG2SGF_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define I2S_BOARD_SIZE 19          /* img2sgf.py:43  */
#define I2S_MAX_GRID 32            /* grid lines kept per axis in i2s_grid_t */
#define I2S_N_CALLS 10             /* HoughCircles calls per image, img2sgf.py:171-186 */
#define I2S_N_UNIQUE 8             /* distinct inputs among them (grey x3 are identical) */

/* enqueue-time errors */
#define I2S_OK 0
#define I2S_E_BADARG (-1)
#define I2S_E_WORKSPACE (-2)
#define I2S_E_CUDA (-3)

/* device-side per-image status bits */
#define I2S_ST_CAND_OVERFLOW 1     /* more accumulator peaks than cand_cap           */
#define I2S_ST_CIRCLE_OVERFLOW 2   /* more circles than the output capacity          */
#define I2S_ST_LINE_OVERFLOW 4     /* more line peaks than line_cap                  */
#define I2S_ST_HYST_NOT_CONVERGED 8/* Canny hysteresis needs more passes             */
#define I2S_ST_GRID_OVERFLOW 16    /* > I2S_MAX_GRID lines on an axis (board not ready anyway) */

/* limits used by the entry points (all data-dependent sizes) */
typedef struct {
    int32_t cand_cap;      /* accumulator peaks per HoughCircles call                   */
    int32_t circle_cap;    /* circles per image in the stacked output (rows of 3 floats) */
    int32_t line_cap;      /* line peaks per direction per image                         */
    int32_t hyst_passes;   /* hysteresis pass budget (cross-tile propagation rounds)     */
} i2s_limits_t;

/* validate_grid() result, img2sgf.py:420-445 (names follow the reference: hsize is the
 * number of VERTICAL lines, hcentres are Y coordinates of horizontal lines, :435-438) */
typedef struct {
    int32_t valid;
    int32_t hsize, vsize;
    int32_t pad_;
    double hspace, vspace;
    double hcentres[I2S_MAX_GRID];   /* hcentres_complete */
    double vcentres[I2S_MAX_GRID];   /* vcentres_complete */
} i2s_grid_t;

/* fixed-size per-image result record (the only thing exchanged between GPUs) */
typedef struct {
    uint8_t board[I2S_BOARD_SIZE * I2S_BOARD_SIZE]; /* full_board, [i*19+j], i = x index, j = y index; 0/1/2 */
    uint8_t valid;         /* validate_grid()[0]                                 */
    uint8_t board_ready;   /* valid && hsize<=19 && vsize<=19 (img2sgf.py:568-574) */
    uint8_t hsize, vsize;
    uint8_t pad_[3];
    int32_t n_black, n_white, n_circles;
    int32_t status;        /* I2S_ST_* bits; 0 = results valid */
} i2s_record_t;            /* 384 bytes */

const char *i2s_last_error(void);
int i2s_version(void);
void i2s_default_limits(i2s_limits_t *lim);

/* cv.cvtColor(rgb, COLOR_BGR2GRAY) -- img2sgf.py:153 */
int i2s_grey(const uint8_t *rgb, uint8_t *grey, int n, int h, int w, void *stream);

/* ImageEnhance.Contrast(..).enhance(f) -- img2sgf.py:142-144 (prologue).  `scratch8n` is a
 * device scratch of n * 8 bytes (the per-image luma sums). */
int i2s_contrast(const uint8_t *rgb, uint8_t *out, void *scratch8n, int n, int h, int w,
                 double factor, void *stream);

/* cv.GaussianBlur(grey,(b,b),b), b in {3,5,7} -- img2sgf.py:175.  Writes all three in one
 * pass over the input: dst3/dst5/dst7 each [n][h][w] (any may be NULL). */
int i2s_gauss357(const uint8_t *src, uint8_t *dst3, uint8_t *dst5, uint8_t *dst7, int n, int h,
                 int w, void *stream);

/* cv.medianBlur(grey,b), b in {3,5,7} -- img2sgf.py:174 */
int i2s_median(const uint8_t *src, uint8_t *dst, int n, int h, int w, int b, void *stream);

/* cv.Canny(rgb, low, high, apertureSize=3, L2gradient=False) -- img2sgf.py:162-165.
 * channels = 3 for the reference call; channels = 1 gives the single-channel Canny that
 * cv.HoughCircles runs internally.  status: n int32 (OR-ed with I2S_ST_* bits). */
size_t i2s_canny_workspace_bytes(int n, int h, int w);
int i2s_canny(const uint8_t *img, int channels, uint8_t *edges, int n, int h, int w, int low,
              int high, int hyst_passes, int32_t *status, void *ws, size_t ws_bytes, void *stream);

/* cv.HoughCircles(img, HOUGH_GRADIENT, 1, 10, [], 100, 30, 1, 30) -- img2sgf.py:180.
 * circles: [n][circle_cap][3] float32 (x,y,r) in OpenCV's output order; counts: [n]. */
size_t i2s_hough_circles_workspace_bytes(int n, int h, int w, const i2s_limits_t *lim);
int i2s_hough_circles(const uint8_t *img, int n, int h, int w, float *circles, int32_t *counts,
                      int32_t *status, const i2s_limits_t *lim, void *ws, size_t ws_bytes,
                      void *stream);

/* the masking loop -- img2sgf.py:169,191-198.  masked may alias edges. */
int i2s_mask_circles(const uint8_t *edges, uint8_t *masked, int n, int h, int w,
                     const float *circles, const int32_t *counts, int circle_cap, void *stream);

/* find_circles: blur pyramid + ten HoughCircles calls stacked in `blurs` order + mask
 * -- img2sgf.py:169-198.  circles [n][circle_cap][3], counts [n], masked [n][h][w]. */
size_t i2s_find_circles_workspace_bytes(int n, int h, int w, const i2s_limits_t *lim);
int i2s_find_circles(const uint8_t *grey, const uint8_t *edges, int n, int h, int w,
                     float *circles, int32_t *counts, uint8_t *masked, int32_t *status,
                     const i2s_limits_t *lim, void *ws, size_t ws_bytes, void *stream);

/* find_lines(threshold, H) and find_lines(threshold, V) -- img2sgf.py:230-255, both
 * directions from ONE read of the masked image (7 angles).  rho: [n][2][line_cap] float32 (dir 0 = H column,
 * dir 1 = V column: V1 rows then V2 rows with rho negated); counts: [n][2]. */
size_t i2s_find_lines_workspace_bytes(int n, int h, int w);
int i2s_find_lines(const uint8_t *masked, int n, int h, int w, int threshold, float *rho,
                   int32_t *counts, int line_cap, int32_t *status, void *ws, size_t ws_bytes,
                   void *stream);

/* find_clusters_fixed_threshold + get_cluster_centres -- img2sgf.py:268-292.
 * rho/counts as produced by i2s_find_lines; centres: [n][2][line_cap] float64 ascending;
 * ncentres: [n][2]. */
int i2s_cluster(const float *rho, const int32_t *counts, int n, int line_cap, double *centres,
                int32_t *ncentres, void *stream);

/* validate_grid -- img2sgf.py:420-445 (complete_grid :335-397, truncate_grid :400-417) */
int i2s_validate_grid(const double *centres, const int32_t *ncentres, int n, int line_cap,
                      i2s_grid_t *grids, int32_t *status, void *stream);

/* identify_board -- img2sgf.py:497-515,537-543 incl. validate_grid's radius filter
 * (:441-443).  brightness: [n][361] float64 in (i,j) scan order (may be NULL). */
int i2s_classify_stones(const uint8_t *grey, int n, int h, int w, const float *circles,
                        const int32_t *counts, int circle_cap, const i2s_grid_t *grids,
                        int black_threshold, i2s_record_t *records, double *brightness,
                        void *stream);

/* the whole path, img2sgf.py:153-198 + 230-292 + 420-445 + 497-543, for n images.
 * Optional taps (may be NULL): grey_out/edges_out/masked_out [n][h][w], circles_out
 * [n][circle_cap][3] + counts_out [n], rho_out [n][2][line_cap] + line_counts_out [n][2],
 * grids_out [n]. */
size_t i2s_pipeline_workspace_bytes(int n, int h, int w, const i2s_limits_t *lim);
int i2s_pipeline(const uint8_t *rgb, int n, int h, int w, int line_threshold, int black_threshold,
                 i2s_record_t *records, uint8_t *grey_out, uint8_t *edges_out, uint8_t *masked_out,
                 float *circles_out, int32_t *counts_out, float *rho_out, int32_t *line_counts_out,
                 i2s_grid_t *grids_out, const i2s_limits_t *lim, void *ws, size_t ws_bytes,
                 void *stream);

/* Optional profiling hooks used by bench.py: CUDA-event timers around each kernel group of the
 * pipeline (off by default) and a counter of this library's kernel launches. */
int i2s_profile_enable(int on);                      /* returns the number of sections */
const char *i2s_profile_section_name(int id);
int i2s_profile_read(double *ms, long long *counts, int nsections);   /* synchronises the events */
long long i2s_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif
