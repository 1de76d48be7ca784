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
 *   - an image plane is u8 with an explicit row pitch in bytes (`pitch` >= w, or >= w * channels
 *     for interleaved RGB; 0 means tightly packed); a batch of n planes is [n][h][pitch].  Any
 *     pitch and alignment is accepted; 16-byte aligned bases with pitch % 16 == 0 take the fast
 *     paths (bulk-copy staging, 128-bit loads);
 *   - outputs are caller-allocated with explicit capacities; counts are written on device;
 *   - `ws` is a caller-allocated device workspace of at least the size reported by the
 *     matching *_workspace_bytes() call; nothing is allocated inside and the library keeps no
 *     state between calls (the optional profiling hooks at the end are the one exception);
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

#define I2S_MAX_CIRCLE_CAP 65534    /* the masking kernel keeps 1 + circle index in 16 bits */

/* limits used by the entry points (all data-dependent sizes) */
typedef struct {
    int32_t cand_cap;      /* accumulator peaks per HoughCircles call (32 .. 16384)       */
    int32_t circle_cap;    /* circles per image in the stacked output (rows of 3 floats), <= I2S_MAX_CIRCLE_CAP */
    int32_t line_cap;      /* line peaks per direction per image (2 .. 4096)             */
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

/* One image of a ragged batch (the reference handles any image size, img2sgf.py:117-204;
 * its test images range from 110x102 to 1265x1245). */
typedef struct {
    int64_t offset;          /* byte offset of the image's first pixel from the batch's base pointer */
    int32_t h, w;            /* rows, columns                                                         */
    int32_t pitch;           /* bytes between rows of this image (>= w * channels)                    */
    int32_t line_threshold;  /* HoughLines vote threshold of this image; 0 = take it from i2s_params_t */
} i2s_image_t;               /* 24 bytes */

/* The input of i2s_pipeline: n images behind one base pointer. */
typedef struct {
    int32_t n;               /* images                                                                 */
    int32_t channels;        /* 3: interleaved u8 in PIL's RGB order (img2sgf.py:150); 1: greyscale
                                source (Image.open(..).convert('RGB') of a mode-"L" file gives R=G=B,
                                img2sgf.py:651: grey == the channel and channel 0 wins every Canny tie,
                                so the single plane gives identical results at a third of the bytes)  */
    int32_t h, w;            /* uniform batch: size of every image; ragged: max over the images        */
    int32_t pitch;           /* uniform batch: bytes between rows (0 = w * channels), image i starts at
                                base + i * h * pitch; ragged batch: unused                            */
    int32_t pad_;
    const i2s_image_t *images;   /* DEVICE array [n] for a ragged batch; NULL for a uniform batch      */
} i2s_batch_t;

/* The sliders / constants the reference reads through Tk getters and globals. */
typedef struct {
    int32_t line_threshold;  /* threshold.get() (img2sgf.py:259,298); 0 = choose_threshold() per image (:606-613) */
    int32_t black_threshold; /* black_stone_threshold (:45,515,541)                                   */
    int32_t canny_low, canny_high;   /* edge_min / edge_max (:47-48,163)                              */
    float contrast_factor;   /* ImageEnhance.Contrast factor 102/(101-c)-1 (:142-144); 1.0 = input is already enhanced */
    float brightness_factor; /* ImageEnhance.Brightness factor 450/(200-b)-2 (:147-149); 1.0 = identity */
} i2s_params_t;

/* Optional outputs of i2s_pipeline (any pointer may be NULL). */
typedef struct {
    int32_t plane_pitch;     /* bytes between rows of grey/edges/masked ([n][h][plane_pitch]); must equal i2s_canvas_pitch(w) */
    int32_t pad_;
    uint8_t *grey, *edges, *masked;       /* grey_image_np, edge_detected_image_np, circles_removed_image_np */
    float *circles; int32_t *counts;      /* stacked circles [n][circle_cap][3] + counts [n]            */
    float *rho; int32_t *line_counts;     /* [n][2][line_cap] + [n][2]                                  */
    i2s_grid_t *grids;                    /* [n]                                                        */
    double *brightness;                   /* stone_brightnesses [n][361]                                */
} i2s_taps_t;

const char *i2s_last_error(void);
int i2s_version(void);
void i2s_default_limits(i2s_limits_t *lim);
void i2s_default_params(i2s_params_t *params);
/* Row pitch (bytes, a multiple of 128) of the planes the library allocates for images w pixels
 * wide; i2s_taps_t planes use it. */
int i2s_canvas_pitch(int w);

/* cv.cvtColor(rgb, COLOR_BGR2GRAY) -- img2sgf.py:153 */
int i2s_grey(const uint8_t *rgb, int rgb_pitch, uint8_t *grey, int pitch, int n, int h, int w, void *stream);

/* ImageEnhance.Contrast(..).enhance(fc) then ImageEnhance.Brightness(..).enhance(fb) --
 * img2sgf.py:142-149 (prologue).  `scratch8n` is a device scratch of n * 8 bytes (luma sums). */
int i2s_enhance(const uint8_t *rgb, int rgb_pitch, uint8_t *out, int out_pitch, void *scratch8n, int n, int h,
                int w, double contrast_factor, double brightness_factor, void *stream);

/* cv.GaussianBlur(grey,(b,b),b), b in {3,5,7} -- img2sgf.py:175.  Writes all three in one
 * pass over the input: dst3/dst5/dst7 each [n][h][pitch] (any may be NULL). */
int i2s_gauss357(const uint8_t *src, uint8_t *dst3, uint8_t *dst5, uint8_t *dst7, int n, int h,
                 int w, int pitch, void *stream);

/* cv.medianBlur(grey,b), b in {1,3,5,7} -- img2sgf.py:174 */
int i2s_median(const uint8_t *src, uint8_t *dst, int n, int h, int w, int pitch, int b, void *stream);

/* cv.Canny(rgb, low, high, apertureSize=3, L2gradient=False) -- img2sgf.py:162-165.
 * channels = 3 for the reference call; channels = 1 gives the single-channel Canny that
 * cv.HoughCircles runs internally.  status: n int32 (OR-ed with I2S_ST_* bits). */
size_t i2s_canny_workspace_bytes(int n, int h, int w);
int i2s_canny(const uint8_t *img, int channels, int img_pitch, uint8_t *edges, int pitch, int n, int h, int w,
              int low, int high, int hyst_passes, int32_t *status, void *ws, size_t ws_bytes, void *stream);

/* cv.HoughCircles(img, HOUGH_GRADIENT, 1, 10, [], 100, 30, 1, 30) -- img2sgf.py:180.
 * circles: [n][circle_cap][3] float32 (x,y,r) in OpenCV's output order; counts: [n]. */
size_t i2s_hough_circles_workspace_bytes(int n, int h, int w, const i2s_limits_t *lim);
int i2s_hough_circles(const uint8_t *img, int pitch, int n, int h, int w, float *circles, int32_t *counts,
                      int32_t *status, const i2s_limits_t *lim, void *ws, size_t ws_bytes,
                      void *stream);

/* the masking loop -- img2sgf.py:169,191-198.  masked may alias edges. */
int i2s_mask_circles(const uint8_t *edges, uint8_t *masked, int pitch, int n, int h, int w,
                     const float *circles, const int32_t *counts, int circle_cap, void *stream);

/* find_circles: blur pyramid + ten HoughCircles calls stacked in `blurs` order + mask
 * -- img2sgf.py:169-198.  circles [n][circle_cap][3], counts [n], masked [n][h][pitch]. */
size_t i2s_find_circles_workspace_bytes(int n, int h, int w, const i2s_limits_t *lim);
int i2s_find_circles(const uint8_t *grey, const uint8_t *edges, int pitch, int n, int h, int w,
                     float *circles, int32_t *counts, uint8_t *masked, int32_t *status,
                     const i2s_limits_t *lim, void *ws, size_t ws_bytes, void *stream);

/* find_lines(threshold, H) and find_lines(threshold, V) -- img2sgf.py:230-255, both
 * directions from ONE read of the masked image (7 angles).  rho: [n][2][line_cap] float32 (dir 0 = H column,
 * dir 1 = V column: V1 rows then V2 rows with rho negated); counts: [n][2]. */
size_t i2s_find_lines_workspace_bytes(int n, int h, int w);
int i2s_find_lines(const uint8_t *masked, int pitch, int n, int h, int w, int threshold, float *rho,
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
int i2s_classify_stones(const uint8_t *grey, int pitch, int n, int h, int w, const float *circles,
                        const int32_t *counts, int circle_cap, const i2s_grid_t *grids,
                        int black_threshold, i2s_record_t *records, double *brightness,
                        void *stream);

/* The whole path, img2sgf.py:142-198 + 230-292 + 420-445 + 497-543, for a batch of images of one
 * size or of different sizes (i2s_batch_t), one record per image.  `workspace_bytes` takes the
 * batch's n and its canvas size (max h, max w). */
size_t i2s_pipeline_workspace_bytes(int n, int h, int w, const i2s_limits_t *lim);
int i2s_pipeline(const uint8_t *src, const i2s_batch_t *batch_host, const i2s_params_t *params_host,
                 i2s_record_t *records, const i2s_taps_t *taps_host, const i2s_limits_t *lim, void *ws,
                 size_t ws_bytes, void *stream);

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
