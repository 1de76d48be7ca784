// pipeline.cu -- the whole hot path for a batch of images (one size or ragged), enqueued on one stream.
// Reference: process_image img2sgf.py:142-198 -> find_grid :546-576 -> identify_board :497-543.
#include "board.cuh"
#include "canny.cuh"
#include "circles.cuh"
#include "lines.cuh"
#include "preproc.cuh"

using namespace i2s;

static_assert(sizeof(i2s_record_t) == 384, "record must be 384 bytes (SURVEY.md section 8e)");
static_assert(sizeof(i2s_image_t) == 24, "i2s_image_t layout");

extern "C" size_t i2s_pipeline_workspace_bytes(int n, int h, int w, const i2s_limits_t *lim)
{
    if (n <= 0 || h <= 0 || w <= 0 || !lim) return 4096;
    const size_t plane = (size_t)h * canvas_pitch(w);
    size_t b = 0;
    b += 3 * align_up((size_t)n * plane, 256);                       // grey, edges, masked
    b += align_up((size_t)n * h * canvas_pitch(3 * w), 256);         // enhanced input (prologue)
    b += align_up((size_t)n * 8, 256);                               // luma sums
    b += align_up((size_t)n * lim->circle_cap * 12, 256);            // stacked circles
    b += align_up((size_t)n * 2 * lim->line_cap * 4, 256);           // rho columns
    b += align_up((size_t)n * 2 * lim->line_cap * 8, 256);           // cluster centres
    b += 4 * align_up((size_t)n * 2 * 4, 256);                       // counts, line counts, ncentres, status
    b += align_up((size_t)n * sizeof(i2s_grid_t), 256);
    b += canny_scratch_bytes(n, h, w) + 256;
    b += find_circles_scratch_bytes(n, h, w, *lim);
    b += lines_scratch_bytes(n, h, w);
    return b + 8192;
}

extern "C" int i2s_pipeline(const uint8_t *src, const i2s_batch_t *batch, const i2s_params_t *params, i2s_record_t *records,
                            const i2s_taps_t *taps, const i2s_limits_t *lim, void *ws, size_t ws_bytes, void *stream)
{
    I2S_ARG(src && batch && params && records && lim && ws);
    const int n = batch->n, h = batch->h, w = batch->w, ch = batch->channels;
    I2S_ARG(n >= 0 && h > 0 && w > 0 && h < 16384 && w < 16384 && (ch == 1 || ch == 3));
    I2S_ARG(params->line_threshold >= 0 && params->canny_low >= 0 && params->canny_high >= 0);
    int rc = check_limits(lim);
    if (rc) return rc;
    if (n == 0) return I2S_OK;
    const int ipitch = batch->pitch ? batch->pitch : w * ch;
    I2S_ARG(batch->images || ipitch >= w * ch);
    if (ws_bytes < i2s_pipeline_workspace_bytes(n, h, w, lim)) {
        set_error("i2s_pipeline: workspace too small (%zu < %zu)", ws_bytes, i2s_pipeline_workspace_bytes(n, h, w, lim));
        return I2S_E_WORKSPACE;
    }
    const int P = canvas_pitch(w);
    const size_t plane = (size_t)h * P;
    const i2s_taps_t none{};
    const i2s_taps_t &t = taps ? *taps : none;
    if (t.grey || t.edges || t.masked) I2S_ARG(t.plane_pitch == P);
    cudaStream_t st = (cudaStream_t)stream;
    Arena ar(ws, ws_bytes);
    uint8_t *grey = t.grey ? t.grey : ar.take<uint8_t>(n * plane);
    uint8_t *edges = t.edges ? t.edges : ar.take<uint8_t>(n * plane);
    uint8_t *masked = t.masked ? t.masked : ar.take<uint8_t>(n * plane);
    float *circles = t.circles ? t.circles : ar.take<float>((size_t)n * lim->circle_cap * 3);
    int32_t *counts = t.counts ? t.counts : ar.take<int32_t>(n);
    float *rho = t.rho ? t.rho : ar.take<float>((size_t)n * 2 * lim->line_cap);
    int32_t *lcounts = t.line_counts ? t.line_counts : ar.take<int32_t>(n * 2);
    double *centres = ar.take<double>((size_t)n * 2 * lim->line_cap);
    int32_t *ncentres = ar.take<int32_t>(n * 2);
    i2s_grid_t *grids = t.grids ? t.grids : ar.take<i2s_grid_t>(n);
    int32_t *status = ar.take<int32_t>(n);
    void *cscratch = ar.take<uint8_t>(canny_scratch_bytes(n, h, w));
    if (!ar.ok()) { set_error("i2s_pipeline: workspace accounting"); return I2S_E_WORKSPACE; }

    const Dims dims{batch->images, w, h};
    MapSet in = MapSet::single(src, ipitch, h, n);
    in.images = batch->images;
    I2S_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t) * n, st));
    if (params->contrast_factor != 1.0f || params->brightness_factor != 1.0f) {                // :142-149
        const int epitch = canvas_pitch(ch * w);
        uint8_t *enh = ar.take<uint8_t>((size_t)n * h * epitch);
        void *sums = ar.take<unsigned long long>(n);
        if (!ar.ok()) { set_error("i2s_pipeline: workspace accounting"); return I2S_E_WORKSPACE; }
        if ((rc = enhance(in, dims, ch, enh, epitch, (size_t)h * epitch, sums, params->contrast_factor,
                          params->brightness_factor, st)))
            return rc;
        in = MapSet::single(enh, epitch, h, n);
    }
    if (ch == 3) {
        // :153 + :162-165 -- the 3-channel Canny writes the greyscale plane from the words it loads
        if ((rc = canny_states(in, dims, 3, edges, P, plane, params->canny_low, params->canny_high, lim->hyst_passes, status,
                               cscratch, st, grey, P, plane)))
            return rc;
    } else {
        // a greyscale source: grey == the plane, and identical channels make the 3-channel Canny the 1-channel one
        if ((rc = to_canvas(in, dims, grey, P, plane, st))) return rc;
        MapSet g = MapSet::single(grey, P, h, n);
        if ((rc = canny_states(g, dims, 1, edges, P, plane, params->canny_low, params->canny_high, lim->hyst_passes, status,
                               cscratch, st, nullptr, 0, 0)))
            return rc;
    }
    if ((rc = states_to_edges(edges, edges, n * plane, st))) return rc;
    if ((rc = find_circles(grey, edges, dims, n, P, circles, counts, masked, status, *lim, ar, st))) return rc;      // :169-198
    if ((rc = find_lines(masked, dims, n, P, plane, params->line_threshold, rho, lcounts, lim->line_cap, status, ar, st)))
        return rc;                                                                             // :230-255
    if ((rc = cluster(rho, lcounts, n, lim->line_cap, centres, ncentres, st))) return rc;     // :268-292
    if ((rc = validate_grid(centres, ncentres, n, lim->line_cap, grids, status, st))) return rc;  // :420-445
    return classify_stones(grey, dims, n, P, plane, circles, counts, lim->circle_cap, grids, params->black_threshold, records,
                           t.brightness, status, st);                                          // :497-543
}
