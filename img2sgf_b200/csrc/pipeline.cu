// pipeline.cu -- the whole hot path for a batch of same-sized images, enqueued on one stream.
// Reference: process_image img2sgf.py:153-198 -> find_grid :546-576 -> identify_board :497-543.
#include "board.cuh"
#include "canny.cuh"
#include "circles.cuh"
#include "lines.cuh"

using namespace i2s;

static_assert(sizeof(i2s_record_t) == 384, "record must be 384 bytes (SURVEY.md section 8e)");

extern "C" size_t i2s_pipeline_workspace_bytes(int n, int h, int w, const i2s_limits_t *lim)
{
    if (n <= 0 || h <= 0 || w <= 0 || !lim) return 4096;
    size_t plane = (size_t)h * w, b = 0;
    b += 3 * align_up((size_t)n * plane, 256);                       // grey, edges, masked
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

extern "C" int i2s_pipeline(const uint8_t *rgb, int n, int h, int w, int line_threshold, int black_threshold,
                            i2s_record_t *records, uint8_t *grey_out, uint8_t *edges_out, uint8_t *masked_out,
                            float *circles_out, int32_t *counts_out, float *rho_out, int32_t *line_counts_out,
                            i2s_grid_t *grids_out, const i2s_limits_t *lim, void *ws, size_t ws_bytes, void *stream)
{
    I2S_ARG(rgb && records && lim && ws && n >= 0 && h > 0 && w > 0 && h < 16384 && w < 16384);
    I2S_ARG(lim->cand_cap >= 32 && lim->cand_cap <= 16384 && lim->circle_cap >= 1 && lim->line_cap >= 2 &&
            lim->line_cap <= 4096 && lim->hyst_passes >= 1);
    if (n == 0) return I2S_OK;
    if (ws_bytes < i2s_pipeline_workspace_bytes(n, h, w, lim)) {
        set_error("i2s_pipeline: workspace too small (%zu < %zu)", ws_bytes, i2s_pipeline_workspace_bytes(n, h, w, lim));
        return I2S_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    Arena ar(ws, ws_bytes);
    const size_t plane = (size_t)h * w;
    uint8_t *grey = grey_out ? grey_out : ar.take<uint8_t>(n * plane);
    uint8_t *edges = edges_out ? edges_out : ar.take<uint8_t>(n * plane);
    uint8_t *masked = masked_out ? masked_out : ar.take<uint8_t>(n * plane);
    float *circles = circles_out ? circles_out : ar.take<float>((size_t)n * lim->circle_cap * 3);
    int32_t *counts = counts_out ? counts_out : ar.take<int32_t>(n);
    float *rho = rho_out ? rho_out : ar.take<float>((size_t)n * 2 * lim->line_cap);
    int32_t *lcounts = line_counts_out ? line_counts_out : ar.take<int32_t>(n * 2);
    double *centres = ar.take<double>((size_t)n * 2 * lim->line_cap);
    int32_t *ncentres = ar.take<int32_t>(n * 2);
    i2s_grid_t *grids = grids_out ? grids_out : ar.take<i2s_grid_t>(n);
    int32_t *status = ar.take<int32_t>(n);
    void *cscratch = ar.take<uint8_t>(canny_scratch_bytes(n, h, w));
    if (!ar.ok()) { set_error("i2s_pipeline: workspace accounting"); return I2S_E_WORKSPACE; }

    int rc;
    I2S_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t) * n, st));
    if ((rc = i2s_grey(rgb, grey, n, h, w, st))) return rc;                                   // :153
    MapSet rgbset = MapSet::single(rgb, n);
    if ((rc = canny_states(rgbset, 3, edges, h, w, 50, 200, lim->hyst_passes, status, cscratch, st))) return rc;  // :162-165
    if ((rc = states_to_edges(edges, edges, n * plane, st))) return rc;
    if ((rc = find_circles(grey, edges, n, h, w, circles, counts, masked, status, *lim, ar, st))) return rc;      // :169-198
    if ((rc = find_lines(masked, n, h, w, line_threshold, rho, lcounts, lim->line_cap, status, ar, st))) return rc;  // :230-255
    if ((rc = cluster(rho, lcounts, n, lim->line_cap, centres, ncentres, st))) return rc;     // :268-292
    if ((rc = validate_grid(centres, ncentres, n, lim->line_cap, grids, status, st))) return rc;  // :420-445
    return classify_stones(grey, n, h, w, circles, counts, lim->circle_cap, grids, black_threshold, records, nullptr,
                           status, st);                                                        // :497-543
}
