// board.cuh -- internal interface of board.cu
#pragma once
#include "common.cuh"

namespace i2s {
int validate_grid(const double *centres, const int32_t *ncentres, int n, int line_cap, i2s_grid_t *grids,
                  int32_t *status, cudaStream_t st);
int classify_stones(const uint8_t *grey, const Dims &dims, int n, int pitch, size_t stride, const float *circles,
                    const int32_t *counts, int circle_cap, const i2s_grid_t *grids, int black_threshold,
                    i2s_record_t *records, double *brightness, const int32_t *status, cudaStream_t st);
}  // namespace i2s
