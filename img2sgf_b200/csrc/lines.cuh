// lines.cuh -- internal interface of lines.cu
#pragma once
#include "common.cuh"

namespace i2s {
size_t lines_scratch_bytes(int n, int h, int w);
int find_lines(const uint8_t *masked, const Dims &dims, int n, int pitch, size_t stride, int threshold, float *rho,
               int32_t *counts, int line_cap, int32_t *status, Arena &ar, cudaStream_t st);
int cluster(const float *rho, const int32_t *counts, int n, int line_cap, double *centres, int32_t *ncentres,
            cudaStream_t st);
}  // namespace i2s
