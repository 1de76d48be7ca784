// circles.cuh -- internal interface of circles.cu
#pragma once
#include "common.cuh"

namespace i2s {
int check_limits(const i2s_limits_t *lim);
size_t circles_scratch_bytes(int maps, int h, int w, const i2s_limits_t &lim);
size_t find_circles_scratch_bytes(int n, int h, int w, const i2s_limits_t &lim);
int hough_circles_maps(const MapSet &ms, const Dims &dims, float *mcirc, int32_t *mcount, int32_t *status,
                       const i2s_limits_t &lim, Arena &ar, cudaStream_t st);
int mask_circles(const uint8_t *edges, uint8_t *masked, const Dims &dims, int n, int pitch, size_t stride,
                 const float *circles, const int32_t *counts, int circle_cap, cudaStream_t st, const int2 *dup = nullptr);
int find_circles(const uint8_t *grey, const uint8_t *edges, const Dims &dims, int n, int pitch, float *circles,
                 int32_t *counts, uint8_t *masked, int32_t *status, const i2s_limits_t &lim, Arena &ar, cudaStream_t st);
}  // namespace i2s
