// canny.cuh -- internal interface of canny.cu (state maps shared with the circle detector)
#pragma once
#include "common.cuh"

namespace i2s {
size_t canny_scratch_bytes(int maps, int h, int w);
int canny_states(const MapSet &ms, const Dims &dims, int channels, uint8_t *state, int spitch, size_t sstride, int low,
                 int high, int passes, int32_t *status, void *scratch, cudaStream_t st, uint8_t *grey, int gpitch,
                 size_t gstride);
int hysteresis(uint8_t *state, int spitch, size_t sstride, const Dims &dims, int maps, int n_images, int passes,
               int32_t *status, void *scratch, cudaStream_t st);
int states_to_edges(const uint8_t *state, uint8_t *edges, size_t total, cudaStream_t st);
}  // namespace i2s
