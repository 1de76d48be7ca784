// preproc.cuh -- internal interface of preproc.cu
#pragma once
#include "common.cuh"

namespace i2s {
int enhance(const MapSet &ms, const Dims &dims, int ch, uint8_t *out, int opitch, size_t ostride, void *scratch8n, float fc,
            float fb, cudaStream_t st);
int to_canvas(const MapSet &ms, const Dims &dims, uint8_t *out, int opitch, size_t ostride, cudaStream_t st);
// src: [n] planes of spitch bytes per row, sstride bytes apart; d3/d5/d7: pitch / stride
int gauss357(const uint8_t *src, int spitch, size_t sstride, uint8_t *d3, uint8_t *d5, uint8_t *d7, int pitch, size_t stride,
             const Dims &dims, int n, cudaStream_t st);
int median357(const uint8_t *src, int spitch, size_t sstride, uint8_t *d3, uint8_t *d5, uint8_t *d7, int pitch, size_t stride,
              const Dims &dims, int n, cudaStream_t st);
}  // namespace i2s
