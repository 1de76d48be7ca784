// profile.cuh -- optional CUDA-event section timers around the kernel groups of the pipeline
// (used by bench.py for the live roofline figure; off by default, zero cost when off).
#pragma once
#include <cuda_runtime.h>

namespace i2s {

enum Section {
    SEC_GREY = 0, SEC_ENHANCE, SEC_SOBEL_NMS_RGB, SEC_SOBEL_NMS, SEC_HYSTERESIS, SEC_STATE_TO_EDGES, SEC_GAUSS, SEC_MEDIAN,
    SEC_EDGE_LIST, SEC_VOTE, SEC_RADIUS, SEC_CIRCLES_FINISH, SEC_STACK, SEC_MASK, SEC_LINE_VOTE, SEC_LINE_PEAKS,
    SEC_CLUSTER, SEC_VALIDATE, SEC_CLASSIFY, SEC_COUNT
};

struct ScopedSection {
    int id; cudaStream_t st; void *stop;
    ScopedSection(int id, cudaStream_t st);
    ~ScopedSection();
};

void count_launch();

}  // namespace i2s
