// sort.cuh -- block-wide bitonic sort of a power-of-two array in shared memory (ascending).
#pragma once
#include <cuda_runtime.h>

namespace i2s {

template <class T> __device__ __forceinline__ void bitonic_sort_block(T *a, int n)
{
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                int p = i ^ j;
                if (p > i) {
                    T x = a[i], y = a[p];
                    bool up = (i & k) == 0;
                    if ((x > y) == up) { a[i] = y; a[p] = x; }
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();
}

}  // namespace i2s
