// common.cuh -- shared helpers for the sm_100a kernels of the img2sgf hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/img2sgf_b200.h"

namespace i2s {

void set_error(const char *fmt, ...);
void count_launch();
bool legacy_enabled(const char *name);
int median357(const uint8_t *src, uint8_t *d3, uint8_t *d5, uint8_t *d7, int n, int h, int w, cudaStream_t st);

#define I2S_CHECK_LAUNCH(what)                                              \
    do {                                                                    \
        i2s::count_launch();                                                \
        cudaError_t e__ = cudaGetLastError();                               \
        if (e__ != cudaSuccess) {                                           \
            i2s::set_error("%s: %s", what, cudaGetErrorString(e__));        \
            return I2S_E_CUDA;                                              \
        }                                                                   \
    } while (0)

#define I2S_CUDA(call)                                                      \
    do {                                                                    \
        cudaError_t e__ = (call);                                           \
        if (e__ != cudaSuccess) {                                           \
            i2s::set_error("%s: %s", #call, cudaGetErrorString(e__));       \
            return I2S_E_CUDA;                                              \
        }                                                                   \
    } while (0)

#define I2S_ARG(cond)                                                       \
    do {                                                                    \
        if (!(cond)) {                                                      \
            i2s::set_error("bad argument: %s (%s:%d)", #cond, __FILE__, __LINE__); \
            return I2S_E_BADARG;                                            \
        }                                                                   \
    } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// bump allocator over the caller's workspace
struct Arena {
    char *base;
    size_t size, off;
    Arena(void *p, size_t n) : base((char *)p), size(n), off(0) {}
    template <class T> T *take(size_t count)
    {
        off = align_up(off, 256);
        T *p = (T *)(base + off);
        off += count * sizeof(T);
        return p;
    }
    bool ok() const { return off <= size; }
};

// A set of `count` image batches, each [n][h][w] (or [n][h][w][3]); "map" m = k * n + i is image i
// of batch k.  Lets one launch cover the grey image, the edge map and the six blurred copies
// (img2sgf.py:171-175) although they live in different buffers.
struct MapSet {
    const uint8_t *src[I2S_N_UNIQUE];
    int count, n;
    __host__ __device__ const uint8_t *plane(int m, size_t plane_bytes) const
    {
        int k = m / n, i = m - k * n;
        return src[k] + (size_t)i * plane_bytes;
    }
    bool aligned4() const
    {
        uintptr_t a = 0;
        for (int k = 0; k < count; k++) a |= (uintptr_t)src[k];
        return (a & 3) == 0;
    }
    bool aligned16() const
    {
        uintptr_t a = 0;
        for (int k = 0; k < count; k++) a |= (uintptr_t)src[k];
        return (a & 15) == 0;
    }
    static MapSet single(const uint8_t *p, int n)
    {
        MapSet ms{};
        ms.src[0] = p; ms.count = 1; ms.n = n;
        return ms;
    }
};

enum Border { BORDER_REPLICATE = 0, BORDER_REFLECT101 = 1, BORDER_ZERO = 2 };

__device__ __forceinline__ int border_index(int p, int len, int mode)
{
    if (mode == BORDER_REPLICATE) return p < 0 ? 0 : (p >= len ? len - 1 : p);
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

// Stage a (tw x th) byte tile whose top-left image coordinate is (x0,y0) into shared
// memory (row pitch `sp` bytes, sp % 4 == 0, x0 % 4 == 0, tw % 4 == 0).  Out-of-image samples
// follow `mode`.  Interior 4-byte groups are fetched with one aligned 32-bit load when the
// image pitch allows it (w % 4 == 0, base 4-aligned); otherwise byte loads.
__device__ __forceinline__ void stage_tile_u8(uint8_t *sm, int sp, const uint8_t *__restrict__ img,
                                              int h, int w, int x0, int y0, int tw, int th, int mode,
                                              bool aligned)
{
    const int groups = tw >> 2;
    for (int idx = threadIdx.x; idx < groups * th; idx += blockDim.x) {
        int ty = idx / groups, g = idx - ty * groups;
        int x = x0 + 4 * g;
        uint32_t v;
        if (mode == BORDER_ZERO) {
            int y = y0 + ty;
            v = 0;
            if (y >= 0 && y < h) {
                const uint8_t *row = img + (size_t)y * w;
                if (aligned && x >= 0 && x + 3 < w) {
                    v = __ldg(reinterpret_cast<const uint32_t *>(row + x));
                } else {
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (x + k >= 0 && x + k < w) v |= (uint32_t)__ldg(row + x + k) << (8 * k);
                }
            }
            *reinterpret_cast<uint32_t *>(sm + ty * sp + 4 * g) = v;
            continue;
        }
        int y = border_index(y0 + ty, h, mode);
        const uint8_t *row = img + (size_t)y * w;
        if (aligned && x >= 0 && x + 3 < w) {
            v = __ldg(reinterpret_cast<const uint32_t *>(row + x));
        } else {
            v = (uint32_t)__ldg(row + border_index(x, w, mode)) |
                ((uint32_t)__ldg(row + border_index(x + 1, w, mode)) << 8) |
                ((uint32_t)__ldg(row + border_index(x + 2, w, mode)) << 16) |
                ((uint32_t)__ldg(row + border_index(x + 3, w, mode)) << 24);
        }
        *reinterpret_cast<uint32_t *>(sm + ty * sp + 4 * g) = v;
    }
}

__device__ __forceinline__ bool ptr_aligned4(const void *p, int w)
{
    return ((reinterpret_cast<uintptr_t>(p) & 3) == 0) && ((w & 3) == 0);
}

}  // namespace i2s
