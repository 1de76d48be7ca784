// common.cuh -- shared helpers for the sm_100a kernels of the img2sgf hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/img2sgf_b200.h"

namespace i2s {

void set_error(const char *fmt, ...);
void count_launch();
int sm_count();                              // SMs of the current device (queried, not assumed)

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

// Row pitch of the library's own planes: 128-byte rows, so every row start is a full cache line and
// bulk copies / 128-bit loads work for any image width (the reference fixtures are 239 .. 1265 wide).
static inline int canvas_pitch(int w) { return (int)align_up((size_t)w, 128); }

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

// Sizes of the images of a batch.  The planes of a batch share one canvas (h rows of `pitch` bytes
// per image); a ragged batch keeps every image in the top-left corner of its canvas slot and the
// kernels look the real size up here.  images == nullptr: every image is w x h.
struct Dims {
    const i2s_image_t *images;
    int w, h;                                  // uniform size, or the canvas (max) size of a ragged batch
    __device__ __forceinline__ int2 of(int i) const      // .x = w, .y = h
    {
        if (!images) return make_int2(w, h);
        const int2 hw = *reinterpret_cast<const int2 *>(&images[i].h);
        return make_int2(hw.y, hw.x);
    }
    static Dims uniform(int h, int w) { return Dims{nullptr, w, h}; }
};

// A set of `count` plane batches ("map" m = k * n + i is image i of batch k): lets one launch cover
// the grey image, the edge map and the six blurred copies (img2sgf.py:171-175) although they live
// in different buffers with different pitches.  A ragged INPUT batch (count == 1, images != nullptr)
// addresses image i at src[0] + images[i].offset with the image's own pitch.
struct MapSet {
    const uint8_t *src[I2S_N_UNIQUE];
    int pitch[I2S_N_UNIQUE];
    size_t stride[I2S_N_UNIQUE];               // bytes between consecutive images of batch k
    int count, n;
    const i2s_image_t *images;
    __device__ __forceinline__ const uint8_t *plane(int m, int &p) const
    {
        const int k = m / n, i = m - k * n;
        if (images) { p = images[i].pitch; return src[0] + images[i].offset; }
        p = pitch[k];
        return src[k] + (size_t)i * stride[k];
    }
    // every plane base and pitch a multiple of `a` (ragged inputs: decided per image in the kernel)
    bool aligned(int a) const
    {
        if (images) return false;
        uintptr_t v = 0;
        for (int k = 0; k < count; k++) v |= (uintptr_t)src[k] | (uintptr_t)pitch[k] | (uintptr_t)stride[k];
        return (v & (uintptr_t)(a - 1)) == 0;
    }
    static MapSet single(const uint8_t *p, int pitch, int h, int n)
    {
        MapSet ms{};
        ms.src[0] = p; ms.pitch[0] = pitch; ms.stride[0] = (size_t)h * pitch; ms.count = 1; ms.n = n;
        return ms;
    }
    void add(const uint8_t *p, int pit, int h)
    {
        src[count] = p; pitch[count] = pit; stride[count] = (size_t)h * pit; count++;
    }
};

static inline bool aligned_to(const void *p, int pitch, int a)
{
    return (((uintptr_t)p | (uintptr_t)pitch) & (uintptr_t)(a - 1)) == 0;
}

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
// image pitch allows it (pitch % 4 == 0, base 4-aligned); otherwise byte loads.
__device__ __forceinline__ void stage_tile_u8(uint8_t *sm, int sp, const uint8_t *__restrict__ img,
                                              int h, int w, int pitch, int x0, int y0, int tw, int th, int mode,
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
                const uint8_t *row = img + (size_t)y * pitch;
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
        const uint8_t *row = img + (size_t)y * pitch;
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

// Store the 4 bytes of `packed` at row[x .. x+3]: one 32-bit store when the plane allows it and the
// word lies inside the row's writable part (`wlim` columns: the image, or the image rounded up into
// the row's padding), else the bytes that lie inside the image.
__device__ __forceinline__ void store4(uint8_t *row, int x, int w, int wlim, bool al, uint32_t packed)
{
    if (al && x + 3 < wlim) *reinterpret_cast<uint32_t *>(row + x) = packed;
    else
        for (int k = 0; k < 4 && x + k < w; k++) row[x + k] = (uint8_t)(packed >> (8 * k));
}

// Columns of a row that kernels may write: the image width rounded up to `a` when the pitch has
// room for it (the padding then holds defined values), else exactly the image.
__host__ __device__ __forceinline__ int write_limit(int w, int pitch, int a)
{
    const int r = (w + a - 1) / a * a;
    return r <= pitch ? r : w;
}

}  // namespace i2s
