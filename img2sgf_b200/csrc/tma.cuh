// tma.cuh -- shared-memory tile staging with the bulk-copy engine (cp.async.bulk, SASS UBLKCP) and an
// mbarrier, falling back to ordinary loads for tiles that touch the left/right image border or when
// a caller-provided plane is not 16-byte aligned / pitched.
#pragma once
#include "common.cuh"

namespace i2s {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}

// Bounded wait: a lost transaction must abort the kernel, never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    for (int spin = 0; spin < (1 << 24); spin++)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}

// Orders earlier generic-proxy accesses to shared memory (ordinary loads/stores/atomics) before later
// async-proxy writes to the same bytes (a bulk copy landing in a buffer the block has just used).
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Stage a (tw x th) byte tile whose top-left image coordinate is (x0,y0) into shared memory with
// row pitch tw.  Requirements for the bulk path (`bulk_ok`: image base 16-byte aligned and
// pitch % 16 == 0 -- true for every plane the library allocates, whatever the image width):
// tw % 16 == 0, x0 % 16 == 0, shared tile 16-byte aligned.  Rows outside the image follow `mode`
// (REPLICATE / REFLECT101) by redirecting the source row; tiles that stick out left or right take
// the ordinary-load path, which also handles those columns.  Ends with the data visible to the
// whole block.
__device__ __forceinline__ void stage_tile_bulk(uint8_t *sm, const uint8_t *__restrict__ img, int h, int w, int pitch,
                                                int x0, int y0, int tw, int th, int mode, bool bulk_ok, bool al,
                                                uint64_t *bar)
{
    if (bulk_ok && x0 >= 0 && x0 + tw <= w) {            // block-uniform
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x < 32) {
            if (threadIdx.x == 0) mbar_arrive_expect_tx(bar, (uint32_t)(tw * th));
            for (int r = threadIdx.x; r < th; r += 32) {
                int y = border_index(y0 + r, h, mode);
                bulk_g2s(sm + r * tw, img + (size_t)y * pitch + x0, (uint32_t)tw, bar);
            }
        }
        mbar_wait(bar, 0);
    } else {
        stage_tile_u8(sm, tw, img, h, w, pitch, x0, y0, tw, th, mode, al);
        __syncthreads();
    }
}

}  // namespace i2s
