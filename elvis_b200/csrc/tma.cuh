// Tensor-memory-accelerator plumbing shared by the kernels that move 2-D byte tiles with TMA
// (cp.async.bulk.tensor): mbarrier helpers, tile loads into shared memory that complete on an mbarrier,
// tile stores from shared memory tracked by bulk groups, and the host-side tensor-map encoder (the driver
// entry point is looked up at run time, so the library does not link libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace elvis {
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}

// box of a 3-D tensor (x = bytes along a row, y = row, t = frame) -> shared memory; completes `bar` with the box's bytes
__device__ __forceinline__ void load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int t, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(t), "r"(bar) : "memory");
}
// shared memory -> box of a 3-D tensor (elements outside the tensor are not written); joins the current bulk group
__device__ __forceinline__ void store_3d(const CUtensorMap* map, int x, int y, int t, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(map), "r"(x), "r"(y), "r"(t), "r"(src) : "memory");
}
__device__ __forceinline__ void store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until all but the N most recent bulk groups of this thread have finished READING shared memory
template <int N> __device__ __forceinline__ void store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void store_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// order this thread's generic-proxy writes to shared memory before later async-proxy (TMA) reads of them
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) { asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory"); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &st) != cudaSuccess ||
            st != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// Can a (W bytes, H rows, T frames) uint8 plane with these strides be described by a tensor map?
inline bool plane_ok(const void* base, int64_t row_stride, int64_t frame_stride, int T) {
    return base && (reinterpret_cast<uintptr_t>(base) & 15) == 0 && row_stride > 0 && row_stride % 16 == 0 &&
           (T <= 1 || (frame_stride > 0 && frame_stride % 16 == 0)) && row_stride < (1ll << 40) && frame_stride < (1ll << 40);
}

// (W, H, T) uint8 tensor with a (box_w bytes x box_h rows x 1 frame) box; zero fill outside the tensor on loads
inline bool make_plane_map(CUtensorMap* m, const uint8_t* base, int W, int H, int T, int64_t row_stride, int64_t frame_stride,
                           int box_w, int box_h, CUtensorMapSwizzle swizzle) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || !plane_ok(base, row_stride, frame_stride, T)) return false;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T};
    cuuint64_t strides[2] = {(cuuint64_t)row_stride, (cuuint64_t)(T > 1 ? frame_stride : row_stride * H)};
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tma
}  // namespace elvis
