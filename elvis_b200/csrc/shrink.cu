// a4-a7: shrink (compact the kept blocks of every block row to the left) and stretch (scatter
// them back, zeros for the removed blocks).  Pure data movement: HBM-bound.
//
// One CTA owns one (frame, block-row).  It first turns the row's mask bytes into a column
// map in shared memory with a warp-ballot prefix sum -- for every DESTINATION block column
// the SOURCE block column, or -1 for "write zeros" -- and then streams the block row with
// the widest vector unit the geometry allows (128-bit for 16-pixel luma blocks and for
// packed 3-channel blocks; 64-bit for 8-pixel chroma blocks).  Both directions are written
// as gathers so that every store of a warp is one contiguous, fully used run of lines.
#include "common.cuh"

namespace elvis {
namespace {

constexpr int kThreads = 256;

struct MoveParams {
    const uint8_t* src;
    uint8_t* dst;
    int64_t src_frame, src_row, dst_frame, dst_row;   // strides in bytes
    const uint8_t* mask;
    int32_t By, Bx;          // block grid of the full-size plane
    int32_t small_bx;        // blocks per row of the shrunk plane
    int32_t bh;              // pixel rows per block
    int32_t upb;             // vector units per block row (block_px * channels / sizeof(V))
    int32_t tx_dim, ty_dim;  // thread tile: tx over destination units, ty over pixel rows
    int32_t n_units;         // T * By (frame, block-row) units, grid-strided
};

template <typename V> __device__ __forceinline__ V zero_v();
template <> __device__ __forceinline__ uint4 zero_v<uint4>() { return make_uint4(0, 0, 0, 0); }
template <> __device__ __forceinline__ uint2 zero_v<uint2>() { return make_uint2(0, 0); }
template <> __device__ __forceinline__ uint32_t zero_v<uint32_t>() { return 0; }
template <> __device__ __forceinline__ uint16_t zero_v<uint16_t>() { return 0; }
template <> __device__ __forceinline__ uint8_t zero_v<uint8_t>() { return 0; }

// STRETCH == false: map[j] (j < small_bx) = j-th kept column of the row, -1 if the row keeps fewer.
// STRETCH == true : map[i] (i < Bx)       = rank of column i among the kept ones, -1 if removed
//                                           or beyond the shrunk width (utils.py:756 guard).
template <bool STRETCH>
__device__ __forceinline__ void build_map(const uint8_t* __restrict__ mrow, int Bx, int small_bx, int16_t* s_map) {
    __shared__ int s_warp[kThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int base = 0;
    for (int c0 = 0; c0 < Bx; c0 += kThreads) {
        const int i = c0 + threadIdx.x;
        const bool kept = i < Bx && mrow[i] == 0;
        const unsigned ballot = __ballot_sync(0xffffffffu, kept);
        if (lane == 0) s_warp[w] = __popc(ballot);
        __syncthreads();
        int before = base, total = base;
#pragma unroll
        for (int q = 0; q < kThreads / 32; ++q) {
            const int n = s_warp[q];
            if (q < w) before += n;
            total += n;
        }
        const int pos = before + __popc(ballot & ((1u << lane) - 1u));
        if (STRETCH) {
            if (i < Bx) s_map[i] = (kept && pos < small_bx) ? (int16_t)pos : (int16_t)-1;
        } else {
            if (kept && pos < small_bx) s_map[pos] = (int16_t)i;
        }
        base = total;
        __syncthreads();
    }
    if (!STRETCH)
        for (int j = base + threadIdx.x; j < small_bx; j += kThreads) s_map[j] = -1;
    __syncthreads();
}

template <typename V, bool STRETCH>
__global__ void __launch_bounds__(kThreads) move_blocks_kernel(const MoveParams p) {
    extern __shared__ int16_t s_map[];
    const int tx = threadIdx.x % p.tx_dim, ty = threadIdx.x / p.tx_dim;
    const int dst_blocks = STRETCH ? p.Bx : p.small_bx;
    const int cols = dst_blocks * p.upb;
    const int64_t srow = p.src_row / (int64_t)sizeof(V), drow = p.dst_row / (int64_t)sizeof(V);

    // grid-stride over (frame, block-row) units: the default grid has one CTA per unit; a capped
    // grid (tuning knob) keeps the kernel's footprint per SM fixed so that it can share the SMs
    // with the scoring kernel of the next clip
    for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
        const int t = unit / p.By;
        const int by = unit - t * p.By;
        build_map<STRETCH>(p.mask + ((int64_t)t * p.By + by) * p.Bx, p.Bx, p.small_bx, s_map);
        if (ty < p.ty_dim) {
            const V* sbase = reinterpret_cast<const V*>(p.src + (int64_t)t * p.src_frame + (int64_t)by * p.bh * p.src_row);
            V* dbase = reinterpret_cast<V*>(p.dst + (int64_t)t * p.dst_frame + (int64_t)by * p.bh * p.dst_row);
            for (int c = tx; c < cols; c += p.tx_dim) {
                const int j = c / p.upb;
                const int u = c - j * p.upb;
                const int sj = s_map[j];
                const V* sp = sbase + (sj < 0 ? 0 : sj * p.upb + u);
                V* dp = dbase + c;
                int r = ty;
                // four rows in flight per thread: all loads issued before the stores
                for (; r + 3 * p.ty_dim < p.bh; r += 4 * p.ty_dim) {
                    V v0 = zero_v<V>(), v1 = zero_v<V>(), v2 = zero_v<V>(), v3 = zero_v<V>();
                    if (sj >= 0) {
                        v0 = ld_stream(sp + (int64_t)r * srow);
                        v1 = ld_stream(sp + (int64_t)(r + p.ty_dim) * srow);
                        v2 = ld_stream(sp + (int64_t)(r + 2 * p.ty_dim) * srow);
                        v3 = ld_stream(sp + (int64_t)(r + 3 * p.ty_dim) * srow);
                    }
                    st_stream(dp + (int64_t)r * drow, v0);
                    st_stream(dp + (int64_t)(r + p.ty_dim) * drow, v1);
                    st_stream(dp + (int64_t)(r + 2 * p.ty_dim) * drow, v2);
                    st_stream(dp + (int64_t)(r + 3 * p.ty_dim) * drow, v3);
                }
                for (; r < p.bh; r += p.ty_dim) {
                    V v = zero_v<V>();
                    if (sj >= 0) v = ld_stream(sp + (int64_t)r * srow);
                    st_stream(dp + (int64_t)r * drow, v);
                }
            }
        }
        __syncthreads();   // the column map is rebuilt for the next unit
    }
}

template <bool STRETCH>
int launch_move(const elvis_plane* src, const elvis_plane* dst, int T, int block_px, int By, int Bx, int small_bx,
                const uint8_t* mask, int ctas_per_sm, cudaStream_t st) {
    const elvis_plane* full = STRETCH ? dst : src;
    const elvis_plane* small = STRETCH ? src : dst;
    if (!plane_ok(src) || !plane_ok(dst) || !mask || T <= 0 || block_px <= 0 || By <= 0 || Bx <= 0 || small_bx < 0)
        return ELVIS_ERR_INVALID_ARG;
    if (src->channels != dst->channels) return ELVIS_ERR_INVALID_ARG;
    if (Bx > 32767 || small_bx > Bx) return ELVIS_ERR_INVALID_ARG;
    if (full->height < By * block_px || full->width < Bx * block_px) return ELVIS_ERR_SHAPE;
    if (small->height < By * block_px || small->width < small_bx * block_px) return ELVIS_ERR_SHAPE;
    if (small_bx == 0 && !STRETCH) return ELVIS_OK;
    const int64_t block_row_bytes = (int64_t)block_px * src->channels;
    int unit = vector_unit(src, block_row_bytes);
    const int unit_dst = vector_unit(dst, block_row_bytes);
    if (unit_dst < unit) unit = unit_dst;

    MoveParams p;
    p.src = static_cast<const uint8_t*>(src->data);
    p.dst = static_cast<uint8_t*>(dst->data);
    p.src_frame = src->frame_stride;
    p.src_row = src->row_stride;
    p.dst_frame = dst->frame_stride;
    p.dst_row = dst->row_stride;
    p.mask = mask;
    p.By = By;
    p.Bx = Bx;
    p.small_bx = small_bx;
    p.bh = block_px;
    p.upb = (int)(block_row_bytes / unit);
    const int cols = (STRETCH ? Bx : small_bx) * p.upb;
    int tx_dim = ((cols + 31) / 32) * 32;
    if (tx_dim > kThreads) tx_dim = kThreads;
    if (tx_dim < 32) tx_dim = 32;
    p.tx_dim = tx_dim;
    p.ty_dim = kThreads / tx_dim;
    if (p.ty_dim > block_px) p.ty_dim = block_px;

    const size_t smem = sizeof(int16_t) * (size_t)(Bx > small_bx ? Bx : small_bx);
    p.n_units = (int)((int64_t)T * By);
    unsigned grid = (unsigned)p.n_units;
    if (ctas_per_sm > 0 && (unsigned)(ctas_per_sm * kNumSMs) < grid) grid = (unsigned)(ctas_per_sm * kNumSMs);
    switch (unit) {
        case 16: move_blocks_kernel<uint4, STRETCH><<<grid, kThreads, smem, st>>>(p); break;
        case 8: move_blocks_kernel<uint2, STRETCH><<<grid, kThreads, smem, st>>>(p); break;
        case 4: move_blocks_kernel<uint32_t, STRETCH><<<grid, kThreads, smem, st>>>(p); break;
        case 2: move_blocks_kernel<uint16_t, STRETCH><<<grid, kThreads, smem, st>>>(p); break;
        default: move_blocks_kernel<uint8_t, STRETCH><<<grid, kThreads, smem, st>>>(p); break;
    }
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

}  // namespace
}  // namespace elvis

extern "C" int elvis_shrink(const elvis_plane* src, const elvis_plane* dst, int32_t n_frames,
                            int32_t block_px, int32_t by, int32_t bx, int32_t out_bx,
                            const uint8_t* mask, int32_t ctas_per_sm, elvis_stream_t stream) {
    return elvis::launch_move<false>(src, dst, n_frames, block_px, by, bx, out_bx, mask, ctas_per_sm, elvis::as_stream(stream));
}

extern "C" int elvis_stretch(const elvis_plane* src, const elvis_plane* dst, int32_t n_frames,
                             int32_t block_px, int32_t by, int32_t bx, int32_t shrunk_bx,
                             const uint8_t* mask, int32_t ctas_per_sm, elvis_stream_t stream) {
    return elvis::launch_move<true>(src, dst, n_frames, block_px, by, bx, shrunk_bx, mask, ctas_per_sm, elvis::as_stream(stream));
}
