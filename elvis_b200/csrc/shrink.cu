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

// Gather one block row: destination unit c of every pixel row <- source unit map[c / upb]*upb + c % upb
// (zeros when the map says -1).  Four rows in flight per thread: all loads issued before the stores.
template <typename V, bool SPARSE = false>      // SPARSE: the source units are a scattered subset (shrink) -> 64-byte L2 fills
__device__ __forceinline__ void copy_block_row(const uint8_t* __restrict__ src, int64_t src_row, uint8_t* __restrict__ dst,
                                               int64_t dst_row, const int16_t* s_map, int cols, int upb, int bh, int tx,
                                               int ty, int tx_dim, int ty_dim) {
    const int64_t srow = src_row / (int64_t)sizeof(V), drow = dst_row / (int64_t)sizeof(V);
    const V* sbase = reinterpret_cast<const V*>(src);
    V* dbase = reinterpret_cast<V*>(dst);
    for (int c = tx; c < cols; c += tx_dim) {
        const int j = c / upb;
        const int u = c - j * upb;
        const int sj = s_map[j];
        const V* sp = sbase + (sj < 0 ? 0 : sj * upb + u);
        V* dp = dbase + c;
        int r = ty;
        for (; r + 3 * ty_dim < bh; r += 4 * ty_dim) {
            V v0 = zero_v<V>(), v1 = zero_v<V>(), v2 = zero_v<V>(), v3 = zero_v<V>();
            if (sj >= 0) {
                if (SPARSE) {
                    v0 = ld_gather(sp + (int64_t)r * srow);
                    v1 = ld_gather(sp + (int64_t)(r + ty_dim) * srow);
                    v2 = ld_gather(sp + (int64_t)(r + 2 * ty_dim) * srow);
                    v3 = ld_gather(sp + (int64_t)(r + 3 * ty_dim) * srow);
                } else {
                    v0 = ld_stream(sp + (int64_t)r * srow);
                    v1 = ld_stream(sp + (int64_t)(r + ty_dim) * srow);
                    v2 = ld_stream(sp + (int64_t)(r + 2 * ty_dim) * srow);
                    v3 = ld_stream(sp + (int64_t)(r + 3 * ty_dim) * srow);
                }
            }
            st_stream(dp + (int64_t)r * drow, v0);
            st_stream(dp + (int64_t)(r + ty_dim) * drow, v1);
            st_stream(dp + (int64_t)(r + 2 * ty_dim) * drow, v2);
            st_stream(dp + (int64_t)(r + 3 * ty_dim) * drow, v3);
        }
        for (; r < bh; r += ty_dim) {
            V v = zero_v<V>();
            if (sj >= 0) v = SPARSE ? ld_gather(sp + (int64_t)r * srow) : ld_stream(sp + (int64_t)r * srow);
            st_stream(dp + (int64_t)r * drow, v);
        }
    }
}

template <typename V, bool STRETCH>
__global__ void __launch_bounds__(kThreads) move_blocks_kernel(const MoveParams p) {
    extern __shared__ int16_t s_map[];
    const int tx = threadIdx.x % p.tx_dim, ty = threadIdx.x / p.tx_dim;
    const int dst_blocks = STRETCH ? p.Bx : p.small_bx;
    const int cols = dst_blocks * p.upb;

    // grid-stride over (frame, block-row) units: the default grid has one CTA per unit; a capped
    // grid (tuning knob) keeps the kernel's footprint per SM fixed so that it can share the SMs
    // with the scoring kernel of the next clip
    for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
        const int t = unit / p.By;
        const int by = unit - t * p.By;
        build_map<STRETCH>(p.mask + ((int64_t)t * p.By + by) * p.Bx, p.Bx, p.small_bx, s_map);
        if (ty < p.ty_dim)
            copy_block_row<V, !STRETCH>(p.src + (int64_t)t * p.src_frame + (int64_t)by * p.bh * p.src_row, p.src_row,
                              p.dst + (int64_t)t * p.dst_frame + (int64_t)by * p.bh * p.dst_row, p.dst_row,
                              s_map, cols, p.upb, p.bh, tx, ty, p.tx_dim, p.ty_dim);
        if (unit + (int)gridDim.x < p.n_units) __syncthreads();   // the column map is rebuilt for the next unit
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


// ---- planar YUV 4:2:0, luma blocks a multiple of 16 pixels: Y, U and V of one (frame, block row)
// in ONE CTA -- the mask row is turned into the column map once, luma moves in 128-bit units and
// the two chroma planes in 64-bit units, and a clip needs one launch instead of three.
struct MoveYuvParams {
    const uint8_t* src[3];
    uint8_t* dst[3];
    int64_t src_frame[3], src_row[3], dst_frame[3], dst_row[3];
    const uint8_t* mask;
    int32_t By, Bx, small_bx, bs, tx_dim, ty_dim, n_units;
};

template <bool STRETCH>
__global__ void __launch_bounds__(kThreads) move_yuv420_kernel(const MoveYuvParams p) {
    extern __shared__ int16_t s_map[];
    const int tx = threadIdx.x % p.tx_dim, ty = threadIdx.x / p.tx_dim;
    const int upb = p.bs / 16;                       // uint4 units per luma block row == uint2 units per chroma block row
    const int cols = (STRETCH ? p.Bx : p.small_bx) * upb;
    for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
        const int t = unit / p.By;
        const int by = unit - t * p.By;
        build_map<STRETCH>(p.mask + ((int64_t)t * p.By + by) * p.Bx, p.Bx, p.small_bx, s_map);
        if (ty < p.ty_dim) {
            copy_block_row<uint4, !STRETCH>(p.src[0] + (int64_t)t * p.src_frame[0] + (int64_t)by * p.bs * p.src_row[0], p.src_row[0],
                                  p.dst[0] + (int64_t)t * p.dst_frame[0] + (int64_t)by * p.bs * p.dst_row[0], p.dst_row[0],
                                  s_map, cols, upb, p.bs, tx, ty, p.tx_dim, p.ty_dim);
#pragma unroll
            for (int c = 1; c < 3; ++c)
                copy_block_row<uint2, !STRETCH>(p.src[c] + (int64_t)t * p.src_frame[c] + (int64_t)by * (p.bs / 2) * p.src_row[c], p.src_row[c],
                                      p.dst[c] + (int64_t)t * p.dst_frame[c] + (int64_t)by * (p.bs / 2) * p.dst_row[c], p.dst_row[c],
                                      s_map, cols, upb, p.bs / 2, tx, ty, p.tx_dim, p.ty_dim);
        }
        if (unit + (int)gridDim.x < p.n_units) __syncthreads();   // the column map is rebuilt for the next unit
    }
}

template <bool STRETCH>
int launch_move_yuv(const elvis_plane* src, const elvis_plane* dst, int T, int bs, int By, int Bx, int small_bx,
                    const uint8_t* mask, int ctas_per_sm, cudaStream_t st) {
    if (!src || !dst || !mask || T <= 0 || By <= 0 || Bx <= 0 || small_bx < 0 || small_bx > Bx || Bx > 32767) return ELVIS_ERR_INVALID_ARG;
    if (bs <= 0 || bs % 16) return ELVIS_ERR_UNSUPPORTED;
    MoveYuvParams p;
    for (int c = 0; c < 3; ++c) {
        const int pb = c == 0 ? bs : bs / 2, al = c == 0 ? 16 : 8;
        const elvis_plane* full = STRETCH ? &dst[c] : &src[c];
        const elvis_plane* small = STRETCH ? &src[c] : &dst[c];
        if (!plane_ok(&src[c]) || !plane_ok(&dst[c]) || src[c].channels != 1 || dst[c].channels != 1) return ELVIS_ERR_INVALID_ARG;
        if (full->height < By * pb || full->width < Bx * pb) return ELVIS_ERR_SHAPE;
        if (small->height < By * pb || small->width < small_bx * pb) return ELVIS_ERR_SHAPE;
        if (vector_unit(&src[c], pb) < al || vector_unit(&dst[c], pb) < al) return ELVIS_ERR_UNSUPPORTED;   // caller falls back per plane
        p.src[c] = static_cast<const uint8_t*>(src[c].data);
        p.dst[c] = static_cast<uint8_t*>(dst[c].data);
        p.src_frame[c] = src[c].frame_stride;
        p.src_row[c] = src[c].row_stride;
        p.dst_frame[c] = dst[c].frame_stride;
        p.dst_row[c] = dst[c].row_stride;
    }
    if (small_bx == 0 && !STRETCH) return ELVIS_OK;
    p.mask = mask;
    p.By = By;
    p.Bx = Bx;
    p.small_bx = small_bx;
    p.bs = bs;
    const int cols = (STRETCH ? Bx : small_bx) * (bs / 16);
    int tx_dim = ((cols + 31) / 32) * 32;
    if (tx_dim > kThreads) tx_dim = kThreads;
    if (tx_dim < 32) tx_dim = 32;
    p.tx_dim = tx_dim;
    p.ty_dim = kThreads / tx_dim;
    if (p.ty_dim > bs / 2) p.ty_dim = bs / 2;
    p.n_units = (int)((int64_t)T * By);
    unsigned grid = (unsigned)p.n_units;
    if (ctas_per_sm > 0 && (unsigned)(ctas_per_sm * kNumSMs) < grid) grid = (unsigned)(ctas_per_sm * kNumSMs);
    const size_t smem = sizeof(int16_t) * (size_t)(Bx > small_bx ? Bx : small_bx);
    move_yuv420_kernel<STRETCH><<<grid, kThreads, smem, st>>>(p);
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

extern "C" int elvis_shrink_yuv420(const elvis_plane* src_yuv, const elvis_plane* dst_yuv, int32_t n_frames,
                                   int32_t block_size, int32_t by, int32_t bx, int32_t out_bx,
                                   const uint8_t* mask, int32_t ctas_per_sm, elvis_stream_t stream) {
    return elvis::launch_move_yuv<false>(src_yuv, dst_yuv, n_frames, block_size, by, bx, out_bx, mask, ctas_per_sm, elvis::as_stream(stream));
}

extern "C" int elvis_stretch_yuv420(const elvis_plane* src_yuv, const elvis_plane* dst_yuv, int32_t n_frames,
                                    int32_t block_size, int32_t by, int32_t bx, int32_t shrunk_bx,
                                    const uint8_t* mask, int32_t ctas_per_sm, elvis_stream_t stream) {
    return elvis::launch_move_yuv<true>(src_yuv, dst_yuv, n_frames, block_size, by, bx, shrunk_bx, mask, ctas_per_sm, elvis::as_stream(stream));
}
