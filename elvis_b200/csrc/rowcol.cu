// row+column shrink variants (utils.py:763-1018).
//
// The reference alternates passes over an ever smaller block grid: a row pass removes the least
// important block (first minimum) of every row and shifts the row left, a column pass does the
// same per column and shifts up, until int(By*Bx*shrink) blocks are gone.  Rows (columns) of one
// pass are independent, passes are sequential -- so ONE CTA simulates one frame: a warp per row
// (column) finds the arg-min with a shuffle reduction over order-preserving keys and shifts its
// row of {key, original position} in place; __syncthreads() separates passes.  Only block
// indices move here; the pixels are moved once at the end by gather_blocks_kernel.
#include "common.cuh"

namespace elvis {
namespace {

constexpr int kPlanThreads = 1024;

__device__ __forceinline__ unsigned long long sortable_key(double v) {
    if (v != v) return ~0ULL;                      // np.argmin returns the first NaN; not supported: NaN ranks last
    v = __dadd_rn(v, 0.0);
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}

// first minimum of n strided elements (key[i * stride]) by the whole warp -> index
__device__ __forceinline__ int warp_argmin(const unsigned long long* key, int n, int64_t stride, int lane) {
    unsigned long long best = ~0ULL;
    int bi = 0x7fffffff;
    for (int i = lane; i < n; i += 32) {
        const unsigned long long k = key[(int64_t)i * stride];
        if (k < best || bi == 0x7fffffff) {
            if (k < best || bi == 0x7fffffff) { best = k; bi = i; }
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        const unsigned long long ok = __shfl_xor_sync(0xffffffffu, best, m);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, m);
        if (ok < best || (ok == best && oi < bi)) { best = ok; bi = oi; }
    }
    return bi;
}

// remove element `at` from a strided vector of length n (shift the tail towards it), whole warp
__device__ __forceinline__ void warp_remove(unsigned long long* key, int32_t* pos, int at, int n, int64_t stride, int lane) {
    for (int c0 = at; c0 < n - 1; c0 += 32) {
        const int c = c0 + lane;
        unsigned long long k = 0;
        int32_t q = 0;
        const bool on = c < n - 1;
        if (on) {
            k = key[(int64_t)(c + 1) * stride];
            q = pos[(int64_t)(c + 1) * stride];
        }
        __syncwarp();
        if (on) {
            key[(int64_t)c * stride] = k;
            pos[(int64_t)c * stride] = q;
        }
        __syncwarp();
    }
}

struct PlanParams {
    const double* importance;
    unsigned long long* keys;
    int32_t* pos;
    uint8_t* mask;
    int32_t* pass_idx;
    int32_t* pass_cnt;
    int32_t* meta;
    int32_t By, Bx, max_passes, lmax;
    int64_t target;
};

__global__ void __launch_bounds__(kPlanThreads) rowcol_plan_kernel(const PlanParams p) {
    const int t = blockIdx.x;
    const int64_t frame = (int64_t)p.By * p.Bx;
    unsigned long long* key = p.keys + t * frame;
    int32_t* pos = p.pos + t * frame;
    uint8_t* mask = p.mask + t * frame;
    int32_t* pidx = p.pass_idx + (int64_t)t * p.max_passes * p.lmax;
    int32_t* pcnt = p.pass_cnt + (int64_t)t * p.max_passes;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = kPlanThreads / 32;
    for (int64_t i = threadIdx.x; i < frame; i += kPlanThreads) {
        key[i] = sortable_key(p.importance[t * frame + i]);
        pos[i] = (int32_t)i;
        mask[i] = 0;
    }
    __syncthreads();
    int by = p.By, bx = p.Bx, n_pass = 0;
    int64_t removed = 0;
    while (removed < p.target && by > 0 && bx > 0 && n_pass + 2 <= p.max_passes) {
        int n = (int)min((int64_t)by, p.target - removed);          // row pass
        for (int r = warp; r < n; r += nwarps) {
            unsigned long long* krow = key + (int64_t)r * p.Bx;
            int32_t* prow = pos + (int64_t)r * p.Bx;
            const int least = warp_argmin(krow, bx, 1, lane);
            if (lane == 0) {
                mask[prow[least]] = 1;
                pidx[(int64_t)n_pass * p.lmax + r] = least;
            }
            __syncwarp();
            warp_remove(krow, prow, least, bx, 1, lane);
        }
        if (threadIdx.x == 0) pcnt[n_pass] = n;
        ++n_pass;
        removed += n;
        if (n == by) --bx;
        __syncthreads();
        if (removed >= p.target || bx <= 0) break;
        n = (int)min((int64_t)bx, p.target - removed);              // column pass
        for (int c = warp; c < n; c += nwarps) {
            const int least = warp_argmin(key + c, by, p.Bx, lane);
            if (lane == 0) {
                mask[pos[(int64_t)least * p.Bx + c]] = 1;
                pidx[(int64_t)n_pass * p.lmax + c] = least;
            }
            __syncwarp();
            warp_remove(key + c, pos + c, least, by, p.Bx, lane);
        }
        if (threadIdx.x == 0) pcnt[n_pass] = n;
        ++n_pass;
        removed += n;
        if (n == bx) --by;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        p.meta[t * 4 + 0] = n_pass;
        p.meta[t * 4 + 1] = by;
        p.meta[t * 4 + 2] = bx;
        p.meta[t * 4 + 3] = (int32_t)removed;
    }
}

// Replays recorded passes in reverse on a grid of shrunk-block indices (utils.py:951-1018):
// a row pass widens the grid by one column, inserting -1 (a black block) at the recorded position
// of every recorded row; rows beyond the record keep their blocks and get -1 appended.
struct ExpandParams {
    const int32_t* pass_idx;
    const int32_t* pass_cnt;
    int32_t* grid;          // (T, gh, gw) pitch gw
    int32_t n_passes, lmax, sby, sbx, gh, gw;
};

__global__ void __launch_bounds__(kPlanThreads) rowcol_expand_kernel(const ExpandParams p) {
    const int t = blockIdx.x;
    int32_t* G = p.grid + (int64_t)t * p.gh * p.gw;
    const int32_t* pidx = p.pass_idx + (int64_t)t * p.n_passes * p.lmax;
    const int32_t* pcnt = p.pass_cnt + (int64_t)t * p.n_passes;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = kPlanThreads / 32;
    for (int i = threadIdx.x; i < p.gh * p.gw; i += kPlanThreads) {
        const int y = i / p.gw, x = i - y * p.gw;
        G[i] = (y < p.sby && x < p.sbx) ? y * p.sbx + x : -1;
    }
    __syncthreads();
    int gy = p.sby, gx = p.sbx;
    for (int ps = p.n_passes - 1; ps >= 0; --ps) {
        const int cnt = pcnt[ps];
        const int32_t* idx = pidx + (int64_t)ps * p.lmax;
        const bool row_pass = (ps & 1) == 0;
        const int n_vec = row_pass ? gy : gx;           // vectors (rows / columns) in the grid
        const int len = row_pass ? gx : gy;             // their current length
        const int64_t stride = row_pass ? 1 : p.gw;
        if (len + 1 > (row_pass ? p.gw : p.gh)) break;  // cannot happen for well-formed input
        for (int v = warp; v < n_vec; v += nwarps) {
            int32_t* vec = row_pass ? G + (int64_t)v * p.gw : G + v;
            if (v < cnt) {
                const int k = min(idx[v], len);
                // shift [k, len) one step away from the origin, highest chunk first
                for (int hi = len; hi > k; hi -= 32) {
                    const int c = hi - 1 - lane;
                    int32_t val = 0;
                    const bool on = c >= k;
                    if (on) val = vec[(int64_t)c * stride];
                    __syncwarp();
                    if (on) vec[(int64_t)(c + 1) * stride] = val;
                    __syncwarp();
                }
                if (lane == 0) vec[(int64_t)k * stride] = -1;
            } else if (lane == 0) {
                vec[(int64_t)len * stride] = -1;
            }
        }
        if (row_pass) ++gx; else ++gy;
        __syncthreads();
    }
}

// inv[orig] = last (row-major) shrunk index whose position_map entry is orig  (utils.py:851-855)
__global__ void __launch_bounds__(256) invert_map_kernel(const int32_t* __restrict__ map, int64_t n_per_frame, int T,
                                                         int32_t* __restrict__ inv, int64_t inv_per_frame) {
    const int64_t total = n_per_frame * T;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int64_t t = i / n_per_frame;
        const int32_t o = map[i];
        if (o >= 0 && o < inv_per_frame) atomicMax(inv + t * inv_per_frame + o, (int32_t)(i - t * n_per_frame));
    }
}

// dst block (j, i) <- src block map[t][j][i] (linear index over a src grid `src_bx` blocks wide), zeros when < 0
struct GatherParams {
    const uint8_t* src;
    uint8_t* dst;
    int64_t src_frame, src_row, dst_frame, dst_row;
    const int32_t* map;
    int32_t map_pitch, map_rows;   // map is (T, map_rows, map_pitch)
    int32_t dby, dbx, src_bx, src_blocks, pb, row_bytes_per_block;
};

template <typename V>
__device__ __forceinline__ V zero_vec();
template <> __device__ __forceinline__ uint8_t zero_vec<uint8_t>() { return 0; }
template <> __device__ __forceinline__ uint16_t zero_vec<uint16_t>() { return 0; }
template <> __device__ __forceinline__ uint32_t zero_vec<uint32_t>() { return 0; }
template <> __device__ __forceinline__ uint2 zero_vec<uint2>() { return make_uint2(0, 0); }
template <> __device__ __forceinline__ uint4 zero_vec<uint4>() { return make_uint4(0, 0, 0, 0); }

// One CTA per (frame, destination block row): thread x walks the row's vector units, thread y the
// pixel rows of the block, so a warp reads/writes contiguous runs of a block row.
template <typename V>
__global__ void __launch_bounds__(256) gather_blocks_kernel(const GatherParams p) {
    const int t = blockIdx.x / p.dby, j = blockIdx.x % p.dby;
    const int32_t* mrow = p.map + ((int64_t)t * p.map_rows + j) * p.map_pitch;
    const int upb = p.row_bytes_per_block / (int)sizeof(V);        // vector units per block row
    const int cols = p.dbx * upb;
    const int64_t srow = p.src_row / (int64_t)sizeof(V), drow = p.dst_row / (int64_t)sizeof(V);
    const V* sframe = reinterpret_cast<const V*>(p.src + (int64_t)t * p.src_frame);
    V* dbase = reinterpret_cast<V*>(p.dst + (int64_t)t * p.dst_frame + (int64_t)j * p.pb * p.dst_row);
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    for (int c = tx; c < cols; c += 64) {
        const int i = c / upb, u = c - i * upb;
        const int32_t s = mrow[i];
        const bool live = s >= 0 && s < p.src_blocks;
        const int sy = live ? s / p.src_bx : 0, sx = live ? s - sy * p.src_bx : 0;
        const V* sp = sframe + (int64_t)sy * p.pb * srow + (int64_t)sx * upb + u;
        for (int r = ty; r < p.pb; r += 4) {
            V v = zero_vec<V>();
            if (live) v = sp[(int64_t)r * srow];
            dbase[(int64_t)r * drow + c] = v;
        }
    }
}


// Row-major refill map of presley.py:787-827 (stretch_video_frames): the i-th kept block of a frame
// (mask == 0, row-major order) takes shrunk block i -- wherever its row and column fall -- as long
// as i < capacity = shrunk_by * shrunk_bx; removed blocks and kept blocks beyond the capacity stay
// black (-1).  One CTA per frame: ballot prefix counts inside a warp, warp totals through smem.
constexpr int kRefillThreads = 256;
__global__ void __launch_bounds__(kRefillThreads) refill_map_kernel(const uint8_t* __restrict__ mask, int n, int capacity,
                                                                    int32_t* __restrict__ map) {
    __shared__ int warp_total[kRefillThreads / 32];
    __shared__ int carry;
    const uint8_t* m = mask + (int64_t)blockIdx.x * n;
    int32_t* out = map + (int64_t)blockIdx.x * n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += kRefillThreads) {
        const int i = base + threadIdx.x;
        const bool kept = i < n && m[i] == 0;
        const unsigned b = __ballot_sync(0xffffffffu, kept);
        if (lane == 0) warp_total[warp] = __popc(b);
        __syncthreads();
        int before = carry + __popc(b & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) before += warp_total[w];
        if (i < n) out[i] = (kept && before < capacity) ? before : -1;
        __syncthreads();
        if (threadIdx.x == kRefillThreads - 1) carry = before + (kept ? 1 : 0);
        __syncthreads();
    }
}

}  // namespace
}  // namespace elvis

using namespace elvis;

extern "C" int elvis_rowcol_plan(const double* importance, int32_t n_frames, int32_t by, int32_t bx, int64_t target_removals,
                                 uint64_t* scratch_keys, int32_t* position, uint8_t* mask, int32_t* pass_indices,
                                 int32_t* pass_counts, int32_t max_passes, int32_t* meta, elvis_stream_t stream) {
    if (!importance || !scratch_keys || !position || !mask || !pass_indices || !pass_counts || !meta) return ELVIS_ERR_INVALID_ARG;
    if (n_frames <= 0 || by <= 0 || bx <= 0 || target_removals < 0 || max_passes < 2) return ELVIS_ERR_INVALID_ARG;
    PlanParams p;
    p.importance = importance;
    p.keys = reinterpret_cast<unsigned long long*>(scratch_keys);
    p.pos = position;
    p.mask = mask;
    p.pass_idx = pass_indices;
    p.pass_cnt = pass_counts;
    p.meta = meta;
    p.By = by;
    p.Bx = bx;
    p.max_passes = max_passes;
    p.lmax = by > bx ? by : bx;
    p.target = target_removals;
    rowcol_plan_kernel<<<n_frames, kPlanThreads, 0, as_stream(stream)>>>(p);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_rowcol_expand(const int32_t* pass_indices, const int32_t* pass_counts, int32_t n_frames, int32_t n_passes,
                                   int32_t lmax, int32_t shrunk_by, int32_t shrunk_bx, int32_t* grid, int32_t grid_h,
                                   int32_t grid_w, elvis_stream_t stream) {
    if (!grid || n_frames <= 0 || n_passes < 0 || shrunk_by < 0 || shrunk_bx < 0 || grid_h < shrunk_by || grid_w < shrunk_bx)
        return ELVIS_ERR_INVALID_ARG;
    if (n_passes > 0 && (!pass_indices || !pass_counts || lmax <= 0)) return ELVIS_ERR_INVALID_ARG;
    ExpandParams p;
    p.pass_idx = pass_indices;
    p.pass_cnt = pass_counts;
    p.grid = grid;
    p.n_passes = n_passes;
    p.lmax = lmax;
    p.sby = shrunk_by;
    p.sbx = shrunk_bx;
    p.gh = grid_h;
    p.gw = grid_w;
    rowcol_expand_kernel<<<n_frames, kPlanThreads, 0, as_stream(stream)>>>(p);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_invert_block_map(const int32_t* map, int32_t n_frames, int64_t entries_per_frame, int32_t* inverse,
                                      int64_t inverse_per_frame, elvis_stream_t stream) {
    if (!map || !inverse || n_frames <= 0 || entries_per_frame <= 0 || inverse_per_frame <= 0) return ELVIS_ERR_INVALID_ARG;
    cudaStream_t st = as_stream(stream);
    cudaError_t e = cudaMemsetAsync(inverse, 0xff, sizeof(int32_t) * (size_t)inverse_per_frame * n_frames, st);
    if (e != cudaSuccess) return cuda_fail(e);
    int64_t grid = (entries_per_frame * n_frames + 255) / 256;
    if (grid > (int64_t)kNumSMs * 8) grid = (int64_t)kNumSMs * 8;
    invert_map_kernel<<<(unsigned)grid, 256, 0, st>>>(map, entries_per_frame, n_frames, inverse, inverse_per_frame);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_gather_blocks(const elvis_plane* src, const elvis_plane* dst, int32_t n_frames, int32_t block_px,
                                   int32_t dst_by, int32_t dst_bx, int32_t src_by, int32_t src_bx, const int32_t* map,
                                   int32_t map_rows, int32_t map_pitch, elvis_stream_t stream) {
    if (!plane_ok(src) || !plane_ok(dst) || !map || n_frames <= 0 || block_px <= 0) return ELVIS_ERR_INVALID_ARG;
    if (dst_by <= 0 || dst_bx <= 0 || src_by <= 0 || src_bx <= 0 || map_rows < dst_by || map_pitch < dst_bx) return ELVIS_ERR_INVALID_ARG;
    if (src->channels != dst->channels) return ELVIS_ERR_INVALID_ARG;
    if (src->height < src_by * block_px || src->width < src_bx * block_px) return ELVIS_ERR_SHAPE;
    if (dst->height < dst_by * block_px || dst->width < dst_bx * block_px) return ELVIS_ERR_SHAPE;
    GatherParams p;
    p.src = static_cast<const uint8_t*>(src->data);
    p.dst = static_cast<uint8_t*>(dst->data);
    p.src_frame = src->frame_stride;
    p.src_row = src->row_stride;
    p.dst_frame = dst->frame_stride;
    p.dst_row = dst->row_stride;
    p.map = map;
    p.map_pitch = map_pitch;
    p.map_rows = map_rows;
    p.dby = dst_by;
    p.dbx = dst_bx;
    p.src_bx = src_bx;
    p.src_blocks = src_by * src_bx;
    p.pb = block_px;
    p.row_bytes_per_block = block_px * src->channels;
    int unit = vector_unit(src, p.row_bytes_per_block);
    const int unit_dst = vector_unit(dst, p.row_bytes_per_block);
    if (unit_dst < unit) unit = unit_dst;
    const unsigned grid = (unsigned)((int64_t)n_frames * dst_by);
    cudaStream_t st = as_stream(stream);
    switch (unit) {
        case 16: gather_blocks_kernel<uint4><<<grid, 256, 0, st>>>(p); break;
        case 8: gather_blocks_kernel<uint2><<<grid, 256, 0, st>>>(p); break;
        case 4: gather_blocks_kernel<uint32_t><<<grid, 256, 0, st>>>(p); break;
        case 2: gather_blocks_kernel<uint16_t><<<grid, 256, 0, st>>>(p); break;
        default: gather_blocks_kernel<uint8_t><<<grid, 256, 0, st>>>(p); break;
    }
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_refill_map(const uint8_t* mask, int32_t n_frames, int64_t blocks_per_frame, int64_t capacity, int32_t* map,
                                elvis_stream_t stream) {
    if (!mask || !map || n_frames <= 0 || blocks_per_frame <= 0 || blocks_per_frame > INT32_MAX || capacity < 0) return ELVIS_ERR_INVALID_ARG;
    const int cap = capacity > blocks_per_frame ? (int)blocks_per_frame : (int)capacity;
    refill_map_kernel<<<n_frames, kRefillThreads, 0, as_stream(stream)>>>(mask, (int)blocks_per_frame, cap, map);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}
