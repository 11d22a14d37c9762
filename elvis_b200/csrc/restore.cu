// the OpenCV client restorer -- per-block unsharp mask driven by the level map
// (elvis.py:2822-2867 restore_blur_opencv_unsharp_mask; utils.py:1253-1392
// restore_with_opencv_lanczos / restore_with_opencv_unsharp, which both run this arithmetic):
//     blurred = cv2.GaussianBlur(tile, (0, 0), sigma = level)        tile = block (+ halo, clamped)
//     out     = cv2.addWeighted(tile, 1 + level/2, blurred, -level/2, 0)
// Bit-exact restatement (oracle/spec_cv.py: gaussian_kernel_q8 / gaussian_blur_sigma / unsharp):
// cv2's u8 Gaussian is a separable 8.8 fixed-point kernel (ksize = 6 sigma + 1, built on the host
// by elvis_b200/_tables.py) with REFLECT_101 at the tile edge and one rounding (acc + 2^15) >> 16;
// 2*out = 2x + level*(x - blurred) is an integer, so addWeighted's float rounding is a
// round-half-to-even of that integer over 2, then saturation.
#include "common.cuh"

namespace elvis {
namespace {

struct RestoreParams {
    const uint8_t* src;
    uint8_t* dst;
    int64_t src_frame, src_row, dst_frame, dst_row;
    int32_t T, By, Bx, pb, C, height, width, halo;
    const int32_t* levels;
    const int32_t* kernels;   // [max_level + 1][kstride]: {ksize, q[0..ksize)}
    int32_t max_level, kstride;
};

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        if (i >= n) i = 2 * (n - 1) - i;
    }
    return i;
}

// one warp per (block, channel); tile (u8) and row-pass results (u16, block columns only) in smem
__global__ void __launch_bounds__(256) unsharp_kernel(const RestoreParams p, int warps_per_cta, int tile_bytes) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int side = p.pb + 2 * p.halo;
    uint8_t* tile = smem + (size_t)w * tile_bytes;
    uint16_t* hp = reinterpret_cast<uint16_t*>(tile + ((side * side + 15) & ~15));

    const int64_t units = (int64_t)p.T * p.By * p.Bx * p.C;
    for (int64_t unit = (int64_t)blockIdx.x * warps_per_cta + w; unit < units; unit += (int64_t)gridDim.x * warps_per_cta) {
        const int c = (int)(unit % p.C);
        int64_t b = unit / p.C;
        const int bx = (int)(b % p.Bx);
        b /= p.Bx;
        const int by = (int)(b % p.By), t = (int)(b / p.By);
        const int level = p.levels[((int64_t)t * p.By + by) * p.Bx + bx];
        const int y = by * p.pb, x = bx * p.pb;
        const uint8_t* sf = p.src + (int64_t)t * p.src_frame + c;
        uint8_t* df = p.dst + (int64_t)t * p.dst_frame + c;
        if (level <= 0) {   // untouched block: copy through
            for (int i = lane; i < p.pb * p.pb; i += 32) {
                const int yy = y + i / p.pb, xx = x + i % p.pb;
                df[(int64_t)yy * p.dst_row + (int64_t)xx * p.C] = sf[(int64_t)yy * p.src_row + (int64_t)xx * p.C];
            }
            continue;
        }
        const int y0 = max(0, y - p.halo), x0 = max(0, x - p.halo);
        const int th = min(p.height, y + p.pb + p.halo) - y0, tw = min(p.width, x + p.pb + p.halo) - x0;
        const int cy = y - y0, cx = x - x0;
        for (int i = lane; i < th * tw; i += 32) {
            const int ty = i / tw, tx = i - ty * tw;
            tile[i] = sf[(int64_t)(y0 + ty) * p.src_row + (int64_t)(x0 + tx) * p.C];
        }
        __syncwarp();
        const int lk = level > p.max_level ? p.max_level : level;
        const int32_t* kr = p.kernels + (size_t)lk * p.kstride;
        const int ksize = kr[0], r = ksize >> 1;
        for (int i = lane; i < th * p.pb; i += 32) {          // row pass on the block's columns, all tile rows
            const int ty = i / p.pb, j = i - ty * p.pb;
            int acc = 0;
            for (int d = 0; d < ksize; ++d) acc += kr[1 + d] * tile[ty * tw + reflect101(cx + j + d - r, tw)];
            hp[i] = (uint16_t)acc;
        }
        __syncwarp();
        for (int i = lane; i < p.pb * p.pb; i += 32) {        // column pass + unsharp on the block's pixels
            const int yy = i / p.pb, j = i - yy * p.pb;
            int acc = 0;
            for (int d = 0; d < ksize; ++d) acc += kr[1 + d] * hp[reflect101(cy + yy + d - r, th) * p.pb + j];
            const int blurred = (acc + 32768) >> 16;
            const int px = tile[(cy + yy) * tw + cx + j];
            const int n = 2 * px + level * (px - blurred);
            const int half = n >> 1;                           // floor(n / 2), also for negative n
            int v = half + ((n & 1) & (half & 1));             // ties to even
            v = v < 0 ? 0 : (v > 255 ? 255 : v);
            df[(int64_t)(y + yy) * p.dst_row + (int64_t)(x + j) * p.C] = (uint8_t)v;
        }
        __syncwarp();
    }
}

// copies the right strip (x >= Bx*pb) and the bottom strip (y >= By*pb) of every frame
__global__ void __launch_bounds__(256) restore_copy_edges_kernel(const RestoreParams p) {
    const int64_t row_bytes = (int64_t)p.width * p.C;
    const int64_t x0 = (int64_t)p.Bx * p.pb * p.C;
    const int y0 = p.By * p.pb;
    const int64_t right = row_bytes - x0;
    const int64_t per_frame = right * y0 + row_bytes * (p.height - y0);
    const int64_t total = per_frame * p.T;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int t = (int)(i / per_frame);
        int64_t r = i - (int64_t)t * per_frame;
        int64_t y, x;
        if (r < right * y0) {
            y = r / right;
            x = x0 + (r - y * right);
        } else {
            r -= right * y0;
            y = y0 + r / row_bytes;
            x = r % row_bytes;
        }
        p.dst[(int64_t)t * p.dst_frame + y * p.dst_row + x] = p.src[(int64_t)t * p.src_frame + y * p.src_row + x];
    }
}

// out[t] = uint8(tb * out[t-1] + (1 - tb) * out[t]) for t >= 1, in place, float64 like NumPy
// (utils.py:1310-1311); one thread per byte position walks the frames.
__global__ void __launch_bounds__(256) temporal_blend_kernel(uint8_t* data, int64_t frame_stride, int64_t row_stride,
                                                              int64_t row_bytes, int rows, int T, double tb, double one_minus_tb) {
    const int64_t n = row_bytes * rows;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const int64_t y = i / row_bytes, x = i - y * row_bytes;
        uint8_t* q = data + y * row_stride + x;
        double prev = (double)q[0];
        for (int t = 1; t < T; ++t) {
            uint8_t* cur = q + (int64_t)t * frame_stride;
            const double v = __dadd_rn(__dmul_rn(tb, prev), __dmul_rn(one_minus_tb, (double)*cur));
            const uint8_t o = (uint8_t)(int)v;   // truncation, as .astype(np.uint8) of a value in [0, 255]
            *cur = o;
            prev = (double)o;
        }
    }
}

}  // namespace
}  // namespace elvis

using namespace elvis;

extern "C" int elvis_restore_unsharp(const elvis_plane* src, const elvis_plane* dst, int32_t n_frames,
                                     int32_t block_px, int32_t by, int32_t bx, const int32_t* levels,
                                     int32_t halo, const int32_t* kernels, int32_t max_level, int32_t kernel_stride,
                                     elvis_stream_t stream) {
    if (!plane_ok(src) || !plane_ok(dst) || !levels || !kernels || n_frames <= 0 || block_px <= 0 || by <= 0 || bx <= 0)
        return ELVIS_ERR_INVALID_ARG;
    if (halo < 0 || max_level < 1 || kernel_stride < 6 * max_level + 2) return ELVIS_ERR_INVALID_ARG;
    if (src->channels != dst->channels || src->height != dst->height || src->width != dst->width) return ELVIS_ERR_INVALID_ARG;
    if (src->height < by * block_px || src->width < bx * block_px) return ELVIS_ERR_SHAPE;
    const int side = block_px + 2 * halo;
    const int tile_bytes = ((side * side + 15) & ~15) + ((side * block_px * 2 + 15) & ~15);
    if (tile_bytes > 48 * 1024) return ELVIS_ERR_UNSUPPORTED;
    RestoreParams p;
    p.src = static_cast<const uint8_t*>(src->data);
    p.dst = static_cast<uint8_t*>(dst->data);
    p.src_frame = src->frame_stride;
    p.src_row = src->row_stride;
    p.dst_frame = dst->frame_stride;
    p.dst_row = dst->row_stride;
    p.T = n_frames;
    p.By = by;
    p.Bx = bx;
    p.pb = block_px;
    p.C = src->channels;
    p.height = src->height;
    p.width = src->width;
    p.halo = halo;
    p.levels = levels;
    p.kernels = kernels;
    p.max_level = max_level;
    p.kstride = kernel_stride;
    cudaStream_t st = as_stream(stream);
    if (p.height != by * block_px || p.width != bx * block_px) {
        restore_copy_edges_kernel<<<kNumSMs * 4, 256, 0, st>>>(p);
        ELVIS_CHECK_LAUNCH();
    }
    int wpc = 8;
    while (wpc > 1 && (size_t)wpc * tile_bytes > 48 * 1024) wpc >>= 1;
    const int64_t units = (int64_t)n_frames * by * bx * p.C;
    int64_t grid = (units + wpc - 1) / wpc;
    if (grid > (int64_t)kNumSMs * 16) grid = (int64_t)kNumSMs * 16;
    unsharp_kernel<<<(unsigned)grid, wpc * 32, (size_t)wpc * tile_bytes, st>>>(p, wpc, tile_bytes);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_temporal_blend(const elvis_plane* clip, int32_t n_frames, double temporal_blend, elvis_stream_t stream) {
    if (!plane_ok(clip) || n_frames <= 0) return ELVIS_ERR_INVALID_ARG;
    if (n_frames < 2 || !(temporal_blend > 0)) return ELVIS_OK;
    const int64_t row_bytes = (int64_t)clip->width * clip->channels;
    const int64_t n = row_bytes * clip->height;
    int64_t grid = (n + 255) / 256;
    if (grid > (int64_t)kNumSMs * 16) grid = (int64_t)kNumSMs * 16;
    temporal_blend_kernel<<<(unsigned)grid, 256, 0, as_stream(stream)>>>(static_cast<uint8_t*>(clip->data), clip->frame_stride,
                                                                       clip->row_stride, row_bytes, clip->height, n_frames,
                                                                       temporal_blend, 1.0 - temporal_blend);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}
