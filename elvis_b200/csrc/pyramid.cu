// the deterministic parts either side of the external neural restorer.
//   * upscale_realesrgan_adaptive (elvis.py:2522-2600): the multi-stage pyramid.  Per stage the
//     frame is upscaled 2x by a pluggable upsampler (external), then every block whose own
//     downscale factor is <= the stage's factor is replaced by the INTER_AREA-downscaled original.
//     Here: the power-of-two INTER_AREA downscale of a whole u8 frame (cv2's integer-ratio rules:
//     (a+b+c+d+2)>>2 for 2x2, round-half-even of sum * float32(1/area) above -- oracle/spec_cv.py
//     resize_area, pinned against cv2) and the per-block merge.
//   * the level-map video side channel (elvis.py:2198-2245): maps -> 8-bit gray for the encoder
//     and back (the encode / decode itself is external).
#include "common.cuh"

namespace elvis {
namespace {

struct DownParams {
    const uint8_t* src;
    uint8_t* dst;
    int64_t src_frame, src_row, dst_frame, dst_row;
    int32_t T, dh, dw, C, f;
    float scale;
};

__global__ void __launch_bounds__(256) area_down_kernel(const DownParams p) {
    const int64_t row_vals = (int64_t)p.dw * p.C;
    const int64_t total = (int64_t)p.T * p.dh * row_vals;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int xc = (int)(i % row_vals);
        const int y = (int)((i / row_vals) % p.dh);
        const int t = (int)(i / (row_vals * p.dh));
        const int x = xc / p.C, c = xc - x * p.C;
        const uint8_t* s = p.src + (int64_t)t * p.src_frame + (int64_t)y * p.f * p.src_row + ((int64_t)x * p.f) * p.C + c;
        int sum = 0;
        for (int dy = 0; dy < p.f; ++dy)
            for (int dx = 0; dx < p.f; ++dx) sum += s[(int64_t)dy * p.src_row + dx * p.C];
        const int v = p.f == 1 ? sum : p.f == 2 ? (sum + 2) >> 2 : __float2int_rn(__fmul_rn((float)sum, p.scale));
        p.dst[(int64_t)t * p.dst_frame + (int64_t)y * p.dst_row + xc] = (uint8_t)min(max(v, 0), 255);
    }
}

struct MergeParams {
    const uint8_t* src;
    uint8_t* dst;
    int64_t src_frame, src_row, dst_frame, dst_row;
    int32_t T, By, Bx, pb, row_bytes;   // row_bytes = pb * channels
    const int32_t* factors;
    int32_t threshold;
};

// one CTA per (frame, block row): copy the blocks whose factor is <= threshold
__global__ void __launch_bounds__(256) merge_blocks_kernel(const MergeParams p) {
    const int t = blockIdx.x / p.By, by = blockIdx.x % p.By;
    const int32_t* frow = p.factors + ((int64_t)t * p.By + by) * p.Bx;
    const int64_t line = (int64_t)p.Bx * p.row_bytes;
    const uint8_t* s0 = p.src + (int64_t)t * p.src_frame + (int64_t)by * p.pb * p.src_row;
    uint8_t* d0 = p.dst + (int64_t)t * p.dst_frame + (int64_t)by * p.pb * p.dst_row;
    for (int64_t e = threadIdx.x; e < line * p.pb; e += 256) {
        const int r = (int)(e / line);
        const int64_t xb = e - r * line;
        if (frow[xb / p.row_bytes] <= p.threshold) d0[(int64_t)r * p.dst_row + xb] = s0[(int64_t)r * p.src_row + xb];
    }
}

// ((maps - min) / (max - min) * 255.0).astype(uint8), float64 (elvis.py:2200-2202)
__global__ void __launch_bounds__(256) levels_to_gray_kernel(const int32_t* __restrict__ maps, int64_t n, int32_t lo, int32_t hi,
                                                             uint8_t* __restrict__ out) {
    const double range = (double)hi - (double)lo;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double v = __dmul_rn(__ddiv_rn((double)maps[i] - (double)lo, range), 255.0);
        out[i] = (uint8_t)(int)v;    // values are within 0..255: truncation
    }
}

// round(float32(img) / 255.0 * (max - min) + min).astype(uint8), float32, half to even (elvis.py:2238-2240)
__global__ void __launch_bounds__(256) gray_to_levels_kernel(const uint8_t* __restrict__ gray, int64_t n, float lo, float range,
                                                             uint8_t* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const float v = __fadd_rn(__fmul_rn(__fdiv_rn((float)gray[i], 255.0f), range), lo);
        out[i] = (uint8_t)__float2int_rn(v);
    }
}

// cv2.resize(map, INTER_NEAREST) of block-level maps: element (dy, dx) <- src(y_idx[dy], x_idx[dx]);
// the index tables hold cv2's min(floor(d * (1 / (dsize / ssize))), ssize - 1)
template <typename E>
__global__ void __launch_bounds__(256) nearest_kernel(const E* __restrict__ src, int T, int sh, int sw, E* __restrict__ dst, int dh,
                                                      int dw, const int32_t* __restrict__ y_idx, const int32_t* __restrict__ x_idx) {
    const int64_t total = (int64_t)T * dh * dw;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int dx = (int)(i % dw);
        const int dy = (int)((i / dw) % dh);
        const int64_t t = i / ((int64_t)dw * dh);
        dst[i] = src[(t * sh + y_idx[dy]) * sw + x_idx[dx]];
    }
}

unsigned grid_of(int64_t n) {
    int64_t g = (n + 255) / 256;
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (unsigned)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace
}  // namespace elvis

using namespace elvis;

extern "C" int elvis_area_downscale(const elvis_plane* src, const elvis_plane* dst, int32_t n_frames, int32_t factor,
                                    elvis_stream_t stream) {
    if (!plane_ok(src) || !plane_ok(dst) || n_frames <= 0 || factor <= 0) return ELVIS_ERR_INVALID_ARG;
    if (src->channels != dst->channels) return ELVIS_ERR_INVALID_ARG;
    if (dst->height <= 0 || dst->width <= 0 || (int64_t)dst->height * factor > src->height || (int64_t)dst->width * factor > src->width)
        return ELVIS_ERR_SHAPE;
    DownParams p;
    p.src = static_cast<const uint8_t*>(src->data);
    p.dst = static_cast<uint8_t*>(dst->data);
    p.src_frame = src->frame_stride;
    p.src_row = src->row_stride;
    p.dst_frame = dst->frame_stride;
    p.dst_row = dst->row_stride;
    p.T = n_frames;
    p.dh = dst->height;
    p.dw = dst->width;
    p.C = src->channels;
    p.f = factor;
    p.scale = 1.0f / (float)(factor * factor);
    area_down_kernel<<<grid_of((int64_t)n_frames * p.dh * p.dw * p.C), 256, 0, as_stream(stream)>>>(p);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_merge_blocks(const elvis_plane* src, const elvis_plane* dst, int32_t n_frames, int32_t block_px, int32_t by,
                                  int32_t bx, const int32_t* factors, int32_t threshold, elvis_stream_t stream) {
    if (!plane_ok(src) || !plane_ok(dst) || !factors || n_frames <= 0 || block_px <= 0 || by <= 0 || bx <= 0) return ELVIS_ERR_INVALID_ARG;
    if (src->channels != dst->channels) return ELVIS_ERR_INVALID_ARG;
    if (src->height < by * block_px || src->width < bx * block_px || dst->height < by * block_px || dst->width < bx * block_px)
        return ELVIS_ERR_SHAPE;
    MergeParams p;
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
    p.row_bytes = block_px * src->channels;
    p.factors = factors;
    p.threshold = threshold;
    merge_blocks_kernel<<<(unsigned)((int64_t)n_frames * by), 256, 0, as_stream(stream)>>>(p);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_levels_to_gray(const int32_t* maps, int64_t n, int32_t min_value, int32_t max_value, uint8_t* gray,
                                    elvis_stream_t stream) {
    if (!maps || !gray || n <= 0 || max_value <= min_value) return ELVIS_ERR_INVALID_ARG;
    levels_to_gray_kernel<<<grid_of(n), 256, 0, as_stream(stream)>>>(maps, n, min_value, max_value, gray);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_gray_to_levels(const uint8_t* gray, int64_t n, float min_value, float max_value, uint8_t* levels,
                                    elvis_stream_t stream) {
    if (!gray || !levels || n <= 0) return ELVIS_ERR_INVALID_ARG;
    gray_to_levels_kernel<<<grid_of(n), 256, 0, as_stream(stream)>>>(gray, n, min_value, max_value - min_value, levels);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_resize_nearest(const void* src, int32_t elem_bytes, int32_t n_maps, int32_t src_h, int32_t src_w, void* dst,
                                    int32_t dst_h, int32_t dst_w, const int32_t* y_idx, const int32_t* x_idx, elvis_stream_t stream) {
    if (!src || !dst || !y_idx || !x_idx || n_maps <= 0 || src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0) return ELVIS_ERR_INVALID_ARG;
    const unsigned grid = grid_of((int64_t)n_maps * dst_h * dst_w);
    cudaStream_t st = as_stream(stream);
    if (elem_bytes == 1)
        nearest_kernel<uint8_t><<<grid, 256, 0, st>>>(static_cast<const uint8_t*>(src), n_maps, src_h, src_w, static_cast<uint8_t*>(dst), dst_h, dst_w, y_idx, x_idx);
    else if (elem_bytes == 4)
        nearest_kernel<uint32_t><<<grid, 256, 0, st>>>(static_cast<const uint32_t*>(src), n_maps, src_h, src_w, static_cast<uint32_t*>(dst), dst_h, dst_w, y_idx, x_idx);
    else if (elem_bytes == 8)
        nearest_kernel<uint64_t><<<grid, 256, 0, st>>>(static_cast<const uint64_t*>(src), n_maps, src_h, src_w, static_cast<uint64_t*>(dst), dst_h, dst_w, y_idx, x_idx);
    else
        return ELVIS_ERR_UNSUPPORTED;
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}
