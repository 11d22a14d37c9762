// The parts the v2 per-block degradations (a8-a12, a14) share: plane geometry and the copy-through of partial blocks.
// The kernels live in blur.cu (rounds of the 5x5 Gaussian of the isolated block), downsample.cu (INTER_AREA down +
// INTER_LINEAR / LANCZOS4 up of the isolated block) and dampen.cu (per-8x8-tile DCT coefficient attenuation).
#include "degrade_common.cuh"

namespace elvis {
namespace {

// ---------------------------------------------------------- copy-through of partial blocks
// copies the right strip (x >= Bx*pb) and the bottom strip (y >= By*pb) of every frame
__global__ void __launch_bounds__(256) copy_edges_kernel(const BlockGeom g) {
    const int64_t row_bytes = (int64_t)g.width * g.C;
    const int64_t x0 = (int64_t)g.Bx * g.pb * g.C;
    const int y0 = g.By * g.pb;
    const int64_t right = row_bytes - x0;                 // bytes per row in the right strip
    const int64_t per_frame = right * y0 + row_bytes * (g.height - y0);
    const int64_t total = per_frame * g.T;
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
        g.dst[(int64_t)t * g.dst_frame + y * g.dst_row + x] = g.src[(int64_t)t * g.src_frame + y * g.src_row + x];
    }
}

}  // namespace

int make_geom(const elvis_plane* src, const elvis_plane* dst, int T, int pb, int By, int Bx, BlockGeom& g) {
    if (!plane_ok(src) || !plane_ok(dst) || T <= 0 || pb <= 0 || By <= 0 || Bx <= 0) return ELVIS_ERR_INVALID_ARG;
    if (src->channels != dst->channels || src->height != dst->height || src->width != dst->width) return ELVIS_ERR_INVALID_ARG;
    if (src->height < By * pb || src->width < Bx * pb) return ELVIS_ERR_SHAPE;
    g.src = static_cast<const uint8_t*>(src->data);
    g.dst = static_cast<uint8_t*>(dst->data);
    g.src_frame = src->frame_stride;
    g.src_row = src->row_stride;
    g.dst_frame = dst->frame_stride;
    g.dst_row = dst->row_stride;
    g.T = T;
    g.By = By;
    g.Bx = Bx;
    g.pb = pb;
    g.C = src->channels;
    g.height = src->height;
    g.width = src->width;
    return ELVIS_OK;
}

int copy_edges(const BlockGeom& g, cudaStream_t st) {
    if (g.height == g.By * g.pb && g.width == g.Bx * g.pb) return ELVIS_OK;
    copy_edges_kernel<<<kNumSMs * 4, 256, 0, st>>>(g);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

}  // namespace elvis
