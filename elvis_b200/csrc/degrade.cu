// a8-a12, a14: the v2 per-block degradations.
//   blur       -- `rounds` successive 5x5 sigma=1 Gaussian blurs of the ISOLATED block
//                 (cv2.GaussianBlur u8 fixed-point path; oracle/spec_cv.py:gaussian_blur5)
//   downsample -- cv2 INTER_AREA down + INTER_LINEAR up of the isolated block
//                 (oracle/spec_cv.py:resize_area / resize_linear), table driven
//   dampen     -- per-8x8-tile DCT coefficient attenuation (oracle/spec_dct_dampen.py)
// Integer paths are bit-exact by construction; every float op on the fractional-area path
// is an explicitly rounded fp32 op in cv2's accumulation order.
#include "common.cuh"
#include "dct8.cuh"

namespace elvis {
namespace {

struct BlockGeom {
    const uint8_t* src;
    uint8_t* dst;
    int64_t src_frame, src_row, dst_frame, dst_row;
    int32_t T, By, Bx, pb, C;
    int32_t height, width;   // full plane, for the copy-through of partial blocks
};

// unit -> (t, by, bx, c)
__device__ __forceinline__ void decode_unit(const BlockGeom& g, int64_t unit, int& t, int& by, int& bx, int& c) {
    c = (int)(unit % g.C);
    int64_t b = unit / g.C;
    bx = (int)(b % g.Bx);
    b /= g.Bx;
    by = (int)(b % g.By);
    t = (int)(b / g.By);
}

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        if (i >= n) i = 2 * (n - 1) - i;
    }
    return i;
}

// ------------------------------------------------------------------------------- blur
// one warp per (block, channel); block in shared memory as u8 plus a u16 row-pass buffer
__global__ void __launch_bounds__(256) blur_kernel(const BlockGeom g, const int32_t* __restrict__ rounds, int warps_per_cta) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int pb = g.pb, n = pb * pb;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // layout: [warps][n] u8 blocks, then (16-byte aligned) [warps][n] u16 row-pass buffers
    uint8_t* a = smem + (size_t)w * n;
    uint16_t* tmp = reinterpret_cast<uint16_t*>(smem + (((size_t)warps_per_cta * n + 15) & ~(size_t)15)) + (size_t)w * n;

    const int64_t units = (int64_t)g.T * g.By * g.Bx * g.C;
    for (int64_t unit = (int64_t)blockIdx.x * warps_per_cta + w; unit < units; unit += (int64_t)gridDim.x * warps_per_cta) {
        int t, by, bx, c;
        decode_unit(g, unit, t, by, bx, c);
        const int r = rounds[((int64_t)t * g.By + by) * g.Bx + bx];
        const uint8_t* sp = g.src + (int64_t)t * g.src_frame + (int64_t)by * pb * g.src_row + ((int64_t)bx * pb) * g.C + c;
        uint8_t* dp = g.dst + (int64_t)t * g.dst_frame + (int64_t)by * pb * g.dst_row + ((int64_t)bx * pb) * g.C + c;
        for (int i = lane; i < n; i += 32) {
            const int y = i / pb, x = i - y * pb;
            a[i] = sp[(int64_t)y * g.src_row + x * g.C];
        }
        __syncwarp();
        for (int k = 0; k < r; ++k) {
            for (int i = lane; i < n; i += 32) {
                const int y = i / pb, x = i - y * pb;
                const uint8_t* row = a + y * pb;
                const int h = 14 * (row[reflect101(x - 2, pb)] + row[reflect101(x + 2, pb)]) +
                              62 * (row[reflect101(x - 1, pb)] + row[reflect101(x + 1, pb)]) + 104 * row[x];
                tmp[i] = (uint16_t)h;
            }
            __syncwarp();
            for (int i = lane; i < n; i += 32) {
                const int y = i / pb, x = i - y * pb;
                const int v = 14 * (tmp[reflect101(y - 2, pb) * pb + x] + tmp[reflect101(y + 2, pb) * pb + x]) +
                              62 * (tmp[reflect101(y - 1, pb) * pb + x] + tmp[reflect101(y + 1, pb) * pb + x]) +
                              104 * tmp[i];
                a[i] = (uint8_t)((v + 32768) >> 16);
            }
            __syncwarp();
        }
        for (int i = lane; i < n; i += 32) {
            const int y = i / pb, x = i - y * pb;
            dp[(int64_t)y * g.dst_row + x * g.C] = a[i];
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------- downsample
// Table blob layout (int32 words), one entry of `level_stride(pb)` words per level:
//   [0] small  [1] area_kind (0 copy, 1 2x2, 2 integer factor, 3 fractional)  [2] factor
//   [3] float bits of float32(1/factor^2)  [4] n_area  [5..7] reserved
//   [8 .. 8+pb]            area_start[0..pb]  (entries of destination index d: [start[d], start[d+1]))
//   then 2*pb entries x {src_index, float-bits alpha}
//   then horizontal linear taps i0[pb] i1[pb] c0[pb] c1[pb], then vertical ones likewise.
__host__ __device__ inline int level_stride(int pb) { return 8 + (pb + 1) + 2 * pb * 2 + 8 * pb; }

__global__ void __launch_bounds__(256) downsample_kernel(const BlockGeom g, const int32_t* __restrict__ levels,
                                                         const int32_t* __restrict__ tables, int n_levels, int warps_per_cta) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int pb = g.pb, n = pb * pb;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // per warp: B (4n bytes, float/int), A (n bytes), S (n bytes)
    int32_t* Bi = reinterpret_cast<int32_t*>(smem) + (size_t)w * n;
    float* Bf = reinterpret_cast<float*>(Bi);
    uint8_t* A = smem + (size_t)warps_per_cta * n * 4 + (size_t)w * n * 2;
    uint8_t* S = A + n;

    const int64_t units = (int64_t)g.T * g.By * g.Bx * g.C;
    for (int64_t unit = (int64_t)blockIdx.x * warps_per_cta + w; unit < units; unit += (int64_t)gridDim.x * warps_per_cta) {
        int t, by, bx, c;
        decode_unit(g, unit, t, by, bx, c);
        int lv = levels[((int64_t)t * g.By + by) * g.Bx + bx];
        lv = lv < 0 ? 0 : (lv >= n_levels ? n_levels - 1 : lv);
        const int32_t* tab = tables + (size_t)lv * level_stride(pb);
        const int small = tab[0], kind = tab[1];
        const uint8_t* sp = g.src + (int64_t)t * g.src_frame + (int64_t)by * pb * g.src_row + ((int64_t)bx * pb) * g.C + c;
        uint8_t* dp = g.dst + (int64_t)t * g.dst_frame + (int64_t)by * pb * g.dst_row + ((int64_t)bx * pb) * g.C + c;
        for (int i = lane; i < n; i += 32) {
            const int y = i / pb, x = i - y * pb;
            A[i] = sp[(int64_t)y * g.src_row + x * g.C];
        }
        __syncwarp();
        if (kind == 0 || small >= pb) {
            for (int i = lane; i < n; i += 32) {
                const int y = i / pb, x = i - y * pb;
                dp[(int64_t)y * g.dst_row + x * g.C] = A[i];
            }
            __syncwarp();
            continue;
        }
        const int ns = small * small;
        if (kind == 1) {
            for (int i = lane; i < ns; i += 32) {
                const int dy = i / small, dx = i - dy * small;
                const uint8_t* q = A + (2 * dy) * pb + 2 * dx;
                S[i] = (uint8_t)((q[0] + q[1] + q[pb] + q[pb + 1] + 2) >> 2);
            }
        } else if (kind == 2) {
            const int f = tab[2];
            const float scale = __int_as_float(tab[3]);
            for (int i = lane; i < ns; i += 32) {
                const int dy = i / small, dx = i - dy * small;
                int s = 0;
                for (int yy = 0; yy < f; ++yy)
                    for (int xx = 0; xx < f; ++xx) s += A[(dy * f + yy) * pb + dx * f + xx];
                S[i] = (uint8_t)__float2int_rn(__fmul_rn((float)s, scale));
            }
        } else {
            const int32_t* start = tab + 8;
            const int32_t* ent = tab + 8 + (pb + 1);
            for (int i = lane; i < pb * small; i += 32) {   // rows: (y, dx)
                const int y = i / small, dx = i - y * small;
                float acc = 0.f;
                for (int e = start[dx]; e < start[dx + 1]; ++e)
                    acc = __fadd_rn(acc, __fmul_rn((float)A[y * pb + ent[2 * e]], __int_as_float(ent[2 * e + 1])));
                Bf[y * small + dx] = acc;
            }
            __syncwarp();
            for (int i = lane; i < ns; i += 32) {           // columns: (dy, dx)
                const int dy = i / small, dx = i - dy * small;
                float acc = 0.f;
                for (int e = start[dy]; e < start[dy + 1]; ++e)
                    acc = __fadd_rn(acc, __fmul_rn(Bf[ent[2 * e] * small + dx], __int_as_float(ent[2 * e + 1])));
                int v = __float2int_rn(acc);
                S[i] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
            }
        }
        __syncwarp();
        // bilinear back up: horizontal pass into Bi[small][pb] (11-bit coefficients)
        const int32_t* lh = tab + 8 + (pb + 1) + 4 * pb;
        const int32_t* lvt = lh + 4 * pb;
        for (int i = lane; i < small * pb; i += 32) {
            const int y = i / pb, d = i - y * pb;
            Bi[i] = S[y * small + lh[d]] * lh[2 * pb + d] + S[y * small + lh[pb + d]] * lh[3 * pb + d];
        }
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            const int d2 = i / pb, d = i - d2 * pb;
            const int r0 = Bi[lvt[d2] * pb + d] >> 4, r1 = Bi[lvt[pb + d2] * pb + d] >> 4;
            int v = (((lvt[2 * pb + d2] * r0) >> 16) + ((lvt[3 * pb + d2] * r1) >> 16) + 2) >> 2;
            v = v < 0 ? 0 : (v > 255 ? 255 : v);
            dp[(int64_t)d2 * g.dst_row + d * g.C] = (uint8_t)v;
        }
        __syncwarp();
    }
}

// ----------------------------------------------------------------------------- dampen
// one thread per (8x8 tile, channel): forward AAN, per-coefficient gain, inverse AAN
template <bool FAST>   // FAST: single channel, 8-byte aligned rows -> 64-bit loads/stores
__global__ void __launch_bounds__(128) dampen_kernel(const BlockGeom g, const float* __restrict__ strength) {
    const int tiles_x = g.Bx * g.pb / 8, tiles_y = g.By * g.pb / 8;
    const int64_t total = (int64_t)g.T * tiles_y * tiles_x * g.C;
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    // consecutive threads -> consecutive tiles of a tile row (coalesced 8-byte row segments)
    int64_t b = id;
    const int c = FAST ? 0 : (int)(b % g.C);
    if (!FAST) b /= g.C;
    const int txi = (int)(b % tiles_x);
    b /= tiles_x;
    const int tyi = (int)(b % tiles_y);
    const int t = (int)(b / tiles_y);
    const float s = fminf(fmaxf(strength[((int64_t)t * g.By + (tyi * 8) / g.pb) * g.Bx + (txi * 8) / g.pb], 0.f), 1.f);

    const uint8_t* sp = g.src + (int64_t)t * g.src_frame + (int64_t)tyi * 8 * g.src_row + ((int64_t)txi * 8) * g.C + c;
    uint8_t* dp = g.dst + (int64_t)t * g.dst_frame + (int64_t)tyi * 8 * g.dst_row + ((int64_t)txi * 8) * g.C + c;

    float x[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        if (FAST) {
            const uint2 v = __ldcs(reinterpret_cast<const uint2*>(sp + (int64_t)r * g.src_row));
            x[r][0] = byte_as_biased_float<0>(v.x) - 8388608.f;
            x[r][1] = byte_as_biased_float<1>(v.x) - 8388608.f;
            x[r][2] = byte_as_biased_float<2>(v.x) - 8388608.f;
            x[r][3] = byte_as_biased_float<3>(v.x) - 8388608.f;
            x[r][4] = byte_as_biased_float<0>(v.y) - 8388608.f;
            x[r][5] = byte_as_biased_float<1>(v.y) - 8388608.f;
            x[r][6] = byte_as_biased_float<2>(v.y) - 8388608.f;
            x[r][7] = byte_as_biased_float<3>(v.y) - 8388608.f;
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) x[r][k] = (float)sp[(int64_t)r * g.src_row + k * g.C];
        }
    }
    fdct8x8(x);
    // gain 2^(-4 s (u+v)/14) / 64: powers of q = 2^(-4 s / 14); the 1/64 undoes the AAN scaling
    float gk[15];
    const float q = exp2f(-4.0f * s / 14.0f);
    gk[0] = 1.0f / 64.0f;
#pragma unroll
    for (int k = 1; k < 15; ++k) gk[k] = gk[k - 1] * q;
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int v = 0; v < 8; ++v) x[u][v] *= gk[u + v];
    idct8x8(x);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        uint32_t o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int v = __float2int_rn(x[r][k]);
            o[k] = (uint32_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
        if (FAST) {
            uint2 v;
            v.x = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
            v.y = o[4] | (o[5] << 8) | (o[6] << 16) | (o[7] << 24);
            __stcs(reinterpret_cast<uint2*>(dp + (int64_t)r * g.dst_row), v);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) dp[(int64_t)r * g.dst_row + k * g.C] = (uint8_t)o[k];
        }
    }
}

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

inline int grid_for_units(int64_t units, int per_cta) {
    int64_t gsz = (units + per_cta - 1) / per_cta;
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (int)(gsz < 1 ? 1 : (gsz > cap ? cap : gsz));
}

}  // namespace
}  // namespace elvis

using namespace elvis;

extern "C" int elvis_degrade_blur(const elvis_plane* src, const elvis_plane* dst, int32_t n_frames,
                                  int32_t block_px, int32_t by, int32_t bx, const int32_t* rounds,
                                  elvis_stream_t stream) {
    BlockGeom g;
    if (int rc = make_geom(src, dst, n_frames, block_px, by, bx, g)) return rc;
    if (!rounds) return ELVIS_ERR_INVALID_ARG;
    if (block_px > 64) return ELVIS_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    if (int rc = copy_edges(g, st)) return rc;
    const int n = block_px * block_px;
    int wpc = 8;
    while (wpc > 1 && (size_t)wpc * n * 3 + 16 > 48 * 1024) wpc >>= 1;
    const size_t smem = (((size_t)wpc * n + 15) & ~(size_t)15) + (size_t)wpc * n * 2;
    const int64_t units = (int64_t)n_frames * by * bx * g.C;
    blur_kernel<<<grid_for_units(units, wpc), wpc * 32, smem, st>>>(g, rounds, wpc);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_degrade_downsample(const elvis_plane* src, const elvis_plane* dst, int32_t n_frames,
                                        int32_t block_px, int32_t by, int32_t bx, const int32_t* levels,
                                        const int32_t* tables, int32_t n_levels, elvis_stream_t stream) {
    BlockGeom g;
    if (int rc = make_geom(src, dst, n_frames, block_px, by, bx, g)) return rc;
    if (!levels || !tables || n_levels <= 0) return ELVIS_ERR_INVALID_ARG;
    if (block_px > 64) return ELVIS_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    if (int rc = copy_edges(g, st)) return rc;
    const int n = block_px * block_px;
    int wpc = 8;
    while (wpc > 1 && (size_t)wpc * n * 6 > 48 * 1024) wpc >>= 1;
    const size_t smem = (size_t)wpc * n * 6;
    const int64_t units = (int64_t)n_frames * by * bx * g.C;
    downsample_kernel<<<grid_for_units(units, wpc), wpc * 32, smem, st>>>(g, levels, tables, n_levels, wpc);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_dct_dampen(const elvis_plane* src, const elvis_plane* dst, int32_t n_frames,
                                int32_t block_px, int32_t by, int32_t bx, const float* strength,
                                elvis_stream_t stream) {
    BlockGeom g;
    if (int rc = make_geom(src, dst, n_frames, block_px, by, bx, g)) return rc;
    if (!strength) return ELVIS_ERR_INVALID_ARG;
    if (block_px % 8) return ELVIS_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    if (int rc = copy_edges(g, st)) return rc;
    const int64_t total = (int64_t)n_frames * (by * block_px / 8) * (bx * block_px / 8) * g.C;
    const bool fast = g.C == 1 && aligned_to(g.src, 8) && aligned_to(g.dst, 8) && g.src_frame % 8 == 0 &&
                      g.dst_frame % 8 == 0 && g.src_row % 8 == 0 && g.dst_row % 8 == 0;
    const unsigned grid = (unsigned)((total + 127) / 128);
    if (fast)
        dampen_kernel<true><<<grid, 128, 0, st>>>(g, strength);
    else
        dampen_kernel<false><<<grid, 128, 0, st>>>(g, strength);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}
