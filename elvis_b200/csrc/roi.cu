// what the server does with the scores right before the external encoder --
// per-block delta-QP side files and the raw 4:2:0 frames (utils.py:453-462, 1026-1092;
// elvis.py:2027-2090).  The file formats are written by the host mirrors; the arithmetic is here:
//   * Kvazaar ROI:  int8 dqp = clip(clip((1 - imp) * 2 * R - R, +-14), -QP, 51 - QP), truncated;
//   * SVT-AV1 ROI:  cv2.resize(float32(imp), 64-px grid, INTER_AREA) -> 8 levels -> QP offsets;
//   * x265 qpfile:  float32(clip(2 s - 1, -1, 1)) resized to the CTU grid with INTER_AREA;
//   * Y4M frames:   cv2.cvtColor(RGB -> YUV_I420), BT.601 fixed point (20 fractional bits), chroma
//                   of the top-left pixel of every 2 x 2 quad.
// cv2's float INTER_AREA is restated operation by operation (oracle/spec_cv.py resize_area_f32,
// pinned against cv2): every product and sum is an explicitly rounded fp32 operation in cv2's
// order, so the results are bit-exact, not merely close.
#include "common.cuh"

namespace elvis {
namespace {

__global__ void __launch_bounds__(256) kvazaar_dqp_kernel(const double* __restrict__ imp, int64_t n, double base_qp, double qp_range,
                                                          int8_t* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        // (1.0 - importance) * 2 * qp_range - qp_range, left to right in float64 (utils.py:1047)
        double d = __dsub_rn(__dmul_rn(__dmul_rn(__dsub_rn(1.0, imp[i]), 2.0), qp_range), qp_range);
        d = fmin(fmax(d, -14.0), 14.0);                       // kvazaar's internal limit
        d = fmin(fmax(d, 0.0 - base_qp), 51.0 - base_qp);     // keep QP + dqp inside HEVC's 0..51
        out[i] = (int8_t)(int)d;                              // astype(np.int8): truncation
    }
}

// mode 0: float32(x) (utils.py:1078).  mode 1: float32(clip(2 x - 1, -1, 1)) (elvis.py:2030)
__global__ void __launch_bounds__(256) roi_prepare_kernel(const double* __restrict__ x, int64_t n, int mode, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        double v = x[i];
        if (mode == 1) v = fmin(fmax(__dsub_rn(__dmul_rn(v, 2.0), 1.0), -1.0), 1.0);
        out[i] = __double2float_rn(v);
    }
}

// levels = clip(int32(r * 8), 0, 7); dqp = R - (levels * 2 R // 7); clip to keep CRF + dqp in 0..63 (utils.py:1082-1088)
__global__ void __launch_bounds__(256) svtav1_offsets_kernel(const float* __restrict__ r, int64_t n, int base_crf, int qp_range,
                                                             int32_t* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        int level = (int)__fmul_rn(r[i], 8.0f);
        level = min(max(level, 0), 7);
        int d = qp_range - (level * 2 * qp_range) / 7;       // operands are non-negative: / is floor division
        out[i] = min(max(d, 0 - base_crf), 63 - base_crf);
    }
}

struct AreaParams {
    const float* src;
    float* dst;
    int32_t T, sh, sw, dh, dw;
    // general path: per destination index, entries [ofs[d], ofs[d + 1]) of (source index, weight)
    const int32_t* x_ofs;
    const int32_t* x_src;
    const float* x_alpha;
    const int32_t* y_ofs;
    const int32_t* y_src;
    const float* y_alpha;
    // integer-ratio path (cv2 resizeAreaFast_): window isy x isx, scale = 1 / (isx isy)
    int32_t fast, isx, isy, simd_cols;
    float scale;
};

__global__ void __launch_bounds__(256) area_f32_kernel(const AreaParams p) {
    const int64_t total = (int64_t)p.T * p.dh * p.dw;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int dx = (int)(i % p.dw);
        const int dy = (int)((i / p.dw) % p.dh);
        const float* s = p.src + (i / ((int64_t)p.dw * p.dh)) * p.sh * p.sw;
        float r;
        if (p.fast) {
            const float* w0 = s + (int64_t)dy * p.isy * p.sw + dx * p.isx;
            if (p.isx == 2 && p.isy == 2 && dx < p.simd_cols) {
                // cv2's 4-lane vector path of the 2 x 2 case: (row0 even + row0 odd) + (row1 even + row1 odd)
                r = __fmul_rn(__fadd_rn(__fadd_rn(w0[0], w0[1]), __fadd_rn(w0[p.sw], w0[p.sw + 1])), 0.25f);
            } else {
                // scalar path: window in row-major order, four at a time ((a + b) + c) + d added to the sum
                const int area = p.isx * p.isy;
                float sum = 0.f;
                int k = 0;
                auto at = [&](int kk) { return w0[(int64_t)(kk / p.isx) * p.sw + kk % p.isx]; };
                for (; k + 4 <= area; k += 4)
                    sum = __fadd_rn(sum, __fadd_rn(__fadd_rn(__fadd_rn(at(k), at(k + 1)), at(k + 2)), at(k + 3)));
                for (; k < area; ++k) sum = __fadd_rn(sum, at(k));
                r = __fmul_rn(sum, p.scale);
            }
        } else {
            // rows in table order: buf = sum_x S * alpha (from 0), result = beta0 buf0 + beta1 buf1 + ...
            r = 0.f;
            for (int ey = p.y_ofs[dy]; ey < p.y_ofs[dy + 1]; ++ey) {
                const float* row = s + (int64_t)p.y_src[ey] * p.sw;
                float buf = 0.f;
                for (int ex = p.x_ofs[dx]; ex < p.x_ofs[dx + 1]; ++ex) buf = __fadd_rn(buf, __fmul_rn(row[p.x_src[ex]], p.x_alpha[ex]));
                const float term = __fmul_rn(p.y_alpha[ey], buf);
                r = (ey == p.y_ofs[dy]) ? term : __fadd_rn(r, term);
            }
        }
        p.dst[i] = r;
    }
}

// BT.601 limited range, cv2's ITUR_BT_601 integer coefficients (20 fractional bits)
constexpr int kShift = 20;
constexpr int kCRY = 269484, kCGY = 528482, kCBY = 102760;
constexpr int kCRU = -155188, kCGU = -305135, kCBU = 460324;
constexpr int kCGV = -385875, kCBV = -74448;

__device__ __forceinline__ uint32_t luma_of(uint32_t r, uint32_t g, uint32_t b) {
    return (uint32_t)((kCRY * (int)r + kCGY * (int)g + kCBY * (int)b + (1 << (kShift - 1)) + (16 << kShift)) >> kShift);
}

struct I420Params {
    const uint8_t* rgb;
    uint8_t* y;
    uint8_t* u;
    uint8_t* v;
    int64_t rgb_frame, rgb_row, y_frame, y_row, u_frame, u_row, v_frame, v_row;
    int32_t T, H, W;
    int32_t words;     // 1: rows are read as 3 x u32 per 4 pixels (needs 4-byte alignment and W % 4 == 0)
};

// one thread: 4 pixels x 2 rows -> 2 x 4 luma bytes, 2 U bytes, 2 V bytes
__global__ void __launch_bounds__(256) rgb_to_i420_kernel(const I420Params p) {
    const int qw = (p.W + 3) / 4, qh = p.H / 2;
    const int64_t total = (int64_t)p.T * qh * qw;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int qx = (int)(i % qw);
        const int qy = (int)((i / qw) % qh);
        const int t = (int)(i / ((int64_t)qw * qh));
        const int x0 = qx * 4, y0 = qy * 2;
        const int n = min(4, p.W - x0);                 // W is even, so n is 4 or 2
        uint8_t px[2][12];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const uint8_t* src = p.rgb + (int64_t)t * p.rgb_frame + (int64_t)(y0 + r) * p.rgb_row + (int64_t)x0 * 3;
            if (p.words && n == 4) {
                const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const uint32_t w = __ldcs(s32 + k);
                    px[r][4 * k] = w & 0xff;
                    px[r][4 * k + 1] = (w >> 8) & 0xff;
                    px[r][4 * k + 2] = (w >> 16) & 0xff;
                    px[r][4 * k + 3] = w >> 24;
                }
            } else {
                for (int k = 0; k < 3 * n; ++k) px[r][k] = src[k];
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            uint8_t* yd = p.y + (int64_t)t * p.y_frame + (int64_t)(y0 + r) * p.y_row + x0;
            uint32_t packed = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) packed |= (k < n ? luma_of(px[r][3 * k], px[r][3 * k + 1], px[r][3 * k + 2]) : 0u) << (8 * k);
            if (p.words && n == 4) {
                *reinterpret_cast<uint32_t*>(yd) = packed;
            } else {
                for (int k = 0; k < n; ++k) yd[k] = (uint8_t)(packed >> (8 * k));
            }
        }
        uint8_t* ud = p.u + (int64_t)t * p.u_frame + (int64_t)qy * p.u_row + x0 / 2;
        uint8_t* vd = p.v + (int64_t)t * p.v_frame + (int64_t)qy * p.v_row + x0 / 2;
        for (int k = 0; k < n / 2; ++k) {
            const int r = px[0][6 * k], g = px[0][6 * k + 1], b = px[0][6 * k + 2];   // top-left pixel of the quad
            ud[k] = (uint8_t)((kCRU * r + kCGU * g + kCBU * b + (1 << (kShift - 1)) + (128 << kShift)) >> kShift);
            vd[k] = (uint8_t)((kCBU * r + kCGV * g + kCBV * b + (1 << (kShift - 1)) + (128 << kShift)) >> kShift);
        }
    }
}

// aligned fast path: one thread = 16 pixels x 2 rows, 3 x 128-bit loads and one 128-bit luma store per
// row, 64-bit chroma stores
__global__ void __launch_bounds__(256) rgb_to_i420_x16_kernel(const I420Params p) {
    const int qw = p.W / 16, qh = p.H / 2;
    const int64_t total = (int64_t)p.T * qh * qw;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int qx = (int)(i % qw);
        const int qy = (int)((i / qw) % qh);
        const int t = (int)(i / ((int64_t)qw * qh));
        const int x0 = qx * 16, y0 = qy * 2;
        uint32_t uu[2] = {0u, 0u}, vv[2] = {0u, 0u};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const uint4* src = reinterpret_cast<const uint4*>(p.rgb + (int64_t)t * p.rgb_frame + (int64_t)(y0 + r) * p.rgb_row + (int64_t)x0 * 3);
            uint32_t w[12];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const uint4 q = __ldcs(src + k);
                w[4 * k] = q.x;
                w[4 * k + 1] = q.y;
                w[4 * k + 2] = q.z;
                w[4 * k + 3] = q.w;
            }
            auto byte_at = [&](int k) { return (w[k >> 2] >> (8 * (k & 3))) & 0xffu; };
            uint32_t yy[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const uint32_t rr = byte_at(3 * k), gg = byte_at(3 * k + 1), bb = byte_at(3 * k + 2);
                yy[k >> 2] |= luma_of(rr, gg, bb) << (8 * (k & 3));
                if (r == 0 && (k & 1) == 0) {          // top-left pixel of the quad
                    const int j = k >> 1;
                    const uint32_t u8v = (uint32_t)((kCRU * (int)rr + kCGU * (int)gg + kCBU * (int)bb + (1 << (kShift - 1)) + (128 << kShift)) >> kShift) & 0xffu;
                    const uint32_t v8v = (uint32_t)((kCBU * (int)rr + kCGV * (int)gg + kCBV * (int)bb + (1 << (kShift - 1)) + (128 << kShift)) >> kShift) & 0xffu;
                    uu[j >> 2] |= u8v << (8 * (j & 3));
                    vv[j >> 2] |= v8v << (8 * (j & 3));
                }
            }
            __stcs(reinterpret_cast<uint4*>(p.y + (int64_t)t * p.y_frame + (int64_t)(y0 + r) * p.y_row + x0), make_uint4(yy[0], yy[1], yy[2], yy[3]));
        }
        __stcs(reinterpret_cast<uint2*>(p.u + (int64_t)t * p.u_frame + (int64_t)qy * p.u_row + x0 / 2), make_uint2(uu[0], uu[1]));
        __stcs(reinterpret_cast<uint2*>(p.v + (int64_t)t * p.v_frame + (int64_t)qy * p.v_row + x0 / 2), make_uint2(vv[0], vv[1]));
    }
}

// cv2.COLOR_RGB2GRAY on uint8 (pinned against cv2 4.13 in tests/test_oracle.py: 15-bit fixed point,
// (9798 R + 19235 G + 3735 B + 2^14) >> 15).  One thread = 4 pixels of a row (12 bytes in, one word out).
struct GrayParams {
    const uint8_t* rgb;
    uint8_t* y;
    int64_t rgb_frame, rgb_row, y_frame, y_row;
    int32_t T, H, W, words;
};

__device__ __forceinline__ uint32_t gray_of(uint32_t r, uint32_t g, uint32_t b) {
    return (9798u * r + 19235u * g + 3735u * b + (1u << 14)) >> 15;
}

__global__ void __launch_bounds__(256) rgb_to_gray_kernel(const GrayParams p) {
    const int qw = (p.W + 3) / 4;
    const int64_t total = (int64_t)p.T * p.H * qw;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int qx = (int)(i % qw);
        const int yy = (int)((i / qw) % p.H);
        const int t = (int)(i / ((int64_t)qw * p.H));
        const int x0 = qx * 4, n = min(4, p.W - x0);
        const uint8_t* src = p.rgb + (int64_t)t * p.rgb_frame + (int64_t)yy * p.rgb_row + (int64_t)x0 * 3;
        uint8_t* dst = p.y + (int64_t)t * p.y_frame + (int64_t)yy * p.y_row + x0;
        if (p.words && n == 4) {
            const uint32_t w0 = __ldcs(reinterpret_cast<const uint32_t*>(src)), w1 = __ldcs(reinterpret_cast<const uint32_t*>(src) + 1),
                           w2 = __ldcs(reinterpret_cast<const uint32_t*>(src) + 2);
            const uint32_t g0 = gray_of(w0 & 0xff, (w0 >> 8) & 0xff, (w0 >> 16) & 0xff);
            const uint32_t g1 = gray_of(w0 >> 24, w1 & 0xff, (w1 >> 8) & 0xff);
            const uint32_t g2 = gray_of((w1 >> 16) & 0xff, w1 >> 24, w2 & 0xff);
            const uint32_t g3 = gray_of((w2 >> 8) & 0xff, (w2 >> 16) & 0xff, w2 >> 24);
            __stcs(reinterpret_cast<uint32_t*>(dst), g0 | (g1 << 8) | (g2 << 16) | (g3 << 24));
        } else {
            for (int k = 0; k < n; ++k) dst[k] = (uint8_t)gray_of(src[3 * k], src[3 * k + 1], src[3 * k + 2]);
        }
    }
}

// Packed 3-channel frames <-> three planes.  The reference's frames are packed (H x W x 3); the fast per-block kernels
// (tensor-core blur, closed-form downsample) work on planes, so the operators split a packed clip, run the plane
// kernels per channel and merge the result -- two HBM-bound passes instead of the generic packed kernels, which are 4-15x
// slower than that (tools/time_packed_degrade.py).  Four pixels per thread: 3 words in, one word per plane out (and back).
struct ChannelParams {
    uint8_t* packed;
    uint8_t* plane[3];
    int64_t packed_frame, packed_row, plane_frame[3], plane_row[3];
    int32_t T, H, W, words;
};

template <bool MERGE>
__global__ void __launch_bounds__(256) channels3_kernel(const ChannelParams p) {
    const int qw = (p.W + 3) / 4;
    const int64_t total = (int64_t)p.T * p.H * qw;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int qx = (int)(i % qw);
        const int yy = (int)((i / qw) % p.H);
        const int t = (int)(i / ((int64_t)qw * p.H));
        const int x0 = qx * 4, n = min(4, p.W - x0);
        uint8_t* pk = p.packed + (int64_t)t * p.packed_frame + (int64_t)yy * p.packed_row + (int64_t)x0 * 3;
        uint8_t* pl[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) pl[c] = p.plane[c] + (int64_t)t * p.plane_frame[c] + (int64_t)yy * p.plane_row[c] + x0;
        if (p.words && n == 4) {
            // packed words of pixels 0..3 with channels (a, b, c):  w0 = a0 b0 c0 a1,  w1 = b1 c1 a2 b2,  w2 = c2 a3 b3 c3
            if (!MERGE) {
                const uint32_t w0 = __ldcs(reinterpret_cast<const uint32_t*>(pk)), w1 = __ldcs(reinterpret_cast<const uint32_t*>(pk) + 1),
                               w2 = __ldcs(reinterpret_cast<const uint32_t*>(pk) + 2);
                __stcs(reinterpret_cast<uint32_t*>(pl[0]), __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210));
                __stcs(reinterpret_cast<uint32_t*>(pl[1]), __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210));
                __stcs(reinterpret_cast<uint32_t*>(pl[2]), __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410));
            } else {
                const uint32_t a = __ldcs(reinterpret_cast<const uint32_t*>(pl[0])), b = __ldcs(reinterpret_cast<const uint32_t*>(pl[1])),
                               c = __ldcs(reinterpret_cast<const uint32_t*>(pl[2]));
                const uint32_t ab = __byte_perm(a, b, 0x5140), ab2 = __byte_perm(a, b, 0x7362);   // a0 b0 a1 b1 | a2 b2 a3 b3
                __stcs(reinterpret_cast<uint32_t*>(pk), __byte_perm(ab, c, 0x2410));                                   // a0 b0 c0 a1
                __stcs(reinterpret_cast<uint32_t*>(pk) + 1, __byte_perm(__byte_perm(ab, ab2, 0x0543), c, 0x2150));     // b1 c1 a2 b2
                __stcs(reinterpret_cast<uint32_t*>(pk) + 2, __byte_perm(ab2, c, 0x7326));                              // c2 a3 b3 c3
            }
        } else {
            for (int k = 0; k < n; ++k)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (MERGE) pk[3 * k + c] = pl[c][k];
                    else pl[c][k] = pk[3 * k + c];
                }
        }
    }
}

// cv2.resize(map, INTER_LINEAR) for float32 / float64 maps (pinned against cv2 4.13 in
// tests/test_oracle.py: oracle/spec_cv.py resize_linear_float).  cv2 interpolates as a fused
// lerp, horizontally then vertically: h = fma(S[x1] - S[x0], fx, S[x0]); out = fma(h1 - h0, fy, h0),
// each subtraction rounded on its own.  The source index / fraction tables come from the host
// (elvis_b200/_tables.py linear_float_index: coordinate = fma(d + 0.5, src/dst, -0.5) in double; the
// fraction is rounded to float for float32 maps).
template <typename F> __device__ __forceinline__ F lerp_rn(F a, F b, F f);
template <> __device__ __forceinline__ float lerp_rn<float>(float a, float b, float f) { return __fmaf_rn(__fsub_rn(b, a), f, a); }
template <> __device__ __forceinline__ double lerp_rn<double>(double a, double b, double f) { return __fma_rn(__dsub_rn(b, a), f, a); }

template <typename F>
__global__ void __launch_bounds__(256) resize_linear_float_kernel(const F* __restrict__ src, int T, int sh, int sw, F* __restrict__ dst,
                                                                  int dh, int dw, const int32_t* __restrict__ yi, const double* __restrict__ yf,
                                                                  const int32_t* __restrict__ xi, const double* __restrict__ xf) {
    const int64_t total = (int64_t)T * dh * dw;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int x = (int)(i % dw);
        const int y = (int)((i / dw) % dh);
        const int t = (int)(i / ((int64_t)dw * dh));
        const int x0 = xi[x], x1 = min(x0 + 1, sw - 1), y0 = yi[y], y1 = min(y0 + 1, sh - 1);
        const F fx = (F)xf[x], fy = (F)yf[y];
        const F* m = src + (int64_t)t * sh * sw;
        const F h0 = lerp_rn<F>(m[(int64_t)y0 * sw + x0], m[(int64_t)y0 * sw + x1], fx);
        const F h1 = lerp_rn<F>(m[(int64_t)y1 * sw + x0], m[(int64_t)y1 * sw + x1], fx);
        dst[i] = lerp_rn<F>(h0, h1, fy);
    }
}

unsigned grid_for(int64_t n) {
    int64_t g = (n + 255) / 256;
    const int64_t cap = (int64_t)kNumSMs * 16;
    if (g > cap) g = cap;
    return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace
}  // namespace elvis

using namespace elvis;

extern "C" int elvis_roi_kvazaar(const double* importance, int64_t n, int32_t base_qp, int32_t qp_range, int8_t* dqp,
                                 elvis_stream_t stream) {
    if (!importance || !dqp || n <= 0) return ELVIS_ERR_INVALID_ARG;
    kvazaar_dqp_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(importance, n, (double)base_qp, (double)qp_range, dqp);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_roi_prepare_f32(const double* x, int64_t n, int32_t mode, float* out, elvis_stream_t stream) {
    if (!x || !out || n <= 0 || (mode != 0 && mode != 1)) return ELVIS_ERR_INVALID_ARG;
    roi_prepare_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(x, n, mode, out);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_roi_svtav1_offsets(const float* resized, int64_t n, int32_t base_crf, int32_t qp_range, int32_t* offsets,
                                        elvis_stream_t stream) {
    if (!resized || !offsets || n <= 0 || qp_range < 0) return ELVIS_ERR_INVALID_ARG;
    svtav1_offsets_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(resized, n, base_crf, qp_range, offsets);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_resize_area_f32(const float* src, int32_t n_maps, int32_t src_h, int32_t src_w, float* dst, int32_t dst_h,
                                     int32_t dst_w, const int32_t* x_ofs, const int32_t* x_src, const float* x_alpha,
                                     const int32_t* y_ofs, const int32_t* y_src, const float* y_alpha, int32_t int_scale_x,
                                     int32_t int_scale_y, int32_t simd_cols, elvis_stream_t stream) {
    if (!src || !dst || n_maps <= 0 || src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0) return ELVIS_ERR_INVALID_ARG;
    if (dst_h > src_h || dst_w > src_w) return ELVIS_ERR_UNSUPPORTED;      // cv2 switches to a bilinear path when enlarging
    AreaParams p;
    p.src = src;
    p.dst = dst;
    p.T = n_maps;
    p.sh = src_h;
    p.sw = src_w;
    p.dh = dst_h;
    p.dw = dst_w;
    p.x_ofs = x_ofs;
    p.x_src = x_src;
    p.x_alpha = x_alpha;
    p.y_ofs = y_ofs;
    p.y_src = y_src;
    p.y_alpha = y_alpha;
    p.fast = int_scale_x > 0 && int_scale_y > 0;
    p.isx = int_scale_x;
    p.isy = int_scale_y;
    p.simd_cols = simd_cols;
    p.scale = 0.f;
    if (p.fast) {
        if ((int64_t)dst_w * int_scale_x > src_w || (int64_t)dst_h * int_scale_y > src_h) return ELVIS_ERR_SHAPE;
        p.scale = 1.0f / (float)(int_scale_x * int_scale_y);
    } else if (!x_ofs || !x_src || !x_alpha || !y_ofs || !y_src || !y_alpha) {
        return ELVIS_ERR_INVALID_ARG;
    }
    area_f32_kernel<<<grid_for((int64_t)n_maps * dst_h * dst_w), 256, 0, as_stream(stream)>>>(p);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_rgb_to_i420(const elvis_plane* rgb, const elvis_plane* y, const elvis_plane* u, const elvis_plane* v,
                                 int32_t n_frames, elvis_stream_t stream) {
    if (!plane_ok(rgb) || !plane_ok(y) || !plane_ok(u) || !plane_ok(v) || n_frames <= 0) return ELVIS_ERR_INVALID_ARG;
    if (rgb->channels != 3 || y->channels != 1 || u->channels != 1 || v->channels != 1) return ELVIS_ERR_INVALID_ARG;
    const int H = rgb->height, W = rgb->width;
    if ((H & 1) || (W & 1)) return ELVIS_ERR_SHAPE;        // cv2 asserts even dimensions for 4:2:0
    if (y->height < H || y->width < W || u->height < H / 2 || u->width < W / 2 || v->height < H / 2 || v->width < W / 2) return ELVIS_ERR_SHAPE;
    I420Params p;
    p.rgb = static_cast<const uint8_t*>(rgb->data);
    p.y = static_cast<uint8_t*>(y->data);
    p.u = static_cast<uint8_t*>(u->data);
    p.v = static_cast<uint8_t*>(v->data);
    p.rgb_frame = rgb->frame_stride;
    p.rgb_row = rgb->row_stride;
    p.y_frame = y->frame_stride;
    p.y_row = y->row_stride;
    p.u_frame = u->frame_stride;
    p.u_row = u->row_stride;
    p.v_frame = v->frame_stride;
    p.v_row = v->row_stride;
    p.T = n_frames;
    p.H = H;
    p.W = W;
    p.words = aligned_to(p.rgb, 4) && rgb->frame_stride % 4 == 0 && rgb->row_stride % 4 == 0 && aligned_to(p.y, 4) &&
              y->frame_stride % 4 == 0 && y->row_stride % 4 == 0;
    auto al = [](const elvis_plane* q, int a) { return aligned_to(q->data, a) && q->frame_stride % a == 0 && q->row_stride % a == 0; };
    if (W % 16 == 0 && al(rgb, 16) && al(y, 16) && al(u, 8) && al(v, 8))
        rgb_to_i420_x16_kernel<<<grid_for((int64_t)n_frames * (H / 2) * (W / 16)), 256, 0, as_stream(stream)>>>(p);
    else
        rgb_to_i420_kernel<<<grid_for((int64_t)n_frames * (H / 2) * ((W + 3) / 4)), 256, 0, as_stream(stream)>>>(p);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_rgb_to_gray(const elvis_plane* rgb, const elvis_plane* y, int32_t n_frames, elvis_stream_t stream) {
    if (!plane_ok(rgb) || !plane_ok(y) || n_frames <= 0) return ELVIS_ERR_INVALID_ARG;
    if (rgb->channels != 3 || y->channels != 1) return ELVIS_ERR_INVALID_ARG;
    if (y->height < rgb->height || y->width < rgb->width) return ELVIS_ERR_SHAPE;
    GrayParams p;
    p.rgb = static_cast<const uint8_t*>(rgb->data);
    p.y = static_cast<uint8_t*>(y->data);
    p.rgb_frame = rgb->frame_stride;
    p.rgb_row = rgb->row_stride;
    p.y_frame = y->frame_stride;
    p.y_row = y->row_stride;
    p.T = n_frames;
    p.H = rgb->height;
    p.W = rgb->width;
    p.words = aligned_to(p.rgb, 4) && rgb->frame_stride % 4 == 0 && rgb->row_stride % 4 == 0 && aligned_to(p.y, 4) &&
              y->frame_stride % 4 == 0 && y->row_stride % 4 == 0;
    rgb_to_gray_kernel<<<grid_for((int64_t)n_frames * p.H * ((p.W + 3) / 4)), 256, 0, as_stream(stream)>>>(p);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_resize_linear_float(const void* src, int32_t dtype, int32_t n_maps, int32_t src_h, int32_t src_w, void* dst,
                                         int32_t dst_h, int32_t dst_w, const int32_t* y_index, const double* y_frac,
                                         const int32_t* x_index, const double* x_frac, elvis_stream_t stream) {
    if (!src || !dst || !y_index || !y_frac || !x_index || !x_frac) return ELVIS_ERR_INVALID_ARG;
    if (n_maps <= 0 || src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0) return ELVIS_ERR_INVALID_ARG;
    const unsigned grid = grid_for((int64_t)n_maps * dst_h * dst_w);
    if (dtype == ELVIS_F32)
        resize_linear_float_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const float*>(src), n_maps, src_h, src_w,
                                                                              static_cast<float*>(dst), dst_h, dst_w, y_index, y_frac, x_index, x_frac);
    else if (dtype == ELVIS_F64)
        resize_linear_float_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const double*>(src), n_maps, src_h, src_w,
                                                                               static_cast<double*>(dst), dst_h, dst_w, y_index, y_frac, x_index, x_frac);
    else
        return ELVIS_ERR_INVALID_ARG;
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

namespace {
int channels3(const elvis_plane* packed, const elvis_plane* planes, int32_t n_frames, bool merge, elvis_stream_t stream) {
    if (!plane_ok(packed) || !planes || n_frames <= 0 || packed->channels != 3) return ELVIS_ERR_INVALID_ARG;
    ChannelParams p;
    p.packed = static_cast<uint8_t*>(packed->data);
    p.packed_frame = packed->frame_stride;
    p.packed_row = packed->row_stride;
    p.T = n_frames;
    p.H = packed->height;
    p.W = packed->width;
    bool words = aligned_to(p.packed, 4) && packed->frame_stride % 4 == 0 && packed->row_stride % 4 == 0;
    for (int c = 0; c < 3; ++c) {
        if (!plane_ok(&planes[c]) || planes[c].channels != 1) return ELVIS_ERR_INVALID_ARG;
        if (planes[c].height != packed->height || planes[c].width != packed->width) return ELVIS_ERR_SHAPE;
        p.plane[c] = static_cast<uint8_t*>(planes[c].data);
        p.plane_frame[c] = planes[c].frame_stride;
        p.plane_row[c] = planes[c].row_stride;
        words = words && aligned_to(p.plane[c], 4) && planes[c].frame_stride % 4 == 0 && planes[c].row_stride % 4 == 0;
    }
    p.words = words;
    const unsigned grid = grid_for((int64_t)n_frames * p.H * ((p.W + 3) / 4));
    if (merge) channels3_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(p);
    else channels3_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(p);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}
}  // namespace

extern "C" int elvis_split_channels3(const elvis_plane* packed, const elvis_plane* planes, int32_t n_frames, elvis_stream_t stream) {
    return channels3(packed, planes, n_frames, false, stream);
}

extern "C" int elvis_merge_channels3(const elvis_plane* planes, const elvis_plane* packed, int32_t n_frames, elvis_stream_t stream) {
    return channels3(packed, planes, n_frames, true, stream);
}
