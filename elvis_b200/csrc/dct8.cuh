// 8-point scaled DCT-II / DCT-III butterflies (Arai-Agui-Nakajima factorisation: 5 multiplies
// and 29 additions per 1-D transform).  The outputs are *scaled*: for the forward transform
//     orthonormal_coefficient[k] = aan_forward_output[k] * kFwdScale[k]
// and the inverse expects inputs pre-multiplied by kInvScale[k].  Both tables were derived
// numerically against scipy.fft.dct(norm="ortho"); callers fold them into
// their per-coefficient weights so the scaling costs nothing at run time.
#pragma once

#ifdef __CUDACC__
#define ELVIS_HD __host__ __device__ __forceinline__
#else
#define ELVIS_HD inline
#endif

namespace elvis {

static const double kFwdScale[8] = {0.3535533905932738,  0.2548977895520796, 0.2705980500730985,
                                    0.30067244346752253, 0.35355339059327373, 0.4499881115682077,
                                    0.6532814824381882,  1.281457723870753};
static const double kInvScale[8] = {0.3535533905932738, 0.4903926402016152, 0.46193976625564337,
                                    0.4157348061512725, 0.35355339059327373, 0.2777851165098011,
                                    0.1913417161825449, 0.09754516100806411};

// forward, in place on 8 values addressed with a compile-time stride
#define ELVIS_FDCT8(d0, d1, d2, d3, d4, d5, d6, d7)                     \
    do {                                                                \
        float t0 = d0 + d7, t7 = d0 - d7;                               \
        float t1 = d1 + d6, t6 = d1 - d6;                               \
        float t2 = d2 + d5, t5 = d2 - d5;                               \
        float t3 = d3 + d4, t4 = d3 - d4;                               \
        float t10 = t0 + t3, t13 = t0 - t3;                             \
        float t11 = t1 + t2, t12 = t1 - t2;                             \
        d0 = t10 + t11;                                                 \
        d4 = t10 - t11;                                                 \
        float z1 = (t12 + t13) * 0.70710678118654752f;                  \
        d2 = t13 + z1;                                                  \
        d6 = t13 - z1;                                                  \
        t10 = t4 + t5;                                                  \
        t11 = t5 + t6;                                                  \
        t12 = t6 + t7;                                                  \
        float z5 = (t10 - t12) * 0.38268343236508977f;                  \
        float z2 = 0.54119610014619698f * t10 + z5;                     \
        float z4 = 1.30656296487637653f * t12 + z5;                     \
        float z3 = t11 * 0.70710678118654752f;                          \
        float z11 = t7 + z3, z13 = t7 - z3;                             \
        d5 = z13 + z2;                                                  \
        d3 = z13 - z2;                                                  \
        d1 = z11 + z4;                                                  \
        d7 = z11 - z4;                                                  \
    } while (0)

// inverse, in place (inputs already multiplied by kInvScale[k])
#define ELVIS_IDCT8(d0, d1, d2, d3, d4, d5, d6, d7)                     \
    do {                                                                \
        float t10 = d0 + d4, t11 = d0 - d4;                             \
        float t13 = d2 + d6;                                            \
        float t12 = (d2 - d6) * 1.41421356237309505f - t13;             \
        float e0 = t10 + t13, e3 = t10 - t13;                           \
        float e1 = t11 + t12, e2 = t11 - t12;                           \
        float z13 = d5 + d3, z10 = d5 - d3;                             \
        float z11 = d1 + d7, z12 = d1 - d7;                             \
        float o7 = z11 + z13;                                           \
        float u11 = (z11 - z13) * 1.41421356237309505f;                 \
        float z5 = (z10 + z12) * 1.84775906502257351f;                  \
        float u10 = 1.08239220029239397f * z12 - z5;                    \
        float u12 = -2.61312592975275306f * z10 + z5;                   \
        float o6 = u12 - o7;                                            \
        float o5 = u11 - o6;                                            \
        float o4 = u10 + o5;                                            \
        d0 = e0 + o7;                                                   \
        d7 = e0 - o7;                                                   \
        d1 = e1 + o6;                                                   \
        d6 = e1 - o6;                                                   \
        d2 = e2 + o5;                                                   \
        d5 = e2 - o5;                                                   \
        d4 = e3 + o4;                                                   \
        d3 = e3 - o4;                                                   \
    } while (0)

// 2-D transform of an 8x8 tile held in registers, x[row][col]
ELVIS_HD void fdct8x8(float (&x)[8][8]) {
#pragma unroll
    for (int r = 0; r < 8; ++r) ELVIS_FDCT8(x[r][0], x[r][1], x[r][2], x[r][3], x[r][4], x[r][5], x[r][6], x[r][7]);
#pragma unroll
    for (int c = 0; c < 8; ++c) ELVIS_FDCT8(x[0][c], x[1][c], x[2][c], x[3][c], x[4][c], x[5][c], x[6][c], x[7][c]);
}

ELVIS_HD void idct8x8(float (&x)[8][8]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) ELVIS_IDCT8(x[0][c], x[1][c], x[2][c], x[3][c], x[4][c], x[5][c], x[6][c], x[7][c]);
#pragma unroll
    for (int r = 0; r < 8; ++r) ELVIS_IDCT8(x[r][0], x[r][1], x[r][2], x[r][3], x[r][4], x[r][5], x[r][6], x[r][7]);
}

}  // namespace elvis
