// The AAN 8-point butterflies of dct8.cuh on PACKED fp32 pairs (Blackwell FADD2 / FMUL2 /
// FFMA2, PTX add/mul/fma.rn.f32x2): one instruction transforms the same position of two
// independent 8-vectors.  The kernels in score.cu are bounded by instruction issue, not by
// HBM (profiles/r1a_*), so halving the FP32 instruction count of the transform is the lever.
// Output scaling is identical to dct8.cuh (kFwdScale).
#pragma once
#include <cuda_runtime.h>

namespace elvis {

__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 f2muls(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
// a * s + c
__device__ __forceinline__ float2 f2fmas(float2 a, float s, float2 c) { return __ffma2_rn(a, make_float2(s, s), c); }

// forward transform of d0..d7 (float2 lvalues), in place
#define ELVIS_FDCT8_X2(d0, d1, d2, d3, d4, d5, d6, d7)                         \
    do {                                                                       \
        float2 t0 = f2add(d0, d7), t7 = f2sub(d0, d7);                         \
        float2 t1 = f2add(d1, d6), t6 = f2sub(d1, d6);                         \
        float2 t2 = f2add(d2, d5), t5 = f2sub(d2, d5);                         \
        float2 t3 = f2add(d3, d4), t4 = f2sub(d3, d4);                         \
        float2 t10 = f2add(t0, t3), t13 = f2sub(t0, t3);                       \
        float2 t11 = f2add(t1, t2), t12 = f2sub(t1, t2);                       \
        d0 = f2add(t10, t11);                                                  \
        d4 = f2sub(t10, t11);                                                  \
        float2 s1 = f2add(t12, t13);                                           \
        d2 = f2fmas(s1, 0.70710678118654752f, t13);                            \
        d6 = f2fmas(s1, -0.70710678118654752f, t13);                           \
        t10 = f2add(t4, t5);                                                   \
        t11 = f2add(t5, t6);                                                   \
        t12 = f2add(t6, t7);                                                   \
        float2 z5 = f2muls(f2sub(t10, t12), 0.38268343236508977f);             \
        float2 z2 = f2fmas(t10, 0.54119610014619698f, z5);                     \
        float2 z4 = f2fmas(t12, 1.30656296487637653f, z5);                     \
        float2 z11 = f2fmas(t11, 0.70710678118654752f, t7);                    \
        float2 z13 = f2fmas(t11, -0.70710678118654752f, t7);                   \
        d5 = f2add(z13, z2);                                                   \
        d3 = f2sub(z13, z2);                                                   \
        d1 = f2add(z11, z4);                                                   \
        d7 = f2sub(z11, z4);                                                   \
    } while (0)

// inverse transform of d0..d7 (float2 lvalues, inputs pre-multiplied by kInvScale), in place
#define ELVIS_IDCT8_X2(d0, d1, d2, d3, d4, d5, d6, d7)                         \
    do {                                                                       \
        float2 t10 = f2add(d0, d4), t11 = f2sub(d0, d4);                       \
        float2 t13 = f2add(d2, d6);                                            \
        float2 t12 = f2sub(f2muls(f2sub(d2, d6), 1.41421356237309505f), t13);  \
        float2 e0 = f2add(t10, t13), e3 = f2sub(t10, t13);                     \
        float2 e1 = f2add(t11, t12), e2 = f2sub(t11, t12);                     \
        float2 z13 = f2add(d5, d3), z10 = f2sub(d5, d3);                       \
        float2 z11 = f2add(d1, d7), z12 = f2sub(d1, d7);                       \
        float2 o7 = f2add(z11, z13);                                           \
        float2 u11 = f2muls(f2sub(z11, z13), 1.41421356237309505f);            \
        float2 z5 = f2muls(f2add(z10, z12), 1.84775906502257351f);             \
        float2 u10 = f2sub(f2muls(z12, 1.08239220029239397f), z5);             \
        float2 u12 = f2fmas(z10, -2.61312592975275306f, z5);                   \
        float2 o6 = f2sub(u12, o7);                                            \
        float2 o5 = f2sub(u11, o6);                                            \
        float2 o4 = f2add(u10, o5);                                            \
        d0 = f2add(e0, o7);                                                    \
        d7 = f2sub(e0, o7);                                                    \
        d1 = f2add(e1, o6);                                                    \
        d6 = f2sub(e1, o6);                                                    \
        d2 = f2add(e2, o5);                                                    \
        d5 = f2sub(e2, o5);                                                    \
        d4 = f2add(e3, o4);                                                    \
        d3 = f2sub(e3, o4);                                                    \
    } while (0)

}  // namespace elvis
