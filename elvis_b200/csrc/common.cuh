// Shared host/device helpers for the elvis_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdlib>
#include <stdint.h>
#include "../../include/elvis_b200.h"

namespace elvis {

extern thread_local int g_last_cuda_error;

inline int cuda_fail(cudaError_t e) {
    g_last_cuda_error = (int)e;
    return ELVIS_ERR_CUDA;
}

// Launch check: catches configuration errors at enqueue time without synchronising.
#define ELVIS_CHECK_LAUNCH()                                   \
    do {                                                       \
        cudaError_t e__ = cudaGetLastError();                  \
        if (e__ != cudaSuccess) return ::elvis::cuda_fail(e__); \
    } while (0)

inline cudaStream_t as_stream(elvis_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline bool aligned_to(const void* p, int64_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// Largest power-of-two vector width (bytes, <= 16) usable for rows of `row_bytes` payload
// addressed as base + frame*frame_stride + row*row_stride + k*unit.
inline int vector_unit(const elvis_plane* p, int64_t row_bytes) {
    int u = 16;
    while (u > 1) {
        if (aligned_to(p->data, u) && p->frame_stride % u == 0 && p->row_stride % u == 0 && row_bytes % u == 0)
            return u;
        u >>= 1;
    }
    return 1;
}

inline int plane_ok(const elvis_plane* p) {
    return p && p->data && p->height > 0 && p->width > 0 && p->channels >= 1 && p->row_stride >= (int64_t)p->width * p->channels;
}

// SM count of the CURRENT device (148 on B200: 2 dies x 74 SMs), queried once per device.
int num_sms();
#define kNumSMs (::elvis::num_sms())

// Per-device "this kernel's opt-in attribute has been set" flag.  Function attributes such as
// cudaFuncAttributeMaxDynamicSharedMemorySize belong to the (function, device) pair, so a
// process-wide once-flag would leave the second GPU of a process unconfigured.  The flag is a
// cache of an idempotent call (benign if two threads race to set it), not library state.
struct PerDeviceOnce {
    std::atomic<uint64_t> lo{0}, hi{0};   // devices 0..127
    template <typename F> cudaError_t run(F&& configure) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= 128) return configure();
        std::atomic<uint64_t>& word = dev < 64 ? lo : hi;
        const uint64_t bit = 1ull << (dev & 63);
        if (word.load(std::memory_order_acquire) & bit) return cudaSuccess;
        e = configure();
        if (e == cudaSuccess) word.fetch_or(bit, std::memory_order_release);
        return e;
    }
};

// streaming 128/64/32-bit accesses: no L1 allocation (every byte is touched once)
template <typename V> __device__ __forceinline__ V ld_stream(const V* p) { return __ldcs(p); }
template <typename V> __device__ __forceinline__ void st_stream(V* p, V v) { __stcs(p, v); }

// Streaming loads for GATHERS of pieces smaller than an L2 line (the kept 16-byte block rows of shrink): with the
// default policy B200's L2 fills the whole 128-byte line around any miss -- reading 16 of every 128 bytes moves all
// 128 from DRAM; the `.L2::64B` prefetch-size hint halves the fill to 64 bytes (tools/micro/gather_granularity.cu under
// ncu: 1.07 GB instead of 2.15 GB of DRAM reads for a 2 GiB span; cudaLimitMaxL2FetchGranularity alone changes
// nothing; PTX has no smaller hint).  Dense reads keep the default, which is what they want.
template <typename V> __device__ __forceinline__ V ld_gather(const V* p) { return __ldcs(p); }
template <> __device__ __forceinline__ uint4 ld_gather<uint4>(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.cs.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
template <> __device__ __forceinline__ uint2 ld_gather<uint2>(const uint2* p) {
    uint2 v;
    asm volatile("ld.global.cs.L2::64B.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

// float(2^23 + b) for byte K of w: byte_perm builds the bit pattern 0x4B0000bb directly, so a
// u8 -> fp32 conversion is one PRMT (and the 2^23 bias cancels in differences).
template <int K> __device__ __forceinline__ float byte_as_biased_float(uint32_t w) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540 + K));
}
// Same, with the 0x4B000000 pattern supplied in a REGISTER (`magic`, a run-time value the
// assembler cannot fold): PRMT takes one immediate, and with a literal pattern ptxas spends it
// on the pattern and re-materialises the selector with a MOV before every PRMT.
template <int K> __device__ __forceinline__ float byte_as_biased_float(uint32_t w, uint32_t magic) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(magic), "n"(0x7540 + K));
    return __uint_as_float(r);
}

// Correctly rounded fp64 division by a divisor that is the same for every element of a launch (the min-max spans
// of elvis.py:864-867 and utils.py:686): a / d == __ddiv_rn(a, d) bit for bit, at 5 fused operations instead of the
// ~30 of a general division.  y = RN(1/d) once; then q0 = RN(a y), two residual corrections q <- RN(q + RN(a - d q) y).
// After the first correction q is a faithful quotient, and by Markstein's theorem (IBM J. R&D 34, 1990; Muller et al.,
// Handbook of Floating-Point Arithmetic, 5.3) one more correction with a correctly rounded reciprocal and an exact
// residual (the FMA) yields the correctly rounded quotient.  The residuals are exact only away from the over- and
// underflow thresholds, so operands outside [2^-500, 2^500] (and NaN / infinity) take the general division; zeros
// keep their sign.  tests/test_gpu_parity.py::test_invariant_division compares the two on adversarial operands.
struct InvariantDivisor {
    double d, y;
    bool fast;
    __device__ explicit InvariantDivisor(double divisor) : d(divisor) {
        fast = divisor >= 0x1p-500 && divisor <= 0x1p500;
        y = fast ? __drcp_rn(divisor) : 0.0;
    }
    __device__ __forceinline__ double divide(double a) const {
        // range tests on the exponent field: integer work, the fp64 pipe is the scarce one
        const unsigned hi = (unsigned)__double2hiint(a);
        if (fast) {
            if (((hi >> 20) & 0x7ffu) - 523u <= 1000u) {                 // 2^-500 <= |a| < 2^501
                double q = __dmul_rn(a, y);
                q = __fma_rn(__fma_rn(-d, q, a), y, q);
                return __fma_rn(__fma_rn(-d, q, a), y, q);
            }
            if (((hi << 1) | (unsigned)__double2loint(a)) == 0u) return a;   // +-0 / d, d > 0
        }
        return __ddiv_rn(a, d);
    }
};

}  // namespace elvis
