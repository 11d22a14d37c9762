// a1 on the 5th-generation tensor cores: the per-tile 8x8 DCT of elvis_score_sc_tc as a
// tcgen05 GEMM (spec: oracle/spec_scoring.py; reference call sites elvis.py:1014-1031,
// presley.py:202).
//
// Formulation.  The 2-D DCT of a flattened tile is a 64 x 64 matrix product, C = X . B^T with
// B[n][p] = a(u,r) a(v,c).  128 tiles form one M = 128 operand, so one frame step of a CTA is
//     D[128 x 64] (fp32, TMEM) = A[128 x 64] (fp16, TMEM) . (Bhi + Blo)^T (fp16, shared)
// issued as 2 x 4 tcgen05.mma (M128 N64 K16) by one thread.  Pixels minus 128 are exact in fp16
// and B is split hi + lo (|B - hi - lo| < 2^-22), so the products are exact and the result has
// fp32-accumulation accuracy.  C_t is computed afresh for every frame -- no running sum, so
// SC / TC do not depend on how the clip is chunked.
//
// Roles.  Worker thread m of 128 owns tile m = TMEM lane m through a run of frames: it streams
// the tile's 8 luma rows through a private cp.async ring (as the CUDA-core kernel does), expands
// the 64 bytes to fp16 with PRMT + one HADD2 per pair, and writes them straight into the A
// operand in TENSOR MEMORY (tcgen05.st) -- A never touches shared memory, whose bandwidth would
// otherwise bound the MMA (SS-mode re-reads A per instruction).  One lane of the warp whose turn
// it is (they rotate per frame; a fifth warp would cost every CTA a sixth warp's registers)
// issues the MMAs once all 128 rows have arrived and commits them to an mbarrier; each worker
// then reads its lane of D back (tcgen05.ld) and does
// the weighted |C| and |C - C_prev| sums that are SC and TC.  A and D are double buffered, so
// the expansion of frame t+1 and the sums of frame t-1 overlap the MMAs of frame t.  Per tile
// the CUDA cores issue ~260 instructions instead of the ~900 of the butterfly kernel.
#include "score_params.cuh"
#include <cuda_fp16.h>

namespace elvis {
namespace {

#include "score_umma_tables.inc"

constexpr int kWorkers = 128;            // worker threads = TMEM lanes = tiles per MMA
constexpr int kUmmaThreads = kWorkers;
constexpr int kUmmaRing = 4;             // cp.async ring depth (frames)
constexpr uint32_t kTmemCols = 256;      // A[2] x 32 + D[2] x 64 columns, rounded to a power of two
constexpr uint32_t kColA = 0, kColD = 64;
constexpr uint32_t kOffB = 0, kOffRing = 16384, kOffBar = kOffRing + kUmmaRing * 8 * kWorkers * 8, kOffTmem = kOffBar + 32;
constexpr uint32_t kUmmaSmem = kOffTmem + 16 + 1024;   // + slack to align the base to 1024 B (128 B swizzle atom)
// instruction descriptor: D fp32, A/B fp16 K-major, N = 64, M = 128
constexpr uint32_t kIdesc = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&a)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(a[14]), "r"(a[15]), "r"(a[16]), "r"(a[17]), "r"(a[18]), "r"(a[19]), "r"(a[20]), "r"(a[21]), "r"(a[22]), "r"(a[23]), "r"(a[24]), "r"(a[25]), "r"(a[26]), "r"(a[27]), "r"(a[28]), "r"(a[29]), "r"(a[30]), "r"(a[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(taddr) : "memory");
}

// D[tmem] (+)= A[tmem] . B[smem]^T
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(kIdesc), "r"(accumulate) : "memory");
}

// K-major SWIZZLE_128B operand: 8-row groups 1024 B apart, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t b_descriptor(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <int R>
__global__ void __launch_bounds__(kUmmaThreads, 2) score_umma_kernel(const ScoreParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    uint2* ring = reinterpret_cast<uint2*>(sm + kOffRing);   // [kUmmaRing][8][kWorkers]
    const uint32_t bar_a = base + kOffBar, bar_d = bar_a + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // one CTA = 4 warp units (each R x 32/R tiles of one block row) of ONE temporal chunk, so that
    // its workers and the MMA warp step through the same frames
    constexpr int TW = 32 / R;
    const int per_chunk = p.By * p.tiles_x;
    const int groups = (per_chunk + 3) / 4;
    const int chunk = blockIdx.x / groups;
    const int group = blockIdx.x - chunk * groups;
    const int t0 = chunk * p.chunk_len;
    const int t1 = min(p.T, t0 + p.chunk_len);
    const bool has_prev = (t0 > 0) || (p.halo != nullptr);
    const int t_start = has_prev ? t0 - 1 : t0;
    const int n_iter = t1 - t_start;

    if (warp == 0) {
        if (lane == 0) {
            mbar_init(bar_a, kWorkers);
            mbar_init(bar_a + 8, kWorkers);
            mbar_init(bar_d, 1);
            mbar_init(bar_d + 8, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + kOffTmem), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 1024; i += kUmmaThreads)
        reinterpret_cast<uint4*>(sm + kOffB)[i] = reinterpret_cast<const uint4*>(&kUmmaB[0][0])[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // B is read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + kOffTmem);

    {
        // ---- workers
        const int unit = group * 4 + warp;
        const bool unit_ok = unit < per_chunk;
        const int by = unit_ok ? unit / p.tiles_x : 0;
        const int tx = unit_ok ? unit - by * p.tiles_x : 0;
        const int tr = lane / TW, tcx = lane % TW;
        const int tile_col = tx * TW + tcx;
        const bool valid = unit_ok && tile_col < p.Bx * R;
        const bool leader = valid && tr == 0 && (tcx % R) == 0;
        const int bxi = tile_col / R;
        const int64_t tile_off = (int64_t)(by * R + tr) * 8 * p.row_stride + (int64_t)tile_col * 8;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        const uint32_t bias = p.magic16;     // 0x64006400: bytes become fp16 1024 + b under PRMT
        const uint64_t b_hi = b_descriptor(base + kOffB), b_lo = b_descriptor(base + kOffB + 8192);

        auto prefetch = [&](int slot, int t) {
            if (valid && t < t1) {
                const uint8_t* src = (t < 0 ? p.halo : p.y + (int64_t)t * p.frame_stride) + tile_off;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const uint32_t dst = smem_u32(&ring[(slot * 8 + r) * kWorkers + threadIdx.x]);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src + r * p.row_stride) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto produce = [&](int it) {
            prefetch((it + kUmmaRing - 1) % kUmmaRing, t_start + it + kUmmaRing - 1);
            asm volatile("cp.async.wait_group 3;" ::: "memory");
            static_assert(kUmmaRing == 4, "wait_group immediate");
            const int slot = it % kUmmaRing;
            uint32_t a[32];
            const __half2 off = __floats2half2_rn(1152.f, 1152.f);   // 1024 (PRMT bias) + 128 (centering)
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const uint2 w = valid ? ring[(slot * 8 + r) * kWorkers + threadIdx.x] : make_uint2(0u, 0u);
                uint32_t e[4];
                asm("prmt.b32 %0, %1, %2, 0x7150;" : "=r"(e[0]) : "r"(w.x), "r"(bias));
                asm("prmt.b32 %0, %1, %2, 0x7352;" : "=r"(e[1]) : "r"(w.x), "r"(bias));
                asm("prmt.b32 %0, %1, %2, 0x7150;" : "=r"(e[2]) : "r"(w.y), "r"(bias));
                asm("prmt.b32 %0, %1, %2, 0x7352;" : "=r"(e[3]) : "r"(w.y), "r"(bias));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const __half2 h = __hsub2(*reinterpret_cast<const __half2*>(&e[j]), off);
                    a[4 * r + j] = *reinterpret_cast<const uint32_t*>(&h);
                }
            }
            tmem_st32(tmem + lane_base + kColA + 32 * (it & 1), a);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            const uint32_t buf = it & 1;
            mbar_arrive(bar_a + 8 * buf);
            if (warp == (it & 3)) {
                // this warp's turn to issue: all 128 rows of A(it) are in tensor memory, and every
                // worker has drained D(it - 2) (its tcgen05.ld precedes this arrive in program order)
                mbar_wait(bar_a + 8 * buf, (it >> 1) & 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t d_tmem = tmem + kColD + 64 * buf, a_tmem = tmem + kColA + 32 * buf;
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_ts(d_tmem, a_tmem + 8 * k, b_hi + 2 * k, k > 0);   // 16 fp16 = 8 columns = 32 B
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_ts(d_tmem, a_tmem + 8 * k, b_lo + 2 * k, 1);
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_d + 8 * buf) : "memory");
                }
                __syncwarp();
            }
        };

        float smin = __int_as_float(0x7f800000), smax = 0.f, tmin = __int_as_float(0x7f800000), tmax = 0.f;
        // c <- coefficients of frame t_start + it; pr = those of the frame before
        auto consume = [&](int it, float2 (&c)[32], const float2 (&pr)[32]) {
            mbar_wait(bar_d + 8 * (it & 1), (it >> 1) & 1);
            tc_fence_after();
            uint32_t v0[32], v1[32];
            const uint32_t d_addr = tmem + lane_base + kColD + 64 * (it & 1);
            tmem_ld32(d_addr, v0);
            tmem_ld32(d_addr + 32, v1);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                c[i] = make_float2(__uint_as_float(v0[2 * i]), __uint_as_float(v0[2 * i + 1]));
                c[16 + i] = make_float2(__uint_as_float(v1[2 * i]), __uint_as_float(v1[2 * i + 1]));
            }
            float s_part[8], d_part[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float s = 0.f, d = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = u * 4 + j;
                    const float2 df = __ffma2_rn(pr[i], make_float2(-1.f, -1.f), c[i]);   // C_t - C_{t-1}, one rounding
                    if (i != 0) {                                                          // DC carries no texture energy
                        s = fmaf(fabsf(c[i].x), kUmmaW[2 * i], s);
                        d = fmaf(fabsf(df.x), kUmmaW[2 * i], d);
                    }
                    s = fmaf(fabsf(c[i].y), kUmmaW[2 * i + 1], s);
                    d = fmaf(fabsf(df.y), kUmmaW[2 * i + 1], d);
                }
                s_part[u] = s;
                d_part[u] = d;
            }
            float s = ((s_part[0] + s_part[1]) + (s_part[2] + s_part[3])) + ((s_part[4] + s_part[5]) + (s_part[6] + s_part[7]));
            float d = ((d_part[0] + d_part[1]) + (d_part[2] + d_part[3])) + ((d_part[4] + d_part[5]) + (d_part[6] + d_part[7]));
#pragma unroll
            for (int m = 1; m < R; m <<= 1) {
                s += __shfl_xor_sync(0xffffffffu, s, m);
                d += __shfl_xor_sync(0xffffffffu, d, m);
            }
#pragma unroll
            for (int m = TW; m < 32; m <<= 1) {
                s += __shfl_xor_sync(0xffffffffu, s, m);
                d += __shfl_xor_sync(0xffffffffu, d, m);
            }
            const int t = t_start + it;
            if (t >= t0 && leader) {
                const float scv = s * p.inv_area;
                const float tcv = (t == 0 && p.halo == nullptr) ? 0.f : d * p.inv_area;
                const int64_t o = ((int64_t)t * p.By + by) * p.Bx + bxi;
                p.sc[o] = scv;
                p.tc[o] = tcv;
                if (t >= p.mm_begin && t < p.mm_end) {
                    smin = fminf(smin, scv);
                    smax = fmaxf(smax, scv);
                    tmin = fminf(tmin, tcv);
                    tmax = fmaxf(tmax, tcv);
                }
            }
        };

        float2 ca[32], cb[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) cb[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int s = 0; s < kUmmaRing - 1; ++s) prefetch(s, t_start + s);
        produce(0);
        for (int it = 0; it < n_iter; it += 2) {
            if (it + 1 < n_iter) produce(it + 1);
            consume(it, ca, cb);
            if (it + 1 >= n_iter) break;
            if (it + 2 < n_iter) produce(it + 2);
            consume(it + 1, cb, ca);
        }

        if (p.mm != nullptr) {
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) {
                smin = fminf(smin, __shfl_xor_sync(0xffffffffu, smin, m));
                smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, m));
                tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, m));
                tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, m));
            }
            if (lane == 0) {
                atomicMin(p.mm + 0, __float_as_uint(smin));
                atomicMax(p.mm + 1, __float_as_uint(smax));
                atomicMin(p.mm + 2, __float_as_uint(tmin));
                atomicMax(p.mm + 3, __float_as_uint(tmax));
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

template <int R>
int launch_umma(const ScoreParams& p, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(score_umma_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUmmaSmem);
        if (e != cudaSuccess) return cuda_fail(e);
        // two CTAs per SM need ~104 KB of shared memory: ask for the large carve-out
        e = cudaFuncSetAttribute(score_umma_kernel<R>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return cuda_fail(e);
        configured = true;
    }
    const int groups = (p.By * p.tiles_x + 3) / 4;
    score_umma_kernel<R><<<groups * p.n_chunks, kUmmaThreads, kUmmaSmem, st>>>(p);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

}  // namespace

int launch_score_umma(ScoreParams p, int block_size, cudaStream_t st) {
    const int R = block_size / 8;
    const int TW = 32 / R;
    p.tiles_x = (p.Bx * R + TW - 1) / TW;
    p.tiles_y = p.By;
    p.magic16 = 0x64006400u;
    switch (R) {
        case 1: return launch_umma<1>(p, st);
        case 2: return launch_umma<2>(p, st);
        default: return launch_umma<4>(p, st);
    }
}

}  // namespace elvis
