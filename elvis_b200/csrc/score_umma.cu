// a1 on the 5th-generation tensor cores: the per-tile 8x8 DCT of elvis_score_sc_tc as a
// tcgen05 GEMM (spec: oracle/spec_scoring.py; reference call sites elvis.py:1014-1031,
// presley.py:202).
//
// Formulation.  The 2-D DCT of a flattened tile is a 64 x 64 matrix product, C = X . B^T with
// B[n][p] = a(u,r) a(v,c).  128 tiles form one M = 128 operand, so one frame step of a group is
//     D[128 x 64] (fp32, TMEM) = A[128 x 64] (fp16, TMEM) . (Bhi + Blo)^T (fp16, shared)
// issued as 2 x 4 tcgen05.mma (M128 N64 K16) by one thread.  Pixels minus 128 are exact in fp16
// and B is split hi + lo (|B - hi - lo| < 2^-22), so the products are exact and the result has
// fp32-accumulation accuracy.  C_t is computed afresh for every frame -- no running sum, so
// SC / TC do not depend on how the clip is chunked.
//
// Roles.  A CTA is one group of 128 tiles (ELVIS_UMMA_GROUPS = 1: two CTAs share an SM and its 512
// TMEM columns; 5 % faster than one CTA of two groups, 356 vs 374 us).  Tile m of the group is
// TMEM lane m; it is served by TWO worker threads (lane m % 32 of warps q and q + 4, q = m / 32
// -- both reach lanes 32q..), which walk it through a run of frames:
//   * both read half of the tile's 8 luma rows from a shared-memory ring, expand the 32 bytes
//     to fp16 (PRMT + one HADD2 per pair) and write them straight into the A operand in TENSOR
//     MEMORY (tcgen05.st) -- A never touches shared memory again, whose bandwidth would
//     otherwise bound the MMA (SS-mode re-reads A per instruction);
//   * after the MMAs each reads its half of the tile's lane of D(t) and of D(t-1) back
//     (tcgen05.ld) and sums w |C_t| (SC) and w |C_t - C_{t-1}| (TC) over its 32 coefficients;
//     the upper-half warp hands its block sums to the lower-half warp through shared memory
//     and a 64-thread named barrier, and that one writes SC and TC.  D is triple buffered, so
//     the previous frame's coefficients are still in tensor memory and no thread carries
//     coefficients in registers between frames.
// The warp after the workers issues the group's MMAs (one lane) once its 8 worker warps have
// published A and commits them to an mbarrier.  The last warp runs the TMA ring, one lane per
// 32-tile unit: one 3-D
// box (8R rows x 256/R bytes x 1 frame, R = block_size / 8) per frame, kUmmaRing frames deep,
// full / empty mbarriers per ring slot.  The frame loop is unrolled by the ring depth so that
// ring slot, A / D buffer and most mbarrier phases are compile-time constants.
#include "score_params.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstring>
#include <type_traits>

namespace elvis {
namespace {

#include "score_umma_tables.inc"

#ifndef ELVIS_UMMA_GROUPS
#define ELVIS_UMMA_GROUPS 1
#endif
constexpr int kGroups = ELVIS_UMMA_GROUPS;   // groups of 128 tiles per CTA, each with its own A / D buffers
constexpr int kCtasPerSm = 2 / kGroups;       // tensor memory holds two groups per SM
constexpr int kUnitsPerCta = kGroups * 4;     // warp units (32 tiles) per CTA
constexpr int kWorkerWarps = 2 * kUnitsPerCta;   // a lower-half and an upper-half warp per unit
constexpr int kMmaWarp = kWorkerWarps;
constexpr int kTmaWarp = kWorkerWarps + 1;
constexpr int kUmmaThreads = (kWorkerWarps + 2) * 32;
#ifndef ELVIS_UMMA_RING
#define ELVIS_UMMA_RING 6
#endif
constexpr int kUmmaRing = ELVIS_UMMA_RING;    // TMA ring depth (frames)
constexpr uint32_t kBoxBytes = 2048;     // one warp unit of one frame: 32 tiles x 64 bytes
constexpr uint32_t kGroupCols = 256;     // per group: A[2] x 32 + D[3] x 64 columns
constexpr uint32_t kTmemCols = kGroups * kGroupCols;
constexpr uint32_t kColA = 0, kColD = 64;
constexpr uint32_t kOffB = 0, kOffRing = 16384;
constexpr uint32_t kOffBar = kOffRing + kUnitsPerCta * kUmmaRing * kBoxBytes;
constexpr uint32_t kGroupBar = 40;       // per group: a_full[2] +0, d_full[3] +16
constexpr uint32_t kUnitBar = 16 * kUmmaRing;   // per unit: ring_full[] +0, ring_empty[] +8 kUmmaRing
constexpr uint32_t kOffUnitBar = kOffBar + kGroups * kGroupBar;
constexpr uint32_t kOffXchg = kOffUnitBar + kUnitsPerCta * kUnitBar;   // per unit: 2 x 32 float2 partial sums
constexpr uint32_t kOffXbar = kOffXchg + kUnitsPerCta * 512;         // per unit: x_full[2] mbarriers of the exchange slots
constexpr uint32_t kOffTmem = kOffXbar + kUnitsPerCta * 16;
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
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int t, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(t), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&a)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(a[14]), "r"(a[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(taddr) : "memory");
}

// 64-thread named barrier of the two warps that share a unit.  Immediate ids: with a register id
// ptxas reserves all 16 barriers of the SM for the CTA and nothing else could share the SM.
__device__ __forceinline__ void pair_barrier(int unit_local) {
    if (kUnitsPerCta > 4 && unit_local >= 4) {
        switch (unit_local) {
            case 4: asm volatile("bar.sync 5, 64;" ::: "memory"); break;
            case 5: asm volatile("bar.sync 6, 64;" ::: "memory"); break;
            case 6: asm volatile("bar.sync 7, 64;" ::: "memory"); break;
            default: asm volatile("bar.sync 8, 64;" ::: "memory"); break;
        }
        return;
    }
    switch (unit_local) {
        case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
        case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
        case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
        default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
    }
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

// XBAR: how the upper-half warp of a unit hands its partial sums to the lower-half warp.  false: a 64-thread named
// barrier per frame (round 1; both warps wait for each other).  true: a one-way mbarrier per exchange slot -- the
// upper half arrives and moves on, only the lower half ever waits.
template <int R, bool XBAR, int ROLLED>
__global__ void __launch_bounds__(kUmmaThreads, kCtasPerSm)
score_umma_kernel(const __grid_constant__ CUtensorMap tm_clip, const __grid_constant__ CUtensorMap tm_halo, const ScoreParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int TW = 32 / R;             // tiles per warp-unit row
    constexpr int kPitch = 256 / R;        // bytes per row of a warp unit's box (8R rows)

    // one CTA = 8 warp units (each R x 32/R tiles of one block row) of ONE temporal chunk, so that
    // all of its warps step through the same frames
    const int per_chunk = p.By * p.tiles_x;
    const int ctas_per_chunk = (per_chunk + kUnitsPerCta - 1) / kUnitsPerCta;
    const int chunk = blockIdx.x / ctas_per_chunk;
    const int first_unit = (blockIdx.x - chunk * ctas_per_chunk) * kUnitsPerCta;
    const int t0 = chunk * p.chunk_len;
    const int t1 = min(p.T, t0 + p.chunk_len);
    const bool has_prev = (t0 > 0) || (p.halo != nullptr);
    const int t_start = has_prev ? t0 - 1 : t0;
    const int n_iter = t1 - t_start;

    if (warp == kMmaWarp) {
        if (lane == 0) {
            for (int g = 0; g < kGroups; ++g) {
                const uint32_t bar = base + kOffBar + kGroupBar * g;
                mbar_init(bar, 256);          // a_full: every thread of the group's 8 worker warps
                mbar_init(bar + 8, 256);
                mbar_init(bar + 16, 1);       // d_full: tcgen05.commit
                mbar_init(bar + 24, 1);
                mbar_init(bar + 32, 1);
            }
            for (int u = 0; u < kUnitsPerCta; ++u)
                for (int s = 0; s < kUmmaRing; ++s) {
                    mbar_init(base + kOffUnitBar + kUnitBar * u + 8 * s, 1);                   // ring_full: expect_tx
                    mbar_init(base + kOffUnitBar + kUnitBar * u + 8 * (kUmmaRing + s), 64);    // ring_empty: both reader warps
                }
            for (int u = 0; u < kUnitsPerCta; ++u) {
                mbar_init(base + kOffXbar + 16 * u, 32);          // x_full[0], x_full[1]: the 32 lanes of the upper-half warp
                mbar_init(base + kOffXbar + 16 * u + 8, 32);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_clip) : "memory");
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

    if (warp == kMmaWarp) {
        // ---- MMA issue: lane g serves group g
        if (lane < kGroups) {
            const int g = lane;
            const uint64_t b_hi = b_descriptor(base + kOffB), b_lo = b_descriptor(base + kOffB + 8192);
            const uint32_t bar = base + kOffBar + kGroupBar * g;
            uint32_t dbuf = 0;
            for (int it = 0; it < n_iter; ++it) {
                const uint32_t buf = it & 1;
                // all 128 rows of A(it) are in tensor memory, and every worker of the group is
                // done with D(it - 3), which it last read as the "previous frame" of it - 2
                // (those tcgen05.ld precede its arrive in program order)
                {   // back off between polls: the spin would take issue slots from the workers
                    uint32_t done = 0;
                    for (;;) {
                        asm volatile(
                            "{\n\t.reg .pred p;\n\t"
                            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                            "selp.u32 %0, 1, 0, p;\n\t}"
                            : "=r"(done) : "r"(bar + 8 * buf), "r"((uint32_t)(it >> 1) & 1u) : "memory");
                        if (done) break;
                        __nanosleep(32);
                    }
                }
                tc_fence_after();
                const uint32_t d_tmem = tmem + kGroupCols * g + kColD + 64 * dbuf, a_tmem = tmem + kGroupCols * g + kColA + 32 * buf;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ts(d_tmem, a_tmem + 8 * k, b_hi + 2 * k, k > 0);   // 16 fp16 = 8 columns = 32 B
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ts(d_tmem, a_tmem + 8 * k, b_lo + 2 * k, 1);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar + 16 + 8 * dbuf) : "memory");
                if (++dbuf == 3) dbuf = 0;
            }
        }
        __syncwarp();
    } else if (warp == kTmaWarp) {
        // ---- TMA ring: lane u feeds unit u (warps past the end of the chunk redo its last unit)
        if (lane < kUnitsPerCta) {
            const int unit = min(first_unit + lane, per_chunk - 1);
            const int by = unit / p.tiles_x, tx = unit - by * p.tiles_x;
            const int x = tx * kPitch, y = by * 8 * R;
            const uint32_t bar_full = base + kOffUnitBar + kUnitBar * lane, bar_empty = bar_full + 8 * kUmmaRing;
            const uint32_t dst0 = base + kOffRing + lane * kUmmaRing * kBoxBytes;
            int slot = 0;
            uint32_t phase = 0;
            for (int f = 0; f < n_iter; ++f) {
                if (f >= kUmmaRing) {            // frame f - kUmmaRing has been read by both warps
                    uint32_t done = 0;
                    for (;;) {
                        asm volatile(
                            "{\n\t.reg .pred p;\n\t"
                            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                            "selp.u32 %0, 1, 0, p;\n\t}"
                            : "=r"(done) : "r"(bar_empty + 8 * slot), "r"(phase ^ 1u) : "memory");
                        if (done) break;
                        __nanosleep(64);
                    }
                }
                const int t = t_start + f;
                mbar_arrive_expect_tx(bar_full + 8 * slot, kBoxBytes);
                tma_load_3d(dst0 + slot * kBoxBytes, t < 0 ? &tm_halo : &tm_clip, x, y, max(t, 0), bar_full + 8 * slot);
                if (++slot == kUmmaRing) {
                    slot = 0;
                    phase ^= 1;
                }
            }
        }
        __syncwarp();
    } else {
        // ---- workers: warp = 8 grp + 4 role + q
        const int grp = warp >> 3, role = (warp >> 2) & 1, q = warp & 3;
        const int unit_local = grp * 4 + q;
        const uint32_t bar_a = base + kOffBar + kGroupBar * grp, bar_d = bar_a + 16;
        const uint32_t bar_full = base + kOffUnitBar + kUnitBar * unit_local, bar_empty = bar_full + 8 * kUmmaRing;
        // warps past the end of the chunk redo its last unit (uniform control flow) and store nothing
        const bool unit_ok = first_unit + unit_local < per_chunk;
        const int unit = min(first_unit + unit_local, per_chunk - 1);
        const int by = unit / p.tiles_x;
        const int tx = unit - by * p.tiles_x;
        const int tr = lane / TW, tcx = lane % TW;
        const int tile_col = tx * TW + tcx;
        const bool valid = unit_ok && tile_col < p.Bx * R;
        const bool leader = valid && tr == 0 && (tcx % R) == 0;   // valid implies unit_ok
        const int bxi = tile_col / R;
        const uint32_t lane_base = ((uint32_t)(q * 32) << 16) + kGroupCols * grp;   // a warp reaches lanes 32 (warp % 4) ..
        const uint32_t bias = p.magic16;     // 0x64006400: bytes become fp16 1024 + b under PRMT
        const uint32_t ring0 = base + kOffRing + unit_local * kUmmaRing * kBoxBytes;
        // rows 4 role .. 4 role + 3 of this thread's tile inside the unit's box
        const uint32_t tile_smem = ring0 + (tr * 8 + 4 * role) * kPitch + tcx * 8;
        float smin = __int_as_float(0x7f800000), smax = 0.f, tmin = __int_as_float(0x7f800000), tmax = 0.f;
        // output cursors of the leader lanes: element (t_start, by, bxi), one frame per step
        const int64_t out_step = (int64_t)p.By * p.Bx;
        float* out_sc = p.sc + ((int64_t)t_start * p.By + by) * p.Bx + bxi;
        float* out_tc = p.tc + ((int64_t)t_start * p.By + by) * p.Bx + bxi;
        float2* const xchg = reinterpret_cast<float2*>(sm + kOffXchg + unit_local * 512) + lane;

        // expand this thread's 4 rows of the frame in ring slot SLOT into its half of A's lane BUF
        // (asynchronous store), then publish it and release the slot
        auto produce = [&](const int slot, const int buf, const uint32_t full_parity) {
            uint32_t a[16];
            const __half2 off = __floats2half2_rn(1152.f, 1152.f);   // 1024 (PRMT bias) + 128 (centering)
            mbar_wait(bar_full + 8 * slot, full_parity);
            const uint32_t src = tile_smem + slot * kBoxBytes;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                uint2 w;
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w.x), "=r"(w.y) : "r"(src + r * kPitch));
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
            tmem_st16(tmem + lane_base + kColA + 32 * buf + 16 * role, a);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            mbar_arrive(bar_a + 8 * buf);
            mbar_arrive(bar_empty + 8 * slot);      // the store consumed every byte this thread read from the slot
        };

        // Each thread covers 32 of the tile's 64 coefficients (rows 4 role .. 4 role + 3): SC part
        // = sum w |C_t|, TC part = sum w |C_t - C_{t-1}| with C_{t-1} = the D buffer written one
        // frame earlier.  2 x 4 independent FMA chains, fixed order => deterministic.
        const uint32_t bar_x = base + kOffXbar + 16 * unit_local;
        auto consume = [&](const int t, const int dbuf, const int dprev, const uint32_t d_parity, const int xslot, const uint32_t x_parity, auto role_c) {
            constexpr int ROLE = decltype(role_c)::value;
            mbar_wait(bar_d + 8 * dbuf, d_parity);
            tc_fence_after();
            uint32_t v[32], q[32];
            tmem_ld32(tmem + lane_base + kColD + 64 * dbuf + 32 * ROLE, v);
            tmem_ld32(tmem + lane_base + kColD + 64 * dprev + 32 * ROLE, q);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float sa[4] = {0.f, 0.f, 0.f, 0.f}, da[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int n = 32 * ROLE + 2 * i;
                const float2 c = make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                const float2 df = __ffma2_rn(make_float2(__uint_as_float(q[2 * i]), __uint_as_float(q[2 * i + 1])), make_float2(-1.f, -1.f), c);   // one rounding
                if (n != 0) {                                    // DC carries no texture energy
                    sa[i & 3] = fmaf(fabsf(c.x), kUmmaW[n], sa[i & 3]);
                    da[i & 3] = fmaf(fabsf(df.x), kUmmaW[n], da[i & 3]);
                }
                sa[(i + 2) & 3] = fmaf(fabsf(c.y), kUmmaW[n + 1], sa[(i + 2) & 3]);
                da[(i + 2) & 3] = fmaf(fabsf(df.y), kUmmaW[n + 1], da[(i + 2) & 3]);
            }
            float s = (sa[0] + sa[1]) + (sa[2] + sa[3]);
            float d = (da[0] + da[1]) + (da[2] + da[3]);
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
            // the upper half hands its sums to the lower half through one of two slots: the write two
            // frames later comes after the next barrier, which the reader reaches after this read
            // The slot written for frame t is reused for frame t + 2.  By then the lower half has read it: its read of
            // frame t precedes (program order) its own A(t + 2) store, which the MMA of t + 2 -- and with it the upper
            // half's D(t + 2) wait -- depends on.
            if (ROLE == 1) xchg[32 * xslot] = make_float2(s, d);
            if (XBAR) {
                if (ROLE == 1) mbar_arrive(bar_x + 8 * xslot);
                else mbar_wait(bar_x + 8 * xslot, x_parity);
            } else {
                pair_barrier(unit_local);
            }
            if (ROLE == 0) {
                const float2 o = xchg[32 * xslot];
                s += o.x;
                d += o.y;
                if (t >= t0 && leader) {
                    const float scv = s * p.inv_area;
                    // the first frame without a halo has no predecessor (and D(-1) is uninitialised)
                    const float tcv = (t == 0 && p.halo == nullptr) ? 0.f : d * p.inv_area;
                    *out_sc = scv;
                    *out_tc = tcv;
                    if (t >= p.mm_begin && t < p.mm_end) {
                        smin = fminf(smin, scv);
                        smax = fmaxf(smax, scv);
                        tmin = fminf(tmin, tcv);
                        tmax = fmaxf(tmax, tcv);
                    }
                }
                out_sc += out_step;
                out_tc += out_step;
            }
        };
        // Frame loop, unrolled by the ring depth (a multiple of 2 and 3): step i of a round works
        // on ring slot (i + 1) % 6, A buffer (i + 1) & 1, D buffer i % 3.  Parities: ring and a_full
        // barriers flip once per round (6 = kUmmaRing uses; 3 uses of each A buffer), d_full
        // completes twice per round, so its parity is (i / 3) & 1.
        static_assert(kUmmaRing == 6, "the unrolled frame loop assumes a ring of 6");
        auto run = [&](auto role_c) {
            produce(0, 0, 0u);
            if (!ROLLED) {
                uint32_t round_parity = 0;
#pragma unroll 1
                for (int it0 = 0; it0 < n_iter; it0 += 6) {
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        const int it = it0 + i;
                        if (it >= n_iter) break;
                        if (it + 1 < n_iter) produce((i + 1) % 6, (i + 1) & 1, i == 5 ? round_parity ^ 1u : round_parity);
                        // exchange slot i & 1 has been used (it >> 1) times before: its phase parity is round ^ (i >> 1)
                        consume(t_start + it, i % 3, (i + 2) % 3, (uint32_t)(i / 3) & 1u, i & 1, round_parity ^ ((uint32_t)(i >> 1) & 1u), role_c);
                    }
                    round_parity ^= 1u;
                }
            } else if (ROLLED == 2) {
                // The same schedule with the ring slot, the D buffer and their parities carried in registers and the
                // loop unrolled by two only (A buffer and exchange slot stay compile-time): a third of the code, so
                // that the two roles' loops fit the instruction caches (ncu: `no_instruction` was the top stall reason).
                int slot = 1, dbuf = 0, dprev = 2;          // next frame to produce: 1; frame 0 lands in D buffer 0
                uint32_t ring_parity = 0, d_parity = 0, x_parity = 0;
#pragma unroll 1
                for (int it0 = 0; it0 < n_iter; it0 += 2) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int it = it0 + i;
                        if (it >= n_iter) break;
                        if (it + 1 < n_iter) produce(slot, (i + 1) & 1, ring_parity);
                        if (++slot == kUmmaRing) {
                            slot = 0;
                            ring_parity ^= 1u;
                        }
                        consume(t_start + it, dbuf, dprev, d_parity, i, x_parity, role_c);
                        dprev = dbuf;
                        if (++dbuf == 3) {
                            dbuf = 0;
                            d_parity ^= 1u;
                        }
                    }
                    x_parity ^= 1u;
                }
            } else {
                // fully rolled: every index in registers
                int slot = 1, dbuf = 0, dprev = 2, abuf = 1, xslot = 0;
                uint32_t ring_parity = 0, d_parity = 0, x_parity = 0;
#pragma unroll 1
                for (int it = 0; it < n_iter; ++it) {
                    if (it + 1 < n_iter) produce(slot, abuf, ring_parity);
                    abuf ^= 1;
                    if (++slot == kUmmaRing) {
                        slot = 0;
                        ring_parity ^= 1u;
                    }
                    consume(t_start + it, dbuf, dprev, d_parity, xslot, x_parity, role_c);
                    dprev = dbuf;
                    if (++dbuf == 3) {
                        dbuf = 0;
                        d_parity ^= 1u;
                    }
                    x_parity ^= (uint32_t)xslot;            // flips after the frame that used slot 1
                    xslot ^= 1;
                }
            }
        };
        if (role == 0) run(std::integral_constant<int, 0>{});
        else run(std::integral_constant<int, 1>{});

        if (p.mm != nullptr && role == 0) {
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) {
                smin = fminf(smin, __shfl_xor_sync(0xffffffffu, smin, m));
                smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, m));
                tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, m));
                tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, m));
            }
            if (lane == 0) {
                // non-negative floats order like their bit patterns
                atomicMin(p.mm + 0, __float_as_uint(smin));
                atomicMax(p.mm + 1, __float_as_uint(smax));
                atomicMin(p.mm + 2, __float_as_uint(tmin));
                atomicMax(p.mm + 3, __float_as_uint(tmax));
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &st) != cudaSuccess ||
            st != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// (W, H, T) uint8 tensor, box = one warp unit (256/R bytes x 8R rows x 1 frame), zero fill out of bounds
bool make_unit_map(CUtensorMap* m, const uint8_t* ptr, int W, int H, int T, int64_t row_stride, int64_t frame_stride, int R) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T};
    cuuint64_t strides[2] = {(cuuint64_t)row_stride, (cuuint64_t)(T > 1 ? frame_stride : row_stride * H)};
    cuuint32_t box[3] = {(cuuint32_t)(256 / R), (cuuint32_t)(8 * R), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(ptr), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int R, bool XBAR, int ROLLED>
int launch_umma(const ScoreParams& p, const CUtensorMap& tm_clip, const CUtensorMap& tm_halo, cudaStream_t st) {
    // no carve-out preference: a kernel that forces its own L1 / shared split cannot overlap with the
    // shrink / stretch kernels of the other stream (measured: pipelined step 1.32 instead of 1.07 ms)
    // ELVIS_UMMA_SMEM_PAD (bytes, experiment): extra dynamic shared memory per CTA -- 50 000 leaves room for ONE scoring
    // CTA per SM, so that half of the register file stays free for the kernels of the other streams
    static const int pad = [] {
        const char* e = getenv("ELVIS_UMMA_SMEM_PAD");
        const int v = e ? atoi(e) : 0;
        return v < 0 ? 0 : (v > 150000 ? 150000 : v);
    }();
    static PerDeviceOnce configured;   // the attribute is per (kernel, device)
    const cudaError_t e = configured.run([] {
        return cudaFuncSetAttribute(score_umma_kernel<R, XBAR, ROLLED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUmmaSmem + pad);
    });
    if (e != cudaSuccess) return cuda_fail(e);
    const int ctas_per_chunk = (p.By * p.tiles_x + kUnitsPerCta - 1) / kUnitsPerCta;
    score_umma_kernel<R, XBAR, ROLLED><<<ctas_per_chunk * p.n_chunks, kUmmaThreads, kUmmaSmem + pad, st>>>(tm_clip, tm_halo, p);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

}  // namespace

int score_umma_units_per_cta() { return kUnitsPerCta; }
int score_umma_ctas_per_sm() { return kCtasPerSm; }

// Returns ELVIS_ERR_UNSUPPORTED when the driver cannot encode tensor maps (the caller then uses
// the CUDA-core kernel).  Plane, strides and halo must be 16-byte aligned.
int launch_score_umma(ScoreParams p, int block_size, int plane_h, int plane_w, cudaStream_t st) {
    const int R = block_size / 8;
    const int TW = 32 / R;
    p.tiles_x = (p.Bx * R + TW - 1) / TW;
    p.tiles_y = p.By;
    p.magic16 = 0x64006400u;
    CUtensorMap tm_clip, tm_halo;
    memset(&tm_clip, 0, sizeof(tm_clip));
    memset(&tm_halo, 0, sizeof(tm_halo));
    if (!make_unit_map(&tm_clip, p.y, plane_w, plane_h, p.T, p.row_stride, p.frame_stride, R)) return ELVIS_ERR_UNSUPPORTED;
    if (!make_unit_map(&tm_halo, p.halo ? p.halo : p.y, plane_w, plane_h, 1, p.row_stride, p.frame_stride, R)) return ELVIS_ERR_UNSUPPORTED;
    const char* xe = getenv("ELVIS_UMMA_XBAR");        // 1 (default): one-way mbarrier exchange; 0: the round-1 named barrier
    const bool xbar = !(xe && xe[0] == '0');
    // frame loop: 1 (default) = not unrolled, every ring index in registers (3.6 % faster: the two roles' loops fit the
    // instruction caches); 2 = unrolled by two; 0 = unrolled by 6 with compile-time indices (round 1)
    const char* re = getenv("ELVIS_UMMA_ROLLED");
    const int rolled = re ? atoi(re) : 1;
    auto go = [&](auto r_c) {
        constexpr int RR = decltype(r_c)::value;
        if (rolled == 2) return xbar ? launch_umma<RR, true, 2>(p, tm_clip, tm_halo, st) : launch_umma<RR, false, 2>(p, tm_clip, tm_halo, st);
        if (rolled == 1) return xbar ? launch_umma<RR, true, 1>(p, tm_clip, tm_halo, st) : launch_umma<RR, false, 1>(p, tm_clip, tm_halo, st);
        return xbar ? launch_umma<RR, true, 0>(p, tm_clip, tm_halo, st) : launch_umma<RR, false, 0>(p, tm_clip, tm_halo, st);
    };
    switch (R) {
        case 1: return go(std::integral_constant<int, 1>{});
        case 2: return go(std::integral_constant<int, 2>{});
        default: return go(std::integral_constant<int, 4>{});
    }
}

}  // namespace elvis
