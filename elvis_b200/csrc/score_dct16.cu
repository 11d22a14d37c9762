// a1 with dct_size = block_size = 16 on the tensor cores (the transform size the reference asks EVCA for at its
// presley / benchmark block size: `python -m evca.main ... -b block_size`, elvis.py:1022-1023; presley.py:202).
// Spec: oracle/spec_scoring.py with n = 16 -- one 16 x 16 orthonormal DCT-II per block, weights
// w(u, v) = exp(|(u v / 256)^2 - 1|), DC excluded.  score_dctn.cu does the same on the CUDA cores (and N = 32).
//
// Formulation.  C = D X D^T for a 16 x 16 block X is two 16 x 16 x 16 products, each a pair of mma.sync.m16n8k16
// (f16 operands, fp32 accumulators) per 8-column half:
//   step 1   T = D X^T.  Pixels minus 128 are exact in fp16; D is split hi + lo (|D - hi - lo| < 2^-23 |D|), so
//            T = (Dhi + Dlo) X^T costs 2 x 2 HMMA and is exact up to the fp32 accumulation.
//   step 2   C = D T^T with T split hi + lo in fp16 as well: Dhi Thi + Dhi Tlo + Dlo Thi, 3 x 2 HMMA (the dropped
//            Dlo Tlo term is < 2^-22 of the result).
// A matrix held in the accumulator layout is, read as a B operand, its own transpose: thread (g, q) = (lane / 4,
// lane % 4) ends step 1 with T[g][8h + 2q + e] and T[g + 8][8h + 2q + e] (h, e in {0, 1}) -- exactly the elements
// B[k][n] = T[n][k] it must supply in step 2 for k in {2q, 2q+1, 2q+8, 2q+9}.  For step 1 the contraction index (the
// pixel column) is permuted so that the same thread's four B elements are four CONSECUTIVE pixels 4q..4q+3 of rows g
// and g + 8 (one 32-bit load each); the permutation is folded into the constant A operand.  No shuffle, no shared
// memory between the two steps.
//
// Roles.  A CTA is one (temporal chunk, block row, group of 8 adjacent blocks): eight consumer warps, one block each,
// walk the chunk's frames with C_{t-1} of their block in registers (8 values per thread); a ninth warp feeds them one
// 16-row x 128-byte TMA box per frame (128-byte swizzle: the consumers' loads are conflict-free) through a ring of
// four slots with full (transaction) and empty (8 arrivals) mbarriers.  C_t is computed afresh for every frame, so
// SC / TC do not depend on the chunking.  The DCT matrix (hi / lo halves) and the weights come from a host table
// passed as a kernel parameter.
#include "score_params.cuh"
#include "tma.cuh"
#include <cmath>
#include <cuda_fp16.h>

namespace elvis {
namespace {

constexpr int kWorkers = 8;                       // consumer warps = blocks per CTA
constexpr int kThreads = (kWorkers + 1) * 32;
constexpr int kRing = 4;
constexpr uint32_t kBox = 2048;                   // 16 rows x 128 bytes

struct Dct16Tables {
    uint16_t hi[256], lo[256];                    // D[v][x] = hi + lo as fp16 bit patterns
    float w[256];                                 // w[u][v], 0 for the DC term
};

const Dct16Tables& host_tables() {
    static const Dct16Tables tab = [] {
        Dct16Tables t;
        const double pi = 3.14159265358979323846;
        for (int v = 0; v < 16; ++v)
            for (int x = 0; x < 16; ++x) {
                const double cv = v == 0 ? std::sqrt(1.0 / 16.0) : std::sqrt(2.0 / 16.0);
                const double d = cv * std::cos((double)((2 * x + 1) * v) * pi / 32.0);
                const __half h = __float2half_rn((float)d);
                const __half l = __float2half_rn((float)(d - (double)__half2float(h)));
                t.hi[v * 16 + x] = *reinterpret_cast<const uint16_t*>(&h);
                t.lo[v * 16 + x] = *reinterpret_cast<const uint16_t*>(&l);
                const double q = (double)(v * x) / 256.0;           // here (v, x) plays (u, v)
                t.w[v * 16 + x] = (v == 0 && x == 0) ? 0.f : (float)std::exp(std::fabs(q * q - 1.0));
            }
        return t;
    }();
    return tab;
}

__device__ __forceinline__ void hmma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// (x, y) -> fp16 pair hi and the fp16 pair of the residuals
__device__ __forceinline__ void split_pair(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x, y);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

__global__ void __launch_bounds__(kThreads, 3) score_dct16_kernel(const __grid_constant__ CUtensorMap tm_clip, const __grid_constant__ CUtensorMap tm_halo,
                                                              const __grid_constant__ Dct16Tables tab, const ScoreParams p) {
    __shared__ __align__(1024) uint8_t s_ring[kRing][kBox];
    __shared__ __align__(8) uint64_t s_full[kRing], s_empty[kRing];
    __shared__ uint16_t s_hi[256], s_lo[256];
    __shared__ float s_w[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 256; i += kThreads) {
        s_hi[i] = tab.hi[i];
        s_lo[i] = tab.lo[i];
        s_w[i] = tab.w[i];
    }
    const uint32_t ring = tma::smem_u32(&s_ring[0][0]), bar_full = tma::smem_u32(&s_full[0]), bar_empty = tma::smem_u32(&s_empty[0]);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kRing; ++s) {
            tma::mbar_init(bar_full + 8u * s, 1);
            tma::mbar_init(bar_empty + 8u * s, kWorkers);
        }
        tma::mbar_init_fence();
    }
    __syncthreads();

    // work decomposition: blockIdx.x = (chunk, block row, group of 8 blocks)
    const int groups_x = p.tiles_x;
    const int per_chunk = p.By * groups_x;
    const int chunk = blockIdx.x / per_chunk;
    const int rem = blockIdx.x - chunk * per_chunk;
    const int by = rem / groups_x, gx = rem - by * groups_x;
    const int t0 = chunk * p.chunk_len;
    const int t1 = min(p.T, t0 + p.chunk_len);
    const bool has_prev = (t0 > 0) || (p.halo != nullptr);
    const int t_start = has_prev ? t0 - 1 : t0;
    const int n_iter = t1 - t_start;

    if (warp == kWorkers) {
        // ---- producer: one box per frame
        if (lane == 0) {
            for (int f = 0; f < n_iter; ++f) {
                const int s = f % kRing;
                if (f >= kRing) tma::mbar_wait(bar_empty + 8u * s, (uint32_t)((f / kRing) - 1) & 1u);
                const int t = t_start + f;
                tma::mbar_arrive_expect_tx(bar_full + 8u * s, kBox);
                tma::load_3d(ring + (uint32_t)s * kBox, t < 0 ? &tm_halo : &tm_clip, gx * 128, by * 16, max(t, 0), bar_full + 8u * s);
            }
        }
        return;
    }

    // ---- consumers
    const int g = lane >> 2, q = lane & 3;
    const int bx = gx * 8 + warp;
    const bool valid = bx < p.Bx;
    // A operands: step 1 with the contraction index permuted (k = 2q, 2q+1, 2q+8, 2q+9 <-> pixel columns 4q .. 4q+3), step 2 natural
    uint32_t a1h[4], a1l[4], a2h[4], a2l[4];
    auto pair16 = [](const uint16_t* m, int row, int col) { return (uint32_t)m[row * 16 + col] | ((uint32_t)m[row * 16 + col + 1] << 16); };
    a1h[0] = pair16(s_hi, g, 4 * q);       a1l[0] = pair16(s_lo, g, 4 * q);
    a1h[1] = pair16(s_hi, g + 8, 4 * q);   a1l[1] = pair16(s_lo, g + 8, 4 * q);
    a1h[2] = pair16(s_hi, g, 4 * q + 2);   a1l[2] = pair16(s_lo, g, 4 * q + 2);
    a1h[3] = pair16(s_hi, g + 8, 4 * q + 2); a1l[3] = pair16(s_lo, g + 8, 4 * q + 2);
    a2h[0] = pair16(s_hi, g, 2 * q);       a2l[0] = pair16(s_lo, g, 2 * q);
    a2h[1] = pair16(s_hi, g + 8, 2 * q);   a2l[1] = pair16(s_lo, g + 8, 2 * q);
    a2h[2] = pair16(s_hi, g, 2 * q + 8);   a2l[2] = pair16(s_lo, g, 2 * q + 8);
    a2h[3] = pair16(s_hi, g + 8, 2 * q + 8); a2l[3] = pair16(s_lo, g + 8, 2 * q + 8);
    // weights of my 8 coefficients: C[g][8 h2 + 2q + e] and C[g + 8][8 h2 + 2q + e]
    float wt[2][4];
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
        wt[h2][0] = s_w[g * 16 + 8 * h2 + 2 * q];
        wt[h2][1] = s_w[g * 16 + 8 * h2 + 2 * q + 1];
        wt[h2][2] = s_w[(g + 8) * 16 + 8 * h2 + 2 * q];
        wt[h2][3] = s_w[(g + 8) * 16 + 8 * h2 + 2 * q + 1];
    }
    // my pixels inside a box: rows g and g + 8 of block `warp` (16-byte chunk warp ^ (row & 7) under the swizzle), columns 4q .. 4q+3
    const uint32_t off0 = (uint32_t)(g * 128 + ((warp ^ g) << 4) + 4 * q), off1 = off0 + 1024u;
    const uint32_t bias = p.magic16;                               // 0x64006400: bytes become fp16 1024 + b under PRMT
    const __half2 off = __floats2half2_rn(1152.f, 1152.f);          // 1024 (PRMT bias) + 128 (centering)

    // One frame: coefficients of my 8 positions into `cur`, partial sums of w |C| and w |C - prev| over them.
    auto frame = [&](const int f, const float (&prev)[2][4], float (&cur)[2][4], float& sacc, float& dacc) {
        const int s = f % kRing;
        tma::mbar_wait(bar_full + 8u * s, (uint32_t)(f / kRing) & 1u);
        uint32_t px[2];
        px[0] = *reinterpret_cast<const uint32_t*>(&s_ring[s][off0]);
        px[1] = *reinterpret_cast<const uint32_t*>(&s_ring[s][off1]);
        // bytes -> centred fp16 pairs: (b0, b1) and (b2, b3) of each word
        uint32_t xb[2][2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t e0, e1;
            asm("prmt.b32 %0, %1, %2, 0x7150;" : "=r"(e0) : "r"(px[h]), "r"(bias));
            asm("prmt.b32 %0, %1, %2, 0x7352;" : "=r"(e1) : "r"(px[h]), "r"(bias));
            const __half2 h0 = __hsub2(*reinterpret_cast<const __half2*>(&e0), off);
            const __half2 h1 = __hsub2(*reinterpret_cast<const __half2*>(&e1), off);
            xb[h][0] = *reinterpret_cast<const uint32_t*>(&h0);
            xb[h][1] = *reinterpret_cast<const uint32_t*>(&h1);
        }
        __syncwarp();                                               // every lane has its pixels in registers
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_empty + 8u * s) : "memory");
        // step 1: T = D X^T, n-tile h = rows 8h .. 8h+7 of the block
        float tt[2][4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int i = 0; i < 4; ++i) tt[h][i] = 0.f;
            hmma_16816(tt[h], a1h, xb[h][0], xb[h][1]);
            hmma_16816(tt[h], a1l, xb[h][0], xb[h][1]);
        }
        // step 2: C = D T^T, n-tile h2 = coefficient columns 8 h2 .. 8 h2 + 7
        sacc = 0.f;
        dacc = 0.f;
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            uint32_t b0h, b0l, b1h, b1l;
            split_pair(tt[0][2 * h2], tt[0][2 * h2 + 1], b0h, b0l);
            split_pair(tt[1][2 * h2], tt[1][2 * h2 + 1], b1h, b1l);
#pragma unroll
            for (int i = 0; i < 4; ++i) cur[h2][i] = 0.f;
            hmma_16816(cur[h2], a2h, b0h, b1h);
            hmma_16816(cur[h2], a2h, b0l, b1l);
            hmma_16816(cur[h2], a2l, b0h, b1h);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                sacc = fmaf(fabsf(cur[h2][i]), wt[h2][i], sacc);
                dacc = fmaf(fabsf(cur[h2][i] - prev[h2][i]), wt[h2][i], dacc);
            }
        }
    };

    // Frames are taken two at a time: the coefficient sets ping-pong between two register arrays, and the four partial
    // sums (SC, TC of both frames) are reduced over the warp TOGETHER -- a transposing butterfly, 6 shuffles instead
    // of 20: after it the lanes with (lane & 24) == 0 / 8 / 16 / 24 hold SC(f), TC(f), SC(f+1), TC(f+1).
    float ca[2][4], cb[2][4];
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
        for (int i = 0; i < 4; ++i) cb[h2][i] = 0.f;
    float vmin = __int_as_float(0x7f800000), vmax = 0.f;            // of the value kind this lane ends up holding
    const int kind = (lane >> 3) & 3;                                // 0: SC(f)  1: TC(f)  2: SC(f+1)  3: TC(f+1)
    const bool is_tc = kind & 1;
    const int64_t out_step = (int64_t)p.By * p.Bx;
    float* out = (is_tc ? p.tc : p.sc) + ((int64_t)(t_start + (kind >> 1)) * p.By + by) * p.Bx + bx;
    const bool writer = valid && (lane & 7) == 0;

    for (int f = 0; f < n_iter; f += 2) {
        float s0, d0, s1 = 0.f, d1 = 0.f;
        frame(f, cb, ca, s0, d0);
        if (f + 1 < n_iter) frame(f + 1, ca, cb, s1, d1);           // warp-uniform
        // level 1 (xor 16): lanes 0..15 keep frame f, lanes 16..31 frame f + 1
        const bool up = lane & 16;
        float ks = up ? s1 : s0, kd = up ? d1 : d0;
        ks += __shfl_xor_sync(0xffffffffu, up ? s0 : s1, 16);
        kd += __shfl_xor_sync(0xffffffffu, up ? d0 : d1, 16);
        // level 2 (xor 8): lanes with bit 3 clear keep SC, the others TC
        const bool tcl = lane & 8;
        float v = tcl ? kd : ks;
        v += __shfl_xor_sync(0xffffffffu, tcl ? ks : kd, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        const int t = t_start + f + (kind >> 1);
        if (writer && t >= t0 && t < t1) {
            const float val = (is_tc && t == 0 && p.halo == nullptr) ? 0.f : v * p.inv_area;
            *out = val;
            if (t >= p.mm_begin && t < p.mm_end) {
                vmin = fminf(vmin, val);
                vmax = fmaxf(vmax, val);
            }
        }
        out += 2 * out_step;
    }
    if (p.mm != nullptr && writer && vmin <= vmax) {
        atomicMin(p.mm + (is_tc ? 2 : 0), __float_as_uint(vmin));   // non-negative floats order like their bit patterns
        atomicMax(p.mm + (is_tc ? 3 : 1), __float_as_uint(vmax));
    }
}

}  // namespace

// dct_size == block_size == 16; plane, strides and halo 16-byte aligned.  ELVIS_ERR_UNSUPPORTED when the driver cannot
// encode the tensor maps (the caller then uses the CUDA-core kernel of score_dctn.cu).  p.n_chunks / p.chunk_len are set by the caller.
int launch_score_dct16(ScoreParams p, int plane_h, int plane_w, cudaStream_t st) {
    CUtensorMap tm_clip, tm_halo;
    memset(&tm_clip, 0, sizeof(tm_clip));
    memset(&tm_halo, 0, sizeof(tm_halo));
    if (!tma::make_plane_map(&tm_clip, p.y, plane_w, plane_h, p.T, p.row_stride, p.frame_stride, 128, 16, CU_TENSOR_MAP_SWIZZLE_128B))
        return ELVIS_ERR_UNSUPPORTED;
    if (!tma::make_plane_map(&tm_halo, p.halo ? p.halo : p.y, plane_w, plane_h, 1, p.row_stride, p.frame_stride, 128, 16, CU_TENSOR_MAP_SWIZZLE_128B))
        return ELVIS_ERR_UNSUPPORTED;
    p.tiles_x = (p.Bx + kWorkers - 1) / kWorkers;
    p.tiles_y = p.By;
    p.magic16 = 0x64006400u;
    const int64_t grid = (int64_t)p.n_chunks * p.By * p.tiles_x;
    if (grid > 0x7fffffffLL) return ELVIS_ERR_UNSUPPORTED;
    score_dct16_kernel<<<(unsigned)grid, kThreads, 0, st>>>(tm_clip, tm_halo, host_tables(), p);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

}  // namespace elvis
