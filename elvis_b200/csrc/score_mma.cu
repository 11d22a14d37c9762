// a1 on the LEGACY tensor-core path (mma.sync): the first tensor-core scoring kernel, for 16x16
// blocks.  Correct, but slower than both other kernels on B200 (744 us vs 508 us CUDA cores vs
// 353 us tcgen05, score_umma.cu) -- kept behind ELVIS_SCORE_IMPL=mma|tma as a measured data point.
//
// Why: ncu on the CUDA-core kernel (score.cu; profiles/r1a_*) shows it issue-bound, not
// HBM-bound -- 16 instructions per pixel, DRAM at 20 % -- which is the condition under which
// north_star allows tensor cores for the DCT.  Here both 1-D passes of every 8x8 transform are
// warp-level MMAs, cutting the CUDA-core work to loads, u8->fp16 conversion, one precision
// split and the weighted-magnitude accumulation (~9 instructions per pixel).
//
// Data flow per CTA (8 consumer warps + 1 producer warp)
//   * the CTA owns a 128-pixel x 48-row luma tile (8 x 3 blocks) and walks it through a chunk
//     of consecutive frames;
//   * the producer thread streams the tile of every frame into a 4-stage shared-memory ring
//     with TMA (cp.async.bulk.tensor, 128-byte swizzle, zero fill outside the plane),
//     full/empty mbarriers per stage -- no register staging, loads run 3 frames ahead;
//   * consumer warp w owns pixel columns [16w, 16w+16): per 8-row group it reads ONE 32-bit
//     word per lane (conflict-free thanks to the swizzle), i.e. 4 pixels = its slice of the
//     B fragment of an MMA whose N axis is the pixel row and whose K axis is the 16 pixels
//     of two side-by-side tiles.
//
// Arithmetic per 8x16 pixel "pair" (tiles L and R), frame difference d = Y_t - Y_{t-1}:
//   pass 1  Rt = blockdiag(D, D) . d^T      HMMA m16n8k16, fp16 in / fp32 out.  d is an exact
//           small integer in fp16; D = hi + lo (two fp16 terms, |residual| < 3e-8) -> 2 MMAs.
//   pass 2  Out^T = Rt . D^T                HMMA m16n8k8 tf32.  The pass-1 accumulator layout
//           IS the pass-2 A layout (k slots permuted), so nothing moves between lanes; Rt is
//           split into tf32 hi + lo and D likewise -> 3 MMAs (hi*hi, hi*lo, lo*hi).
//   acc += Out (running C_t);  SC += w|acc|,  TC += w|Out|   (same temporal streaming and
//   accuracy argument as score.cu: the transform input is the exact integer difference).
// Relative error of SC/TC against the float64 spec is ~1e-6 (tests: 1e-4).
#include "score_params.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

namespace elvis {
namespace {

#include "dct8_tables.inc"

constexpr int kTileW = 128;                 // pixels (= bytes) per tile row: one swizzle span
constexpr int kTileBlocksY = 3;
constexpr int kTileH = 16 * kTileBlocksY;   // 48 rows
constexpr int kRowGroups = kTileH / 8;      // 6 MMA row groups
constexpr int kStages = 4;
constexpr int kStageBytes = kTileW * kTileH;   // 6144, a multiple of 1024 (swizzle atom alignment)
constexpr int kConsumerWarps = 8;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*alignment slack*/ + 2 * kStages * 8;

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
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int t, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(t), "r"(bar)
        : "memory");
}

__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <bool TMA>
__global__ void __launch_bounds__((kConsumerWarps + (TMA ? 1 : 0)) * 32, 2)
score_mma_kernel(const __grid_constant__ CUtensorMap tm_clip, const __grid_constant__ CUtensorMap tm_halo, const ScoreParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t ring = smem_u32(smem);
    const uint32_t full_bar = ring + kStages * kStageBytes;
    const uint32_t empty_bar = full_bar + kStages * 8;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_chunk = p.tiles_x * p.tiles_y;
    const int chunk = blockIdx.x / per_chunk;
    const int rem = blockIdx.x - chunk * per_chunk;
    const int ty = rem / p.tiles_x;
    const int tx = rem - ty * p.tiles_x;

    const int t0 = chunk * p.chunk_len;
    const int t1 = min(p.T, t0 + p.chunk_len);
    const bool has_prev = (t0 > 0) || (p.halo != nullptr);
    const int t_start = has_prev ? t0 - 1 : t0;   // one priming frame: acc = DCT(Y_{t0-1})
    const int n_iter = t1 - t_start;

    if (TMA) {
        if (threadIdx.x == 0) {
            for (int s = 0; s < kStages; ++s) {
                mbar_init(full_bar + 8 * s, 1);
                mbar_init(empty_bar + 8 * s, kConsumerWarps);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (warp == kConsumerWarps) {   // producer warp: one thread feeds the ring
            if (lane == 0) {
                for (int it = 0; it < n_iter; ++it) {
                    const int s = it % kStages;
                    const uint32_t ph = (it / kStages) & 1;
                    mbar_wait(empty_bar + 8 * s, ph ^ 1);
                    mbar_arrive_expect_tx(full_bar + 8 * s, kStageBytes);
                    const int t = t_start + it;
                    if (t < 0)
                        tma_load_3d(ring + s * kStageBytes, &tm_halo, tx * kTileW, ty * kTileH, 0, full_bar + 8 * s);
                    else
                        tma_load_3d(ring + s * kStageBytes, &tm_clip, tx * kTileW, ty * kTileH, t, full_bar + 8 * s);
                }
            }
            return;
        }
    }

    // ---------------- consumers ----------------
    const int g = lane >> 2, q = lane & 3;
    // pass-1 A operand: blockdiag(D, D) with the K axis permuted to the order in which a lane's
    // four bytes arrive (k = 2q, 2q+1, 2q+8, 2q+9 <-> bytes 4q .. 4q+3 of the 16-pixel row)
    uint32_t a_hi[4], a_lo[4];
    {
        const int xb = 4 * (q & 1);
        const bool left = q < 2;
        auto pack = [&](const uint16_t* tab, int x) { return (uint32_t)tab[g * 8 + x] | ((uint32_t)tab[g * 8 + x + 1] << 16); };
        const uint32_t hA = pack(c_d_f16_hi, xb), hB = pack(c_d_f16_hi, xb + 2);
        const uint32_t lA = pack(c_d_f16_lo, xb), lB = pack(c_d_f16_lo, xb + 2);
        a_hi[0] = left ? hA : 0u;  a_hi[1] = left ? 0u : hA;  a_hi[2] = left ? hB : 0u;  a_hi[3] = left ? 0u : hB;
        a_lo[0] = left ? lA : 0u;  a_lo[1] = left ? 0u : lA;  a_lo[2] = left ? lB : 0u;  a_lo[3] = left ? 0u : lB;
    }
    // pass-2 B operand: D[u = g][r], k slot q <-> r = 2q, k slot q+4 <-> r = 2q+1
    const uint32_t b_hi0 = c_d_tf32_hi[g * 8 + 2 * q], b_hi1 = c_d_tf32_hi[g * 8 + 2 * q + 1];
    const uint32_t b_lo0 = c_d_tf32_lo[g * 8 + 2 * q], b_lo1 = c_d_tf32_lo[g * 8 + 2 * q + 1];
    // this lane's outputs are Out[u = 2q, 2q+1][v = g] of both tiles
    const float w0 = c_score_w[(2 * q) * 8 + g] * p.inv_area, w1 = c_score_w[(2 * q + 1) * 8 + g] * p.inv_area;

    float acc[kRowGroups][4];
    uint32_t prev[kRowGroups][2];
#pragma unroll
    for (int rg = 0; rg < kRowGroups; ++rg) {
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[rg][i] = 0.f;
        prev[rg][0] = prev[rg][1] = 0x64006400u;   // half2(1024, 1024) == pixel value 0
    }

    const int bxi = tx * (kTileW / 16) + warp;
    const bool col_ok = bxi < p.Bx;
    const uint32_t lds_off = (uint32_t)(g * kTileW + ((warp ^ g) << 4) + 4 * q);   // swizzled 16-byte chunk
    const int64_t g_off = (int64_t)(ty * kTileH + g) * p.row_stride + (int64_t)tx * kTileW + warp * 16 + 4 * q;
    float smin = __int_as_float(0x7f800000), smax = 0.f, tmin = __int_as_float(0x7f800000), tmax = 0.f;

    for (int it = 0; it < n_iter; ++it) {
        const int t = t_start + it;
        uint32_t word[kRowGroups];
        if (TMA) {
            const int s = it % kStages;
            mbar_wait(full_bar + 8 * s, (it / kStages) & 1);
            const uint32_t base = ring + s * kStageBytes + lds_off;
#pragma unroll
            for (int rg = 0; rg < kRowGroups; ++rg)
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(word[rg]) : "r"(base + rg * 8 * kTileW));
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_bar + 8 * s);
        } else {
            const uint8_t* fp = (t < 0 ? p.halo : p.y + (int64_t)t * p.frame_stride) + g_off;
#pragma unroll
            for (int rg = 0; rg < kRowGroups; ++rg) {
                const bool ok = col_ok && (ty * kTileH + rg * 8) < p.By * 16;
                word[rg] = ok ? __ldg(reinterpret_cast<const uint32_t*>(fp + (int64_t)rg * 8 * p.row_stride)) : 0u;
            }
        }

        float v[8];   // {sc block 0..2, 0, tc block 0..2, 0}
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
#pragma unroll
        for (int rg = 0; rg < kRowGroups; ++rg) {
            // u8 -> fp16 without arithmetic: bytes {b, 0x64} are the half 1024 + b
            const uint32_t h0 = __byte_perm(word[rg], 0x64646464u, 0x4140);
            const uint32_t h1 = __byte_perm(word[rg], 0x64646464u, 0x4342);
            __half2 d0 = __hsub2(*reinterpret_cast<const __half2*>(&h0), *reinterpret_cast<const __half2*>(&prev[rg][0]));
            __half2 d1 = __hsub2(*reinterpret_cast<const __half2*>(&h1), *reinterpret_cast<const __half2*>(&prev[rg][1]));
            prev[rg][0] = h0;
            prev[rg][1] = h1;
            float r[4] = {0.f, 0.f, 0.f, 0.f};
            mma_f16(r, a_hi, *reinterpret_cast<uint32_t*>(&d0), *reinterpret_cast<uint32_t*>(&d1));
            mma_f16(r, a_lo, *reinterpret_cast<uint32_t*>(&d0), *reinterpret_cast<uint32_t*>(&d1));
            // tf32 split of the pass-1 result; registers (r0, r2, r1, r3) are the pass-2 A fragment
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                hi[i] = __float_as_uint(r[i]) & 0xffffe000u;
                lo[i] = __float_as_uint(r[i] - __uint_as_float(hi[i]));
            }
            float o[4] = {0.f, 0.f, 0.f, 0.f};
            mma_tf32(o, hi[0], hi[2], hi[1], hi[3], b_hi0, b_hi1);
            mma_tf32(o, hi[0], hi[2], hi[1], hi[3], b_lo0, b_lo1);
            mma_tf32(o, lo[0], lo[2], lo[1], lo[3], b_hi0, b_hi1);
            float s = 0.f, d = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[rg][i] += o[i];
                const float w = (i & 1) ? w1 : w0;
                s = fmaf(fabsf(acc[rg][i]), w, s);
                d = fmaf(fabsf(o[i]), w, d);
            }
            v[rg >> 1] += s;
            v[4 + (rg >> 1)] += d;
        }

        if (t < t0) continue;   // priming frame: only the accumulators matter (warp-uniform)

        // transposing warp reduction: 8 values x 32 lanes -> value (lane >> 2) in every lane, fixed order
        float w4[4], w2[2], w1v;
        {
            const bool b = lane & 16;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float send = b ? v[i] : v[i + 4], keep = b ? v[i + 4] : v[i];
                w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
        }
        {
            const bool b = lane & 8;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float send = b ? w4[i] : w4[i + 2], keep = b ? w4[i + 2] : w4[i];
                w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
        }
        {
            const bool b = lane & 4;
            const float send = b ? w2[0] : w2[1], keep = b ? w2[1] : w2[0];
            w1v = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        w1v += __shfl_xor_sync(0xffffffffu, w1v, 2);
        w1v += __shfl_xor_sync(0xffffffffu, w1v, 1);
        // lane holds: kind = lane bit 4 (0 sc, 1 tc), block j = (lane >> 2) & 3 (3 = padding)
        const int j = (lane >> 2) & 3;
        const int byi = ty * kTileBlocksY + j;
        if ((lane & 3) == 0 && j < kTileBlocksY && col_ok && byi < p.By) {
            const int64_t o = ((int64_t)t * p.By + byi) * p.Bx + bxi;
            const bool in_mm = t >= p.mm_begin && t < p.mm_end;
            if (lane & 16) {
                const float tcv = (t == 0 && p.halo == nullptr) ? 0.f : w1v;
                p.tc[o] = tcv;
                if (in_mm) {
                    tmin = fminf(tmin, tcv);
                    tmax = fmaxf(tmax, tcv);
                }
            } else {
                p.sc[o] = w1v;
                if (in_mm) {
                    smin = fminf(smin, w1v);
                    smax = fmaxf(smax, w1v);
                }
            }
        }
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

// (W, H, T) uint8 tensor, box = 128 x 48 x 1, 128-byte swizzle, zero fill out of bounds
bool make_map(CUtensorMap* m, const uint8_t* base, int W, int H, int T, int64_t row_stride, int64_t frame_stride) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T};
    cuuint64_t strides[2] = {(cuuint64_t)row_stride, (cuuint64_t)(T > 1 ? frame_stride : row_stride * H)};
    cuuint32_t box[3] = {kTileW, kTileH, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

int launch_score_mma(ScoreParams p, int plane_h, int plane_w, bool use_tma, cudaStream_t st) {
    p.tiles_x = (p.Bx + kTileW / 16 - 1) / (kTileW / 16);
    p.tiles_y = (p.By + kTileBlocksY - 1) / kTileBlocksY;
    const unsigned grid = (unsigned)((int64_t)p.tiles_x * p.tiles_y * p.n_chunks);
    CUtensorMap tm_clip, tm_halo;
    memset(&tm_clip, 0, sizeof(tm_clip));
    memset(&tm_halo, 0, sizeof(tm_halo));
    if (use_tma) {
        bool ok = make_map(&tm_clip, p.y, plane_w, plane_h, p.T, p.row_stride, p.frame_stride);
        ok = ok && make_map(&tm_halo, p.halo ? p.halo : p.y, plane_w, plane_h, 1, p.row_stride, p.frame_stride);
        if (!ok) use_tma = false;   // driver without tensor-map support: same arithmetic, direct loads
    }
    if (use_tma) {
        static PerDeviceOnce configured;   // the attribute is per (kernel, device)
        const cudaError_t e = configured.run([] {
            return cudaFuncSetAttribute(score_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        });
        if (e != cudaSuccess) return cuda_fail(e);
        score_mma_kernel<true><<<grid, (kConsumerWarps + 1) * 32, kSmemBytes, st>>>(tm_clip, tm_halo, p);
    } else {
        score_mma_kernel<false><<<grid, kConsumerWarps * 32, 0, st>>>(tm_clip, tm_halo, p);
    }
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

}  // namespace elvis
