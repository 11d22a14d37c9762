// a8, a10, a12 (downsample) and the LANCZOS4 restorer of 8f rank 1: cv2 INTER_AREA down + INTER_LINEAR (or LANCZOS4) up
// of the isolated block (oracle/spec_cv.py:resize_area / resize_linear / resize_lanczos4).  Integer paths are bit-exact
// by construction; every float op on the fractional-area path is an explicitly rounded fp32 op in cv2's accumulation order.
#include "degrade_common.cuh"
#include "down_pow2.cuh"
#include "tma.cuh"
#include <cstddef>
#include <cstring>

namespace elvis {
namespace {

// ------------------------------------------------------------------------- downsample
// Table blob layout (int32 words), one entry of `level_stride(pb)` words per level:
//   [0] small  [1] area_kind (0 copy, 1 2x2, 2 integer factor, 3 fractional)  [2] factor
//   [3] float bits of float32(1/factor^2)  [4] n_area  [5..7] reserved
//   [8 .. 8+pb]            area_start[0..pb]  (entries of destination index d: [start[d], start[d+1]))
//   then 2*pb entries x {src_index, float-bits alpha}
//   then horizontal linear taps i0[pb] i1[pb] c0[pb] c1[pb], then vertical ones likewise.
// With LANCZOS the two bilinear tap blocks are replaced by ONE 8-tap table (cv2 INTER_LANCZOS4 for
// u8: idx[pb][8] clamped source indices, coef[pb][8] 11-bit integer weights, used for both axes;
// result = (sum + 2^21) >> 22, saturated) -- oracle/spec_cv.py:lanczos4_taps / resize_lanczos4.
__host__ __device__ inline int level_stride(int pb, bool lanczos = false) { return 8 + (pb + 1) + 2 * pb * 2 + (lanczos ? 16 : 8) * pb; }

template <bool LANCZOS>
__global__ void __launch_bounds__(256) downsample_kernel(const BlockGeom g, const int32_t* __restrict__ levels,
                                                         const int32_t* __restrict__ tables, int n_levels, int warps_per_cta,
                                                         const uint32_t only_mask = 0u) {   // non-zero: only blocks whose level bit is set
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
        if (only_mask && !((only_mask >> lv) & 1u)) continue;      // the closed-form kernel handles this block (warp-uniform)
        const int32_t* tab = tables + (size_t)lv * level_stride(pb, LANCZOS);
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
        const int32_t* lh = tab + 8 + (pb + 1) + 4 * pb;
        if (LANCZOS) {
            const int32_t* idx = lh;
            const int32_t* coef = lh + 8 * pb;
            for (int i = lane; i < small * pb; i += 32) {        // horizontal 8-tap pass into Bi[small][pb]
                const int y = i / pb, d = i - y * pb;
                int acc = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) acc += S[y * small + idx[d * 8 + k]] * coef[d * 8 + k];
                Bi[i] = acc;
            }
            __syncwarp();
            for (int i = lane; i < n; i += 32) {                 // vertical 8-tap pass, (sum + 2^21) >> 22
                const int d2 = i / pb, d = i - d2 * pb;
                int acc = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) acc += Bi[idx[d2 * 8 + k] * pb + d] * coef[d2 * 8 + k];
                int v = (acc + (1 << 21)) >> 22;
                v = v < 0 ? 0 : (v > 255 ? 255 : v);
                dp[(int64_t)d2 * g.dst_row + d * g.C] = (uint8_t)v;
            }
        } else {
            // bilinear back up: horizontal pass into Bi[small][pb] (11-bit coefficients)
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
        }
        __syncwarp();
    }
}

// ------------------------------------------- downsample, table driven, planar 16 / 8 blocks
// The arithmetic of downsample_kernel<false> above (same tables, same operation order: bit-identical) for single-channel
// planes with aligned rows and 16- or 8-pixel blocks -- the shape in which fractional reductions actually occur (utils'
// 16 -> 5, presley's bs // 3 and bs // 5): the block comes in and goes out as 8- / 4-byte row pieces through shared
// memory instead of byte by byte, the block size is a compile-time constant, and the (row, column) pairs of the
// reduced image advance incrementally instead of being divided out of a linear index.
template <int PB>
__global__ void __launch_bounds__(256) downsample_planar_kernel(const BlockGeom g, const int32_t* __restrict__ levels,
                                                                const int32_t* __restrict__ tables, int n_levels, const uint32_t only_mask) {
    constexpr int kWarps = 8, N = PB * PB;
    __shared__ __align__(16) int32_t s_b[kWarps][N];          // horizontal passes: float (area) / int (bilinear)
    __shared__ __align__(16) uint8_t s_a[kWarps][N];          // the block, then the result
    __shared__ __align__(16) uint8_t s_s[kWarps][N];          // the reduced image
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int32_t* Bi = s_b[w];
    float* Bf = reinterpret_cast<float*>(Bi);
    uint8_t* A = s_a[w];
    uint8_t* S = s_s[w];
    // row piece of this lane: 16-pixel blocks: 8 bytes of row lane / 2; 8-pixel blocks: 4 bytes of row lane / 2 (lanes 0..15)
    constexpr int kPiece = PB / 2;
    const bool mover = PB == 16 || lane < 16;
    const int mrow = lane >> 1, mcol = (lane & 1) * kPiece;

    const int64_t n_blocks = (int64_t)g.T * g.By * g.Bx;
    for (int64_t b = (int64_t)blockIdx.x * kWarps + w; b < n_blocks; b += (int64_t)gridDim.x * kWarps) {
        int lv = levels[b];
        lv = lv < 0 ? 0 : (lv >= n_levels ? n_levels - 1 : lv);
        if (only_mask && !((only_mask >> lv) & 1u)) continue;      // the closed-form kernel handles this block (warp-uniform)
        const int bx = (int)(b % g.Bx);
        const int64_t q = b / g.Bx;
        const int by = (int)(q % g.By), t = (int)(q / g.By);
        const int32_t* tab = tables + (size_t)lv * level_stride(PB);
        const int small = tab[0], kind = tab[1];
        const uint8_t* sp = g.src + (int64_t)t * g.src_frame + ((int64_t)by * PB + mrow) * g.src_row + (int64_t)bx * PB + mcol;
        uint8_t* dp = g.dst + (int64_t)t * g.dst_frame + ((int64_t)by * PB + mrow) * g.dst_row + (int64_t)bx * PB + mcol;
        if (mover) {
            if (PB == 16) *reinterpret_cast<uint2*>(A + mrow * PB + mcol) = __ldcs(reinterpret_cast<const uint2*>(sp));
            else *reinterpret_cast<uint32_t*>(A + mrow * PB + mcol) = __ldcs(reinterpret_cast<const uint32_t*>(sp));
        }
        __syncwarp();
        if (kind != 0 && small < PB) {
            const int ns = small * small;
            const int q32 = 32 / small, r32 = 32 - q32 * small;         // how (row, column) of the reduced image advance per 32 lanes
            if (kind == 1) {
                int dy = lane / small, dx = lane - dy * small;
                for (int i = lane; i < ns; i += 32) {
                    const uint8_t* p = A + (2 * dy) * PB + 2 * dx;
                    S[i] = (uint8_t)((p[0] + p[1] + p[PB] + p[PB + 1] + 2) >> 2);
                    dx += r32; dy += q32;
                    if (dx >= small) { dx -= small; ++dy; }
                }
            } else if (kind == 2) {
                const int f = tab[2];
                const float scale = __int_as_float(tab[3]);
                int dy = lane / small, dx = lane - dy * small;
                for (int i = lane; i < ns; i += 32) {
                    int sum = 0;
                    for (int yy = 0; yy < f; ++yy)
                        for (int xx = 0; xx < f; ++xx) sum += A[(dy * f + yy) * PB + dx * f + xx];
                    S[i] = (uint8_t)__float2int_rn(__fmul_rn((float)sum, scale));
                    dx += r32; dy += q32;
                    if (dx >= small) { dx -= small; ++dy; }
                }
            } else {
                const int32_t* start = tab + 8;
                const int32_t* ent = tab + 8 + (PB + 1);
                int y = lane / small, dx = lane - y * small;
                for (int i = lane; i < PB * small; i += 32) {       // rows: (y, dx)
                    float acc = 0.f;
                    for (int e = start[dx]; e < start[dx + 1]; ++e)
                        acc = __fadd_rn(acc, __fmul_rn((float)A[y * PB + ent[2 * e]], __int_as_float(ent[2 * e + 1])));
                    Bf[i] = acc;                                    // i == y * small + dx
                    dx += r32; y += q32;
                    if (dx >= small) { dx -= small; ++y; }
                }
                __syncwarp();
                int dy = lane / small;
                dx = lane - dy * small;
                for (int i = lane; i < ns; i += 32) {               // columns: (dy, dx)
                    float acc = 0.f;
                    for (int e = start[dy]; e < start[dy + 1]; ++e)
                        acc = __fadd_rn(acc, __fmul_rn(Bf[ent[2 * e] * small + dx], __int_as_float(ent[2 * e + 1])));
                    const int v = __float2int_rn(acc);
                    S[i] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
                    dx += r32; dy += q32;
                    if (dx >= small) { dx -= small; ++dy; }
                }
            }
            __syncwarp();
            // bilinear back up: horizontal pass into Bi[small][PB] (11-bit coefficients), then the vertical one into A
            const int32_t* lh = tab + 8 + (PB + 1) + 4 * PB;
            const int32_t* lvt = lh + 4 * PB;
            for (int i = lane; i < small * PB; i += 32) {
                const int y = i / PB, d = i % PB;
                Bi[i] = S[y * small + lh[d]] * lh[2 * PB + d] + S[y * small + lh[PB + d]] * lh[3 * PB + d];
            }
            __syncwarp();
            for (int i = lane; i < N; i += 32) {
                const int d2 = i / PB, d = i % PB;
                const int r0 = Bi[lvt[d2] * PB + d] >> 4, r1 = Bi[lvt[PB + d2] * PB + d] >> 4;
                int v = (((lvt[2 * PB + d2] * r0) >> 16) + ((lvt[3 * PB + d2] * r1) >> 16) + 2) >> 2;
                A[i] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
            }
            __syncwarp();
        }
        if (mover) {
            if (PB == 16) __stcs(reinterpret_cast<uint2*>(dp), *reinterpret_cast<const uint2*>(A + mrow * PB + mcol));
            else __stcs(reinterpret_cast<uint32_t*>(dp), *reinterpret_cast<const uint32_t*>(A + mrow * PB + mcol));
        }
        __syncwarp();
    }
}

// ----------------------------------------------------------------- downsample, fast path
// Planar planes, 16- or 8-pixel blocks, power-of-two reductions (the elvis 1x/2x/4x/8x(/16x)
// pyramid).  Same lane geometry as blur_fast_kernel (8 pixels of one row per lane).
//   area     per-lane byte sums (packed 16-bit adds) + xor-shuffle sums over the f rows;
//            f = 2: (s + 2) >> 2, f >= 4: round-half-even(s / f^2)  (cv2's two integer paths)
//   linear   every 11-bit horizontal coefficient of these ratios is a multiple of 64, so the two
//            taps of an output pixel are byte weights over the <= 8-byte source row and
//            S[i0]*a0 + S[i1]*a1 == 64 * (dp4a(row.lo, wx) + dp4a(row.hi, wy)); the vertical
//            pass is cv2's  ((b0*(R0>>4))>>16) + ((b1*(R1>>4))>>16) + 2 >> 2.
// Weight vectors come from elvis_b200/_tables.py (fast part of the blob).
template <int PB, bool ALIGNED>
__global__ void __launch_bounds__(256) downsample_fast_kernel(const BlockGeom g, const int32_t* __restrict__ levels,
                                                              const int32_t* __restrict__ tables, int n_levels) {
    constexpr int kWarps = 8;
    constexpr int kBlocks = PB == 16 ? 1 : 4;
    constexpr int kLevelStride = 8 + (PB + 1) + 4 * PB + 8 * PB;
    constexpr int kFastStride = 6 * PB;
    __shared__ __align__(8) uint8_t s_small[kWarps][kBlocks][8 * 8];   // reduced image, 8-byte row pitch
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int blk = PB == 16 ? 0 : lane >> 3;
    const int r = PB == 16 ? lane >> 1 : lane & 7;
    const int h = PB == 16 ? lane & 1 : 0;
    uint8_t* S = s_small[w][blk];
    const int32_t* fast_base = tables + (((size_t)n_levels * kLevelStride + 3) & ~(size_t)3);   // 16-byte aligned

    const int64_t n_blocks = (int64_t)g.T * g.By * g.Bx;
    const int64_t stride = (int64_t)gridDim.x * kWarps * kBlocks;
    for (int64_t b0 = ((int64_t)blockIdx.x * kWarps + w) * kBlocks; b0 < n_blocks; b0 += stride) {
        const int64_t b = b0 + blk;
        const bool live = b < n_blocks;
        int L = 0;            // log2 of the reduction factor; 0 = copy
        int lv = 0;
        const uint8_t* sp = g.src;
        uint8_t* dp = g.dst;
        if (live) {
            const int bx = (int)(b % g.Bx);
            const int64_t q = b / g.Bx;
            const int by = (int)(q % g.By), t = (int)(q / g.By);
            lv = levels[b];
            lv = lv < 0 ? 0 : (lv >= n_levels ? n_levels - 1 : lv);
            const int small = __ldg(tables + (size_t)lv * kLevelStride);
            L = small >= PB ? 0 : 31 - __clz(PB / small);
            sp += (int64_t)t * g.src_frame + ((int64_t)by * PB + r) * g.src_row + (int64_t)bx * PB + 8 * h;
            dp += (int64_t)t * g.dst_frame + ((int64_t)by * PB + r) * g.dst_row + (int64_t)bx * PB + 8 * h;
        }
        uint2 px = make_uint2(0u, 0u);
        if (live) {
            if (ALIGNED) {
                px = __ldcs(reinterpret_cast<const uint2*>(sp));
            } else {
                px.x = sp[0] | (sp[1] << 8) | (sp[2] << 16) | ((uint32_t)sp[3] << 24);
                px.y = sp[4] | (sp[5] << 8) | (sp[6] << 16) | ((uint32_t)sp[7] << 24);
            }
        }
        const unsigned any = __ballot_sync(0xffffffffu, L > 0);
        if (any) {
            // ---- area: horizontal sums inside the lane
            const uint32_t e0 = (px.x & 0x00ff00ffu) + ((px.x >> 8) & 0x00ff00ffu);   // (b0+b1, b2+b3)
            const uint32_t e1 = (px.y & 0x00ff00ffu) + ((px.y >> 8) & 0x00ff00ffu);   // (b4+b5, b6+b7)
            int hs[4];
            if (L == 1) {
                hs[0] = e0 & 0xffff; hs[1] = e0 >> 16; hs[2] = e1 & 0xffff; hs[3] = e1 >> 16;
            } else if (L == 2) {
                hs[0] = (e0 & 0xffff) + (e0 >> 16); hs[1] = (e1 & 0xffff) + (e1 >> 16); hs[2] = hs[3] = 0;
            } else {
                hs[0] = (e0 & 0xffff) + (e0 >> 16) + (e1 & 0xffff) + (e1 >> 16); hs[1] = hs[2] = hs[3] = 0;
            }
            if (PB == 16) {        // factor 16: the two halves of the row
                const int o = __shfl_xor_sync(0xffffffffu, hs[0], 1);
                if (L == 4) hs[0] += o;
            }
            // ---- vertical sums over the f rows of the cell (row bits of the lane index)
            constexpr int kRowBit = PB == 16 ? 2 : 1;
#pragma unroll
            for (int sft = 0; sft < (PB == 16 ? 4 : 3); ++sft) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int o = __shfl_xor_sync(0xffffffffu, hs[i], kRowBit << sft);
                    if (sft < L) hs[i] += o;
                }
            }
            // ---- rounding and store of the reduced image (one writer per cell)
            if (L > 0 && (r & ((1 << L) - 1)) == 0 && !(L == 4 && h == 1)) {
                const int k = 2 * L;
                const int cnt = L >= 3 ? 1 : (8 >> L);                 // cells of this lane in the row
                const int col0 = L >= 3 ? (L == 4 ? 0 : h) : (8 * h) >> L;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (i < cnt) {
                        int v;
                        if (L == 1) {
                            v = (hs[i] + 2) >> 2;
                        } else {
                            v = hs[i] >> k;
                            const int rem = hs[i] & ((1 << k) - 1), half = 1 << (k - 1);
                            v += (rem > half) || (rem == half && (v & 1));
                        }
                        S[(r >> L) * 8 + col0 + i] = (uint8_t)v;
                    }
                }
            }
            __syncwarp();
            // ---- bilinear back up
            if (L > 0) {
                const int32_t* ft = fast_base + (size_t)lv * kFastStride;
                const int4 vt = __ldg(reinterpret_cast<const int4*>(ft + 2 * PB) + r);           // i0, i1, b0, b1
                const uint2 r0 = *reinterpret_cast<const uint2*>(S + vt.x * 8);
                const uint2 r1 = *reinterpret_cast<const uint2*>(S + vt.y * 8);
                const int4* wv = reinterpret_cast<const int4*>(ft) + 4 * h;                       // {wx,wy} x 8 pixels
                uint32_t o[8];
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const int4 wq = __ldg(wv + j4);                                               // pixels 2*j4, 2*j4+1
                    const uint32_t wxy[4] = {(uint32_t)wq.x, (uint32_t)wq.y, (uint32_t)wq.z, (uint32_t)wq.w};
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const uint32_t d0 = __dp4a(r0.y, wxy[2 * e + 1], __dp4a(r0.x, wxy[2 * e], 0u));   // R0 / 64
                        const uint32_t d1 = __dp4a(r1.y, wxy[2 * e + 1], __dp4a(r1.x, wxy[2 * e], 0u));   // R1 / 64
                        const int v = (int)((((uint32_t)vt.z * (d0 * 4u)) >> 16) + (((uint32_t)vt.w * (d1 * 4u)) >> 16) + 2u) >> 2;
                        o[2 * j4 + e] = (uint32_t)(v > 255 ? 255 : v);
                    }
                }
                px.x = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
                px.y = o[4] | (o[5] << 8) | (o[6] << 16) | (o[7] << 24);
            }
            __syncwarp();
        }
        if (live) {
            if (ALIGNED) {
                __stcs(reinterpret_cast<uint2*>(dp), px);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dp[j] = (uint8_t)(px.x >> (8 * j));
                    dp[4 + j] = (uint8_t)(px.y >> (8 * j));
                }
            }
        }
    }
}

// ------------------------------------------------- downsample, power-of-two closed form
// The default for planar planes with 16- / 8-pixel blocks and power-of-two reductions (down_pow2.cuh):
// one warp = one 16 x 16 block or two 8 x 8 blocks, the level is uniform per lane group, so the
// per-level code is straight-line packed 16-bit integer arithmetic with a handful of shuffles
// (about 11 instructions per pixel against 30 of the table-driven kernel above).
template <int PB, bool ALIGNED>
__global__ void __launch_bounds__(256) downsample_pow2_kernel(const BlockGeom g, const int32_t* __restrict__ levels,
                                                              const int32_t* __restrict__ tables, int n_levels,
                                                              const uint32_t skip_mask = 0u) {   // levels left to the generic kernel
    constexpr int kWarps = 8;
    constexpr int kGroup = 2 * PB;
    constexpr int kBlocks = 32 / kGroup;
    constexpr int kLevelStride = 8 + (PB + 1) + 4 * PB + 8 * PB;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int gl = lane % kGroup, base = lane - gl, blk = lane / kGroup;
    const int r = gl >> 1, h = gl & 1;
    constexpr int kBytes = PB / 2;                          // bytes of a row this lane owns
    const int64_t n_blocks = (int64_t)g.T * g.By * g.Bx;
    const int64_t stride = (int64_t)gridDim.x * kWarps * kBlocks;
    for (int64_t b0 = ((int64_t)blockIdx.x * kWarps + w) * kBlocks; b0 < n_blocks; b0 += stride) {
        const int64_t b = b0 + blk;
        bool live = b < n_blocks;
        int L = 0;
        const uint8_t* sp = g.src;
        uint8_t* dp = g.dst;
        if (live) {
            int lv = levels[b];
            lv = lv < 0 ? 0 : (lv >= n_levels ? n_levels - 1 : lv);
            live = !((skip_mask >> lv) & 1u);
            const int small = __ldg(tables + (size_t)lv * kLevelStride);
            L = (!live || small >= PB) ? 0 : 31 - __clz(PB / small);
        }
        if (live) {
            const int bx = (int)(b % g.Bx);
            const int64_t q = b / g.Bx;
            const int by = (int)(q % g.By), t = (int)(q / g.By);
            sp += (int64_t)t * g.src_frame + ((int64_t)by * PB + r) * g.src_row + (int64_t)bx * PB + kBytes * h;
            dp += (int64_t)t * g.dst_frame + ((int64_t)by * PB + r) * g.dst_row + (int64_t)bx * PB + kBytes * h;
        }
        uint32_t p0 = 0u, p1 = 0u;
        if (live) {
            if (ALIGNED) {
                if (PB == 16) {
                    const uint2 v = __ldcs(reinterpret_cast<const uint2*>(sp));
                    p0 = v.x;
                    p1 = v.y;
                } else {
                    p0 = __ldcs(reinterpret_cast<const uint32_t*>(sp));
                }
            } else {
                p0 = sp[0] | (sp[1] << 8) | (sp[2] << 16) | ((uint32_t)sp[3] << 24);
                if (PB == 16) p1 = sp[4] | (sp[5] << 8) | (sp[6] << 16) | ((uint32_t)sp[7] << 24);
            }
        }
        // every lane group runs the code of every level present in the warp (shuffles are warp-wide) and keeps its own
        const int La = __shfl_sync(0xffffffffu, L, 0);
        const int Lb = kBlocks == 2 ? __shfl_sync(0xffffffffu, L, 16) : La;
        if (La > 0) {
            uint32_t q0 = p0, q1 = p1;
            down_up_pow2_level<PB>(q0, q1, La, gl, base);
            if (L == La) {
                p0 = q0;
                p1 = q1;
            }
        }
        if (Lb > 0 && Lb != La) {
            uint32_t q0 = p0, q1 = p1;
            down_up_pow2_level<PB>(q0, q1, Lb, gl, base);
            if (L == Lb) {
                p0 = q0;
                p1 = q1;
            }
        }
        if (live) {
            if (ALIGNED) {
                if (PB == 16) __stcs(reinterpret_cast<uint2*>(dp), make_uint2(p0, p1));
                else __stcs(reinterpret_cast<uint32_t*>(dp), p0);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dp[j] = (uint8_t)(p0 >> (8 * j));
                    if (PB == 16) dp[4 + j] = (uint8_t)(p1 >> (8 * j));
                }
            }
        }
    }
}

// Planar YUV 4:2:0 with 16 x 16 luma blocks, Y, U and V in ONE launch: a warp takes the luma block and then
// its two 8 x 8 chroma blocks (lanes 0..15 U, 16..31 V), which share the block's level -- no divergence.
struct YuvGeom {
    const uint8_t* src[3];
    uint8_t* dst[3];
    int64_t src_frame[3], src_row[3], dst_frame[3], dst_row[3];
    int32_t T, By, Bx;
};

// Shared-memory tile of eight horizontally adjacent blocks: 16 luma rows x 128 bytes, 8 + 8 chroma rows x 64 bytes and
// the eight levels.  Row pitches are padded (144 / 72 bytes) so that the warp-per-block reads below -- 8 bytes of luma
// per lane, rows 16 bytes apart in the tile; 4 bytes of chroma -- are free of bank conflicts, V sits 16 banks after U.
struct __align__(16) DownTile {
    uint8_t y[16][144];
    uint8_t u[8][72];
    uint8_t v[8][72];
    int32_t lv[8];
};

// WHY the staging: a warp that reads its own block straight from global memory touches 16 different 128-byte lines
// with every load or store instruction (16 rows x 16 bytes), and L1 looks its tags up one line at a time -- ncu showed
// the register-prefetch version of this kernel (146 instructions per block, issue 41 %) waiting on exactly that
// (37 % of all stall samples on the first use of a prefetched value; 11 sectors per request).  Here the CTA moves
// whole tiles with coalesced 16-byte cp.async copies (one per thread, three tiles ahead) and coalesced 16-byte
// stores (4 lines per warp instruction), and the warps talk to shared memory only.
__global__ void __launch_bounds__(256) downsample_pow2_yuv420_kernel(const YuvGeom g, const int32_t* __restrict__ levels, int max_level) {
    constexpr int kStages = 4;
    __shared__ DownTile s_in[kStages];
    __shared__ DownTile s_out[2];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int t = blockIdx.x / g.By, by = blockIdx.x - t * g.By;
    const int n_groups = (g.Bx + 7) / 8;
    const int32_t* lv = levels + (int64_t)blockIdx.x * g.Bx;

    // ---- mover role of this thread: one 16-byte luma piece (threads 0..127) or one 8-byte chroma piece (128..255) of a tile
    const bool mv_luma = tid < 128;
    const int mk = mv_luma ? tid : (tid - 128) & 63;
    const int mrow = mk >> 3, mseg = mk & 7;                  // row inside the tile, block inside the group
    const int mpl = mv_luma ? 0 : 1 + ((tid - 128) >> 6);     // plane
    const int mbytes = mv_luma ? 16 : 8;
    const int64_t mrow_px = mv_luma ? (int64_t)by * 16 + mrow : (int64_t)by * 8 + mrow;
    const uint8_t* msrc = g.src[mpl] + (int64_t)t * g.src_frame[mpl] + mrow_px * g.src_row[mpl] + mseg * mbytes;
    uint8_t* mdst = g.dst[mpl] + (int64_t)t * g.dst_frame[mpl] + mrow_px * g.dst_row[mpl] + mseg * mbytes;
    // byte offset of this thread's piece inside a tile, shared-memory addresses of the rings, running global pointers
    const uint32_t slot_off = mv_luma ? (uint32_t)(mrow * 144 + 16 * mseg)
                                      : (uint32_t)((mpl == 1 ? offsetof(DownTile, u) : offsetof(DownTile, v)) + mrow * 72 + 8 * mseg);
    const uint32_t in_base = (uint32_t)__cvta_generic_to_shared(&s_in[0]);
    const uint8_t* const out_base = reinterpret_cast<const uint8_t*>(&s_out[0]);
    const int mstep = mv_luma ? 128 : 64;
    const int m_last = g.Bx - mseg;                           // this thread moves a piece of group grp iff 8 grp < m_last
    const uint8_t* src_next = msrc;                           // piece of the next group to be requested
    int grp_next = 0;
    auto issue = [&]() {
        if (grp_next < n_groups) {
            const uint32_t tile = in_base + (uint32_t)(grp_next % kStages) * (uint32_t)sizeof(DownTile);
            if (8 * grp_next < m_last) {
                if (mv_luma) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tile + slot_off), "l"(src_next) : "memory");
                else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(tile + slot_off), "l"(src_next) : "memory");
            }
            if (tid < 8 && grp_next * 8 + tid < g.Bx)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tile + (uint32_t)offsetof(DownTile, lv) + 4u * tid),
                             "l"(lv + grp_next * 8 + tid) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");      // always: keeps the group count uniform
        src_next += mstep;
        ++grp_next;
    };

    // ---- worker role: warp w owns block 8 grp + w; luma lane = (row, 8-pixel half), chroma lanes 0..15 U / 16..31 V
    const int yr = lane >> 1, yh = lane & 1;
    const int cpl = lane >> 4, gl = lane & 15, cr = gl >> 1, ch = gl & 1;

#pragma unroll
    for (int sgi = 0; sgi < kStages - 1; ++sgi) issue();
    for (int grp = 0; grp < n_groups; ++grp) {
        issue();
        asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 1) : "memory");
        __syncthreads();                                          // tile grp is complete for every warp
        const DownTile& in = s_in[grp % kStages];
        DownTile& out = s_out[grp & 1];
        const int bx = grp * 8 + w;
        if (bx < g.Bx) {                                          // warp-uniform
            const uint2 y = *reinterpret_cast<const uint2*>(&in.y[yr][16 * w + 8 * yh]);
            uint32_t c0 = *reinterpret_cast<const uint32_t*>(cpl ? &in.v[cr][8 * w + 4 * ch] : &in.u[cr][8 * w + 4 * ch]);
            int L = in.lv[w];
            L = L < 0 ? 0 : (L > max_level ? max_level : L);
            uint32_t p0 = y.x, p1 = y.y, c1 = 0u;
            if (L > 0) {
                down_up_pow2_level<16>(p0, p1, L > 4 ? 4 : L, lane, 0);
                down_up_pow2_level<8>(c0, c1, L > 3 ? 3 : L, gl, lane & 16);
            }
            *reinterpret_cast<uint2*>(&out.y[yr][16 * w + 8 * yh]) = make_uint2(p0, p1);
            *reinterpret_cast<uint32_t*>(cpl ? &out.v[cr][8 * w + 4 * ch] : &out.u[cr][8 * w + 4 * ch]) = c0;
        }
        __syncthreads();                                          // the output tile is complete; s_in[grp % kStages] is free again
        if (8 * grp < m_last) {
            const uint8_t* from = out_base + (grp & 1) * sizeof(DownTile) + slot_off;
            if (mv_luma) __stcs(reinterpret_cast<uint4*>(mdst), *reinterpret_cast<const uint4*>(from));
            else __stcs(reinterpret_cast<uint2*>(mdst), *reinterpret_cast<const uint2*>(from));
        }
        mdst += mstep;
        // s_out[grp & 1] is rewritten two iterations later, after two more barriers
    }
}

// The same kernel with the tile traffic handed to the TMA unit (default when the planes can be described by tensor
// maps: 16-byte aligned bases and strides).  One thread issues three box loads per tile of 8 blocks -- 16 rows x 128 bytes
// of luma with the 128-byte swizzle, 8 rows x 64 bytes of U and of V with the 64-byte swizzle -- onto an mbarrier of a
// 4-deep ring, and three box stores of the finished tile; nobody else executes a mover instruction (the cp.async version
// above spends 70 of its 216 instructions per block on moving).  The swizzles replace the padded pitches: warp w reads 16-byte
// chunk w of every luma row, which the hardware has placed at chunk w ^ (row & 7), so the sixteen rows of a block
// spread over all banks; chroma likewise with chunk (w >> 1) ^ ((row >> 1) & 3).  Partial tiles at the right edge need no code:
// loads zero-fill and stores clip at the tensor bounds.
struct DownMaps {
    CUtensorMap in[3], out[3];
};

__global__ void __launch_bounds__(256) downsample_pow2_yuv420_tma_kernel(const __grid_constant__ DownMaps m, const int By, const int Bx,
                                                                         const int32_t* __restrict__ levels, const int max_level) {
    constexpr int kStages = 4;
    constexpr uint32_t kTile = 3072, kOffU = 2048, kOffV = 2560;     // luma 16 x 128, U 8 x 64, V 8 x 64
    __shared__ __align__(1024) uint8_t s_in[kStages * kTile];
    __shared__ __align__(1024) uint8_t s_out[2 * kTile];
    __shared__ __align__(8) uint64_t s_full[kStages];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int t = blockIdx.x / By, by = blockIdx.x - t * By;
    const int n_groups = (Bx + 7) / 8;
    const int32_t* lv = levels + (int64_t)blockIdx.x * Bx;
    const uint32_t in_base = tma::smem_u32(s_in), out_base = tma::smem_u32(s_out), bar = tma::smem_u32(s_full);

    auto issue_load = [&](int grp) {                                  // thread 0 only
        const uint32_t dst = in_base + (uint32_t)(grp % kStages) * kTile, b = bar + 8u * (uint32_t)(grp % kStages);
        tma::mbar_arrive_expect_tx(b, kTile);
        tma::load_3d(dst, &m.in[0], grp * 128, by * 16, t, b);
        tma::load_3d(dst + kOffU, &m.in[1], grp * 64, by * 8, t, b);
        tma::load_3d(dst + kOffV, &m.in[2], grp * 64, by * 8, t, b);
    };
    if (tid == 0) {
#pragma unroll
        for (int sgi = 0; sgi < kStages; ++sgi) tma::mbar_init(bar + 8u * sgi, 1);
        tma::mbar_init_fence();
        for (int grp = 0; grp < kStages && grp < n_groups; ++grp) issue_load(grp);
    }
    __syncthreads();                                                  // the barriers are initialised for everyone

    // worker role: warp w owns block 8 grp + w; luma lane = (row, 8-pixel half), chroma lanes 0..15 U / 16..31 V
    const int yr = lane >> 1, yh = lane & 1;
    const int cpl = lane >> 4, gl = lane & 15, cr = gl >> 1, ch = gl & 1;
    const uint32_t y_off = (uint32_t)(yr * 128 + (((w ^ (yr & 7)) << 4) | (yh << 3)));
    const uint32_t c_off = (cpl ? kOffV : kOffU) + (uint32_t)(cr * 64 + ((((w >> 1) ^ ((cr >> 1) & 3)) << 4) | ((w & 1) << 3) | (ch << 2)));
    int L_next = w < Bx ? lv[w] : 0;
    for (int grp = 0; grp < n_groups; ++grp) {
        const int slot = grp % kStages;
        const int bx = grp * 8 + w;
        int L = L_next;
        L_next = bx + 8 < Bx ? lv[bx + 8] : 0;
        tma::mbar_wait(bar + 8u * slot, (uint32_t)(grp / kStages) & 1u);
        const uint8_t* in = s_in + slot * kTile;
        uint8_t* out = s_out + (grp & 1) * kTile;
        if (bx < Bx) {                                                // warp-uniform
            const uint2 y = *reinterpret_cast<const uint2*>(in + y_off);
            uint32_t c0 = *reinterpret_cast<const uint32_t*>(in + c_off);
            L = L < 0 ? 0 : (L > max_level ? max_level : L);
            uint32_t p0 = y.x, p1 = y.y, c1 = 0u;
            if (L > 0) {
                down_up_pow2_level<16>(p0, p1, L > 4 ? 4 : L, lane, 0);
                down_up_pow2_level<8>(c0, c1, L > 3 ? 3 : L, gl, lane & 16);
            }
            *reinterpret_cast<uint2*>(out + y_off) = make_uint2(p0, p1);
            *reinterpret_cast<uint32_t*>(out + c_off) = c0;
        }
        tma::fence_proxy_async();                                     // my tile writes, before the TMA store reads them
        if (tid == 0) tma::store_wait_read<0>();                      // the store of tile grp - 1 has left s_out[(grp + 1) & 1]
        __syncthreads();                                              // out tile complete; in slot read by everyone; other out tile free
        if (tid == 0) {
            const uint32_t src = out_base + (uint32_t)(grp & 1) * kTile;
            tma::store_3d(&m.out[0], grp * 128, by * 16, t, src);
            tma::store_3d(&m.out[1], grp * 64, by * 8, t, src + kOffU);
            tma::store_3d(&m.out[2], grp * 64, by * 8, t, src + kOffV);
            tma::store_commit();
            if (grp + kStages < n_groups) issue_load(grp + kStages);
        }
    }
    if (tid == 0) tma::store_wait<0>();                               // shared memory must outlive the last store's reads
}

// Experimental variant (ELVIS_DOWNSAMPLE_TMA=2): the same tiles, but one pipeline PER WARP as in the blur kernel -- a warp
// takes a whole tile of eight blocks (three box loads onto its own mbarrier, eight blocks in place, three box stores, three
// tile buffers) and no CTA-wide barrier couples warps whose blocks need different amounts of work.
__global__ void __launch_bounds__(128) downsample_pow2_yuv420_warp_kernel(const __grid_constant__ DownMaps m, const int T, const int By, const int Bx,
                                                                          const int32_t* __restrict__ levels, const int max_level) {
    constexpr int kWarps = 4, kBufs = 3;
    constexpr uint32_t kTile = 3072, kOffU = 2048, kOffV = 2560;
    __shared__ __align__(1024) uint8_t s_buf[kWarps][kBufs][kTile];
    __shared__ __align__(8) uint64_t s_full[kWarps][kBufs];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int groups = (Bx + 7) / 8;
    const int64_t n_strips = (int64_t)T * By * groups;
    const int64_t stride = (int64_t)gridDim.x * kWarps, first = (int64_t)blockIdx.x * kWarps + w;
    const uint32_t buf_base = tma::smem_u32(&s_buf[w][0][0]), bar = tma::smem_u32(&s_full[w][0]);
    auto issue = [&](int64_t s, int b) {                              // lane 0 only
        const int gx = (int)(s % groups);
        const int64_t row = s / groups;                              // t * By + by
        const int by = (int)(row % By), t = (int)(row / By);
        const uint32_t dst = buf_base + (uint32_t)b * kTile, bb = bar + 8u * b;
        tma::mbar_arrive_expect_tx(bb, kTile);
        tma::load_3d(dst, &m.in[0], gx * 128, by * 16, t, bb);
        tma::load_3d(dst + kOffU, &m.in[1], gx * 64, by * 8, t, bb);
        tma::load_3d(dst + kOffV, &m.in[2], gx * 64, by * 8, t, bb);
    };
    auto level_of = [&](int64_t s) -> int {                          // lane j < 8: level of block j of strip s
        if (s >= n_strips || lane >= 8) return 0;
        const int gx = (int)(s % groups);
        const int bx = gx * 8 + lane;
        return bx < Bx ? levels[(s / groups) * Bx + bx] : 0;
    };
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < kBufs; ++b) tma::mbar_init(bar + 8u * b, 1);
        tma::mbar_init_fence();
        if (first < n_strips) issue(first, 0);
        if (first + stride < n_strips) issue(first + stride, 1);
    }
    __syncwarp();
    const int yr = lane >> 1, yh = lane & 1;
    const int cpl = lane >> 4, gl = lane & 15, cr = gl >> 1, ch = gl & 1;
    const uint32_t y_row = (uint32_t)(yr * 128 + (yh << 3)), y_x = (uint32_t)(yr & 7);
    const uint32_t c_row = (cpl ? kOffV : kOffU) + (uint32_t)(cr * 64 + (ch << 2)), c_x = (uint32_t)((cr >> 1) & 3);
    int lv_next = level_of(first);
    int it = 0;
    for (int64_t s = first; s < n_strips; s += stride, ++it) {
        const int b = it % kBufs;
        const int lv_mine = lv_next;
        lv_next = level_of(s + stride);
        tma::mbar_wait(bar + 8u * b, (uint32_t)(it / kBufs) & 1u);
        uint8_t* buf = &s_buf[w][b][0];
        const int gx = (int)(s % groups);
        const int n_here = min(8, Bx - gx * 8);
#pragma unroll 1
        for (int j = 0; j < n_here; ++j) {
            int L = __shfl_sync(0xffffffffu, lv_mine, j);
            L = L < 0 ? 0 : (L > max_level ? max_level : L);
            if (L > 0) {                                              // warp-uniform
                uint2* py = reinterpret_cast<uint2*>(buf + y_row + (((uint32_t)j ^ y_x) << 4));
                uint32_t* pc = reinterpret_cast<uint32_t*>(buf + c_row + ((((uint32_t)j >> 1) ^ c_x) << 4) + (((uint32_t)j & 1u) << 3));
                const uint2 y = *py;
                uint32_t p0 = y.x, p1 = y.y, c0 = *pc, c1 = 0u;
                down_up_pow2_level<16>(p0, p1, L > 4 ? 4 : L, lane, 0);
                down_up_pow2_level<8>(c0, c1, L > 3 ? 3 : L, gl, lane & 16);
                *py = make_uint2(p0, p1);
                *pc = c0;
            }
        }
        tma::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            const int64_t row = s / groups;
            const int by = (int)(row % By), t = (int)(row / By);
            const uint32_t src = buf_base + (uint32_t)b * kTile;
            tma::store_3d(&m.out[0], gx * 128, by * 16, t, src);
            tma::store_3d(&m.out[1], gx * 64, by * 8, t, src + kOffU);
            tma::store_3d(&m.out[2], gx * 64, by * 8, t, src + kOffV);
            tma::store_commit();
            if (s + 2 * stride < n_strips) {
                tma::store_wait_read<1>();
                issue(s + 2 * stride, (it + 2) % kBufs);
            }
        }
    }
    if (lane == 0) tma::store_wait<0>();
}

}  // namespace
}  // namespace elvis

using namespace elvis;

extern "C" int elvis_degrade_downsample(const elvis_plane* src, const elvis_plane* dst, int32_t n_frames,
                                        int32_t block_px, int32_t by, int32_t bx, const int32_t* levels,
                                        const int32_t* tables, int32_t n_levels, int32_t fast_tables_ok,
                                        elvis_stream_t stream) {
    BlockGeom g;
    if (int rc = make_geom(src, dst, n_frames, block_px, by, bx, g)) return rc;
    if (!levels || !tables || n_levels <= 0) return ELVIS_ERR_INVALID_ARG;
    if (block_px > 64) return ELVIS_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    if (int rc = copy_edges(g, st)) return rc;
    const int n = block_px * block_px;
    // fast_tables_ok: 1 = every level is a power-of-two reduction; an even value > 1 = MIXED, bit l + 1 set for every level l
    // WITHOUT fast tables (e.g. utils' 16 -> 5): those blocks go through the generic kernel, all others through the closed form
    const uint32_t slow_mask = (fast_tables_ok > 1 && !(fast_tables_ok & 1)) ? ((uint32_t)fast_tables_ok >> 1) : 0u;
    if (g.C == 1 && (block_px == 16 || block_px == 8) && slow_mask && !getenv("ELVIS_DOWNSAMPLE_GENERIC") && !getenv("ELVIS_DOWNSAMPLE_TABLE")) {
        const bool al4 = aligned_to(g.src, block_px / 2) && aligned_to(g.dst, block_px / 2) && g.src_frame % (block_px / 2) == 0 &&
                         g.dst_frame % (block_px / 2) == 0 && g.src_row % (block_px / 2) == 0 && g.dst_row % (block_px / 2) == 0;
        const int64_t blocks = (int64_t)n_frames * by * bx;
        const int grid2 = grid_for_units(blocks, 8 * (block_px == 16 ? 1 : 2));
        if (block_px == 16) {
            if (al4) downsample_pow2_kernel<16, true><<<grid2, 256, 0, st>>>(g, levels, tables, n_levels, slow_mask);
            else downsample_pow2_kernel<16, false><<<grid2, 256, 0, st>>>(g, levels, tables, n_levels, slow_mask);
        } else {
            if (al4) downsample_pow2_kernel<8, true><<<grid2, 256, 0, st>>>(g, levels, tables, n_levels, slow_mask);
            else downsample_pow2_kernel<8, false><<<grid2, 256, 0, st>>>(g, levels, tables, n_levels, slow_mask);
        }
        ELVIS_CHECK_LAUNCH();
        if (al4) {
            if (block_px == 16) downsample_planar_kernel<16><<<grid_for_units(blocks, 8), 256, 0, st>>>(g, levels, tables, n_levels, slow_mask);
            else downsample_planar_kernel<8><<<grid_for_units(blocks, 8), 256, 0, st>>>(g, levels, tables, n_levels, slow_mask);
        } else {
            int wpc = 8;
            while (wpc > 1 && (size_t)wpc * n * 6 > 48 * 1024) wpc >>= 1;
            downsample_kernel<false><<<grid_for_units(blocks, wpc), wpc * 32, (size_t)wpc * n * 6, st>>>(g, levels, tables, n_levels, wpc, slow_mask);
        }
        ELVIS_CHECK_LAUNCH();
        return ELVIS_OK;
    }
    if (g.C == 1 && (block_px == 16 || block_px == 8) && fast_tables_ok == 1 && !getenv("ELVIS_DOWNSAMPLE_GENERIC")) {
        const bool al = aligned_to(g.src, 8) && aligned_to(g.dst, 8) && g.src_frame % 8 == 0 && g.dst_frame % 8 == 0 &&
                        g.src_row % 8 == 0 && g.dst_row % 8 == 0;
        const int64_t blocks = (int64_t)n_frames * by * bx;
        if (!getenv("ELVIS_DOWNSAMPLE_TABLE")) {     // closed-form kernel (default); the table-driven one stays selectable
            const bool al4 = block_px == 16 ? al : (aligned_to(g.src, 4) && aligned_to(g.dst, 4) && g.src_frame % 4 == 0 &&
                                                    g.dst_frame % 4 == 0 && g.src_row % 4 == 0 && g.dst_row % 4 == 0);
            const int grid2 = grid_for_units(blocks, 8 * (block_px == 16 ? 1 : 2));
            if (block_px == 16) {
                if (al4) downsample_pow2_kernel<16, true><<<grid2, 256, 0, st>>>(g, levels, tables, n_levels);
                else downsample_pow2_kernel<16, false><<<grid2, 256, 0, st>>>(g, levels, tables, n_levels);
            } else {
                if (al4) downsample_pow2_kernel<8, true><<<grid2, 256, 0, st>>>(g, levels, tables, n_levels);
                else downsample_pow2_kernel<8, false><<<grid2, 256, 0, st>>>(g, levels, tables, n_levels);
            }
            ELVIS_CHECK_LAUNCH();
            return ELVIS_OK;
        }
        const int grid = grid_for_units(blocks, 8 * (block_px == 16 ? 1 : 4));
        if (block_px == 16) {
            if (al) downsample_fast_kernel<16, true><<<grid, 256, 0, st>>>(g, levels, tables, n_levels);
            else downsample_fast_kernel<16, false><<<grid, 256, 0, st>>>(g, levels, tables, n_levels);
        } else {
            if (al) downsample_fast_kernel<8, true><<<grid, 256, 0, st>>>(g, levels, tables, n_levels);
            else downsample_fast_kernel<8, false><<<grid, 256, 0, st>>>(g, levels, tables, n_levels);
        }
        ELVIS_CHECK_LAUNCH();
        return ELVIS_OK;
    }
    if (g.C == 1 && (block_px == 16 || block_px == 8) && !getenv("ELVIS_DOWNSAMPLE_GENERIC")) {
        const int a = block_px / 2;
        if (aligned_to(g.src, a) && aligned_to(g.dst, a) && g.src_frame % a == 0 && g.dst_frame % a == 0 && g.src_row % a == 0 && g.dst_row % a == 0) {
            const int64_t blocks = (int64_t)n_frames * by * bx;
            if (block_px == 16) downsample_planar_kernel<16><<<grid_for_units(blocks, 8), 256, 0, st>>>(g, levels, tables, n_levels, 0u);
            else downsample_planar_kernel<8><<<grid_for_units(blocks, 8), 256, 0, st>>>(g, levels, tables, n_levels, 0u);
            ELVIS_CHECK_LAUNCH();
            return ELVIS_OK;
        }
    }
    int wpc = 8;
    while (wpc > 1 && (size_t)wpc * n * 6 > 48 * 1024) wpc >>= 1;
    const size_t smem = (size_t)wpc * n * 6;
    const int64_t units = (int64_t)n_frames * by * bx * g.C;
    downsample_kernel<false><<<grid_for_units(units, wpc), wpc * 32, smem, st>>>(g, levels, tables, n_levels, wpc);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_restore_lanczos(const elvis_plane* src, const elvis_plane* dst, int32_t n_frames,
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
    const int64_t units = (int64_t)n_frames * by * bx * g.C;
    downsample_kernel<true><<<grid_for_units(units, wpc), wpc * 32, (size_t)wpc * n * 6, st>>>(g, levels, tables, n_levels, wpc);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

// Fused planar 4:2:0 form of the power-of-two downsample: level l reduces the 16 x 16 luma block by
// 2^min(l, max_level, 4) per axis and its two 8 x 8 chroma blocks by 2^min(l, max_level, 3).
extern "C" int elvis_degrade_downsample_pow2_yuv420(const elvis_plane* src_yuv, const elvis_plane* dst_yuv, int32_t n_frames,
                                                    int32_t block_size, int32_t by, int32_t bx, const int32_t* levels,
                                                    int32_t max_level, elvis_stream_t stream) {
    if (!src_yuv || !dst_yuv || !levels || n_frames <= 0 || by <= 0 || bx <= 0 || max_level < 0) return ELVIS_ERR_INVALID_ARG;
    if (block_size != 16) return ELVIS_ERR_UNSUPPORTED;
    YuvGeom g;
    for (int i = 0; i < 3; ++i) {
        const elvis_plane *s = src_yuv + i, *d = dst_yuv + i;
        if (!plane_ok(s) || !plane_ok(d) || s->channels != 1 || d->channels != 1) return ELVIS_ERR_INVALID_ARG;
        const int pb = i == 0 ? 16 : 8;
        // whole blocks only (the per-plane entry point copies partial blocks through)
        if (s->height != by * pb || s->width != bx * pb || d->height != s->height || d->width != s->width) return ELVIS_ERR_UNSUPPORTED;
        const int a = i == 0 ? 16 : 8;          // the tile movers copy 16-byte luma and 8-byte chroma pieces
        if (!aligned_to(s->data, a) || !aligned_to(d->data, a) || s->frame_stride % a || d->frame_stride % a || s->row_stride % a ||
            d->row_stride % a)
            return ELVIS_ERR_UNSUPPORTED;
        g.src[i] = static_cast<const uint8_t*>(s->data);
        g.dst[i] = static_cast<uint8_t*>(d->data);
        g.src_frame[i] = s->frame_stride;
        g.src_row[i] = s->row_stride;
        g.dst_frame[i] = d->frame_stride;
        g.dst_row[i] = d->row_stride;
    }
    g.T = n_frames;
    g.By = by;
    g.Bx = bx;
    if ((int64_t)n_frames * by > 0x7fffffffLL) return ELVIS_ERR_UNSUPPORTED;
    const unsigned grid = (unsigned)((int64_t)n_frames * by);
    // TMA version when every plane can be described by a tensor map (ELVIS_DOWNSAMPLE_TMA=0: the cp.async movers)
    const char* use_tma = getenv("ELVIS_DOWNSAMPLE_TMA");
    if (!(use_tma && use_tma[0] == '0')) {
        DownMaps m;
        bool ok = true;
        for (int i = 0; i < 3 && ok; ++i) {
            const int pb = i == 0 ? 16 : 8, box_w = i == 0 ? 128 : 64;
            const CUtensorMapSwizzle sw = i == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
            ok = tma::make_plane_map(&m.in[i], g.src[i], bx * pb, by * pb, n_frames, g.src_row[i], g.src_frame[i], box_w, pb, sw) &&
                 tma::make_plane_map(&m.out[i], g.dst[i], bx * pb, by * pb, n_frames, g.dst_row[i], g.dst_frame[i], box_w, pb, sw);
        }
        if (ok && use_tma && use_tma[0] == '2') {
            const int64_t strips = (int64_t)n_frames * by * ((bx + 7) / 8);
            downsample_pow2_yuv420_warp_kernel<<<grid_for_units(strips, 4), 128, 0, as_stream(stream)>>>(m, n_frames, by, bx, levels, max_level);
            ELVIS_CHECK_LAUNCH();
            return ELVIS_OK;
        }
        if (ok) {
            downsample_pow2_yuv420_tma_kernel<<<grid, 256, 0, as_stream(stream)>>>(m, by, bx, levels, max_level);
            ELVIS_CHECK_LAUNCH();
            return ELVIS_OK;
        }
    }
    downsample_pow2_yuv420_kernel<<<grid, 256, 0, as_stream(stream)>>>(g, levels, max_level);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}
