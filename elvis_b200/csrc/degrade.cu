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
#include "dct8_packed.cuh"
#include "down_pow2.cuh"
#include "tma.cuh"
#include <cstddef>
#include <cstring>
#include <cuda_fp16.h>

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


// ----------------------------------------------------------------------- blur, fast path
// Planar planes with 16- or 8-pixel blocks (luma / 4:2:0 chroma of 16x16 blocks).  A group of
// G lanes owns one block (G = 32 for PB = 16, 8 for PB = 8 -> four blocks per warp); every
// lane produces 8 pixels per pass:
//   row pass     lane = (row, 8-pixel half): the row is read with one 128/64-bit LDS, the 12-byte
//                tap window (block-edge reflection folded into PRMT selectors) is walked with
//                two IDP.4A per pixel: (14,62,104,62).(x-2..x+1) + 14*x(+2).  The 16-bit results
//                are stored TRANSPOSED (column major, two halo rows per side holding the
//                reflected rows), so that
//   column pass  lane = (column, 8-row half): 12 vertically consecutive 16-bit values arrive as
//                six 32-bit pairs (LDS.128 + LDS.64) and each output is three IDP.2A
//                (pair . two 8-bit taps) on top of the rounding constant; >> 16 gives the u8.
// About 11 instructions per pixel and round, all integer, bit-exact with cv2 by construction.
template <int PB> struct BlurGeom;
template <> struct BlurGeom<16> { static constexpr int G = 32, kBlocks = 1, kPitch = 48; };
template <> struct BlurGeom<8>  { static constexpr int G = 8,  kBlocks = 4, kPitch = 32; };

template <int PB, bool ALIGNED>
__global__ void __launch_bounds__(256) blur_fast_kernel(const BlockGeom g, const int32_t* __restrict__ rounds) {
    using GG = BlurGeom<PB>;
    constexpr int kWarps = 8;
    constexpr int kABytes = GG::kBlocks * PB * PB;            // u8 blocks, row major
    constexpr int kTBytes = GG::kBlocks * PB * GG::kPitch;    // u16 row-pass results, column major + halo
    __shared__ __align__(16) uint8_t s_a[kWarps][kABytes];
    __shared__ __align__(16) uint8_t s_t[kWarps][kTBytes];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int blk = PB == 16 ? 0 : lane >> 3;
    const int r = PB == 16 ? lane >> 1 : lane & 7;            // row pass: my row ...
    const int h = PB == 16 ? lane & 1 : 0;                    // ... and 8-pixel half
    const int x = PB == 16 ? lane & 15 : lane & 7;            // column pass: my column ...
    const int yh = PB == 16 ? lane >> 4 : 0;                  // ... and 8-row half
    uint8_t* a = s_a[w] + blk * PB * PB;
    uint8_t* tm = s_t[w] + blk * PB * GG::kPitch;
    const uint32_t kTaps4 = 14u | (62u << 8) | (104u << 16) | (62u << 24);

    const int64_t n_blocks = (int64_t)g.T * g.By * g.Bx;
    const int64_t stride = (int64_t)gridDim.x * kWarps * GG::kBlocks;
    for (int64_t b0 = ((int64_t)blockIdx.x * kWarps + w) * GG::kBlocks; b0 < n_blocks; b0 += stride) {
        const int64_t b = b0 + blk;
        const bool live = b < n_blocks;
        int nr = 0;
        const uint8_t* sp = g.src;
        uint8_t* dp = g.dst;
        if (live) {
            const int bx = (int)(b % g.Bx);
            const int64_t q = b / g.Bx;
            const int by = (int)(q % g.By), t = (int)(q / g.By);
            nr = rounds[b];
            const int64_t off_s = (int64_t)t * g.src_frame + ((int64_t)by * PB + r) * g.src_row + (int64_t)bx * PB + 8 * h;
            const int64_t off_d = (int64_t)t * g.dst_frame + ((int64_t)by * PB + r) * g.dst_row + (int64_t)bx * PB + 8 * h;
            sp += off_s;
            dp += off_d;
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
        int max_r = nr;
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) max_r = max(max_r, __shfl_xor_sync(0xffffffffu, max_r, m));
        if (max_r > 0) {
            *reinterpret_cast<uint2*>(a + r * PB + 8 * h) = px;
            __syncwarp();
            for (int k = 0; k < max_r; ++k) {
                const bool act = k < nr;
                // ---- row pass
                if (act) {
                    uint32_t W0, W1, W2;
                    if (PB == 16) {
                        const uint4 row = *reinterpret_cast<const uint4*>(a + r * 16);
                        // h == 0: pixels -2..9 = (b2,b1,b0,b1 | b2..b5 | b6..b9); h == 1: pixels 6..17 = (b6..b9 | b10..b13 | b14,b15,b14,b13)
                        const uint32_t A0 = h ? row.y : row.x, B0 = h ? row.z : row.x;
                        const uint32_t A1 = h ? row.z : row.x, B1 = h ? row.w : row.y;
                        const uint32_t A2 = h ? row.w : row.y, B2 = h ? row.w : row.z;
                        W0 = __byte_perm(A0, B0, h ? 0x5432 : 0x1012);
                        W1 = __byte_perm(A1, B1, 0x5432);
                        W2 = __byte_perm(A2, B2, h ? 0x1232 : 0x5432);
                    } else {
                        const uint2 row = *reinterpret_cast<const uint2*>(a + r * 8);
                        W0 = __byte_perm(row.x, row.x, 0x1012);      // b2 b1 b0 b1
                        W1 = __byte_perm(row.x, row.y, 0x5432);      // b2 b3 b4 b5
                        W2 = __byte_perm(row.y, row.y, 0x1232);      // b6 b7 b6 b5
                    }
                    uint32_t o[8];
                    o[0] = __dp4a(W0, kTaps4, __dp4a(W1, 14u, 0u));
                    o[1] = __dp4a(__byte_perm(W0, W1, 0x4321), kTaps4, __dp4a(W1, 14u << 8, 0u));
                    o[2] = __dp4a(__byte_perm(W0, W1, 0x5432), kTaps4, __dp4a(W1, 14u << 16, 0u));
                    o[3] = __dp4a(__byte_perm(W0, W1, 0x6543), kTaps4, __dp4a(W1, 14u << 24, 0u));
                    o[4] = __dp4a(W1, kTaps4, __dp4a(W2, 14u, 0u));
                    o[5] = __dp4a(__byte_perm(W1, W2, 0x4321), kTaps4, __dp4a(W2, 14u << 8, 0u));
                    o[6] = __dp4a(__byte_perm(W1, W2, 0x5432), kTaps4, __dp4a(W2, 14u << 16, 0u));
                    o[7] = __dp4a(__byte_perm(W1, W2, 0x6543), kTaps4, __dp4a(W2, 14u << 24, 0u));
                    // transposed store: column c = 8h + j, stored row r + 2; reflected halo rows
                    uint16_t* tcol = reinterpret_cast<uint16_t*>(tm) + (8 * h) * (GG::kPitch / 2) + (r + 2);
                    int dup = -1;                                   // halo slot that mirrors my row
                    if (r == 1) dup = 1; else if (r == 2) dup = 0;
                    else if (r == PB - 2) dup = PB + 2; else if (r == PB - 3) dup = PB + 3;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        tcol[j * (GG::kPitch / 2)] = (uint16_t)o[j];
                        if (dup >= 0) tcol[j * (GG::kPitch / 2) + (dup - (r + 2))] = (uint16_t)o[j];
                    }
                }
                __syncwarp();
                // ---- column pass
                if (act) {
                    const uint8_t* col = tm + x * GG::kPitch + 16 * yh;     // stored rows 8yh .. 8yh+11
                    const uint4 q0 = *reinterpret_cast<const uint4*>(col);
                    const uint2 q1 = *reinterpret_cast<const uint2*>(col + 16);
                    const uint32_t P[6] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y};
                    uint8_t* acol = a + (8 * yh) * PB + x;
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                        const int p = i >> 1;
                        // even row: taps (14,62 | 104,62 | 14,-); odd row: (-,14 | 62,104 | 62,14)
                        uint32_t ve = __dp2a_lo(P[p], 14u | (62u << 8), 32768u);
                        ve = __dp2a_lo(P[p + 1], 104u | (62u << 8), ve);
                        ve = __dp2a_lo(P[p + 2], 14u, ve);
                        uint32_t vo = __dp2a_lo(P[p], 14u << 8, 32768u);
                        vo = __dp2a_lo(P[p + 1], 62u | (104u << 8), vo);
                        vo = __dp2a_lo(P[p + 2], 62u | (14u << 8), vo);
                        acol[i * PB] = (uint8_t)(ve >> 16);
                        acol[(i + 1) * PB] = (uint8_t)(vo >> 16);
                    }
                }
                __syncwarp();
            }
            px = *reinterpret_cast<const uint2*>(a + r * PB + 8 * h);
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

// ---------------------------------------------------------------- blur on the tensor cores
// One blur round of an isolated block is Z = round((G X G^T) / 2^16) with G the 5-tap operator along an
// axis (taps 14 62 104 62 14, reflect-101 folded into the edge rows) -- two 16 x 16 x 16 integer matrix
// products.  A warp owns one 16 x 16 tile (one luma block, or 2 x 2 blocks of 8 x 8 with a block-diagonal
// G) and keeps it in the operand layout of mma.sync.m16n8k16 (u8 x u8 -> s32) for all rounds:
//   * the tile is held "transposed for free": a matrix M in the accumulator (C) layout is, read as a
//     B operand, M^T with the K index permuted; the permutation is absorbed into which pixel columns a
//     thread owns, so thread (g, q) = (lane / 4, lane % 4) simply owns pixels 4q..4q+3 of two tile rows
//     (one 32-bit word each), loads them as the B operand and stores the result words as they come;
//   * step 1: M1 = G X^T (2 IMMA, one per n-tile); step 2: Z = G M1^T = G X G^T with the 16-bit M1
//     split into high and low bytes (2 + 2 IMMA, the high product shifted left by 8; the rounding
//     constant 2^15 enters as the initial accumulator 128 of the high product); byte 2 of every
//     accumulator is the blurred pixel.  A = G never changes: two registers per thread.
// 6 IMMA + ~30 integer instructions per thread and round for 8 pixels (the dp4a kernel above: ~90),
// all exact: bit-identical to cv2's fixed-point GaussianBlur (oracle/spec_cv.py, tools/emu/check_blur_imma.py).
__device__ __forceinline__ void imma_16816(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0, const int (&c)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.u8.u8.s32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%7, %8, %9, %10};"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
                 : "r"(a0), "r"(a1), "r"(b0), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]));
}

// entry (m, c) of the per-axis operator of a 16-wide tile made of PB-wide blocks
__device__ __forceinline__ uint32_t blur_operator_entry(int PB, int m, int c) {
    const int blk = (m / PB) * PB, ml = m - blk;
    uint32_t s = 0;
#pragma unroll
    for (int d = -2; d <= 2; ++d) {
        const uint32_t tap = d == 0 ? 104u : ((d == 1 || d == -1) ? 62u : 14u);
        if (blk + reflect101(ml + d, PB) == c) s += tap;
    }
    return s;
}

// tile-row owned by accumulator-layout row i (the same map orders the columns: 4 q + j <-> layout 2 q + j, 8 + 2 q + j - 2)
__device__ __forceinline__ int imma_tile_row(int i) { return 4 * ((i & 7) >> 1) + (i & 1) + (i >= 8 ? 2 : 0); }

// `nr` blur rounds (per thread: the rounds of the block its 8 pixels belong to) on the warp's tile, held as the words
// w0 / w1 of the thread's two rows; every lane of the warp must call it (the IMMAs are warp-wide).
template <int PB>
__device__ __forceinline__ void blur_imma_rounds(uint32_t& w0, uint32_t& w1, const uint32_t a0, const uint32_t a1, const int nr) {
    const int zero4[4] = {0, 0, 0, 0}, half4[4] = {128, 128, 128, 128};
    // rounds the warp has to run: the tile's own count (one block), or the maximum over its four blocks -- the block of a
    // lane is given by bit 4 (rows) and bit 1 (columns) of the lane index
    int max_r = nr;
    if (PB == 8) {
        max_r = max(max_r, __shfl_xor_sync(0xffffffffu, max_r, 16));
        max_r = max(max_r, __shfl_xor_sync(0xffffffffu, max_r, 2));
    }
    for (int k = 0; k < max_r; ++k) {
        int m1a[4], m1b[4];                               // M1 = G X^T: n-tile 0 (from my first row) and 1 (second row)
        imma_16816(m1a, a0, a1, w0, zero4);
        imma_16816(m1b, a0, a1, w1, zero4);
        uint32_t za[2], zb[2];                             // byte pairs of my first / second row from n-tile `half`
#pragma unroll
        for (int half = 0; half < 2; ++half) {            // layout rows g / g + 8 of M1 feed n-tile `half` of step 2
            const uint32_t q0 = (uint32_t)m1a[2 * half], q1 = (uint32_t)m1a[2 * half + 1];
            const uint32_t q2 = (uint32_t)m1b[2 * half], q3 = (uint32_t)m1b[2 * half + 1];
            // the four 16-bit values as two registers of halves, then their high / low bytes: 4 PRMT (6 when each is built apart)
            const uint32_t p01 = __byte_perm(q0, q1, 0x5410), p23 = __byte_perm(q2, q3, 0x5410);
            const uint32_t hi = __byte_perm(p01, p23, 0x7531);
            const uint32_t lo = __byte_perm(p01, p23, 0x6420);
            int acc[4], acl[4];                            // two independent products: shorter dependent chain per round
            imma_16816(acc, a0, a1, hi, half4);
            imma_16816(acl, a0, a1, lo, zero4);
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = acc[i] * 256 + acl[i];
            // byte 2 of every accumulator is the pixel: acc[0..1] belong to my first row, acc[2..3] to my second
            za[half] = __byte_perm((uint32_t)acc[0], (uint32_t)acc[1], 0x0062);
            zb[half] = __byte_perm((uint32_t)acc[2], (uint32_t)acc[3], 0x0062);
        }
        // words of my two rows: columns 2q, 2q+1 of n-tile 0, then of n-tile 1
        const uint32_t n0 = __byte_perm(za[0], za[1], 0x5410), n1 = __byte_perm(zb[0], zb[1], 0x5410);
        if (k < nr) {
            w0 = n0;
            w1 = n1;
        }
    }
}

// The same tile arithmetic with TMA doing the moving (default when the plane can be described by a tensor map).  Every
// WARP runs its own pipeline -- no CTA-wide barrier, because the tiles of a CTA need anything from 0 to 10 rounds.  A warp
// takes strips of eight horizontally adjacent tiles: one 16-row x 128-byte box (128-byte swizzle) loaded onto an mbarrier,
// the eight tiles blurred in place (words read from and written back to shared memory), one box store; three strip
// buffers per warp, so the strip after next is already in flight while a strip is being worked on.  The direct
// version below loads each tile when it is needed -- 16 rows of 16 bytes, sixteen 128-byte lines per instruction -- and so
// exposes the memory latency once per tile, which is what bounds the few-rounds case (presley: 0..4 rounds, ~400 cycles
// per tile and scheduler for ~270 of work).  (Boxes of a single tile, 16 x 16 bytes, were measured slower than the direct
// loads: the TMA unit's cost is per box row, not per byte.)
template <int PB>
__global__ void __launch_bounds__(128) blur_imma_tma_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                                                            const int T, const int By, const int Bx, const int32_t* __restrict__ rounds) {
    constexpr int kWarps = 4, kBufs = 3;
    constexpr int kPerTile = 16 / PB;
    constexpr uint32_t kStrip = 2048;
    __shared__ __align__(1024) uint8_t s_buf[kWarps][kBufs][kStrip];
    __shared__ __align__(8) uint64_t s_full[kWarps][kBufs];
    __shared__ int32_t s_nr[kWarps][8][32];                              // rounds per tile of the current strip, one column per lane
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const int r0 = imma_tile_row(gq), r1 = imma_tile_row(gq + 8), c0 = 4 * tq;
    uint32_t a0 = 0u, a1 = 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a0 |= blur_operator_entry(PB, r0, c0 + i) << (8 * i);
        a1 |= blur_operator_entry(PB, r1, c0 + i) << (8 * i);
    }
    const int tiles_x = (Bx + kPerTile - 1) / kPerTile, tiles_y = (By + kPerTile - 1) / kPerTile;
    const int strips_x = (tiles_x + 7) / 8;
    const int64_t n_strips = (int64_t)T * tiles_y * strips_x;
    const int64_t stride = (int64_t)gridDim.x * kWarps, first = (int64_t)blockIdx.x * kWarps + w;
    const uint32_t buf_base = tma::smem_u32(&s_buf[w][0][0]), bar = tma::smem_u32(&s_full[w][0]);
    const uint32_t row0 = (uint32_t)(r0 * 128 + c0), row1 = (uint32_t)(r1 * 128 + c0);
    const int x0 = r0 & 7, x1 = r1 & 7;                                  // swizzle: 16-byte chunk j of row r sits at chunk j ^ (r & 7)

    struct Strip { int sx, ty, t; };
    auto strip_of = [&](int64_t s) {
        const int64_t q = s / strips_x;
        return Strip{(int)(s - q * strips_x), (int)(q % tiles_y), (int)(q / tiles_y)};
    };
    auto issue = [&](int64_t s, int b) {                                 // lane 0 only
        const Strip p = strip_of(s);
        tma::mbar_arrive_expect_tx(bar + 8u * b, kStrip);
        tma::load_3d(buf_base + (uint32_t)b * kStrip, &tm_in, p.sx * 128, p.ty * 16, p.t, bar + 8u * b);
    };
    // rounds of the blocks my 8 pixels belong to in the eight tiles of a strip (in every tile they lie in ONE block:
    // rows r0, r1 share a half, columns 4q..4q+3 too)
    auto load_rounds = [&](int64_t s, int (&nr)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) nr[j] = 0;
        if (s >= n_strips) return;
        const Strip p = strip_of(s);
        const int byq = p.ty * kPerTile + (PB == 8 ? (gq >= 4) : 0);
        if (byq >= By) return;
        const int32_t* row = rounds + ((int64_t)p.t * By + byq) * Bx;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int bxq = (p.sx * 8 + j) * kPerTile + (PB == 8 ? (tq >= 2) : 0);
            if (bxq < Bx) nr[j] = row[bxq];
        }
    };
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < kBufs; ++b) tma::mbar_init(bar + 8u * b, 1);
        tma::mbar_init_fence();
        if (first < n_strips) issue(first, 0);
        if (first + stride < n_strips) issue(first + stride, 1);
    }
    __syncwarp();
    int nr_next[8];
    load_rounds(first, nr_next);
    int it = 0;
    for (int64_t s = first; s < n_strips; s += stride, ++it) {
        const int b = it % kBufs;
#pragma unroll
        for (int j = 0; j < 8; ++j) s_nr[w][j][lane] = nr_next[j];       // read back by this lane only
        load_rounds(s + stride, nr_next);                                // in flight while this strip is worked on
        tma::mbar_wait(bar + 8u * b, (uint32_t)(it / kBufs) & 1u);
        uint8_t* buf = &s_buf[w][b][0];
        const int n_here = min(8, tiles_x - strip_of(s).sx * 8);         // tiles of this strip inside the plane (warp-uniform)
#pragma unroll 1
        for (int j = 0; j < n_here; ++j) {
            uint32_t* p0 = reinterpret_cast<uint32_t*>(buf + row0 + ((j ^ x0) << 4));
            uint32_t* p1 = reinterpret_cast<uint32_t*>(buf + row1 + ((j ^ x1) << 4));
            uint32_t w0 = *p0, w1 = *p1;
            blur_imma_rounds<PB>(w0, w1, a0, a1, s_nr[w][j][lane]);
            *p0 = w0;
            *p1 = w1;
        }
        tma::fence_proxy_async();                                        // my in-place writes, before the box store reads them
        __syncwarp();
        if (lane == 0) {
            const Strip p = strip_of(s);
            tma::store_3d(&tm_out, p.sx * 128, p.ty * 16, p.t, buf_base + (uint32_t)b * kStrip);
            tma::store_commit();
            // buffer (it + 2) % 3 == (it - 1) % 3 was stored from at the end of the previous strip: once that store has
            // read it, it takes the strip after next
            if (s + 2 * stride < n_strips) {
                tma::store_wait_read<1>();
                issue(s + 2 * stride, (it + 2) % kBufs);
            }
        }
    }
    if (lane == 0) tma::store_wait<0>();                                 // shared memory must outlive the last stores' reads
}

template <int PB, bool ALIGNED>
__global__ void __launch_bounds__(256) blur_imma_kernel(const BlockGeom g, const int32_t* __restrict__ rounds) {
    constexpr int kWarps = 8;
    constexpr int kPerTile = 16 / PB;                       // blocks per tile side: 1 (luma) or 2 (chroma)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const int r0 = imma_tile_row(gq), r1 = imma_tile_row(gq + 8), c0 = 4 * tq;
    uint32_t a0 = 0u, a1 = 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a0 |= blur_operator_entry(PB, r0, c0 + i) << (8 * i);
        a1 |= blur_operator_entry(PB, r1, c0 + i) << (8 * i);
    }
    const int tiles_x = (g.Bx + kPerTile - 1) / kPerTile, tiles_y = (g.By + kPerTile - 1) / kPerTile;
    const int64_t n_tiles = (int64_t)g.T * tiles_y * tiles_x;
    const int64_t stride = (int64_t)gridDim.x * kWarps;
    for (int64_t tile = (int64_t)blockIdx.x * kWarps + w; tile < n_tiles; tile += stride) {
        const int tx = (int)(tile % tiles_x);
        const int64_t q = tile / tiles_x;
        const int ty = (int)(q % tiles_y), t = (int)(q / tiles_y);
        // the thread's 8 pixels lie in ONE block of the tile: rows r0, r1 share a half, columns 4q..4q+3 too
        const int byq = ty * kPerTile + (PB == 8 ? (gq >= 4) : 0), bxq = tx * kPerTile + (PB == 8 ? (tq >= 2) : 0);
        const bool live = byq < g.By && bxq < g.Bx;
        int nr = 0;
        const uint8_t* sp = g.src;
        uint8_t* dp = g.dst;
        if (live) {
            nr = rounds[((int64_t)t * g.By + byq) * g.Bx + bxq];
            sp += (int64_t)t * g.src_frame + (int64_t)ty * 16 * g.src_row + (int64_t)tx * 16 + c0;
            dp += (int64_t)t * g.dst_frame + (int64_t)ty * 16 * g.dst_row + (int64_t)tx * 16 + c0;
        }
        uint32_t w0 = 0u, w1 = 0u;
        if (live) {
            const uint8_t *p0 = sp + (int64_t)r0 * g.src_row, *p1 = sp + (int64_t)r1 * g.src_row;
            if (ALIGNED) {
                w0 = __ldcs(reinterpret_cast<const uint32_t*>(p0));
                w1 = __ldcs(reinterpret_cast<const uint32_t*>(p1));
            } else {
                w0 = p0[0] | (p0[1] << 8) | (p0[2] << 16) | ((uint32_t)p0[3] << 24);
                w1 = p1[0] | (p1[1] << 8) | (p1[2] << 16) | ((uint32_t)p1[3] << 24);
            }
        }
        blur_imma_rounds<PB>(w0, w1, a0, a1, nr);
        if (live) {
            uint8_t *p0 = dp + (int64_t)r0 * g.dst_row, *p1 = dp + (int64_t)r1 * g.dst_row;
            if (ALIGNED) {
                __stcs(reinterpret_cast<uint32_t*>(p0), w0);
                __stcs(reinterpret_cast<uint32_t*>(p1), w1);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    p0[j] = (uint8_t)(w0 >> (8 * j));
                    p1[j] = (uint8_t)(w1 >> (8 * j));
                }
            }
        }
    }
}

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
                                                              const int32_t* __restrict__ tables, int n_levels) {
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
        const bool live = b < n_blocks;
        int L = 0;
        const uint8_t* sp = g.src;
        uint8_t* dp = g.dst;
        if (live) {
            const int bx = (int)(b % g.Bx);
            const int64_t q = b / g.Bx;
            const int by = (int)(q % g.By), t = (int)(q / g.By);
            int lv = levels[b];
            lv = lv < 0 ? 0 : (lv >= n_levels ? n_levels - 1 : lv);
            const int small = __ldg(tables + (size_t)lv * kLevelStride);
            L = small >= PB ? 0 : 31 - __clz(PB / small);
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

// The same per-tile transform for planar, 8-byte aligned planes with half the floating-point instructions
// where it is free: the tile lives in registers as packed fp32 pairs (row r, columns 2j / 2j+1); the passes
// along a row are scalar butterflies on the halves, the passes down the columns are ONE packed butterfly per
// column pair (FADD2 / FMUL2 / FFMA2), the gains multiply pairs, and the bytes come back through the
// magic-number rounding (x + 1.5 * 2^23: round-half-even like rint, full-rate FADD2 instead of the
// quarter-rate F2I) and cvt.pack.sat (clamp + pack, two pixels per instruction).
__device__ __forceinline__ uint32_t pack_sat_u8x4(int a, int b, int c, int d) {     // bytes (a, b, c, d), each clamped to 0..255
    uint32_t lo, r;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(lo) : "r"(d), "r"(c), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(a), "r"(lo));
    return r;
}

__global__ void __launch_bounds__(128) dampen_packed_kernel(const BlockGeom g, const float* __restrict__ strength, const uint32_t magic) {
    const int tiles_x = g.Bx * g.pb / 8, tiles_y = g.By * g.pb / 8;
    const int64_t total = (int64_t)g.T * tiles_y * tiles_x;
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    int64_t b = id;
    const int txi = (int)(b % tiles_x);
    b /= tiles_x;
    const int tyi = (int)(b % tiles_y);
    const int t = (int)(b / tiles_y);
    const float s = fminf(fmaxf(strength[((int64_t)t * g.By + (tyi * 8) / g.pb) * g.Bx + (txi * 8) / g.pb], 0.f), 1.f);
    const uint8_t* sp = g.src + (int64_t)t * g.src_frame + (int64_t)tyi * 8 * g.src_row + (int64_t)txi * 8;
    uint8_t* dp = g.dst + (int64_t)t * g.dst_frame + (int64_t)tyi * 8 * g.dst_row + (int64_t)txi * 8;

    float2 x[8][4];
    const float2 bias = make_float2(-8388608.f, -8388608.f);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const uint2 v = __ldcs(reinterpret_cast<const uint2*>(sp + (int64_t)r * g.src_row));
        x[r][0] = __fadd2_rn(make_float2(byte_as_biased_float<0>(v.x, magic), byte_as_biased_float<1>(v.x, magic)), bias);
        x[r][1] = __fadd2_rn(make_float2(byte_as_biased_float<2>(v.x, magic), byte_as_biased_float<3>(v.x, magic)), bias);
        x[r][2] = __fadd2_rn(make_float2(byte_as_biased_float<0>(v.y, magic), byte_as_biased_float<1>(v.y, magic)), bias);
        x[r][3] = __fadd2_rn(make_float2(byte_as_biased_float<2>(v.y, magic), byte_as_biased_float<3>(v.y, magic)), bias);
    }
    // forward: along the rows (scalar), then down the columns (packed)
#pragma unroll
    for (int r = 0; r < 8; ++r) ELVIS_FDCT8(x[r][0].x, x[r][0].y, x[r][1].x, x[r][1].y, x[r][2].x, x[r][2].y, x[r][3].x, x[r][3].y);
#pragma unroll
    for (int j = 0; j < 4; ++j) ELVIS_FDCT8_X2(x[0][j], x[1][j], x[2][j], x[3][j], x[4][j], x[5][j], x[6][j], x[7][j]);
    // gain 2^(-4 s (u+v)/14) / 64: powers of q = 2^(-4 s / 14); the 1/64 undoes the AAN scaling
    float gk[15];
    const float q = exp2f(-4.0f * s / 14.0f);
    gk[0] = 1.0f / 64.0f;
#pragma unroll
    for (int k = 1; k < 15; ++k) gk[k] = gk[k - 1] * q;
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int j = 0; j < 4; ++j) x[u][j] = __fmul2_rn(x[u][j], make_float2(gk[u + 2 * j], gk[u + 2 * j + 1]));
    // inverse: down the columns (packed), then along the rows (scalar)
#pragma unroll
    for (int j = 0; j < 4; ++j) ELVIS_IDCT8_X2(x[0][j], x[1][j], x[2][j], x[3][j], x[4][j], x[5][j], x[6][j], x[7][j]);
#pragma unroll
    for (int r = 0; r < 8; ++r) ELVIS_IDCT8(x[r][0].x, x[r][0].y, x[r][1].x, x[r][1].y, x[r][2].x, x[r][2].y, x[r][3].x, x[r][3].y);
    // round half to even and clamp: x + 1.5 * 2^23 leaves rint(x) in the low mantissa bits (two's complement around 0x4B400000)
    const float2 rnd = make_float2(12582912.f, 12582912.f);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int n[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 y = __fadd2_rn(x[r][j], rnd);
            n[2 * j] = __float_as_int(y.x) - 0x4B400000;
            n[2 * j + 1] = __float_as_int(y.y) - 0x4B400000;
        }
        __stcs(reinterpret_cast<uint2*>(dp + (int64_t)r * g.dst_row), make_uint2(pack_sat_u8x4(n[0], n[1], n[2], n[3]), pack_sat_u8x4(n[4], n[5], n[6], n[7])));
    }
}

// Two horizontally adjacent tiles per thread, packed ELEMENT-WISE: x[r][c] = (tile A [r][c], tile B [r][c]).  All four
// passes are then packed butterflies with no register transposes at all (the row passes of dampen_packed_kernel are
// scalar because its pairs run along a row): half the floating-point instructions of the scalar kernel per tile, at
// the price of 128 live registers for the two tiles.
__global__ void __launch_bounds__(128) dampen_pair_kernel(const BlockGeom g, const float* __restrict__ strength, const uint32_t magic) {
    const int pairs_x = g.Bx * g.pb / 16, tiles_y = g.By * g.pb / 8;     // width in tile pairs (Bx * pb is a multiple of 16 here)
    const int64_t total = (int64_t)g.T * tiles_y * pairs_x;
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    int64_t b = id;
    const int pxi = (int)(b % pairs_x);
    b /= pairs_x;
    const int tyi = (int)(b % tiles_y);
    const int t = (int)(b / tiles_y);
    const float* srow = strength + ((int64_t)t * g.By + (tyi * 8) / g.pb) * g.Bx;
    const float sa = fminf(fmaxf(srow[(pxi * 16) / g.pb], 0.f), 1.f), sb = fminf(fmaxf(srow[(pxi * 16 + 8) / g.pb], 0.f), 1.f);
    const uint8_t* sp = g.src + (int64_t)t * g.src_frame + (int64_t)tyi * 8 * g.src_row + (int64_t)pxi * 16;
    uint8_t* dp = g.dst + (int64_t)t * g.dst_frame + (int64_t)tyi * 8 * g.dst_row + (int64_t)pxi * 16;

    float2 x[8][8];
    const float2 bias = make_float2(-8388608.f, -8388608.f);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const uint4 v = __ldcs(reinterpret_cast<const uint4*>(sp + (int64_t)r * g.src_row));
        x[r][0] = __fadd2_rn(make_float2(byte_as_biased_float<0>(v.x, magic), byte_as_biased_float<0>(v.z, magic)), bias);
        x[r][1] = __fadd2_rn(make_float2(byte_as_biased_float<1>(v.x, magic), byte_as_biased_float<1>(v.z, magic)), bias);
        x[r][2] = __fadd2_rn(make_float2(byte_as_biased_float<2>(v.x, magic), byte_as_biased_float<2>(v.z, magic)), bias);
        x[r][3] = __fadd2_rn(make_float2(byte_as_biased_float<3>(v.x, magic), byte_as_biased_float<3>(v.z, magic)), bias);
        x[r][4] = __fadd2_rn(make_float2(byte_as_biased_float<0>(v.y, magic), byte_as_biased_float<0>(v.w, magic)), bias);
        x[r][5] = __fadd2_rn(make_float2(byte_as_biased_float<1>(v.y, magic), byte_as_biased_float<1>(v.w, magic)), bias);
        x[r][6] = __fadd2_rn(make_float2(byte_as_biased_float<2>(v.y, magic), byte_as_biased_float<2>(v.w, magic)), bias);
        x[r][7] = __fadd2_rn(make_float2(byte_as_biased_float<3>(v.y, magic), byte_as_biased_float<3>(v.w, magic)), bias);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) ELVIS_FDCT8_X2(x[r][0], x[r][1], x[r][2], x[r][3], x[r][4], x[r][5], x[r][6], x[r][7]);
#pragma unroll
    for (int c = 0; c < 8; ++c) ELVIS_FDCT8_X2(x[0][c], x[1][c], x[2][c], x[3][c], x[4][c], x[5][c], x[6][c], x[7][c]);
    // gains 2^(-4 s (u+v)/14) / 64 of the two tiles, packed
    float2 gk[15];
    const float2 q = make_float2(exp2f(-4.0f * sa / 14.0f), exp2f(-4.0f * sb / 14.0f));
    gk[0] = make_float2(1.0f / 64.0f, 1.0f / 64.0f);
#pragma unroll
    for (int k = 1; k < 15; ++k) gk[k] = __fmul2_rn(gk[k - 1], q);
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int v = 0; v < 8; ++v) x[u][v] = __fmul2_rn(x[u][v], gk[u + v]);
#pragma unroll
    for (int c = 0; c < 8; ++c) ELVIS_IDCT8_X2(x[0][c], x[1][c], x[2][c], x[3][c], x[4][c], x[5][c], x[6][c], x[7][c]);
#pragma unroll
    for (int r = 0; r < 8; ++r) ELVIS_IDCT8_X2(x[r][0], x[r][1], x[r][2], x[r][3], x[r][4], x[r][5], x[r][6], x[r][7]);
    const float2 rnd = make_float2(12582912.f, 12582912.f);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int na[8], nb[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float2 y = __fadd2_rn(x[r][c], rnd);
            na[c] = __float_as_int(y.x) - 0x4B400000;
            nb[c] = __float_as_int(y.y) - 0x4B400000;
        }
        __stcs(reinterpret_cast<uint4*>(dp + (int64_t)r * g.dst_row),
               make_uint4(pack_sat_u8x4(na[0], na[1], na[2], na[3]), pack_sat_u8x4(na[4], na[5], na[6], na[7]),
                          pack_sat_u8x4(nb[0], nb[1], nb[2], nb[3]), pack_sat_u8x4(nb[4], nb[5], nb[6], nb[7])));
    }
}

// ------------------------------------------------------------- dampen on the tensor cores
// The gain 2^(-4 s (u + v) / 14) = q^u q^v is separable, so dampening an 8 x 8 tile is X' = M X M^T with the
// 8 x 8 operator M(s) = A^T diag(q^u) A (A = orthonormal DCT-II, q = 2^(-4 s / 14)): per block two small
// matrix products instead of a forward and an inverse DCT.  A warp owns a 16 x 16 tile -- one luma block
// (2 x 2 transform tiles, operator diag(M, M)) or two 8 x 8 blocks of a plane with 8-pixel blocks placed on
// the diagonal quadrants (operator diag(M(s_a), M(s_b))) -- in the accumulator layout of
// mma.sync.m16n8k16 (f16 x f16 -> f32), chained exactly like the blur above: step 1 M X^T, step 2
// M (M X^T)^T.  Pixels minus 128 are exact in f16 (M preserves constants, 128 is added back); M and the
// intermediate are split hi + lo in f16 (lo x lo dropped, 2^-22 relative), fp32 accumulation: the result
// is within 1e-4 of the float64 reconstruction before rounding (tools/emu/check_dampen_hmma.py), the
// bar being 0.0255.  10 HMMA per tile; M is built cooperatively (two entries per lane, shared memory).
__device__ __forceinline__ void hmma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, const float (&c)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%10, %11, %12, %13};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}

__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// hi + lo split of two floats: (hi pair, lo pair) as packed f16
__device__ __forceinline__ void split_half2(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x, y);
    const float2 back = __half22float2(h);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = pack_half2(x - back.x, y - back.y);
}

template <int PB>
__global__ void __launch_bounds__(256) dampen_hmma_kernel(const BlockGeom g, const float* __restrict__ strength) {
    constexpr int kWarps = 8;
    constexpr int kSlots = PB == 16 ? 1 : 2;               // operators per tile
    __shared__ __align__(8) __half s_m[kWarps][2][kSlots][8][8];   // [hi / lo][slot][row][col]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const int r0 = imma_tile_row(gq), r1 = imma_tile_row(gq + 8), c0 = 4 * tq;
    const int hr = gq >= 4, hc = tq >= 2;                  // which half of the tile my rows / columns lie in
    const bool diag = hr == hc;                            // my 8 pixels sit in a diagonal quadrant
    // coefficients of this lane's two operator entries: M[i][j] = sum_u q^u A[u][i] A[u][j]
    float kc[2][8];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int i = (lane >> 3) + 4 * e, j = lane & 7;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float cu = u == 0 ? 0.125f : 0.25f;          // c(u)^2: 1/8 for u = 0, 1/4 otherwise
            kc[e][u] = cu * cospif((float)((2 * i + 1) * u) / 16.0f) * cospif((float)((2 * j + 1) * u) / 16.0f);
        }
    }
    const int tiles_x = PB == 16 ? g.Bx : (g.Bx + 1) / 2;
    const int64_t n_tiles = (int64_t)g.T * g.By * tiles_x;
    const int64_t stride = (int64_t)gridDim.x * kWarps;
    const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
    for (int64_t tile = (int64_t)blockIdx.x * kWarps + w; tile < n_tiles; tile += stride) {
        const int tx = (int)(tile % tiles_x);
        const int64_t q = tile / tiles_x;
        const int by = (int)(q % g.By), t = (int)(q / g.By);
        // ---- operators of the tile, two entries per lane and slot
#pragma unroll
        for (int slot = 0; slot < kSlots; ++slot) {
            const int bxs = PB == 16 ? tx : 2 * tx + slot;
            float sv = bxs < g.Bx ? strength[((int64_t)t * g.By + by) * g.Bx + bxs] : 0.f;
            sv = fminf(fmaxf(sv, 0.f), 1.f);
            const float qq = exp2f(-4.0f * sv / 14.0f);
            float pw = 1.f, m0 = 0.f, m1 = 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                m0 = fmaf(kc[0][u], pw, m0);
                m1 = fmaf(kc[1][u], pw, m1);
                pw *= qq;
            }
            const __half h0 = __float2half_rn(m0), h1 = __float2half_rn(m1);
            s_m[w][0][slot][lane >> 3][lane & 7] = h0;
            s_m[w][0][slot][4 + (lane >> 3)][lane & 7] = h1;
            s_m[w][1][slot][lane >> 3][lane & 7] = __float2half_rn(m0 - __half2float(h0));
            s_m[w][1][slot][4 + (lane >> 3)][lane & 7] = __float2half_rn(m1 - __half2float(h1));
        }
        __syncwarp();
        // ---- my A fragments: rows r0 / r1, columns c0..c0+3 of diag(M, M'); zero off the diagonal quadrants
        uint32_t ah[4] = {0u, 0u, 0u, 0u}, al[4] = {0u, 0u, 0u, 0u};
        if (diag) {
            const int slot = PB == 16 ? 0 : hr;
            const uint2 h0 = *reinterpret_cast<const uint2*>(&s_m[w][0][slot][r0 & 7][c0 & 7]);
            const uint2 h1 = *reinterpret_cast<const uint2*>(&s_m[w][0][slot][r1 & 7][c0 & 7]);
            const uint2 l0 = *reinterpret_cast<const uint2*>(&s_m[w][1][slot][r0 & 7][c0 & 7]);
            const uint2 l1 = *reinterpret_cast<const uint2*>(&s_m[w][1][slot][r1 & 7][c0 & 7]);
            ah[0] = h0.x; ah[1] = h1.x; ah[2] = h0.y; ah[3] = h1.y;      // a0 (r0, c0..1)  a1 (r1, c0..1)  a2 (r0, c0+2..3)  a3 (r1, c0+2..3)
            al[0] = l0.x; al[1] = l1.x; al[2] = l0.y; al[3] = l1.y;
        }
        __syncwarp();                                      // the table is free for the next tile
        // ---- my pixels: luma -- every thread; 8-pixel blocks -- the diagonal quadrants hold blocks 2 tx and 2 tx + 1
        const int bxq = PB == 16 ? tx : 2 * tx + hr;
        const bool live = PB == 16 ? true : (diag && bxq < g.Bx);
        const int64_t col = PB == 16 ? (int64_t)tx * 16 + c0 : (int64_t)bxq * 8 + (c0 & 7);
        const int rr0 = PB == 16 ? r0 : (r0 & 7), rr1 = PB == 16 ? r1 : (r1 & 7);
        uint32_t w0 = 0x80808080u, w1 = 0x80808080u;       // 128: zero after centring
        if (live) {
            const uint8_t* sp = g.src + (int64_t)t * g.src_frame + (int64_t)by * PB * g.src_row + col;
            w0 = __ldcs(reinterpret_cast<const uint32_t*>(sp + (int64_t)rr0 * g.src_row));
            w1 = __ldcs(reinterpret_cast<const uint32_t*>(sp + (int64_t)rr1 * g.src_row));
        }
        // bytes -> f16 pairs minus 128: PRMT builds 0x6400 | byte = 1024 + byte, exact subtraction of 1152
        const __half2 off = __floats2half2_rn(1152.f, 1152.f);
        auto centred = [&](uint32_t word, int pair) -> uint32_t {
            const uint32_t e = __byte_perm(word, 0x64646464u, pair ? 0x7372 : 0x7170);
            const __half2 hv = __hsub2(*reinterpret_cast<const __half2*>(&e), off);
            return *reinterpret_cast<const uint32_t*>(&hv);
        };
        // ---- step 1: M1 = M X^T, n-tile 0 from my first row, n-tile 1 from my second row
        float m1a[4], m1b[4];
        {
            const uint32_t b0 = centred(w0, 0), b1 = centred(w0, 1);
            hmma_16816(m1a, ah, b0, b1, zero4);
            hmma_16816(m1a, al, b0, b1, m1a);
        }
        {
            const uint32_t b0 = centred(w1, 0), b1 = centred(w1, 1);
            hmma_16816(m1b, ah, b0, b1, zero4);
            hmma_16816(m1b, al, b0, b1, m1b);
        }
        // ---- step 2: Z = M M1^T with M1 split hi + lo
        float z[2][4];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t yh0, yl0, yh1, yl1;
            split_half2(m1a[2 * half], m1a[2 * half + 1], yh0, yl0);
            split_half2(m1b[2 * half], m1b[2 * half + 1], yh1, yl1);
            hmma_16816(z[half], ah, yh0, yh1, zero4);
            hmma_16816(z[half], al, yh0, yh1, z[half]);
            hmma_16816(z[half], ah, yl0, yl1, z[half]);
        }
        // ---- my first row: (z[0][0], z[0][1], z[1][0], z[1][1]); second row: (z[0][2], z[0][3], z[1][2], z[1][3])
        if (live) {
            auto to_byte = [](float v) -> uint32_t {
                const int i = __float2int_rn(v + 128.f);
                return (uint32_t)(i < 0 ? 0 : (i > 255 ? 255 : i));
            };
            const uint32_t o0 = to_byte(z[0][0]) | (to_byte(z[0][1]) << 8) | (to_byte(z[1][0]) << 16) | (to_byte(z[1][1]) << 24);
            const uint32_t o1 = to_byte(z[0][2]) | (to_byte(z[0][3]) << 8) | (to_byte(z[1][2]) << 16) | (to_byte(z[1][3]) << 24);
            uint8_t* dp = g.dst + (int64_t)t * g.dst_frame + (int64_t)by * PB * g.dst_row + col;
            __stcs(reinterpret_cast<uint32_t*>(dp + (int64_t)rr0 * g.dst_row), o0);
            __stcs(reinterpret_cast<uint32_t*>(dp + (int64_t)rr1 * g.dst_row), o1);
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
    const char* blur_impl = getenv("ELVIS_BLUR_IMPL");       // imma (default) | dp4a | generic
    if (g.C == 1 && (block_px == 16 || block_px == 8) && !(blur_impl && (!strcmp(blur_impl, "dp4a") || !strcmp(blur_impl, "generic"))) &&
        !getenv("ELVIS_BLUR_GENERIC")) {
        const bool al4 = aligned_to(g.src, 4) && aligned_to(g.dst, 4) && g.src_frame % 4 == 0 && g.dst_frame % 4 == 0 &&
                         g.src_row % 4 == 0 && g.dst_row % 4 == 0;
        const int per = 16 / block_px;
        const int64_t tiles = (int64_t)n_frames * ((by + per - 1) / per) * ((bx + per - 1) / per);
        const int grid = grid_for_units(tiles, 8);
        const char* use_tma = getenv("ELVIS_BLUR_TMA");            // 0: direct loads and stores
        CUtensorMap tm_in, tm_out;
        if (!(use_tma && use_tma[0] == '0') &&
            tma::make_plane_map(&tm_in, g.src, bx * block_px, by * block_px, n_frames, g.src_row, g.src_frame, 128, 16, CU_TENSOR_MAP_SWIZZLE_128B) &&
            tma::make_plane_map(&tm_out, g.dst, bx * block_px, by * block_px, n_frames, g.dst_row, g.dst_frame, 128, 16, CU_TENSOR_MAP_SWIZZLE_128B)) {
            const int tiles_x = (bx + per - 1) / per;
            const int64_t strips = (int64_t)n_frames * ((by + per - 1) / per) * ((tiles_x + 7) / 8);
            const int sgrid = grid_for_units(strips, 4);           // 4 warps per CTA, up to 16 CTAs per SM's worth of strips
            if (block_px == 16) blur_imma_tma_kernel<16><<<sgrid, 128, 0, st>>>(tm_in, tm_out, n_frames, by, bx, rounds);
            else blur_imma_tma_kernel<8><<<sgrid, 128, 0, st>>>(tm_in, tm_out, n_frames, by, bx, rounds);
            ELVIS_CHECK_LAUNCH();
            return ELVIS_OK;
        }
        if (block_px == 16) {
            if (al4) blur_imma_kernel<16, true><<<grid, 256, 0, st>>>(g, rounds);
            else blur_imma_kernel<16, false><<<grid, 256, 0, st>>>(g, rounds);
        } else {
            if (al4) blur_imma_kernel<8, true><<<grid, 256, 0, st>>>(g, rounds);
            else blur_imma_kernel<8, false><<<grid, 256, 0, st>>>(g, rounds);
        }
        ELVIS_CHECK_LAUNCH();
        return ELVIS_OK;
    }
    if (g.C == 1 && (block_px == 16 || block_px == 8) && !(blur_impl && !strcmp(blur_impl, "generic")) && !getenv("ELVIS_BLUR_GENERIC")) {
        const bool al = aligned_to(g.src, 8) && aligned_to(g.dst, 8) && g.src_frame % 8 == 0 && g.dst_frame % 8 == 0 &&
                        g.src_row % 8 == 0 && g.dst_row % 8 == 0;
        const int64_t blocks = (int64_t)n_frames * by * bx;
        const int per_cta = 8 * (block_px == 16 ? 1 : 4);
        const int grid = grid_for_units(blocks, per_cta);
        if (block_px == 16) {
            if (al) blur_fast_kernel<16, true><<<grid, 256, 0, st>>>(g, rounds);
            else blur_fast_kernel<16, false><<<grid, 256, 0, st>>>(g, rounds);
        } else {
            if (al) blur_fast_kernel<8, true><<<grid, 256, 0, st>>>(g, rounds);
            else blur_fast_kernel<8, false><<<grid, 256, 0, st>>>(g, rounds);
        }
        ELVIS_CHECK_LAUNCH();
        return ELVIS_OK;
    }
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
                                        const int32_t* tables, int32_t n_levels, int32_t fast_tables_ok,
                                        elvis_stream_t stream) {
    BlockGeom g;
    if (int rc = make_geom(src, dst, n_frames, block_px, by, bx, g)) return rc;
    if (!levels || !tables || n_levels <= 0) return ELVIS_ERR_INVALID_ARG;
    if (block_px > 64) return ELVIS_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    if (int rc = copy_edges(g, st)) return rc;
    const int n = block_px * block_px;
    if (g.C == 1 && (block_px == 16 || block_px == 8) && fast_tables_ok && !getenv("ELVIS_DOWNSAMPLE_GENERIC")) {
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
    int wpc = 8;
    while (wpc > 1 && (size_t)wpc * n * 6 > 48 * 1024) wpc >>= 1;
    const size_t smem = (size_t)wpc * n * 6;
    const int64_t units = (int64_t)n_frames * by * bx * g.C;
    downsample_kernel<false><<<grid_for_units(units, wpc), wpc * 32, smem, st>>>(g, levels, tables, n_levels, wpc);
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
    const bool al4 = aligned_to(g.src, 4) && aligned_to(g.dst, 4) && g.src_frame % 4 == 0 && g.dst_frame % 4 == 0 &&
                     g.src_row % 4 == 0 && g.dst_row % 4 == 0;
    // ELVIS_DAMPEN_IMPL=hmma selects the tensor-core variant (correct, but measured 2.8x slower than the CUDA-core
    // kernel on B200: 0.99 vs 0.35 ms per 30 4K frames -- one dependent chain per tile)
    const char* dampen_impl = getenv("ELVIS_DAMPEN_IMPL");
    if (g.C == 1 && al4 && (block_px == 16 || block_px == 8) && dampen_impl && !strcmp(dampen_impl, "hmma")) {
        const int64_t tiles = (int64_t)n_frames * by * (block_px == 16 ? bx : (bx + 1) / 2);
        const int grid = grid_for_units(tiles, 8);
        if (block_px == 16) dampen_hmma_kernel<16><<<grid, 256, 0, st>>>(g, strength);
        else dampen_hmma_kernel<8><<<grid, 256, 0, st>>>(g, strength);
        ELVIS_CHECK_LAUNCH();
        return ELVIS_OK;
    }
    const int64_t total = (int64_t)n_frames * (by * block_px / 8) * (bx * block_px / 8) * g.C;
    const bool fast = g.C == 1 && aligned_to(g.src, 8) && aligned_to(g.dst, 8) && g.src_frame % 8 == 0 &&
                      g.dst_frame % 8 == 0 && g.src_row % 8 == 0 && g.dst_row % 8 == 0;
    const unsigned grid = (unsigned)((total + 127) / 128);
    const bool al16 = aligned_to(g.src, 16) && aligned_to(g.dst, 16) && g.src_frame % 16 == 0 && g.dst_frame % 16 == 0 &&
                      g.src_row % 16 == 0 && g.dst_row % 16 == 0 && (bx * block_px) % 16 == 0;
    if (fast && al16 && dampen_impl && !strcmp(dampen_impl, "pair")) {   // two tiles per thread, element-wise packed (experimental)
        const int64_t pairs = total / 2;
        dampen_pair_kernel<<<(unsigned)((pairs + 127) / 128), 128, 0, st>>>(g, strength, 0x4B000000u);
    } else if (fast && !(dampen_impl && !strcmp(dampen_impl, "scalar")))      // packed-fp32 kernel (default); ELVIS_DAMPEN_IMPL=scalar: the round-1 kernel
        dampen_packed_kernel<<<grid, 128, 0, st>>>(g, strength, 0x4B000000u);
    else if (fast)
        dampen_kernel<true><<<grid, 128, 0, st>>>(g, strength);
    else
        dampen_kernel<false><<<grid, 128, 0, st>>>(g, strength);
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
