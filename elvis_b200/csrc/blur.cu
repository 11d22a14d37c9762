// a9, a11, a12 (blur): `rounds` successive 5x5 sigma=1 Gaussian blurs of the ISOLATED block (cv2.GaussianBlur u8
// fixed-point path; oracle/spec_cv.py:gaussian_blur5).  Integer arithmetic, bit-exact by construction.
#include "degrade_common.cuh"
#include "tma.cuh"
#include <cstring>

namespace elvis {
namespace {

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
__global__ void __launch_bounds__(128, 8) blur_imma_tma_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                                                            const int T, const int By, const int Bx, const int32_t* __restrict__ rounds) {
    constexpr int kWarps = 4, kBufs = 3;
    constexpr int kPerTile = 16 / PB;
    constexpr uint32_t kStrip = 2048;
    __shared__ __align__(1024) uint8_t s_buf[kWarps][kBufs][kStrip];
    __shared__ __align__(8) uint64_t s_full[kWarps][kBufs];
    __shared__ int32_t s_nr[kWarps][8][4];                               // rounds per tile of the current strip and block of the tile (quadrant)
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
    const int quad = PB == 8 ? ((gq >= 4) * 2 + (tq >= 2)) : 0;          // which block of the tile my 8 pixels belong to

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
        for (int j = 0; j < 8; ++j) s_nr[w][j][quad] = nr_next[j];       // every lane of a quadrant writes the same value
        load_rounds(s + stride, nr_next);                                // in flight while this strip is worked on
        tma::mbar_wait(bar + 8u * b, (uint32_t)(it / kBufs) & 1u);
        uint8_t* buf = &s_buf[w][b][0];
        const int n_here = min(8, tiles_x - strip_of(s).sx * 8);         // tiles of this strip inside the plane (warp-uniform)
#pragma unroll 1
        for (int j = 0; j < n_here; ++j) {
            uint32_t* p0 = reinterpret_cast<uint32_t*>(buf + row0 + ((j ^ x0) << 4));
            uint32_t* p1 = reinterpret_cast<uint32_t*>(buf + row1 + ((j ^ x1) << 4));
            uint32_t w0 = *p0, w1 = *p1;
            blur_imma_rounds<PB>(w0, w1, a0, a1, s_nr[w][j][quad]);
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
