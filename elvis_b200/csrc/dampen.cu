// a14: per-8x8-tile DCT coefficient attenuation (oracle/spec_dct_dampen.py; absent from the reference, README.md:11,44).
#include "degrade_common.cuh"
#include "dct8.cuh"
#include "dct8_packed.cuh"
#include <cstring>
#include <cuda_fp16.h>

namespace elvis {
namespace {

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

}  // namespace
}  // namespace elvis

using namespace elvis;

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
