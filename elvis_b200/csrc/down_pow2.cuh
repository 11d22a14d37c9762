// a8 / a12 for power-of-two reductions (the elvis 1x/2x/4x/8x(/16x) pyramid, elvis.py:2141-2169): cv2's
// INTER_AREA down + INTER_LINEAR up of an isolated PB x PB block in closed form, bit-exact with
// oracle/spec_cv.py:down_up (checked on the CPU against the restated cv2 arithmetic before the
// kernel was written, and on the GPU by tests/test_gpu_parity.py).
//
// For a reduction by f = 2^L (small image PB/f square, W = 2f):
//   area    S = cell sums over f x f;  L = 1: (s + 2) >> 2,  L >= 2: round-half-even(s / 4^L)
//   H pass  t[R][d] = c0(d) S[R][i0(d)] + c1(d) S[R][i1(d)],   c1(d) = (2d + 1 - f) mod W, c0 = W - c1,
//           i0 = clamp(floor((2d + 1 - f) / W)), i1 = clamp(i0' + 1)            (cv2's 11-bit weights / 2^(10-L))
//   V pass  out[y][d] = ((c0(y) t[i0(y)][d] >> 2L) + (c1(y) t[i1(y)][d] >> 2L) + 2) >> 2
// which is cv2's ((b0 (R0 >> 4)) >> 16) + ((b1 (R1 >> 4)) >> 16) + 2 >> 2 with the common powers of
// two cancelled.  Everything fits 16 bits, so two pixels travel per 32-bit register.
//
// Lane layout (shared with the blur kernel): a group of 2 PB lanes holds one block, lane g of the
// group owns PB/2 consecutive pixels of row r = g >> 1 (half h = g & 1): two words for PB = 16, one
// word for PB = 8.  A warp is one luma block or two chroma blocks; the level is uniform per group.
#pragma once
#include <stdint.h>

namespace elvis {

__device__ __forceinline__ uint32_t pk_lo(uint32_t x) { return x & 0xffffu; }
__device__ __forceinline__ uint32_t pk_pair(uint32_t lo, uint32_t hi) { return __byte_perm(lo, hi, 0x5410); }   // (lo.lo16, hi.lo16)

// ((a * ka) >> s) + ((b * kb) >> s) + 2 on both 16-bit halves; bits 8..15 of each half are scratch afterwards
template <int S2L>
__device__ __forceinline__ uint32_t v_pair(uint32_t a, uint32_t ka, uint32_t b, uint32_t kb) {
    constexpr uint32_t kMask = (0xffffu >> S2L) * 0x00010001u;
    const uint32_t x = ((a * ka) >> S2L) & kMask;
    const uint32_t y = ((b * kb) >> S2L) & kMask;
    return (x + y + 0x00020002u) >> 2;
}

// round-half-even(s / 2^K) on both 16-bit halves (K >= 2): (s + 2^(K-1) - 1 + ((s >> K) & 1)) >> K
template <int K>
__device__ __forceinline__ uint32_t rhe_pair(uint32_t s) {
    constexpr uint32_t kHalf = ((1u << (K - 1)) - 1u) * 0x00010001u;
    constexpr uint32_t kMask = (0xffffu >> K) * 0x00010001u;
    const uint32_t odd = (s >> K) & 0x00010001u;
    return ((s + kHalf + odd) >> K) & kMask;
}

// The degraded pixels of this lane (p0 = pixels 0..3 of its row segment, p1 = pixels 4..7 for PB = 16).
// g = lane index inside the group, `base` = first lane of the group inside the warp.  All lanes of the
// warp must call with the same <PB, L> being valid for their group (callers branch per group).
template <int PB, int L>
__device__ __forceinline__ void down_up_pow2(uint32_t& p0, uint32_t& p1, const int g, const int base) {
    static_assert((PB == 16 && L >= 1 && L <= 4) || (PB == 8 && L >= 1 && L <= 3), "unsupported reduction");
    constexpr uint32_t kFull = 0xffffffffu;
    const int h = g & 1, r = g >> 1;
    constexpr int f = 1 << L;
    // ---- horizontal byte sums inside the lane: pairs (16-bit halves)
    uint32_t e0 = (p0 & 0x00ff00ffu) + ((p0 >> 8) & 0x00ff00ffu);       // (b0+b1, b2+b3)
    uint32_t e1 = PB == 16 ? (p1 & 0x00ff00ffu) + ((p1 >> 8) & 0x00ff00ffu) : 0u;

    if constexpr ((PB == 16 && L == 4) || (PB == 8 && L == 3)) {
        // one cell: the whole block becomes its rounded mean
        uint32_t s = (e0 & 0xffffu) + (e0 >> 16) + (e1 & 0xffffu) + (e1 >> 16);
#pragma unroll
        for (int m = 1; m < 2 * PB; m <<= 1) s += __shfl_xor_sync(kFull, s, m);
        constexpr int K = 2 * L;
        const uint32_t v = (s + ((1u << (K - 1)) - 1u) + ((s >> K) & 1u)) >> K;
        p0 = v * 0x01010101u;
        p1 = p0;
        return;
    } else {
        // ---- cells of this lane, two per register
        uint32_t E0, E1 = 0u;        // PB 16: L1 (S0,S1),(S2,S3); L2 (S0,S1); L3 (S0,-).  PB 8: L1 (S0,S1); L2 (S0,-)
        if constexpr (L == 1) {
            e0 += __shfl_xor_sync(kFull, e0, 2);
            if (PB == 16) e1 += __shfl_xor_sync(kFull, e1, 2);
            E0 = ((e0 + 0x00020002u) >> 2) & 0x00ff00ffu;
            E1 = ((e1 + 0x00020002u) >> 2) & 0x00ff00ffu;
        } else if constexpr (L == 2) {
            uint32_t s = PB == 16 ? pk_pair((e0 & 0xffffu) + (e0 >> 16), (e1 & 0xffffu) + (e1 >> 16)) : (e0 & 0xffffu) + (e0 >> 16);
            s += __shfl_xor_sync(kFull, s, 2);
            s += __shfl_xor_sync(kFull, s, 4);
            E0 = rhe_pair<4>(s);
        } else {                     // PB 16, L 3
            uint32_t s = (e0 & 0xffffu) + (e0 >> 16) + (e1 & 0xffffu) + (e1 >> 16);
            s += __shfl_xor_sync(kFull, s, 2);
            s += __shfl_xor_sync(kFull, s, 4);
            s += __shfl_xor_sync(kFull, s, 8);
            E0 = rhe_pair<6>(s);
        }
        // ---- the neighbouring half's adjacent cell (the outer side clamps to the own edge cell)
        constexpr int kCells = (PB / 2) >> L;                     // cells per lane: 4, 2, 1 (PB 16) / 2, 1 (PB 8)
        const uint32_t first = pk_lo(E0);
        const uint32_t last = kCells == 4 ? E1 >> 16 : (kCells == 2 ? E0 >> 16 : pk_lo(E0));
        const uint32_t nb = __shfl_xor_sync(kFull, h ? first : last, 1);
        const uint32_t left = h ? nb : first, right = h ? last : nb;

        // ---- H pass: t for the lane's pixels, two per register (pairing chosen so that both halves share weights)
        uint32_t t[PB / 4];
        if constexpr (PB == 16 && L == 1) {
            // even pixel 2m: S[m-1] + 3 S[m]; odd pixel 2m+1: 3 S[m] + S[m+1].   t = {(t0,t2), (t4,t6), (t1,t3), (t5,t7)}
            const uint32_t T0 = E0 * 3u, T1 = E1 * 3u;
            const uint32_t mid = __byte_perm(E0, E1, 0x5432);            // (S1, S2)
            t[0] = pk_pair(left, E0) + T0;                                  // (left + 3 S0, S0 + 3 S1)
            t[1] = mid + T1;                                                // (S1 + 3 S2, S2 + 3 S3)
            t[2] = T0 + mid;                                                // (3 S0 + S1, 3 S1 + S2)
            t[3] = T1 + __byte_perm(E1, right, 0x5432);                   // (3 S2 + S3, 3 S3 + right)
        } else if constexpr (PB == 16 && L == 2) {
            // weights on (a, b): j = 0..7 -> (3,5) (1,7) (7,1) (5,3) (3,5) (1,7) (7,1) (5,3); pairs (j, j + 4)
            const uint32_t A = pk_pair(left, E0);                           // (left, S0)
            const uint32_t B = __byte_perm(E0, right, 0x5432);            // (S1, right)
            t[0] = A * 3u + E0 * 5u;                                        // (t0, t4)
            t[1] = A + E0 * 7u;                                             // (t1, t5)
            t[2] = E0 * 7u + B;                                             // (t2, t6)
            t[3] = E0 * 5u + B * 3u;                                        // (t3, t7)
        } else if constexpr (PB == 16 && L == 3) {
            // t_j = c0 left + c1 S0 (j = 0..3), t_j = c0' S0 + c1' right (j = 4..7) with weights (7,9) (5,11) (3,13) (1,15) |
            // (15,1) (13,3) (11,5) (9,7): mirror pairs (j, 7 - j) share their weights
            const uint32_t LR = pk_pair(left, right), SS = first * 0x00010001u;
            t[0] = LR * 7u + SS * 9u;                                       // (t0, t7)
            t[1] = LR * 5u + SS * 11u;                                      // (t1, t6)
            t[2] = LR * 3u + SS * 13u;                                      // (t2, t5)
            t[3] = LR + SS * 15u;                                           // (t3, t4)
        } else if constexpr (PB == 8 && L == 1) {
            // t0 = left + 3 S0, t1 = 3 S0 + S1, t2 = S0 + 3 S1, t3 = 3 S1 + right.   t = {(t0,t2), (t1,t3)}
            const uint32_t T0 = E0 * 3u;
            t[0] = pk_pair(left, E0) + T0;
            t[1] = T0 + __byte_perm(E0, right, 0x5432);
        } else {                     // PB 8, L 2: t0 = 3 left + 5 S0, t1 = left + 7 S0, t2 = 7 S0 + right, t3 = 5 S0 + 3 right
            const uint32_t LR = pk_pair(left, right), SS = first * 0x00010001u;
            t[0] = LR * 3u + SS * 5u;                                       // (t0, t3)
            t[1] = LR + SS * 7u;                                            // (t1, t2)
        }
        // ---- V pass: the other small row sits f rows away (lane distance 2 f), clamped at the block edge
        const int q = r & (f - 1);
        const bool up = q < f / 2;
        int src = g + (up ? -2 * f : 2 * f);
        if (src < 0 || src >= 2 * PB) src = g;
        const uint32_t c1 = (uint32_t)((2 * r + 1 - f) & (2 * f - 1));
        const uint32_t k_other = up ? 2 * f - c1 : c1, k_own = 2 * f - k_other;
        uint32_t v[PB / 4];
#pragma unroll
        for (int i = 0; i < PB / 4; ++i) {
            const uint32_t o = __shfl_sync(kFull, t[i], base + src);
            v[i] = v_pair<2 * L>(t[i], k_own, o, k_other);
        }
        // ---- bytes 0 and 2 of every v register are pixels; put them back in order
        if constexpr (PB == 16 && L == 1) {
            p0 = __byte_perm(v[0], v[2], 0x6240);                         // o0 o1 o2 o3
            p1 = __byte_perm(v[1], v[3], 0x6240);
        } else if constexpr (PB == 16 && L == 2) {
            const uint32_t x = __byte_perm(v[0], v[1], 0x6240);           // o0 o1 o4 o5
            const uint32_t y = __byte_perm(v[2], v[3], 0x6240);           // o2 o3 o6 o7
            p0 = __byte_perm(x, y, 0x5410);
            p1 = __byte_perm(x, y, 0x7632);
        } else if constexpr (PB == 16) {
            const uint32_t x = __byte_perm(v[0], v[1], 0x2640);           // o0 o1 o6 o7
            const uint32_t y = __byte_perm(v[2], v[3], 0x2640);           // o2 o3 o4 o5
            p0 = __byte_perm(x, y, 0x5410);
            p1 = __byte_perm(x, y, 0x3276);
        } else if constexpr (L == 1) {
            p0 = __byte_perm(v[0], v[1], 0x6240);
        } else {
            p0 = __byte_perm(v[0], v[1], 0x2640);                         // (o0, o3), (o1, o2) -> o0 o1 o2 o3
        }
    }
}

// run-time level for one group; `valid` lanes outside any block still take part in the shuffles
template <int PB>
__device__ __forceinline__ void down_up_pow2_level(uint32_t& p0, uint32_t& p1, const int L, const int g, const int base) {
    if (PB == 16) {
        switch (L) {
            case 1: down_up_pow2<16, 1>(p0, p1, g, base); break;
            case 2: down_up_pow2<16, 2>(p0, p1, g, base); break;
            case 3: down_up_pow2<16, 3>(p0, p1, g, base); break;
            case 4: down_up_pow2<16, 4>(p0, p1, g, base); break;
            default: break;
        }
    } else {
        switch (L) {
            case 1: down_up_pow2<8, 1>(p0, p1, g, base); break;
            case 2: down_up_pow2<8, 2>(p0, p1, g, base); break;
            case 3: down_up_pow2<8, 3>(p0, p1, g, base); break;
            default: break;
        }
    }
}

}  // namespace elvis
