// a1: per-block spatial (SC) and temporal (TC) DCT-energy features -- the arithmetic the
// reference delegates to the external EVCA package (elvis.py:1014-1031, presley.py:202).
// Spec: oracle/spec_scoring.py.  This file holds the entry point elvis_score_sc_tc, which picks
// the kernel, and the CUDA-core kernel: the path for small clips, planes that are not 16-byte
// aligned and drivers without tensor maps.  Large aligned clips go to the tcgen05 kernel in
// score_umma.cu; score_mma.cu is the earlier mma.sync attempt, kept selectable.
//
// Mapping.  One thread owns one 8x8 luma tile and walks it through a run of consecutive
// frames; the tile never leaves registers, so there is no shared-memory staging and no
// transposition.  A warp covers R x (32/R) tiles (R = block_size / 8), i.e. one block row of
// 32/R^2 blocks, so that every row load of the warp is R fully used 128-byte lines and the
// per-block sum is an R*R-lane xor-shuffle reduction.
//
// Temporal streaming.  DCT is linear, so C_t = C_{t-1} + DCT(Y_t - Y_{t-1}).  Each frame costs
// ONE 8x8 transform (of the exact integer frame difference): its weighted magnitude is TC, and
// the running sum of the differences is C_t, whose weighted magnitude is SC.  Transforming the
// small integer difference keeps TC accurate to fp32 rounding of *its own* magnitude
// (|C_t| - |C_{t-1}| computed from two large fp32 coefficients would not).  Every luma byte is
// read from HBM once per chunk of frames; a chunk primes its accumulator with one extra
// transform of the frame before it.
#include "score_params.cuh"
#include "dct8.cuh"
#include "dct8_packed.cuh"
#include <cstring>

namespace elvis {

namespace {

__device__ constexpr float kW[8][8] = {
#include "score_weights.inc"
};

constexpr int kScoreThreads = 128;
#ifndef ELVIS_SCORE_MIN_CTAS
#define ELVIS_SCORE_MIN_CTAS 2
#endif
constexpr int kScoreMinCtas = ELVIS_SCORE_MIN_CTAS;   // resident CTAs per SM the register budget is sized for
#ifndef ELVIS_PACKED_PAIRS
#define ELVIS_PACKED_PAIRS 4
#endif
constexpr int kPackedColumnPairs = ELVIS_PACKED_PAIRS;   // column pairs of the vertical pass done with FADD2/FFMA2
constexpr int kRing = 3;   // cp.async ring depth: two frames in flight ahead of the one being transformed

template <bool ALIGNED>
__device__ __forceinline__ void load_tile(uint2 (&dst)[8], const uint8_t* p, int64_t row_stride, bool valid) {
    if (!valid) {
#pragma unroll
        for (int r = 0; r < 8; ++r) dst[r] = make_uint2(0u, 0u);
        return;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const uint8_t* q = p + r * row_stride;
        if (ALIGNED) {
            dst[r] = __ldcs(reinterpret_cast<const uint2*>(q));
        } else {
            uint32_t lo = q[0] | (q[1] << 8) | (q[2] << 16) | ((uint32_t)q[3] << 24);
            uint32_t hi = q[4] | (q[5] << 8) | (q[6] << 16) | ((uint32_t)q[7] << 24);
            dst[r] = make_uint2(lo, hi);
        }
    }
}

template <int R, bool ALIGNED>
__global__ void __launch_bounds__(kScoreThreads, kScoreMinCtas) score_kernel(const ScoreParams p) {
    constexpr int TW = 32 / R;   // tiles per warp-tile row
    const int lane = threadIdx.x & 31;
    const int unit = blockIdx.x * (kScoreThreads / 32) + (threadIdx.x >> 5);
    const int per_chunk = p.By * p.tiles_x;
    if (unit >= per_chunk * p.n_chunks) return;
    const int chunk = unit / per_chunk;
    const int rem = unit - chunk * per_chunk;
    const int by = rem / p.tiles_x;
    const int tx = rem - by * p.tiles_x;

    const int tr = lane / TW, tcx = lane % TW;
    const int tile_col = tx * TW + tcx;
    const bool valid = tile_col < p.Bx * R;
    const bool leader = valid && tr == 0 && (tcx % R) == 0;
    const int bxi = tile_col / R;
    const int64_t tile_off = (int64_t)(by * R + tr) * 8 * p.row_stride + (int64_t)tile_col * 8;

    const int t0 = chunk * p.chunk_len;
    const int t1 = min(p.T, t0 + p.chunk_len);
    const bool has_prev = (t0 > 0) || (p.halo != nullptr);
    const int t_start = has_prev ? t0 - 1 : t0;

    auto frame_ptr = [&](int t) -> const uint8_t* {
        return (t < 0 ? p.halo : p.y + (int64_t)t * p.frame_stride) + tile_off;
    };

    // running coefficients C_t, packed over column pairs: acc[u][j] = (C[u][2j], C[u][2j+1])
    float2 acc[8][4];
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[u][j] = make_float2(0.f, 0.f);

    // Luma rows reach the thread through a private 3-slot ring in shared memory filled with
    // cp.async two frames ahead.  (Prefetching into registers does not survive ptxas: it sinks
    // the loads to the end of the loop body to save registers, which leaves one frame of
    // loads in flight per warp and the kernel latency bound -- profiles/r1b.)  Every thread
    // reads back only what it copied itself, so no barrier is needed, only wait_group.
    __shared__ uint2 s_ring[kRing][8][kScoreThreads];
    uint2 cur[8], prv[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) prv[r] = make_uint2(0u, 0u);
    auto prefetch = [&](int slot, int t) {
        if (ALIGNED && valid && t < t1) {
            const uint8_t* src = frame_ptr(t);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s_ring[slot][r][threadIdx.x]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src + r * p.row_stride) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");   // always: keeps the group count uniform
    };
    if (ALIGNED) {
        prefetch(0, t_start);
        prefetch(1, t_start + 1);
    }

    float smin = __int_as_float(0x7f800000), smax = 0.f, tmin = __int_as_float(0x7f800000), tmax = 0.f;
    const uint32_t magic = p.magic;

    for (int t = t_start, it = 0; t < t1; ++t, ++it) {
        if (ALIGNED) {
            prefetch((it + 2) % kRing, t + 2);
            asm volatile("cp.async.wait_group 2;" ::: "memory");
            const int slot = it % kRing;
#pragma unroll
            for (int r = 0; r < 8; ++r) cur[r] = valid ? s_ring[slot][r][threadIdx.x] : make_uint2(0u, 0u);
        } else {
            load_tile<false>(cur, frame_ptr(t), p.row_stride, valid);
        }

        // frame difference, packed over ROW pairs: y[i][c] = (d[2i][c], d[2i+1][c]).  The 2^23
        // bias of the byte->float trick cancels in the subtraction.
        float2 y[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint2 ca = cur[2 * i], cb = cur[2 * i + 1], pa = prv[2 * i], pb = prv[2 * i + 1];
#define ELVIS_DIFF(C, K, W)                                                                                   \
    y[i][C] = f2sub(make_float2(byte_as_biased_float<K>(ca.W, magic), byte_as_biased_float<K>(cb.W, magic)),  \
                    make_float2(byte_as_biased_float<K>(pa.W, magic), byte_as_biased_float<K>(pb.W, magic)))
            ELVIS_DIFF(0, 0, x); ELVIS_DIFF(1, 1, x); ELVIS_DIFF(2, 2, x); ELVIS_DIFF(3, 3, x);
            ELVIS_DIFF(4, 0, y); ELVIS_DIFF(5, 1, y); ELVIS_DIFF(6, 2, y); ELVIS_DIFF(7, 3, y);
#undef ELVIS_DIFF
        }
        // horizontal pass: SCALAR butterflies (they can issue to both FMA sub-pipes, whereas the
        // packed instructions only run on the heavy one), written straight into the column-pair
        // registers of the vertical pass -- no register transposition in between
        float2 x[8][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float a0 = y[i][0].x, a1 = y[i][1].x, a2 = y[i][2].x, a3 = y[i][3].x, a4 = y[i][4].x, a5 = y[i][5].x, a6 = y[i][6].x, a7 = y[i][7].x;
            ELVIS_FDCT8(a0, a1, a2, a3, a4, a5, a6, a7);
            x[2 * i][0] = make_float2(a0, a1); x[2 * i][1] = make_float2(a2, a3);
            x[2 * i][2] = make_float2(a4, a5); x[2 * i][3] = make_float2(a6, a7);
            float b0 = y[i][0].y, b1 = y[i][1].y, b2 = y[i][2].y, b3 = y[i][3].y, b4 = y[i][4].y, b5 = y[i][5].y, b6 = y[i][6].y, b7 = y[i][7].y;
            ELVIS_FDCT8(b0, b1, b2, b3, b4, b5, b6, b7);
            x[2 * i + 1][0] = make_float2(b0, b1); x[2 * i + 1][1] = make_float2(b2, b3);
            x[2 * i + 1][2] = make_float2(b4, b5); x[2 * i + 1][3] = make_float2(b6, b7);
        }
        // vertical pass: two columns per instruction.  x[u][j] = (dC[u][2j], dC[u][2j+1]), AAN-scaled
#pragma unroll
        for (int j = 0; j < kPackedColumnPairs; ++j) ELVIS_FDCT8_X2(x[0][j], x[1][j], x[2][j], x[3][j], x[4][j], x[5][j], x[6][j], x[7][j]);
#pragma unroll
        for (int j = kPackedColumnPairs; j < 4; ++j) {   // remaining columns: scalar (both FMA sub-pipes)
            ELVIS_FDCT8(x[0][j].x, x[1][j].x, x[2][j].x, x[3][j].x, x[4][j].x, x[5][j].x, x[6][j].x, x[7][j].x);
            ELVIS_FDCT8(x[0][j].y, x[1][j].y, x[2][j].y, x[3][j].y, x[4][j].y, x[5][j].y, x[6][j].y, x[7][j].y);
        }

        // per-row partial sums keep 16 independent FMA chains in flight; fixed order => deterministic
        float s_part[8], d_part[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            float s = 0.f, d = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[u][j] = f2add(acc[u][j], x[u][j]);
                if (u != 0 || j != 0) {   // DC (u = v = 0) carries no texture energy
                    s = fmaf(fabsf(acc[u][j].x), kW[u][2 * j], s);
                    d = fmaf(fabsf(x[u][j].x), kW[u][2 * j], d);
                }
                s = fmaf(fabsf(acc[u][j].y), kW[u][2 * j + 1], s);
                d = fmaf(fabsf(x[u][j].y), kW[u][2 * j + 1], d);
            }
            s_part[u] = s;
            d_part[u] = d;
        }
        float s = ((s_part[0] + s_part[1]) + (s_part[2] + s_part[3])) + ((s_part[4] + s_part[5]) + (s_part[6] + s_part[7]));
        float d = ((d_part[0] + d_part[1]) + (d_part[2] + d_part[3])) + ((d_part[4] + d_part[5]) + (d_part[6] + d_part[7]));

        // sum the block's R x R tiles
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
#pragma unroll
        for (int r = 0; r < 8; ++r) prv[r] = cur[r];
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
            // non-negative floats order like their bit patterns
            atomicMin(p.mm + 0, __float_as_uint(smin));
            atomicMax(p.mm + 1, __float_as_uint(smax));
            atomicMin(p.mm + 2, __float_as_uint(tmin));
            atomicMax(p.mm + 3, __float_as_uint(tmax));
        }
    }
}

__global__ void score_minmax_init(unsigned* mm) {
    if (threadIdx.x < 4) mm[threadIdx.x] = (threadIdx.x & 1) ? 0u : 0x7f800000u;
}

template <int R>
int launch_score(const ScoreParams& p, bool aligned, cudaStream_t st) {
    const int units = p.By * p.tiles_x * p.n_chunks;
    const int grid = (units + (kScoreThreads / 32) - 1) / (kScoreThreads / 32);
    if (aligned)
        score_kernel<R, true><<<grid, kScoreThreads, 0, st>>>(p);
    else
        score_kernel<R, false><<<grid, kScoreThreads, 0, st>>>(p);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

}  // namespace

int launch_score_simt(ScoreParams p, int block_size, bool aligned8, cudaStream_t st) {
    const int R = block_size / 8;
    const int TW = 32 / R;
    p.tiles_x = (p.Bx * R + TW - 1) / TW;
    p.tiles_y = p.By;
    switch (R) {
        case 1: return launch_score<1>(p, aligned8, st);
        case 2: return launch_score<2>(p, aligned8, st);
        default: return launch_score<4>(p, aligned8, st);
    }
}

// chunking heuristic: enough (spatial tile x chunk) work units for >= 6 waves of resident
// workers (measured on B200: 60-frame chunks beat 24-frame ones by 3 %, each chunk pays one
// extra priming transform), and chunks no shorter than 16 frames.
static int pick_chunks(int T, long tiles, long resident, int override_len) {
    if (override_len > 0) return (T + override_len - 1) / override_len;
    const long target = 6L * resident;
    long n = (target + tiles - 1) / tiles;
    const long max_chunks = (T + 15) / 16;
    if (n > max_chunks) n = max_chunks;
    if (n < 1) n = 1;
    return (int)n;
}

}  // namespace elvis

extern "C" int elvis_score_sc_tc(const elvis_plane* y, int32_t n_frames, const uint8_t* prev_halo,
                                 int32_t block_size, int32_t dct_size, float* sc, float* tc,
                                 float* minmax, int32_t mm_begin, int32_t mm_end, elvis_stream_t stream) {
    using namespace elvis;
    if (!plane_ok(y) || !sc || !tc || n_frames <= 0) return ELVIS_ERR_INVALID_ARG;
    if (y->channels != 1) return ELVIS_ERR_UNSUPPORTED;
    if (block_size != 8 && block_size != 16 && block_size != 32) return ELVIS_ERR_UNSUPPORTED;
    if (dct_size != 8 && dct_size != block_size) return ELVIS_ERR_UNSUPPORTED;   // 8 x 8 tiles, or one transform per block
    const int By = y->height / block_size, Bx = y->width / block_size;
    if (By <= 0 || Bx <= 0) return ELVIS_ERR_SHAPE;

    ScoreParams p;
    p.y = static_cast<const uint8_t*>(y->data);
    p.halo = prev_halo;
    p.frame_stride = y->frame_stride;
    p.row_stride = y->row_stride;
    p.T = n_frames;
    p.By = By;
    p.Bx = Bx;
    p.sc = sc;
    p.tc = tc;
    p.mm = reinterpret_cast<unsigned*>(minmax);
    p.mm_begin = mm_begin;
    p.mm_end = mm_end;
    p.inv_area = 1.0f / (float)(block_size * block_size);
    p.magic = 0x4B000000u;
    p.magic16 = 0x64006400u;

    // Implementation choice (profiles/).  Default: the tcgen05 kernel (score_umma.cu:
    // TMA ring, A operand and accumulators in tensor memory) when plane, strides and halo are
    // 16-byte aligned and the clip is big enough to fill the machine; otherwise the CUDA-core
    // kernel of this file (packed-fp32 butterflies, cp.async ring; any alignment).
    // ELVIS_SCORE_IMPL = umma | simt | mma | tma forces one (mma / tma: the legacy mma.sync kernel
    // with direct loads / with a TMA ring, 16x16 blocks only); the tests exercise every path.
    auto al = [&](int a) {
        return aligned_to(p.y, a) && p.frame_stride % a == 0 && p.row_stride % a == 0 && (!prev_halo || aligned_to(prev_halo, a));
    };
    if (dct_size != 8) {       // one block_size x block_size transform per block (score_dctn.cu)
        p.n_chunks = pick_chunks(n_frames, (long)By * Bx, (long)kNumSMs * 16, getenv("ELVIS_SCORE_CHUNK") ? atoi(getenv("ELVIS_SCORE_CHUNK")) : 0);
        p.chunk_len = (n_frames + p.n_chunks - 1) / p.n_chunks;
        p.n_chunks = (n_frames + p.chunk_len - 1) / p.chunk_len;
        cudaStream_t st = as_stream(stream);
        if (minmax) {
            score_minmax_init<<<1, 32, 0, st>>>(p.mm);
            ELVIS_CHECK_LAUNCH();
        }
        // 16 x 16: tensor-core kernel (score_dct16.cu) when the planes are TMA-capable; ELVIS_SCORE_DCT16=simt forces the CUDA cores
        const char* e16 = getenv("ELVIS_SCORE_DCT16");
        if (block_size == 16 && al(16) && !(e16 && !strcmp(e16, "simt"))) {
            ScoreParams q = p;
            q.n_chunks = pick_chunks(n_frames, (long)By * ((Bx + 7) / 8), (long)kNumSMs * 3, getenv("ELVIS_SCORE_CHUNK") ? atoi(getenv("ELVIS_SCORE_CHUNK")) : 0);
            q.chunk_len = (n_frames + q.n_chunks - 1) / q.n_chunks;
            q.n_chunks = (n_frames + q.chunk_len - 1) / q.chunk_len;
            const int rc = launch_score_dct16(q, y->height, y->width, st);
            if (rc != ELVIS_ERR_UNSUPPORTED) return rc;
        }
        return launch_score_dctn(p, block_size, st);
    }
    enum { SIMT, MMA_DIRECT, MMA_TMA, UMMA } impl = SIMT;
    const long warp_units = (long)By * ((Bx * (block_size / 8) * (block_size / 8) + 31) / 32);
    if (al(16) && warp_units * n_frames >= 8L * kNumSMs * 4) impl = UMMA;
    if (const char* e = getenv("ELVIS_SCORE_IMPL")) {
        if (!strcmp(e, "mma") && block_size == 16 && al(4)) impl = MMA_DIRECT;
        else if (!strcmp(e, "tma") && block_size == 16 && al(16)) impl = MMA_TMA;
        else if (!strcmp(e, "umma") && al(16)) impl = UMMA;
        else if (!strcmp(e, "simt")) impl = SIMT;
    }
    int override_len = 0;
    if (const char* e = getenv("ELVIS_SCORE_CHUNK")) override_len = atoi(e);
    long tiles, resident;
    if (impl == SIMT) {
        const int R = block_size / 8;
        tiles = (long)By * ((Bx * R * R + 31) / 32);
        resident = (long)kNumSMs * 8;               // warps
    } else if (impl == UMMA) {
        const int R = block_size / 8;
        const int upc = score_umma_units_per_cta();
        tiles = ((long)By * ((Bx * R * R + 31) / 32) + upc - 1) / upc;
        resident = (long)kNumSMs * score_umma_ctas_per_sm();
    } else {
        tiles = (long)((Bx + 7) / 8) * ((By + 2) / 3);
        resident = (long)kNumSMs * 2;               // CTAs
    }
    p.n_chunks = pick_chunks(n_frames, tiles, resident, override_len);
    p.chunk_len = (n_frames + p.n_chunks - 1) / p.n_chunks;
    p.n_chunks = (n_frames + p.chunk_len - 1) / p.chunk_len;

    cudaStream_t st = as_stream(stream);
    if (minmax) {
        score_minmax_init<<<1, 32, 0, st>>>(p.mm);
        ELVIS_CHECK_LAUNCH();
    }
    if (impl == SIMT) return launch_score_simt(p, block_size, al(8), st);
    if (impl == UMMA) {
        const int rc = launch_score_umma(p, block_size, y->height, y->width, st);
        if (rc != ELVIS_ERR_UNSUPPORTED) return rc;
        return launch_score_simt(p, block_size, al(8), st);
    }
    return launch_score_mma(p, y->height, y->width, impl == MMA_TMA, st);
}
