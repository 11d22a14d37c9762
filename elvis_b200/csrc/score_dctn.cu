// a1 with dct_size = block_size (16 or 32): the transform size the reference asks EVCA for
// (`python -m evca.main ... -b block_size`, elvis.py:1022-1023; EVCAConfig(block_size=bs),
// presley.py:202).  Spec: oracle/spec_scoring.py with n = block_size -- one N x N orthonormal
// DCT-II per block, weights w(u, v) = exp(|(u v / N^2)^2 - 1|), DC excluded.
//
// Mapping.  One warp owns one N x N block and walks it through a run of frames.  The DCT is
// computed separably in fp32 on the CUDA cores, C = A X A^T, with the basis row A[a][.] of a lane
// held in registers:
//   stage 1   lane (a, h) computes Y[r][a] = sum_c X[r][c] A[a][c] for its N / H rows r (H = 32 / N
//             lanes share a column); the pixel rows come from shared memory as broadcast 128-bit
//             loads (every lane of a half reads the same address);
//   stage 2   the same lane folds its rows into all N coefficients of column a,
//             C[u][a] += A[u][r] Y[r][a] with A rows broadcast from shared memory; for N = 16 the
//             two halves are combined with one shuffle per coefficient.
// Y never leaves registers.  C_t is computed afresh for every frame and C_{t-1} is kept in
// registers, so SC / TC do not depend on the temporal chunking.  About 2 N FMA per pixel: an optional
// mode (the benchmarked configuration uses the 8 x 8 transform north_star names, score_umma.cu).
#include "score_params.cuh"

namespace elvis {
namespace {

constexpr int kDctnWarps = 4;

template <int N> struct DctnGeom {
    static constexpr int H = 32 / N;            // lanes per basis row: 2 (N = 16) or 1 (N = 32)
    static constexpr int PER = N / H;           // rows (stage 1) / coefficients (epilogue) per lane
    static constexpr int kPitch = N + 4;        // floats per shared-memory row: keeps 16-byte alignment, spreads the banks
    static constexpr int kRowBytes = N / (32 / N);   // pixel bytes one lane loads per frame: 8 (N = 16) or 32 (N = 32)
};

template <int N>
__global__ void __launch_bounds__(kDctnWarps * 32) score_dctn_kernel(const ScoreParams p) {
    using G = DctnGeom<N>;
    __shared__ __align__(16) float sA[N][G::kPitch];              // orthonormal DCT-II basis A[v][x]
    __shared__ float sW[N][N + 1];                                // weights w[u][v]
    __shared__ __align__(16) float sX[kDctnWarps][N][G::kPitch];  // the warp's block of the current frame
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // tables, once per CTA (double precision, rounded to float)
    for (int i = threadIdx.x; i < N * N; i += kDctnWarps * 32) {
        const int v = i / N, x = i - v * N;
        const double cv = v == 0 ? sqrt(1.0 / N) : sqrt(2.0 / N);
        sA[v][x] = (float)(cv * cospi((double)((2 * x + 1) * v) / (double)(2 * N)));
        const double q = (double)(v * x) / (double)(N * N);       // here (v, x) plays (u, v)
        sW[v][x] = i == 0 ? 0.f : (float)exp(fabs(q * q - 1.0));
    }
    __syncthreads();

    const int64_t unit = (int64_t)blockIdx.x * kDctnWarps + warp;
    const int64_t per_chunk = (int64_t)p.By * p.Bx;
    if (unit >= per_chunk * p.n_chunks) return;
    const int chunk = (int)(unit / per_chunk);
    const int rem = (int)(unit - (int64_t)chunk * per_chunk);
    const int by = rem / p.Bx, bx = rem - by * p.Bx;
    const int a = lane % N, h = lane / N;

    float areg[N];
#pragma unroll
    for (int c = 0; c < N; ++c) areg[c] = sA[a][c];

    const int t0 = chunk * p.chunk_len;
    const int t1 = min(p.T, t0 + p.chunk_len);
    const bool has_prev = (t0 > 0) || (p.halo != nullptr);
    const int t_start = has_prev ? t0 - 1 : t0;

    // pixel loads: lane l covers kRowBytes consecutive bytes of the block
    constexpr int kLanesPerRow = N / G::kRowBytes;               // 2 (N = 16) or 1 (N = 32)
    const int lr = lane / kLanesPerRow, lc = (lane % kLanesPerRow) * G::kRowBytes;
    const int64_t px_off = ((int64_t)by * N + lr) * p.row_stride + (int64_t)bx * N + lc;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(p.y) | (uintptr_t)p.frame_stride | (uintptr_t)p.row_stride |
                          (p.halo ? reinterpret_cast<uintptr_t>(p.halo) : 0)) & (G::kRowBytes == 8 ? 7 : 15)) == 0;
    uint32_t raw[G::kRowBytes / 4], nxt[G::kRowBytes / 4];
    auto load_frame = [&](int t, uint32_t (&dst)[G::kRowBytes / 4]) {
        const uint8_t* src = (t < 0 ? p.halo : p.y + (int64_t)t * p.frame_stride) + px_off;
        if (vec_ok) {
            if (G::kRowBytes == 8) {
                const uint2 v = __ldcs(reinterpret_cast<const uint2*>(src));
                dst[0] = v.x;
                dst[1] = v.y;
            } else {
#pragma unroll
                for (int k = 0; k < G::kRowBytes / 16; ++k) {
                    const uint4 v = __ldcs(reinterpret_cast<const uint4*>(src) + k);
                    dst[4 * k] = v.x;
                    dst[4 * k + 1] = v.y;
                    dst[4 * k + 2] = v.z;
                    dst[4 * k + 3] = v.w;
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < G::kRowBytes / 4; ++k)
                dst[k] = src[4 * k] | (src[4 * k + 1] << 8) | (src[4 * k + 2] << 16) | ((uint32_t)src[4 * k + 3] << 24);
        }
    };

    float cprev[G::PER];
#pragma unroll
    for (int i = 0; i < G::PER; ++i) cprev[i] = 0.f;
    float smin = __int_as_float(0x7f800000), smax = 0.f, tmin = __int_as_float(0x7f800000), tmax = 0.f;
    float (*X)[G::kPitch] = sX[warp];
    const int64_t out_step = (int64_t)p.By * p.Bx;
    float* out_sc = p.sc + ((int64_t)t_start * p.By + by) * p.Bx + bx;
    float* out_tc = p.tc + ((int64_t)t_start * p.By + by) * p.Bx + bx;

    load_frame(t_start, raw);
    for (int t = t_start; t < t1; ++t) {
        if (t + 1 < t1) load_frame(t + 1, nxt);       // in flight while this frame is transformed
        // pixels - 128 -> fp32 rows in shared memory (PRMT builds 2^23 + byte; bias and centering are subtracted
        // exactly; centering only moves the excluded DC coefficient and keeps the fp32 sums small)
#pragma unroll
        for (int k = 0; k < G::kRowBytes / 4; ++k) {
            float4 f;
            f.x = byte_as_biased_float<0>(raw[k], p.magic) - 8388736.f;
            f.y = byte_as_biased_float<1>(raw[k], p.magic) - 8388736.f;
            f.z = byte_as_biased_float<2>(raw[k], p.magic) - 8388736.f;
            f.w = byte_as_biased_float<3>(raw[k], p.magic) - 8388736.f;
            *reinterpret_cast<float4*>(&X[lr][lc + 4 * k]) = f;
        }
        __syncwarp();
        // stage 1: Y[r][a] for this lane's rows
        float y[G::PER];
#pragma unroll
        for (int i = 0; i < G::PER; ++i) {
            const float* row = X[h * G::PER + i];
            float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
            for (int c = 0; c < N; c += 8) {
                const float4 x0 = *reinterpret_cast<const float4*>(row + c);
                const float4 x1 = *reinterpret_cast<const float4*>(row + c + 4);
                acc0 = fmaf(x0.x, areg[c], acc0);
                acc1 = fmaf(x0.y, areg[c + 1], acc1);
                acc0 = fmaf(x0.z, areg[c + 2], acc0);
                acc1 = fmaf(x0.w, areg[c + 3], acc1);
                acc0 = fmaf(x1.x, areg[c + 4], acc0);
                acc1 = fmaf(x1.y, areg[c + 5], acc1);
                acc0 = fmaf(x1.z, areg[c + 6], acc0);
                acc1 = fmaf(x1.w, areg[c + 7], acc1);
            }
            y[i] = acc0 + acc1;
        }
        __syncwarp();                                 // every lane is done with X before the next frame overwrites it
        // stage 2 + epilogue: coefficients C[u][a]; this lane accounts for u in [h PER, (h + 1) PER)
        float s = 0.f, d = 0.f;
#pragma unroll
        for (int u = 0; u < N; ++u) {
            const float* arow = &sA[u][h * G::PER];
            float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
            for (int i = 0; i < G::PER; i += 4) {
                const float4 av = *reinterpret_cast<const float4*>(arow + i);
                acc0 = fmaf(av.x, y[i], acc0);
                acc1 = fmaf(av.y, y[i + 1], acc1);
                acc0 = fmaf(av.z, y[i + 2], acc0);
                acc1 = fmaf(av.w, y[i + 3], acc1);
            }
            float c = acc0 + acc1;
            if (G::H == 2) c += __shfl_xor_sync(0xffffffffu, c, 16);
            if (u / G::PER == h || G::H == 1) {       // compile-time for H = 1; one predicated half for H = 2
                const int i = u % G::PER;
                const float w = sW[u][a];
                s = fmaf(fabsf(c), w, s);
                d = fmaf(fabsf(c - cprev[i]), w, d);
                cprev[i] = c;
            }
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, m);
            d += __shfl_xor_sync(0xffffffffu, d, m);
        }
        if (t >= t0 && lane == 0) {
            const float scv = s * p.inv_area;
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
#pragma unroll
        for (int k = 0; k < G::kRowBytes / 4; ++k) raw[k] = nxt[k];
    }
    if (p.mm != nullptr && lane == 0 && smin <= smax) {
        atomicMin(p.mm + 0, __float_as_uint(smin));   // non-negative floats order like their bit patterns
        atomicMax(p.mm + 1, __float_as_uint(smax));
        atomicMin(p.mm + 2, __float_as_uint(tmin));
        atomicMax(p.mm + 3, __float_as_uint(tmax));
    }
}

}  // namespace

// dct_size == block_size in {16, 32}
int launch_score_dctn(ScoreParams p, int block_size, cudaStream_t st) {
    const int64_t units = (int64_t)p.By * p.Bx * p.n_chunks;
    const unsigned grid = (unsigned)((units + kDctnWarps - 1) / kDctnWarps);
    if (block_size == 16)
        score_dctn_kernel<16><<<grid, kDctnWarps * 32, 0, st>>>(p);
    else if (block_size == 32)
        score_dctn_kernel<32><<<grid, kDctnWarps * 32, 0, st>>>(p);
    else
        return ELVIS_ERR_UNSUPPORTED;
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

}  // namespace elvis
