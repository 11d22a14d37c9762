// a4 / a6: per block-row top-k removal mask, and a13: side-channel bit packers.
//
// Scores become order-preserving 64-bit keys; together with the column index they form a
// strict total order "key, then column", which realises the tie rule of the oracle (lowest
// column first == np.argmin's rule, utils.py:721, and the stable reading of
// np.argsort(-row), elvis.py:1401).  The removed blocks of a row are its k smallest elements
// in that order -- a SELECTION problem, not a sort:
//   * rows of up to 512 blocks: one WARP per (frame, block-row), the row held in registers
//     (32 columns per register slot), randomised quickselect with ballot/shuffle counting --
//     about 2 ln(Bx) rounds of Bx/32 compares per lane (~100 instructions each), no shared memory, no barriers;
//   * longer rows: one CTA per row, shared-memory bitonic sort of (key, column) pairs.
#include "common.cuh"

namespace elvis {
namespace {

// IEEE double -> uint64 whose unsigned order equals the float order (-0 == +0, NaN last)
__device__ __forceinline__ unsigned long long sortable_key(double v) {
    v = __dadd_rn(v, 0.0);   // -0.0 -> +0.0 so that signed zeros tie like they compare
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}

__device__ __forceinline__ unsigned long long row_key(double v, int polarity) {
    // NaN ranks last in either polarity (np.argsort's rule), but ahead of the padding
    return (v != v) ? ~0ULL - 1 : sortable_key(polarity == ELVIS_REMOVE_HIGH ? -v : v);
}

constexpr int select_threads(int n) { return n / 2 > 512 ? 512 : (n / 2 < 32 ? 32 : n / 2); }

// Min-max normalisation folded into the selection (elvis_normalize_select_rows): the row is normalised as it is
// loaded, written back, and ranked on the normalised values -- the same numbers, one pass over the scores less.
struct RowNormalizer {
    double lo;
    InvariantDivisor span;
    bool on;
    __device__ explicit RowNormalizer(const double* mm) : lo(mm ? mm[0] : 0.0), span(mm ? mm[1] - mm[0] : 1.0), on(mm && mm[1] > mm[0]) {}
    __device__ __forceinline__ double load(double* p) const {
        double v = *p;
        if (on) {
            v = span.divide(__dsub_rn(v, lo));
            *p = v;
        }
        return v;
    }
};

template <int N>   // N = padded row length (power of two)
__global__ void __launch_bounds__(select_threads(N)) select_rows_kernel(
        double* __restrict__ scores, const double* __restrict__ minmax, int by, int bx, const int32_t* __restrict__ k_per_row, int k_uniform,
        int polarity, uint8_t* __restrict__ mask) {
    __shared__ unsigned long long s_key[N];
    __shared__ uint16_t s_col[N];
    const RowNormalizer norm(minmax);
    const int64_t row = blockIdx.x;
    const int k = k_per_row ? k_per_row[row % by] : k_uniform;
    double* src = scores + row * bx;
    uint8_t* dst = mask + row * bx;
    if (k <= 0 || k >= bx) {   // nothing or everything removed: no ranking needed
        for (int i = threadIdx.x; i < bx; i += blockDim.x) {
            norm.load(src + i);
            dst[i] = k >= bx ? 1 : 0;
        }
        return;
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        unsigned long long key = ~0ULL;
        if (i < bx) {
            key = row_key(norm.load(src + i), polarity);
        }
        s_key[i] = key;
        s_col[i] = (uint16_t)i;
    }
    __syncthreads();
    for (int size = 2; size <= N; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < N / 2; t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool ascending = (lo & size) == 0;
                const unsigned long long ka = s_key[lo], kb = s_key[hi];
                const uint16_t ca = s_col[lo], cb = s_col[hi];
                const bool a_after_b = ka > kb || (ka == kb && ca > cb);
                if (a_after_b == ascending) {
                    s_key[lo] = kb;
                    s_key[hi] = ka;
                    s_col[lo] = cb;
                    s_col[hi] = ca;
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int c = s_col[i];
        if (c < bx) dst[c] = i < k ? 1 : 0;
    }
}


// One warp per row; slot j of lane l holds column j*32 + l.
template <int E>
__global__ void __launch_bounds__(256) select_rows_warp_kernel(
        double* __restrict__ scores, const double* __restrict__ minmax, int64_t rows, int by, int bx,
        const int32_t* __restrict__ k_per_row, int k_uniform, int polarity, uint8_t* __restrict__ mask) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const RowNormalizer norm(minmax);
    const int k = k_per_row ? k_per_row[row % by] : k_uniform;
    double* src = scores + row * bx;
    uint8_t* dst = mask + row * bx;

    unsigned long long key[E];
    unsigned valid = 0;
#pragma unroll
    for (int j = 0; j < E; ++j) {
        const int c = j * 32 + lane;
        key[j] = ~0ULL;
        if (c < bx) {
            key[j] = row_key(norm.load(src + c), polarity);
            valid |= 1u << j;
        }
    }
    unsigned selected = 0;
    if (k >= bx) {
        selected = valid;
    } else if (k > 0) {
        // Invariant: the k smallest are `selected` plus the `need` smallest of the `total` candidates in `active`.
        unsigned active = valid;
        int need = k, total = bx;
        unsigned h = (unsigned)row * 2654435761u + 0x7F4A7C15u;
        while (total > need) {
            // pivot: a pseudo-random (reproducible) candidate -- the first lane at or after a random position that
            // still has candidates, and in it the first candidate slot at or after a random slot
            h = h * 1664525u + 1013904223u;
            const unsigned has = __ballot_sync(0xffffffffu, active != 0);
            const int s1 = (int)(h >> 27);
            const int src_lane = (__ffs(__funnelshift_r(has, has, s1)) - 1 + s1) & 31;
            const int s2 = (int)((h >> 20) & (E - 1));
            int pslot = (__ffs((active | (active << E)) >> s2) - 1 + s2) & (E - 1);
            unsigned long long pick[E / 2];
#pragma unroll
            for (int j = 0; j < E / 2; ++j) pick[j] = (pslot & (E / 2)) ? key[j + E / 2] : key[j];
#pragma unroll
            for (int w = E / 4; w >= 1; w >>= 1)
#pragma unroll
                for (int j = 0; j < w; ++j) pick[j] = (pslot & w) ? pick[j + w] : pick[j];
            const unsigned long long pk = __shfl_sync(0xffffffffu, pick[0], src_lane);
            pslot = __shfl_sync(0xffffffffu, pslot, src_lane);
            // candidates at or before the pivot in (key, column) order: key < pk + [column <= pivot column]
            // (a pivot is never the padding key ~0, so pk + 1 does not wrap)
            const unsigned col_le = ((1u << pslot) - 1u) | ((unsigned)(lane <= src_lane) << pslot);
            const unsigned long long pk1 = pk + 1;
            unsigned le = 0;
#pragma unroll
            for (int j = 0; j < E; ++j) le |= (unsigned)(key[j] < (((col_le >> j) & 1u) ? pk1 : pk)) << j;
            le &= active;
            const int c_le = __reduce_add_sync(0xffffffffu, __popc(le));
            if (c_le <= need) {          // all of them belong to the k smallest
                selected |= le;
                need -= c_le;
                total -= c_le;
                active &= ~le;
            } else {                     // the k-th smallest lies strictly before the pivot
                active = le;
                if (lane == src_lane) active &= ~(1u << pslot);
                total = c_le - 1;
            }
        }
        if (need > 0) selected |= active;    // total == need: every remaining candidate is removed
    }
#pragma unroll
    for (int j = 0; j < E; ++j) {
        const int c = j * 32 + lane;
        if (c < bx) dst[c] = (selected >> j) & 1u;
    }
}

template <int E>
int launch_select_warp(double* scores, const double* mm, int64_t rows, int by, int bx, const int32_t* kpr, int ku, int pol, uint8_t* mask, cudaStream_t st) {
    const unsigned grid = (unsigned)((rows + 7) / 8);
    select_rows_warp_kernel<E><<<grid, 256, 0, st>>>(scores, mm, rows, by, bx, kpr, ku, pol, mask);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

// np.packbits over a flat array: bit 7 of byte j is value 8j
__global__ void __launch_bounds__(256) pack_bits_kernel(const uint8_t* __restrict__ m, int64_t n, uint8_t* __restrict__ out) {
    const int64_t nbytes = (n + 7) / 8;
    for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < nbytes; j += (int64_t)gridDim.x * 256) {
        unsigned b = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int64_t i = j * 8 + q;
            if (i < n && m[i]) b |= 0x80u >> q;
        }
        out[j] = (uint8_t)b;
    }
}

__global__ void __launch_bounds__(256) unpack_bits_kernel(const uint8_t* __restrict__ p, int64_t n, uint8_t* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
        out[i] = (p[i >> 3] >> (7 - (i & 7))) & 1;
}

__global__ void __launch_bounds__(256) pack2_kernel(const int32_t* __restrict__ lv, int64_t rows, int bx, uint8_t* __restrict__ out) {
    const int pb = (bx + 3) / 4;
    const int64_t total = rows * pb;
    for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < total; j += (int64_t)gridDim.x * 256) {
        const int64_t r = j / pb;
        const int c = (int)(j - r * pb) * 4;
        unsigned b = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (c + q < bx) b |= (unsigned)min(max(lv[r * bx + c + q], 0), 3) << (2 * q);   // saturating: level 4 (16x at bs 16) is kept as 8x
        out[j] = (uint8_t)b;
    }
}

__global__ void __launch_bounds__(256) unpack2_kernel(const uint8_t* __restrict__ p, int64_t rows, int bx, int32_t* __restrict__ out) {
    const int pb = (bx + 3) / 4;
    const int64_t total = rows * bx;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int64_t r = i / bx;
        const int c = (int)(i - r * bx);
        out[i] = (p[r * pb + (c >> 2)] >> (2 * (c & 3))) & 3;
    }
}

inline int grid_for(int64_t n) {
    int64_t g = (n + 255) / 256;
    const int64_t cap = (int64_t)kNumSMs * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

template <int N>
int launch_select(double* scores, const double* mm, int64_t rows, int by, int bx, const int32_t* kpr, int ku, int pol, uint8_t* mask, cudaStream_t st) {
    select_rows_kernel<N><<<(unsigned)rows, select_threads(N), 0, st>>>(scores, mm, by, bx, kpr, ku, pol, mask);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

}  // namespace
}  // namespace elvis

using namespace elvis;

namespace {
// scores are written only when minmax is given
int select_rows(double* scores, const double* minmax, int32_t n_frames, int32_t by, int32_t bx, const int32_t* k_per_row,
                int32_t k_uniform, int32_t polarity, uint8_t* mask, elvis_stream_t stream) {
    if (!scores || !mask || n_frames <= 0 || by <= 0 || bx <= 0) return ELVIS_ERR_INVALID_ARG;
    if (polarity != ELVIS_REMOVE_HIGH && polarity != ELVIS_REMOVE_LOW) return ELVIS_ERR_INVALID_ARG;
    if (bx > 4096) return ELVIS_ERR_UNSUPPORTED;
    const int64_t rows = (int64_t)n_frames * by;
    cudaStream_t st = as_stream(stream);
    const bool force_sort = getenv("ELVIS_SELECT_SORT") != nullptr;   // test hook: exercise the CTA path on short rows
    if (!force_sort) {
        if (bx <= 64) return launch_select_warp<2>(scores, minmax, rows, by, bx, k_per_row, k_uniform, polarity, mask, st);
        if (bx <= 128) return launch_select_warp<4>(scores, minmax, rows, by, bx, k_per_row, k_uniform, polarity, mask, st);
        if (bx <= 256) return launch_select_warp<8>(scores, minmax, rows, by, bx, k_per_row, k_uniform, polarity, mask, st);
        if (bx <= 512) return launch_select_warp<16>(scores, minmax, rows, by, bx, k_per_row, k_uniform, polarity, mask, st);
    }
    if (bx <= 64) return launch_select<64>(scores, minmax, rows, by, bx, k_per_row, k_uniform, polarity, mask, st);
    if (bx <= 128) return launch_select<128>(scores, minmax, rows, by, bx, k_per_row, k_uniform, polarity, mask, st);
    if (bx <= 256) return launch_select<256>(scores, minmax, rows, by, bx, k_per_row, k_uniform, polarity, mask, st);
    if (bx <= 512) return launch_select<512>(scores, minmax, rows, by, bx, k_per_row, k_uniform, polarity, mask, st);
    if (bx <= 1024) return launch_select<1024>(scores, minmax, rows, by, bx, k_per_row, k_uniform, polarity, mask, st);
    if (bx <= 2048) return launch_select<2048>(scores, minmax, rows, by, bx, k_per_row, k_uniform, polarity, mask, st);
    return launch_select<4096>(scores, minmax, rows, by, bx, k_per_row, k_uniform, polarity, mask, st);
}
}  // namespace

extern "C" int elvis_select_rows(const double* scores, int32_t n_frames, int32_t by, int32_t bx,
                                 const int32_t* k_per_row, int32_t k_uniform, int32_t polarity,
                                 uint8_t* mask, elvis_stream_t stream) {
    return select_rows(const_cast<double*>(scores), nullptr, n_frames, by, bx, k_per_row, k_uniform, polarity, mask, stream);
}

extern "C" int elvis_normalize_select_rows(double* scores, const double* minmax, int32_t n_frames, int32_t by, int32_t bx,
                                           const int32_t* k_per_row, int32_t k_uniform, int32_t polarity,
                                           uint8_t* mask, elvis_stream_t stream) {
    if (!minmax) return ELVIS_ERR_INVALID_ARG;
    return select_rows(scores, minmax, n_frames, by, bx, k_per_row, k_uniform, polarity, mask, stream);
}

extern "C" int elvis_pack_mask_bits(const uint8_t* mask, int64_t n, uint8_t* packed, elvis_stream_t stream) {
    if (!mask || !packed || n <= 0) return ELVIS_ERR_INVALID_ARG;
    pack_bits_kernel<<<grid_for((n + 7) / 8), 256, 0, as_stream(stream)>>>(mask, n, packed);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_unpack_mask_bits(const uint8_t* packed, int64_t n, uint8_t* mask, elvis_stream_t stream) {
    if (!mask || !packed || n <= 0) return ELVIS_ERR_INVALID_ARG;
    unpack_bits_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(packed, n, mask);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_pack_levels_2bit(const int32_t* levels, int64_t rows, int32_t bx, uint8_t* packed, elvis_stream_t stream) {
    if (!levels || !packed || rows <= 0 || bx <= 0) return ELVIS_ERR_INVALID_ARG;
    pack2_kernel<<<grid_for(rows * ((bx + 3) / 4)), 256, 0, as_stream(stream)>>>(levels, rows, bx, packed);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_unpack_levels_2bit(const uint8_t* packed, int64_t rows, int32_t bx, int32_t* levels, elvis_stream_t stream) {
    if (!levels || !packed || rows <= 0 || bx <= 0) return ELVIS_ERR_INVALID_ARG;
    unpack2_kernel<<<grid_for(rows * bx), 256, 0, as_stream(stream)>>>(packed, rows, bx, levels);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}
