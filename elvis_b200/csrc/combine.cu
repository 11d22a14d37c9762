// a2 / a3 / level maps: block-level float64 arithmetic (tens of thousands of values per
// frame).  Every operation is an explicitly rounded IEEE double op (__dmul_rn, __dadd_rn,
// __dsub_rn, __ddiv_rn -- never contracted into FMAs) issued in the order NumPy evaluates
// the reference expressions, so that given the same SC/TC inputs the scores are BIT-EXACT
// against elvis.py:1173-1220 and utils.py:665-688 -- which is what makes the removal masks
// reproducible.
#include "common.cuh"

namespace elvis {
namespace {

constexpr int kThreads = 256;

template <typename T> __device__ __forceinline__ double ld(const void* p, int64_t i) {
    return (double)static_cast<const T*>(p)[i];
}

__device__ __forceinline__ void atomic_min_f64(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (v < __longlong_as_double((long long)old)) {
        unsigned long long assumed = old;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ void atomic_max_f64(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (v > __longlong_as_double((long long)old)) {
        unsigned long long assumed = old;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}

// NumPy's min/max propagate NaN; scores are finite by contract, NaNs are ignored here.
__device__ __forceinline__ void block_minmax_commit(double lo, double hi, double* out) {
    __shared__ double s_lo[kThreads / 32], s_hi[kThreads / 32];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, m));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, m));
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        s_lo[w] = lo;
        s_hi[w] = hi;
    }
    __syncthreads();
    if (w == 0) {
        lo = l < kThreads / 32 ? s_lo[l] : __longlong_as_double(0x7ff0000000000000LL);
        hi = l < kThreads / 32 ? s_hi[l] : __longlong_as_double(0xfff0000000000000LL);
#pragma unroll
        for (int m = 4; m > 0; m >>= 1) {
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, m));
            hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, m));
        }
        if (l == 0) {
            atomic_min_f64(out + 0, lo);
            atomic_max_f64(out + 1, hi);
        }
    }
}

__global__ void minmax_init(double* out) {
    if (threadIdx.x == 0) {
        out[0] = __longlong_as_double(0x7ff0000000000000LL);
        out[1] = __longlong_as_double(0xfff0000000000000LL);
    }
}

template <typename T> __global__ void __launch_bounds__(kThreads) minmax_kernel(const T* x, int64_t n, double* out) {
    double lo = __longlong_as_double(0x7ff0000000000000LL), hi = __longlong_as_double(0xfff0000000000000LL);
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const double v = (double)x[i];
        lo = fmin(lo, v);
        hi = fmax(hi, v);
    }
    block_minmax_commit(lo, hi, out);
}

// normalize_array applied to one value (elvis.py:864-867); sp = InvariantDivisor(hi - lo)
__device__ __forceinline__ double norm01(double v, double lo, double hi, const InvariantDivisor& sp) {
    return hi > lo ? sp.divide(__dsub_rn(v, lo)) : v;
}

struct CombineParams {
    const void* sc;
    const void* tc;
    const void* norm;
    const uint8_t* background;
    int32_t t_begin, t_count, is_first, is_last, smooth, chunk_len;
    int64_t frame;   // By * Bx
    double alpha, one_minus_alpha, beta, one_minus_beta;
    double* out;
    double* out_minmax;
};

struct NormRange {
    double lo, hi;
    InvariantDivisor span;
};
__device__ __forceinline__ NormRange norm_range(double lo, double hi) { return NormRange{lo, hi, InvariantDivisor(__dsub_rn(hi, lo))}; }

// un-smoothed removability of one block from its raw SC (frame t) and TC (frame t + 1) values
__device__ __forceinline__ double removability_of(const CombineParams& p, double sc_raw, double tc_raw, bool last_frame, bool in_background,
                                                  const NormRange& scn, const NormRange& tcn) {
    const double s = norm01(sc_raw, scn.lo, scn.hi, scn.span);
    double r;
    if (last_frame) {
        r = s;                                                       // elvis.py:1183
    } else {
        const double tn = norm01(tc_raw, tcn.lo, tcn.hi, tcn.span);
        r = __dadd_rn(__dmul_rn(p.alpha, s), __dmul_rn(p.one_minus_alpha, tn));   // elvis.py:1180
    }
    if (in_background) r = __dmul_rn(r, 10.0);                       // elvis.py:1195
    return r;
}

// raw inputs of one block in frame t (extended-range index): SC(t), TC(t + 1) unless t is the clip's last frame, background flag
template <typename T> struct RawBlock {
    T sc, tc;
    bool bg;
    __device__ __forceinline__ void load(const CombineParams& p, int t, int64_t i, bool last_frame) {
        const int64_t o = (int64_t)t * p.frame + i;
        sc = static_cast<const T*>(p.sc)[o];
        tc = last_frame ? T(0) : static_cast<const T*>(p.tc)[o + p.frame];
        bg = p.background && p.background[o];
    }
};

// One thread owns one block position and walks a chunk of consecutive frames, so that the
// un-smoothed value of frame t is computed once and carried in a register into frame t+1.
// Each value costs two correctly rounded fp64 divisions by the two spans (InvariantDivisor);
// the raw inputs of frame t+1 are loaded before frame t is computed, so the load latency
// overlaps the dependent fp64 chain.
template <typename T> __global__ void __launch_bounds__(kThreads) combine_kernel(const CombineParams p) {
    const T* nm = static_cast<const T*>(p.norm);
    const NormRange scn = norm_range((double)nm[0], (double)nm[1]), tcn = norm_range((double)nm[2], (double)nm[3]);
    double lo = __longlong_as_double(0x7ff0000000000000LL), hi = __longlong_as_double(0xfff0000000000000LL);
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int tl0 = blockIdx.y * p.chunk_len;
    const int tl1 = min(p.t_count, tl0 + p.chunk_len);
    if (i < p.frame && tl0 < tl1) {
        RawBlock<T> next, cur;
        next.load(p, p.t_begin + tl0, i, p.is_last && tl0 == p.t_count - 1);
        double r_prev = 0.0;
        if (p.smooth && !(p.is_first && tl0 == 0)) {
            cur.load(p, p.t_begin + tl0 - 1, i, false);
            r_prev = removability_of(p, (double)cur.sc, (double)cur.tc, false, cur.bg, scn, tcn);
        }
        for (int tl = tl0; tl < tl1; ++tl) {
            const bool clip_last = p.is_last && tl == p.t_count - 1;
            const bool clip_first = p.is_first && tl == 0;
            cur = next;
            if (tl + 1 < tl1) next.load(p, p.t_begin + tl + 1, i, p.is_last && tl + 1 == p.t_count - 1);
            const double r = removability_of(p, (double)cur.sc, (double)cur.tc, clip_last, cur.bg, scn, tcn);
            double v = r;
            if (p.smooth && !clip_first)                                 // elvis.py:1206-1213
                v = __dadd_rn(__dmul_rn(p.beta, r), __dmul_rn(p.one_minus_beta, r_prev));
            r_prev = r;
            p.out[(int64_t)tl * p.frame + i] = v;
            lo = fmin(lo, v);
            hi = fmax(hi, v);
        }
    }
    block_minmax_commit(lo, hi, p.out_minmax);
}

__global__ void __launch_bounds__(kThreads) normalize_kernel(double* x, int64_t n, const double* mm) {
    const double lo = mm[0], hi = mm[1];
    if (!(hi > lo)) return;
    const InvariantDivisor span(__dsub_rn(hi, lo));
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        x[i] = span.divide(__dsub_rn(x[i], lo));
}

// ---- a3 ----------------------------------------------------------------------------------
struct ImportanceParams {
    const void* sc;
    const void* tc;
    const void* fg;
    int32_t t_begin, t_count, is_first, is_last;
    int64_t frame;
    double alpha, one_minus_alpha, beta, one_minus_beta;
    double* out;
};

template <typename T> __device__ __forceinline__ double complexity_at(const ImportanceParams& p, int t, int64_t i, bool last_frame) {
    const int64_t o = (int64_t)t * p.frame + i;
    const double s = ld<T>(p.sc, o);
    if (last_frame) return s;                                                   // utils.py:671
    return __dadd_rn(__dmul_rn(p.alpha, s), __dmul_rn(p.one_minus_alpha, ld<T>(p.tc, o + p.frame)));   // utils.py:670
}

// one CTA per frame: two passes over the frame's blocks (values recomputed, not staged)
template <typename T> __global__ void __launch_bounds__(kThreads) importance_kernel(const ImportanceParams p) {
    __shared__ double s_mm[2];
    const int tl = blockIdx.x;
    const int t = p.t_begin + tl;
    const bool clip_last = p.is_last && tl == p.t_count - 1;
    const bool clip_first = p.is_first && tl == 0;
    if (threadIdx.x == 0) {
        s_mm[0] = __longlong_as_double(0x7ff0000000000000LL);
        s_mm[1] = __longlong_as_double(0xfff0000000000000LL);
    }
    __syncthreads();
    auto value = [&](int64_t i) -> double {
        double c = complexity_at<T>(p, t, i, clip_last);
        if (!clip_first)                                                       // utils.py:675-676
            c = __dadd_rn(__dmul_rn(p.beta, c), __dmul_rn(p.one_minus_beta, complexity_at<T>(p, t - 1, i, false)));
        double f = 1.0;
        if (p.fg) {                                                            // utils.py:679-681
            f = ld<T>(p.fg, (int64_t)t * p.frame + i);
            if (f < 0.5) f = -1.0;
        }
        return __dmul_rn(c, f);
    };
    double lo = __longlong_as_double(0x7ff0000000000000LL), hi = __longlong_as_double(0xfff0000000000000LL);
    for (int64_t i = threadIdx.x; i < p.frame; i += kThreads) {
        const double v = value(i);
        lo = fmin(lo, v);
        hi = fmax(hi, v);
    }
    block_minmax_commit(lo, hi, s_mm);
    __syncthreads();
    const double mn = s_mm[0];
    const InvariantDivisor den(__dadd_rn(__dsub_rn(s_mm[1], mn), 1e-8));      // utils.py:686
    for (int64_t i = threadIdx.x; i < p.frame; i += kThreads)
        p.out[(int64_t)tl * p.frame + i] = den.divide(__dsub_rn(value(i), mn));
}

// ---- level maps --------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) levels_kernel(const double* s, int64_t n, int rule, int param, int32_t* out) {
    const double pm = (double)param;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const double v = s[i];
        int32_t lv;
        if (rule == ELVIS_LEVELS_ROUND) {
            lv = (int32_t)rint(__dmul_rn(v, pm));                               // np.round: half to even
        } else if (rule == ELVIS_LEVELS_INVERTED_ROUND) {
            lv = (int32_t)rint(__dmul_rn(__dsub_rn(1.0, v), pm));
            lv = max(0, min(param, lv));
        } else {
            int32_t b = (int32_t)floor(__dmul_rn(__dsub_rn(1.0, v), pm));
            b = max(0, min(param - 1, b));
            lv = b == 0 ? 0 : b + 1;
        }
        out[i] = lv;
    }
}

inline int grid_for(int64_t n) {
    int64_t g = (n + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)kNumSMs * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace
}  // namespace elvis

using namespace elvis;

extern "C" int elvis_minmax(const void* x, int32_t dtype, int64_t n, double* out, elvis_stream_t stream) {
    if (!x || !out || n <= 0) return ELVIS_ERR_INVALID_ARG;
    if (dtype != ELVIS_F32 && dtype != ELVIS_F64) return ELVIS_ERR_INVALID_ARG;
    cudaStream_t st = as_stream(stream);
    minmax_init<<<1, 32, 0, st>>>(out);
    ELVIS_CHECK_LAUNCH();
    if (dtype == ELVIS_F32)
        minmax_kernel<float><<<grid_for(n), kThreads, 0, st>>>(static_cast<const float*>(x), n, out);
    else
        minmax_kernel<double><<<grid_for(n), kThreads, 0, st>>>(static_cast<const double*>(x), n, out);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_combine_removability(const void* sc, const void* tc, int32_t dtype, const void* norm,
                                          int32_t t_ext, int32_t by, int32_t bx,
                                          int32_t t_begin, int32_t t_count, int32_t is_first, int32_t is_last,
                                          const uint8_t* background, double alpha, double beta, int32_t smooth,
                                          double* out, double* out_minmax, elvis_stream_t stream) {
    if (!sc || !tc || !norm || !out || !out_minmax) return ELVIS_ERR_INVALID_ARG;
    if (dtype != ELVIS_F32 && dtype != ELVIS_F64) return ELVIS_ERR_INVALID_ARG;
    if (t_ext <= 0 || by <= 0 || bx <= 0 || t_count <= 0 || t_begin < 0 || t_begin + t_count > t_ext) return ELVIS_ERR_INVALID_ARG;
    if (!is_last && t_begin + t_count >= t_ext) return ELVIS_ERR_INVALID_ARG;     // needs TC of the next frame
    if (smooth && !is_first && t_begin < 1) return ELVIS_ERR_INVALID_ARG;         // needs the previous frame
    CombineParams p;
    p.sc = sc;
    p.tc = tc;
    p.norm = norm;
    p.background = background;
    p.t_begin = t_begin;
    p.t_count = t_count;
    p.is_first = is_first;
    p.is_last = is_last;
    p.smooth = smooth;
    p.frame = (int64_t)by * bx;
    p.alpha = alpha;
    p.one_minus_alpha = 1.0 - alpha;
    p.beta = beta;
    p.one_minus_beta = 1.0 - beta;
    p.out = out;
    p.out_minmax = out_minmax;
    cudaStream_t st = as_stream(stream);
    minmax_init<<<1, 32, 0, st>>>(out_minmax);
    ELVIS_CHECK_LAUNCH();
    // enough CTAs for ~8 per SM; every chunk recomputes one carried value, so keep chunks >= 8 frames
    const int gx = (int)((p.frame + kThreads - 1) / kThreads);
    int chunks = (8 * kNumSMs + gx - 1) / gx;
    if (chunks > (t_count + 7) / 8) chunks = (t_count + 7) / 8;
    if (chunks < 1) chunks = 1;
    p.chunk_len = (t_count + chunks - 1) / chunks;
    chunks = (t_count + p.chunk_len - 1) / p.chunk_len;
    const dim3 grid((unsigned)gx, (unsigned)chunks);
    if (dtype == ELVIS_F32)
        combine_kernel<float><<<grid, kThreads, 0, st>>>(p);
    else
        combine_kernel<double><<<grid, kThreads, 0, st>>>(p);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_normalize(double* x, int64_t n, const double* minmax, elvis_stream_t stream) {
    if (!x || !minmax || n <= 0) return ELVIS_ERR_INVALID_ARG;
    normalize_kernel<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(x, n, minmax);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_importance_scores(const void* sc, const void* tc, const void* foreground, int32_t dtype,
                                       int32_t t_ext, int32_t by, int32_t bx,
                                       int32_t t_begin, int32_t t_count, int32_t is_first, int32_t is_last,
                                       double alpha, double beta, double* out, elvis_stream_t stream) {
    if (!sc || !tc || !out) return ELVIS_ERR_INVALID_ARG;
    if (dtype != ELVIS_F32 && dtype != ELVIS_F64) return ELVIS_ERR_INVALID_ARG;
    if (t_ext <= 0 || by <= 0 || bx <= 0 || t_count <= 0 || t_begin < 0 || t_begin + t_count > t_ext) return ELVIS_ERR_INVALID_ARG;
    if (!is_last && t_begin + t_count >= t_ext) return ELVIS_ERR_INVALID_ARG;
    if (!is_first && t_begin < 1) return ELVIS_ERR_INVALID_ARG;
    ImportanceParams p;
    p.sc = sc;
    p.tc = tc;
    p.fg = foreground;
    p.t_begin = t_begin;
    p.t_count = t_count;
    p.is_first = is_first;
    p.is_last = is_last;
    p.frame = (int64_t)by * bx;
    p.alpha = alpha;
    p.one_minus_alpha = 1.0 - alpha;
    p.beta = beta;
    p.one_minus_beta = 1.0 - beta;
    p.out = out;
    cudaStream_t st = as_stream(stream);
    if (dtype == ELVIS_F32)
        importance_kernel<float><<<t_count, kThreads, 0, st>>>(p);
    else
        importance_kernel<double><<<t_count, kThreads, 0, st>>>(p);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_levels_from_scores(const double* scores, int64_t n, int32_t rule, int32_t param,
                                        int32_t* levels, elvis_stream_t stream) {
    if (!scores || !levels || n <= 0 || param < 0) return ELVIS_ERR_INVALID_ARG;
    if (rule < ELVIS_LEVELS_ROUND || rule > ELVIS_LEVELS_INVERTED_BINS) return ELVIS_ERR_INVALID_ARG;
    if (rule == ELVIS_LEVELS_INVERTED_BINS && param < 1) return ELVIS_ERR_INVALID_ARG;
    levels_kernel<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(scores, n, rule, param, levels);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}
