// Host build of the device arithmetic in elvis_b200/csrc/dct8.cuh + score_weights.inc, used by
// tests/test_host_arith.py to check the AAN butterflies, their scale tables and the folded
// score weights against the NumPy spec without a GPU.  Not part of the product.
#include <cstdint>
#include <cmath>
#include <cstdio>
#include "../elvis_b200/csrc/dct8.cuh"

static const float kW[8][8] = {
#include "../elvis_b200/csrc/score_weights.inc"
};

extern "C" {

// y: (T, 8, 8) uint8 tiles of one position over T frames -> sc[T], tc[T] (un-normalised sums,
// i.e. before the division by bs^2), computed exactly like score_kernel does per thread.
void host_score_tile(const uint8_t* y, int T, float* sc, float* tc) {
    float acc[8][8] = {};
    for (int t = 0; t < T; ++t) {
        float x[8][8];
        for (int r = 0; r < 8; ++r)
            for (int c = 0; c < 8; ++c)
                x[r][c] = (float)y[(t * 8 + r) * 8 + c] - (t ? (float)y[((t - 1) * 8 + r) * 8 + c] : 0.f);
        elvis::fdct8x8(x);
        float s = 0.f, d = 0.f;
        for (int u = 0; u < 8; ++u)
            for (int v = 0; v < 8; ++v) {
                if (u == 0 && v == 0) continue;
                acc[u][v] += x[u][v];
                s = fmaf(fabsf(acc[u][v]), kW[u][v], s);
                d = fmaf(fabsf(x[u][v]), kW[u][v], d);
            }
        sc[t] = s;
        tc[t] = t ? d : 0.f;
    }
}

// in: (8, 8) uint8, gains g[15] indexed by u+v (already including the 1/64) -> out (8, 8) float
void host_dampen_tile(const uint8_t* in, const float* g, float* out) {
    float x[8][8];
    for (int r = 0; r < 8; ++r)
        for (int c = 0; c < 8; ++c) x[r][c] = (float)in[r * 8 + c];
    elvis::fdct8x8(x);
    for (int u = 0; u < 8; ++u)
        for (int v = 0; v < 8; ++v) x[u][v] *= g[u + v];
    elvis::idct8x8(x);
    for (int r = 0; r < 8; ++r)
        for (int c = 0; c < 8; ++c) out[r * 8 + c] = x[r][c];
}
}
