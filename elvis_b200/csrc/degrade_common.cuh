// Shared by the v2 per-block degradation kernels (blur.cu, downsample.cu, dampen.cu): the geometry of a plane cut into
// blocks, the copy-through of partial blocks at the right / bottom edge, grid sizing.
#pragma once
#include "common.cuh"

namespace elvis {

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

// tile-row owned by accumulator-layout row i (the same map orders the columns: 4 q + j <-> layout 2 q + j, 8 + 2 q + j - 2)
__device__ __forceinline__ int imma_tile_row(int i) { return 4 * ((i & 7) >> 1) + (i & 1) + (i >= 8 ? 2 : 0); }

// src / dst planes of equal shape cut into By x Bx blocks of pb pixels (defined in degrade.cu)
int make_geom(const elvis_plane* src, const elvis_plane* dst, int T, int pb, int By, int Bx, BlockGeom& g);
// copies the right strip (x >= Bx*pb) and the bottom strip (y >= By*pb) of every frame, if there is one
int copy_edges(const BlockGeom& g, cudaStream_t st);

inline int grid_for_units(int64_t units, int per_cta) {
    int64_t gsz = (units + per_cta - 1) / per_cta;
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (int)(gsz < 1 ? 1 : (gsz > cap ? cap : gsz));
}

}  // namespace elvis
