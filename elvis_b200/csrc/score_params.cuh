// Parameters shared by the two implementations of elvis_score_sc_tc.
#pragma once
#include "common.cuh"

namespace elvis {

struct ScoreParams {
    const uint8_t* y;
    const uint8_t* halo;
    int64_t frame_stride, row_stride;
    int32_t T, By, Bx;
    int32_t tiles_x, tiles_y;      // spatial work tiles (meaning depends on the kernel)
    int32_t chunk_len, n_chunks;   // temporal chunks
    float* sc;
    float* tc;
    unsigned* mm;   // {sc_min, sc_max, tc_min, tc_max} as float bits (all values are >= +0)
    int32_t mm_begin, mm_end;
    float inv_area;
    uint32_t magic;     // 0x4B000000, passed at run time (see byte_as_biased_float)
    uint32_t magic16;   // 0x64006400: the fp16 analogue, used by the tcgen05 kernel
};

// implemented in score.cu (CUDA cores, any supported block size) and score_mma.cu
// (tensor cores + TMA, 16x16 blocks)
int launch_score_simt(ScoreParams p, int block_size, bool aligned8, cudaStream_t st);
int launch_score_mma(ScoreParams p, int plane_h, int plane_w, bool use_tma, cudaStream_t st);
// score_umma.cu: tcgen05 tensor cores, TMA ring, A operand in tensor memory (planes 16-byte aligned)
int launch_score_umma(ScoreParams p, int block_size, int plane_h, int plane_w, cudaStream_t st);
// score_dctn.cu: dct_size == block_size in {16, 32}, CUDA cores, one warp per block
int launch_score_dctn(ScoreParams p, int block_size, cudaStream_t st);
// score_dct16.cu: dct_size == block_size == 16 on the tensor cores (mma.sync f16 hi/lo split), TMA ring (16-byte aligned planes)
int launch_score_dct16(ScoreParams p, int plane_h, int plane_w, cudaStream_t st);
int score_umma_units_per_cta();
int score_umma_ctas_per_sm();

}  // namespace elvis
