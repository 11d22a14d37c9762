/*
 * elvis_b200 -- C ABI of the B200 (sm_100a) implementation of the ELVIS / PRESLEY
 * pre-/post-processing hot path.
 *
 * The reference (emanuele-artioli/elvis) is pure Python and has no FFI layer; its boundary
 * for this path is a set of Python function signatures in elvis.py / utils.py /
 * presley.py.  Each entry point below names the reference function (file:line under
 * /root/reference) whose arithmetic it replaces; elvis_b200/{elvis,utils,presley}.py keep
 * the reference's Python signatures and call these through ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name starts with `host_`;
 *   - nothing here allocates, frees or synchronises; all work is enqueued on `stream`
 *     (a cudaStream_t passed as void*; NULL = the legacy default stream);
 *   - all entry points are re-entrant and may be used on several devices of one process: the
 *     only process-wide state is a per-device cache of idempotent facts (SM count, opt-in
 *     kernel attributes already set), kernels are launched on the CURRENT device;
 *   - the return value is ELVIS_OK (0) or a negative ELVIS_ERR_* code; for ELVIS_ERR_CUDA
 *     the failing cudaError_t is returned by elvis_last_cuda_error() (thread local);
 *   - clips are batched: leading dimension T (frames); the reference's per-frame calls are
 *     the T = 1 case;
 *   - block maps (scores, masks, levels) are dense row-major (T, By, Bx).
 */
#ifndef ELVIS_B200_H
#define ELVIS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ELVIS_B200_ABI_VERSION 7

#define ELVIS_OK               0
#define ELVIS_ERR_INVALID_ARG (-1)   /* NULL pointer, non-positive size, out-of-range parameter   */
#define ELVIS_ERR_UNSUPPORTED (-2)   /* a size/configuration this build has no kernel for          */
#define ELVIS_ERR_CUDA        (-3)   /* a CUDA runtime call failed; see elvis_last_cuda_error()    */
#define ELVIS_ERR_SHAPE       (-4)   /* dimensions not divisible by the block size (elvis.py:1376) */

#if defined(__GNUC__)
#define ELVIS_API __attribute__((visibility("default")))
#else
#define ELVIS_API
#endif

typedef void* elvis_stream_t;

/* One 8-bit image plane of a clip.  Planar YUV 4:2:0 is three of these (Y with the full
 * block size, U and V with half of it); the reference's packed H x W x 3 BGR/RGB frames
 * (elvis.py:4341, presley.py:184) are one plane with channels = 3. */
typedef struct elvis_plane {
    void*   data;          /* device pointer to frame 0, row 0                       */
    int64_t frame_stride;  /* bytes between consecutive frames                       */
    int64_t row_stride;    /* bytes between consecutive pixel rows                   */
    int32_t height;        /* pixel rows                                             */
    int32_t width;         /* pixels per row                                         */
    int32_t channels;      /* interleaved bytes per pixel: 1 (planar) or 3 (packed)  */
    int32_t reserved;
} elvis_plane;

/* element type selector for block-level float inputs */
#define ELVIS_F32 0
#define ELVIS_F64 1

/* which end of the ranking is removed (SURVEY.md appendix "Score polarity") */
#define ELVIS_REMOVE_HIGH 0          /* elvis.py:1401  argsort(-row)[:k]            */
#define ELVIS_REMOVE_LOW  1          /* utils.py:721   k successive argmin passes   */

/* level-map rules */
#define ELVIS_LEVELS_ROUND          0  /* rint(score * param)                     elvis.py:2146, 2176      */
#define ELVIS_LEVELS_INVERTED_ROUND 1  /* clip(rint((1-imp)*param), 0, param)     utils.py:1196-1197, presley.py:968-975 */
#define ELVIS_LEVELS_INVERTED_BINS  2  /* b = clip(floor((1-imp)*param), 0, param-1); b ? b+1 : 0   utils.py:1138-1148 */

ELVIS_API int         elvis_abi_version(void);
ELVIS_API const char* elvis_error_string(int code);
ELVIS_API int         elvis_last_cuda_error(void);

/* ---- a1: per-block SC / TC features (replaces the external EVCA call, elvis.py:1014-1031,
 * presley.py:202; arithmetic defined by oracle/spec_scoring.py -- parity unpinned).
 * y: luma plane, channels must be 1; only the top-left (By*bs) x (Bx*bs) region is read,
 * By = height / bs, Bx = width / bs.  prev_halo: one luma frame (same row_stride) that
 * precedes frame 0, or NULL (then TC[0] = 0).  block_size in {8, 16, 32}; dct_size = 8 (every
 * block is tiled into 8 x 8 transforms -- north_star's definition, the benchmarked mode) or
 * dct_size = block_size (one transform per block, what the reference's `evca.main -b block_size`
 * call asks EVCA for, elvis.py:1022-1023).  sc, tc: float32 (T, By, Bx).  minmax: NULL or 4 floats {sc_min, sc_max, tc_min,
 * tc_max} over frames [mm_begin, mm_end) -- overwritten, not accumulated.
 * Two kernels sit behind this entry point: the tcgen05 / TMA kernel (plane, strides and halo
 * 16-byte aligned, clip large enough to fill the GPU) and a CUDA-core kernel (anything else);
 * both meet the same tolerance against the spec.  ELVIS_SCORE_IMPL=umma|simt forces one. */
ELVIS_API int elvis_score_sc_tc(const elvis_plane* y, int32_t n_frames, const uint8_t* prev_halo,
                      int32_t block_size, int32_t dct_size, float* sc, float* tc,
                      float* minmax, int32_t mm_begin, int32_t mm_end, elvis_stream_t stream);

/* min and max of n block values -> out[0], out[1] (float64). */
ELVIS_API int elvis_minmax(const void* x, int32_t dtype, int64_t n, double* out, elvis_stream_t stream);

/* ---- a2: in-tree tail of calculate_removability_scores (elvis.py:1173-1215).
 * sc/tc: (t_ext, By, Bx) of `dtype`, covering local frames [0, t_ext).  The call produces
 * the un-normalised (smoothed) score for frames [t_begin, t_begin + t_count) into
 * out (t_count, By, Bx) float64 and their min/max into out_minmax[2] (overwritten).
 * norm: {sc_min, sc_max, tc_min, tc_max} of the WHOLE clip (same dtype as sc/tc).
 * is_first / is_last: local frame t_begin is the clip's first frame / local frame
 * t_begin + t_count - 1 is the clip's last frame (elvis.py:1183, 1206); when they are 0
 * the halo frames t_begin - 1 / t_begin + t_count must be present in sc/tc.
 * background: NULL or uint8 (t_ext, By, Bx), non-zero where the block is background
 * (elvis.py:1193-1195).  smooth = (beta < 1 && clip length >= 2) (elvis.py:1202). */
ELVIS_API int elvis_combine_removability(const void* sc, const void* tc, int32_t dtype, const void* norm,
                               int32_t t_ext, int32_t by, int32_t bx,
                               int32_t t_begin, int32_t t_count, int32_t is_first, int32_t is_last,
                               const uint8_t* background, double alpha, double beta, int32_t smooth,
                               double* out, double* out_minmax, elvis_stream_t stream);

/* final normalize_array (elvis.py:864-867, 1218): x = (x - min) / (max - min) when
 * max > min, unchanged otherwise.  minmax: 2 float64 on the device. */
ELVIS_API int elvis_normalize(double* x, int64_t n, const double* minmax, elvis_stream_t stream);

/* ---- a3: calculate_importance_scores (utils.py:665-688 == presley.py:129-152).
 * Same extended-range convention as above.  foreground: NULL (all foreground) or
 * (t_ext, By, Bx) of `dtype`.  out: float64 (t_count, By, Bx), per-frame normalised. */
ELVIS_API int elvis_importance_scores(const void* sc, const void* tc, const void* foreground, int32_t dtype,
                            int32_t t_ext, int32_t by, int32_t bx,
                            int32_t t_begin, int32_t t_count, int32_t is_first, int32_t is_last,
                            double alpha, double beta, double* out, elvis_stream_t stream);

/* ---- a4/a6: per block-row top-k removal mask (elvis.py:1399-1415; utils.py:716-733).
 * scores: float64 (T, By, Bx).  Row (t, by) removes k = k_per_row[by] blocks when
 * k_per_row != NULL, else k_uniform; ties are broken lowest column first.  mask: uint8
 * (T, By, Bx), 1 = removed.  Bx <= 4096. */
ELVIS_API int elvis_select_rows(const double* scores, int32_t n_frames, int32_t by, int32_t bx,
                      const int32_t* k_per_row, int32_t k_uniform, int32_t polarity,
                      uint8_t* mask, elvis_stream_t stream);

/* elvis_normalize followed by elvis_select_rows in one pass over the scores (elvis.py:1218 then 1399-1415): every
 * score is normalised in place with `minmax` (2 float64 on the device, the same rule and the same bits as
 * elvis_normalize) as its row is loaded for ranking, and the mask is selected on the normalised values. */
ELVIS_API int elvis_normalize_select_rows(double* scores, const double* minmax, int32_t n_frames, int32_t by, int32_t bx,
                                const int32_t* k_per_row, int32_t k_uniform, int32_t polarity,
                                uint8_t* mask, elvis_stream_t stream);

/* ---- a4/a6: shrink -- left-compact the kept blocks of every block row
 * (elvis.py:1418-1425; utils.py:727-735).  Plane block = block_px x block_px pixels.
 * dst must hold (By*block_px) rows of out_bx*block_px pixels; a row that keeps fewer than
 * out_bx blocks is zero padded, one that keeps more is truncated (utils.py:735).
 * ctas_per_sm: 0 = one CTA per (frame, block row) (fastest on an otherwise idle GPU);
 * n > 0 caps the grid at n CTAs per SM (grid-stride) so that the kernel can share the SMs
 * with a concurrently running scoring kernel (see elvis_b200.pipeline.ElvisV1Pipelined). */
ELVIS_API int elvis_shrink(const elvis_plane* src, const elvis_plane* dst, int32_t n_frames,
                 int32_t block_px, int32_t by, int32_t bx, int32_t out_bx,
                 const uint8_t* mask, int32_t ctas_per_sm, elvis_stream_t stream);

/* ---- a5/a7: stretch -- scatter shrunk blocks back to the mask == 0 positions of their
 * row, zeros elsewhere (elvis.py:1436-1455; utils.py:739-759).  shrunk_bx = blocks per
 * row of src. */
ELVIS_API int elvis_stretch(const elvis_plane* src, const elvis_plane* dst, int32_t n_frames,
                  int32_t block_px, int32_t by, int32_t bx, int32_t shrunk_bx,
                  const uint8_t* mask, int32_t ctas_per_sm, elvis_stream_t stream);

/* ---- a4-a7 for planar YUV 4:2:0 clips: Y, U and V in one launch (the block mask computed from
 * luma applies to the co-located (block_size/2)^2 chroma blocks).  src_yuv / dst_yuv: arrays of
 * three planes {Y, U, V}.  Needs block_size % 16 == 0, 16-byte aligned luma and 8-byte aligned
 * chroma planes; returns ELVIS_ERR_UNSUPPORTED otherwise (use the per-plane entry points). */
ELVIS_API int elvis_shrink_yuv420(const elvis_plane* src_yuv, const elvis_plane* dst_yuv, int32_t n_frames,
                        int32_t block_size, int32_t by, int32_t bx, int32_t out_bx,
                        const uint8_t* mask, int32_t ctas_per_sm, elvis_stream_t stream);
ELVIS_API int elvis_stretch_yuv420(const elvis_plane* src_yuv, const elvis_plane* dst_yuv, int32_t n_frames,
                         int32_t block_size, int32_t by, int32_t bx, int32_t shrunk_bx,
                         const uint8_t* mask, int32_t ctas_per_sm, elvis_stream_t stream);

/* ---- a8-a12: level maps from scores.  scores float64 (n), levels int32 (n). */
ELVIS_API int elvis_levels_from_scores(const double* scores, int64_t n, int32_t rule, int32_t param,
                             int32_t* levels, elvis_stream_t stream);

/* ---- a9/a11/a12: per-block repeated 5x5 sigma=1 Gaussian blur (elvis.py:2183-2191,
 * utils.py:1200-1210, presley.py:986-990).  rounds: int32 (T, By, Bx), values <= 0 copy.
 * block_px <= 64.  Rows/columns beyond By*block_px / Bx*block_px are copied through
 * (utils.py:1215-1216).  src and dst must not overlap. */
ELVIS_API int elvis_degrade_blur(const elvis_plane* src, const elvis_plane* dst, int32_t n_frames,
                       int32_t block_px, int32_t by, int32_t bx, const int32_t* rounds,
                       elvis_stream_t stream);

/* ---- a8/a10/a12: per-block AREA down + LINEAR up (elvis.py:2154-2164,
 * utils.py:1151-1161, presley.py:978-983).  levels: int32 (T, By, Bx) indexing
 * `tables`, the device copy of the blob built by elvis_b200/_tables.py
 * (n_levels entries; layout documented there); a level whose small size equals block_px
 * copies the block.  block_px <= 64.  fast_tables_ok: 1 when the builder marked every level as a
 * power-of-two reduction (closed-form integer kernel for planar 8/16-pixel blocks); 0: the generic
 * table-driven kernel; an even value > 1: MIXED -- bit l + 1 is set for every level l that is NOT a
 * power-of-two reduction (n_levels <= 30): those blocks take the generic kernel, the others the closed
 * form (utils.py:1142-1148: 16 -> 8, 5, 4). */
ELVIS_API int elvis_degrade_downsample(const elvis_plane* src, const elvis_plane* dst, int32_t n_frames,
                             int32_t block_px, int32_t by, int32_t bx, const int32_t* levels,
                             const int32_t* tables, int32_t n_levels, int32_t fast_tables_ok,
                             elvis_stream_t stream);

/* a8 for planar YUV 4:2:0 clips, power-of-two levels, Y, U and V in ONE launch (src_yuv / dst_yuv: arrays
 * of three planes {Y, U, V}): level l of a 16 x 16 block reduces the luma block by 2^min(l, max_level, 4)
 * per axis and its two 8 x 8 chroma blocks by 2^min(l, max_level, 3) (cv2 INTER_AREA down, INTER_LINEAR
 * up, elvis.py:2158-2163).  Needs block_size == 16, planes made of whole blocks, 16-byte aligned luma and
 * 8-byte aligned chroma; returns ELVIS_ERR_UNSUPPORTED otherwise (use the per-plane entry point). */
ELVIS_API int elvis_degrade_downsample_pow2_yuv420(const elvis_plane* src_yuv, const elvis_plane* dst_yuv, int32_t n_frames,
                                         int32_t block_size, int32_t by, int32_t bx, const int32_t* levels,
                                         int32_t max_level, elvis_stream_t stream);

/* ---- a14: DCT-coefficient dampening (README.md:11,44 only; defined by
 * oracle/spec_dct_dampen.py -- parity unpinned).  strength: float32 (T, By, Bx).
 * block_px must be a multiple of 8. */
ELVIS_API int elvis_dct_dampen(const elvis_plane* src, const elvis_plane* dst, int32_t n_frames,
                     int32_t block_px, int32_t by, int32_t bx, const float* strength,
                     elvis_stream_t stream);

/* ---- 8f rank 1: OpenCV client restorer -- per-block unsharp mask (elvis.py:2822-2867
 * restore_blur_opencv_unsharp_mask; utils.py:1253-1392 restore_with_opencv_lanczos /
 * restore_with_opencv_unsharp).  levels: int32 (T, By, Bx); a block with level L > 0 becomes
 * addWeighted(tile, 1 + L/2, GaussianBlur(tile, (0,0), L), -L/2) where tile = the block plus
 * `halo` pixels of context clamped to the frame (utils.py:1227-1250); other pixels are copied.
 * kernels: device table [max_level + 1][kernel_stride] of {ksize, q[ksize]} 8.8 fixed-point
 * Gaussian kernels (elvis_b200/_tables.py:gaussian_kernels); levels above max_level use the
 * last row.  src and dst must not overlap. */
ELVIS_API int elvis_restore_unsharp(const elvis_plane* src, const elvis_plane* dst, int32_t n_frames,
                          int32_t block_px, int32_t by, int32_t bx, const int32_t* levels,
                          int32_t halo, const int32_t* kernels, int32_t max_level, int32_t kernel_stride,
                          elvis_stream_t stream);

/* restore_downsample_opencv_lanczos (elvis.py:2773-2820): per block with level > 0, INTER_AREA
 * down to the level's small size and INTER_LANCZOS4 back up.  tables: the blob of
 * elvis_b200/_tables.py:build(..., lanczos=True). */
ELVIS_API int elvis_restore_lanczos(const elvis_plane* src, const elvis_plane* dst, int32_t n_frames,
                          int32_t block_px, int32_t by, int32_t bx, const int32_t* levels,
                          const int32_t* tables, int32_t n_levels, elvis_stream_t stream);

/* In-place temporal blending of a clip (utils.py:1310-1311): for t >= 1,
 * frame[t] = uint8(tb * frame[t-1] + (1 - tb) * frame[t]) in float64, frame[t-1] already blended. */
ELVIS_API int elvis_temporal_blend(const elvis_plane* clip, int32_t n_frames, double temporal_blend,
                         elvis_stream_t stream);

/* ---- a13: side-channel packers.  Bit packing is np.packbits-compatible (elvis.py:4414):
 * flat over n values, MSB first, last byte zero padded.  The 2-bit level packer is new
 * (README.md:50 TODO): per row of bx levels, ceil(bx/4) bytes, level i in bits
 * 2*(i%4).. of byte i/4; levels saturate to 0..3 (the reference's level 4 = 16x at block
 * size 16, elvis.py:2146, is kept as 8x). */
ELVIS_API int elvis_pack_mask_bits(const uint8_t* mask, int64_t n, uint8_t* packed, elvis_stream_t stream);
ELVIS_API int elvis_unpack_mask_bits(const uint8_t* packed, int64_t n, uint8_t* mask, elvis_stream_t stream);
ELVIS_API int elvis_pack_levels_2bit(const int32_t* levels, int64_t rows, int32_t bx, uint8_t* packed,
                           elvis_stream_t stream);
ELVIS_API int elvis_unpack_levels_2bit(const uint8_t* packed, int64_t rows, int32_t bx, int32_t* levels,
                             elvis_stream_t stream);

/* ---- 8f rank 2: row+column shrink (utils.py:763-1018).
 * elvis_rowcol_plan simulates shrink_frame_position_map / shrink_frame_removal_indices
 * (utils.py:763-830, 874-949) on block indices only: alternating row / column passes, each removing
 * the first-minimum block of every row (column) and shifting the rest, until target_removals
 * blocks are gone.  Per frame it writes
 *   position      (by, bx) int32, pitch bx: original linear block index y*bx+x now sitting at each
 *                 cell of the final (meta[1], meta[2]) grid (cells outside it are stale),
 *   mask          (by, bx) uint8: 1 = removed,
 *   pass_indices  (max_passes, max(by,bx)) int32 and pass_counts (max_passes): the reference's
 *                 removal_indices list, even entries row passes, odd entries column passes,
 *   meta          {n_passes, final_by, final_bx, removed}.
 * scratch_keys: n_frames*by*bx uint64.  max_passes >= by + bx + 2 always suffices. */
ELVIS_API int elvis_rowcol_plan(const double* importance, int32_t n_frames, int32_t by, int32_t bx, int64_t target_removals,
                      uint64_t* scratch_keys, int32_t* position, uint8_t* mask, int32_t* pass_indices,
                      int32_t* pass_counts, int32_t max_passes, int32_t* meta, elvis_stream_t stream);

/* stretch_frame_removal_indices (utils.py:951-1018): replays the recorded passes in reverse on a
 * grid of shrunk-block indices; grid (n_frames, grid_h, grid_w) int32 receives, for the final
 * (shrunk_by + #column passes, shrunk_bx + #row passes) region, the shrunk linear index
 * y*shrunk_bx+x to copy or -1 for a black block.  pass_indices is (n_frames, n_passes, lmax). */
ELVIS_API int elvis_rowcol_expand(const int32_t* pass_indices, const int32_t* pass_counts, int32_t n_frames, int32_t n_passes,
                        int32_t lmax, int32_t shrunk_by, int32_t shrunk_bx, int32_t* grid, int32_t grid_h,
                        int32_t grid_w, elvis_stream_t stream);

/* stretch_frame_position_map's scatter order (utils.py:851-855): inverse[t][orig] = the LAST
 * row-major entry index i with map[t][i] == orig, -1 where no entry points. */
ELVIS_API int elvis_invert_block_map(const int32_t* map, int32_t n_frames, int64_t entries_per_frame, int32_t* inverse,
                           int64_t inverse_per_frame, elvis_stream_t stream);

/* dst block (j, i) of frame t <- src block map[t][j][i] (linear index over a grid src_bx wide), or
 * zeros when the entry is negative.  map is (n_frames, map_rows, map_pitch) int32. */
ELVIS_API int elvis_gather_blocks(const elvis_plane* src, const elvis_plane* dst, int32_t n_frames, int32_t block_px,
                        int32_t dst_by, int32_t dst_bx, int32_t src_by, int32_t src_bx, const int32_t* map,
                        int32_t map_rows, int32_t map_pitch, elvis_stream_t stream);

/* ---- 8f rank 3: the step after scoring on the server -- delta-QP side files and raw 4:2:0 frames
 * (utils.py:453-462, 1026-1092; elvis.py:2027-2090).  The host mirrors write the files. */

/* create_kvazaar_roi_file (utils.py:1046-1052): dqp = int8(clip(clip((1 - imp) * 2 * qp_range - qp_range,
 * -14, 14), 0 - base_qp, 51 - base_qp)), float64 arithmetic, truncating cast. */
ELVIS_API int elvis_roi_kvazaar(const double* importance, int64_t n, int32_t base_qp, int32_t qp_range, int8_t* dqp,
                      elvis_stream_t stream);

/* float64 -> float32 input of the two resized ROI maps.  mode 0: float32(x) (utils.py:1078);
 * mode 1: float32(clip(2 x - 1, -1, 1)) (elvis.py:2030). */
ELVIS_API int elvis_roi_prepare_f32(const double* x, int64_t n, int32_t mode, float* out, elvis_stream_t stream);

/* cv2.resize(float32 map, INTER_AREA), shrinking only, bit-exact (oracle/spec_cv.py resize_area_f32):
 * n_maps maps of src_h x src_w -> dst_h x dst_w.  General ratios: per destination column / row the
 * entries [ofs[d], ofs[d+1]) of (source index, weight) of cv2's decimation table
 * (elvis_b200/_tables.py:area_f32_tables).  When both ratios are integers pass them as
 * int_scale_x/y (> 0) and null tables: cv2 then sums the window and multiplies by 1/area;
 * simd_cols = number of leading destination columns cv2 computes with its 4-lane 2x2 vector path. */
ELVIS_API int elvis_resize_area_f32(const float* src, int32_t n_maps, int32_t src_h, int32_t src_w, float* dst, int32_t dst_h,
                          int32_t dst_w, const int32_t* x_ofs, const int32_t* x_src, const float* x_alpha,
                          const int32_t* y_ofs, const int32_t* y_src, const float* y_alpha, int32_t int_scale_x,
                          int32_t int_scale_y, int32_t simd_cols, elvis_stream_t stream);

/* create_svtav1_roi_file (utils.py:1081-1088): levels = clip(int32(r * 8), 0, 7);
 * offset = clip(qp_range - levels * 2 * qp_range // 7, 0 - base_crf, 63 - base_crf). */
ELVIS_API int elvis_roi_svtav1_offsets(const float* resized, int64_t n, int32_t base_crf, int32_t qp_range, int32_t* offsets,
                             elvis_stream_t stream);

/* write_y4m's cv2.cvtColor(frame, COLOR_RGB2YUV_I420) (utils.py:453-462): packed RGB clip -> Y, U, V
 * planes (BT.601 limited range, 20-bit fixed point, chroma from the top-left pixel of each 2x2 quad).
 * Height and width must be even (ELVIS_ERR_SHAPE otherwise, as cv2 asserts). */
ELVIS_API int elvis_rgb_to_i420(const elvis_plane* rgb, const elvis_plane* y, const elvis_plane* u, const elvis_plane* v,
                      int32_t n_frames, elvis_stream_t stream);

/* ---- 8f rank 4: the deterministic parts either side of the external neural restorer. */

/* cv2.resize(frame, (w / factor, h / factor), INTER_AREA) for uint8 frames and an integer factor
 * (the pyramid of upscale_realesrgan_adaptive, elvis.py:2567, 2583): factor 2 -> (a+b+c+d+2)>>2,
 * larger -> round-half-even(sum * float32(1 / factor^2)).  dst gives the output size. */
ELVIS_API int elvis_area_downscale(const elvis_plane* src, const elvis_plane* dst, int32_t n_frames, int32_t factor,
                         elvis_stream_t stream);

/* The per-block restore of one pyramid stage (elvis.py:2586-2591): dst block (j, i) <- src block
 * (j, i) where factors[t][j][i] <= threshold; other blocks of dst are left as they are. */
ELVIS_API int elvis_merge_blocks(const elvis_plane* src, const elvis_plane* dst, int32_t n_frames, int32_t block_px, int32_t by,
                       int32_t bx, const int32_t* factors, int32_t threshold, elvis_stream_t stream);

/* Level maps <-> the 8-bit gray frames of the map video (elvis.py:2200-2202 and 2238-2240; the
 * video encode / decode is external): gray = uint8((m - min) / (max - min) * 255.0) in float64;
 * level = uint8(round_half_even(float32(g) / 255 * (max - min) + min)) in float32. */
ELVIS_API int elvis_levels_to_gray(const int32_t* maps, int64_t n, int32_t min_value, int32_t max_value, uint8_t* gray,
                         elvis_stream_t stream);
ELVIS_API int elvis_gray_to_levels(const uint8_t* gray, int64_t n, float min_value, float max_value, uint8_t* levels,
                         elvis_stream_t stream);

/* cv2.resize(map, (dst_w, dst_h), INTER_NEAREST) of n_maps dense block-level maps with 1-, 4- or 8-byte
 * elements (masks, level maps, scores whose grid differs from the frame's: elvis.py:1189-1193,
 * utils.py:1291-1296).  y_idx / x_idx: cv2's source index per destination row / column,
 * min(floor(d * (1.0 / (dsize / ssize))), ssize - 1) in double (elvis_b200/_tables.py:nearest_index). */
ELVIS_API int elvis_resize_nearest(const void* src, int32_t elem_bytes, int32_t n_maps, int32_t src_h, int32_t src_w, void* dst,
                         int32_t dst_h, int32_t dst_w, const int32_t* y_idx, const int32_t* x_idx, elvis_stream_t stream);

/* cv2.resize(map, (dst_w, dst_h), INTER_LINEAR) of n_maps dense float32 / float64 block-level maps
 * (dtype = ELVIS_F32 / ELVIS_F64): the importance map whose grid differs from the frame's
 * (utils.py:1127-1128, 1197-1198) and the x265 QP map when the CTU grid is finer than the block
 * grid (elvis.py:2068-2073).  cv2's fused lerp, horizontal then vertical; y_/x_index and
 * y_/x_frac are the source index and fraction per destination row / column
 * (elvis_b200/_tables.py:linear_float_index). */
ELVIS_API int elvis_resize_linear_float(const void* src, int32_t dtype, int32_t n_maps, int32_t src_h, int32_t src_w, void* dst,
                              int32_t dst_h, int32_t dst_w, const int32_t* y_index, const double* y_frac,
                              const int32_t* x_index, const double* x_frac, elvis_stream_t stream);

/* Luma of packed RGB frames for the scoring stage when the caller holds RGB frames
 * (presley.py:184-202: `analyze_frames(np.array(frames), ...)`): cv2.COLOR_RGB2GRAY's 15-bit
 * fixed point, (9798 R + 19235 G + 3735 B + 2^14) >> 15.  rgb: channels = 3; y: channels = 1. */
ELVIS_API int elvis_rgb_to_gray(const elvis_plane* rgb, const elvis_plane* y, int32_t n_frames, elvis_stream_t stream);

/* Packed 3-channel frames (the reference's H x W x 3 layout) <-> three single-channel planes of the same size.
 * packed: channels = 3; planes: array of 3 planes with channels = 1 (any strides).  The operators use the pair to run the
 * fast planar per-block kernels on packed clips (elvis.py:2141-2196 / utils.py:1101-1217 take packed frames). */
ELVIS_API int elvis_split_channels3(const elvis_plane* packed, const elvis_plane* planes, int32_t n_frames, elvis_stream_t stream);
ELVIS_API int elvis_merge_channels3(const elvis_plane* planes, const elvis_plane* packed, int32_t n_frames, elvis_stream_t stream);

/* Row-major refill map of stretch_video_frames (presley.py:806-819): map[t][i] = rank of block i
 * among the kept blocks (mask == 0) of frame t in row-major order, or -1 for removed blocks and for
 * kept blocks whose rank is >= capacity (= shrunk_by * shrunk_bx).  Feed it to elvis_gather_blocks. */
ELVIS_API int elvis_refill_map(const uint8_t* mask, int32_t n_frames, int64_t blocks_per_frame, int64_t capacity, int32_t* map,
                     elvis_stream_t stream);

/* ---- e: frame sharding over peer memory (NVLink / NVSwitch; SURVEY.md 8e, the split of elvis.py:264-278).
 * These replace the NCCL send/recv halo exchange and the two NCCL all-reduces of the sharded scorer
 * (elvis_b200/sharding.py keeps both transports).  Memory that other ranks touch comes from
 * elvis_peer_alloc (cudaMalloc + zero fill + CUDA IPC handle; this one call synchronises the device) and
 * is mapped by the peers with elvis_peer_open; the 64-byte handles travel over the caller's own channel
 * (torch.distributed).  Sequence numbers are 32-bit counters that only grow (modular comparison).
 *   elvis_peer_put     copy-engine peer copy of `bytes` into a peer's buffer, then a release store of `seq`
 *                      into a flag word in the peer's memory (stream ordered);
 *   elvis_peer_signal  the release store alone (e.g. "your halo slot is free again");
 *   elvis_peer_wait    enqueue a wait until a flag word in LOCAL memory has reached `seq`;
 *   elvis_peer_allreduce_minmax   in-place all-reduce of n <= 8 interleaved {min, max, ...} values (float32 or
 *                      float64, dtype = ELVIS_F32 / ELVIS_F64) through per-rank mailboxes of
 *                      elvis_peer_mailbox_bytes() bytes: host_mailboxes[r] = rank r's mailbox as mapped in this
 *                      process (own included), slot = call counter % 4, seq = call counter (>= 1).  Calls of one
 *                      rank must be stream ordered.
 * Waits give up after 4 s and store 1 into *error_word (device memory, may be NULL) instead of hanging. */
ELVIS_API int64_t elvis_peer_mailbox_bytes(void);
ELVIS_API int elvis_peer_alloc(int64_t bytes, void** device_ptr, void* host_handle_64_bytes);
ELVIS_API int elvis_peer_open(const void* host_handle_64_bytes, void** device_ptr);
ELVIS_API int elvis_peer_close(void* device_ptr);
ELVIS_API int elvis_peer_free(void* device_ptr);
ELVIS_API int elvis_peer_put(void* peer_dst, const void* src, int64_t bytes, uint32_t* peer_flag, uint32_t seq, elvis_stream_t stream);
ELVIS_API int elvis_peer_signal(uint32_t* peer_flag, uint32_t seq, elvis_stream_t stream);
ELVIS_API int elvis_peer_wait(const uint32_t* local_flag, uint32_t seq, int32_t* error_word, elvis_stream_t stream);
ELVIS_API int elvis_peer_allreduce_minmax(void* inout, int32_t dtype, int32_t n_values, int32_t rank, int32_t world,
                                void* const* host_mailboxes, int32_t slot, uint32_t seq, int32_t* error_word,
                                elvis_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ELVIS_B200_H */
