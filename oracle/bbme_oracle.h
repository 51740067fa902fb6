/*
 * bbme_oracle.h -- CPU oracle for the block-matching hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C, single-threaded restatement of the reference's
 * algorithm (reference = /root/reference, cited as file:line).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (blockbasedmotionestimation_b200/) never includes, links or calls anything in oracle/.
 *
 * Parity pinning: see the header comment of bbme_oracle.c.
 */
#ifndef BBME_ORACLE_H
#define BBME_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_LEVELS 16

enum {
  ORC_OK = 0,
  ORC_E_ARG = -1,       /* bad argument */
  ORC_E_NOPAD = -2,     /* reference would print "Could not find any multiples..." and exit(1), motion_framework.cpp:21-26 */
  ORC_E_ODD_PAD = -3,   /* (padded - orig) odd: reference reads out of bounds */
  ORC_E_ONE_BLOCK = -4, /* < 2 blocks on an axis at some level: reference reads out of bounds (motion_framework.cpp:475) */
  ORC_E_NOMEM = -5,
  ORC_E_IO = -6,
  ORC_E_FORMAT = -7
};

typedef struct {
  int padded_w, padded_h, pad_x, pad_y; /* motion_framework.h:16-19 */
  int num_levels;
  int level_w[ORC_MAX_LEVELS], level_h[ORC_MAX_LEVELS];
} orc_shape;

typedef struct {
  double t_ctor_s;            /* pad + pyramid (MF::MF) */
  double t_run_s;             /* calcMotionBlockMatching, what main_class.cpp:47-55 times */
  uint64_t search_sad_calls;  /* in-bounds spiral positions evaluated */
  uint64_t search_absdiffs;   /* pixel |a-b| in the search */
  uint64_t reg_sad_calls;     /* in-bounds regularisation candidates */
  uint64_t reg_absdiffs;
  uint64_t level_search_absdiffs[ORC_MAX_LEVELS];
  uint64_t level_reg_absdiffs[ORC_MAX_LEVELS];
} orc_stats;

/* MF::MF padding search, motion_framework.cpp:15-54. */
int orc_plan_shape(int w, int h, int levels, const int* block_size, orc_shape* out);

/* cv::copyMakeBorder(BORDER_CONSTANT, 0), motion_framework.cpp:60-61. dst is (w+2*pad_x) x (h+2*pad_y), dense. */
void orc_pad_image(const uint8_t* src, int w, int h, size_t pitch, int pad_x, int pad_y, uint8_t* dst);

/* cv::pyrDown(src, dst, Size(sw/2, sh/2)), motion_framework.cpp:89-90. Dense buffers. */
void orc_pyrdown(const uint8_t* src, int sw, int sh, uint8_t* dst);

/* Literal spiral walk of find_min_block_spiral (motion_framework.cpp:326-411) without bounds
 * skipping: writes (dx,dy) pairs in visit order (first entry is the centre), returns their number. */
int orc_spiral_walk(int shift, int* dxdy, int cap_pairs);

/* Closed-form visit rank used by the GPU kernel's tie-break; here to be checked against the walk. */
int orc_spiral_rank(int dx, int dy);

/* MF::calcLevelBM, motion_framework.cpp:226-244.  flow is dense w*h float2 (u,v); reads the
 * prediction at block corners and writes the result there. */
void orc_search_level(const uint8_t* im1, const uint8_t* im2, int w, int h, int block_size, int search_size,
                      float* flow, orc_stats* st, int level);

/* MF::calcLevelBM with MF::find_min_block (motion_framework.cpp:246-294, the raster-scan search with the L1-distance tie-break
 * that the commented line :235 would call).  Pinned against the reference's own function through oracle/_ref. */
void orc_search_level_raster(const uint8_t* im1, const uint8_t* im2, int w, int h, int block_size, int search_size, float* flow);

/* MF::draw_MVimage, motion_framework.cpp:887-905: motion-compensated frame from image 2 and the field at block corners;
 * blocks whose source leaves the image keep the bytes already in `out`. */
void orc_compensate(const uint8_t* im2, int w, int h, int block_size, const float* flow, uint8_t* out);

/* One MF::regularize_MVs sweep in place, motion_framework.cpp:424-530. */
void orc_regularize_sweep(const uint8_t* im1, const uint8_t* im2, int w, int h, int block_size, float lambda,
                          int lambda_multiplier, float* flow, orc_stats* st, int level);

/* MF::divide_blocks, motion_framework.cpp:845-862 (block_size is the size BEFORE halving). */
void orc_divide_blocks(int w, int h, int block_size, float* flow);

/* MF::copyMVs, motion_framework.cpp:828-843.  coarse is cw x ch, fine is 2cw x 2ch. */
void orc_copy_mvs(const float* coarse, int cw, int ch, int coarse_block_size, float* fine);

/* MF::copy_to_all_pixels, motion_framework.cpp:815-826. */
void orc_copy_to_all_pixels(int w, int h, int block_size, float* flow);

/* MF::MF + MF::calcMotionBlockMatching on one pair.  flow_out: padded_h x padded_w x 2 floats.
 * sweeps = regularisation sweeps per block size (reference hard-codes 2, motion_framework.cpp:143,184). */
int orc_estimate(const uint8_t* im1, const uint8_t* im2, int w, int h, size_t pitch, int levels,
                 const int* search_size, const int* block_size, int sweeps, float* flow_out, orc_stats* st);

/* The whole path with MF::find_min_block (raster scan) as the per-level search instead of the spiral search. */
int orc_estimate_raster(const uint8_t* im1, const uint8_t* im2, int w, int h, size_t pitch, int levels,
                        const int* search_size, const int* block_size, int sweeps, float* flow_out);

/* Same, additionally copying the state after each stage into caller buffers (any may be NULL):
 *   pyr1/pyr2[l]   : level images (dense level_w x level_h)
 *   after_search[l], after_reg[l]: dense level flow after calcLevelBM / after the whole schedule. */
int orc_estimate_debug(const uint8_t* im1, const uint8_t* im2, int w, int h, size_t pitch, int levels,
                       const int* search_size, const int* block_size, int sweeps, float* flow_out, orc_stats* st,
                       uint8_t* const* pyr1, uint8_t* const* pyr2, float* const* after_search,
                       float* const* after_reg);

/* n independent pairs over `threads` host threads (pthread); the reference itself is single-threaded
 * (calcLevelBM_Parallel is disabled, motion_framework.cpp:127,170) so this is "one worker per core". */
int orc_estimate_many(int n, const uint8_t* const* im1, const uint8_t* const* im2, int w, int h, size_t pitch,
                      int levels, const int* search_size, const int* block_size, int sweeps, float* const* flow_out,
                      int threads, orc_stats* st_sum);

/* main()'s quarter-pel wrapper around the path (main_class.cpp:32-33): cv::resize(img, img, Size(), f, f, INTER_LINEAR)
 * for 8-bit images.  OpenCV is not in the reference tree; this restates its published fixed-point algorithm (imgproc
 * resize.cpp: 11-bit coefficients cvRound(w * 2048), horizontal pass in int32, vertical pass
 * (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2; x taps clamped with the weight zeroed, y taps
 * clamped by row index) and is pinned bit-for-bit against the container's cv2 4.13.0 (tests/golden/resize_cv2.npz).
 * dst is (factor * w) x (factor * h), dense. */
void orc_resize_linear(const uint8_t* src, int w, int h, int factor, uint8_t* dst);

/* main_class.cpp:58-70: strip the padding, keep every factor-th pixel, divide the vectors by factor.
 * out is (h / factor) x (w / factor) x 2 floats with (w, h) the size MF was constructed on. */
void orc_strip_subsample(const float* padded_flow, int padded_w, int padded_h, int pad_x, int pad_y, int factor,
                         float* out);

/* Flow::ReadFlowFile / WriteFlowFile / CalculateMSE, rw_flow.cpp:50-136,139-200,309-332. */
int orc_flo_read_header(const char* path, int* w, int* h);
int orc_flo_read(const char* path, float* data, int w, int h);
int orc_flo_write(const char* path, const float* data, int w, int h);
double orc_aee(const float* gt, const float* flow, int w, int h);

#ifdef __cplusplus
}
#endif
#endif
