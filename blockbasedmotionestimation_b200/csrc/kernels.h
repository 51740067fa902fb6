// kernels.h -- host-callable launchers of the sm_100a kernels (implemented in kernels.cu / search_tma.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace bbme {

// per-pair control words of the regularisation (device memory, kCtrWords uint32 per pair): the epoch of the de-duplication
// stamps (persists across launches) and the fix-up statistics
constexpr int kCtrWords = 8;
enum { CTR_EPOCH = 2, CTR_ROUNDS = 3, CTR_BLOCKS = 4 };

struct RegArgs {
  ImgView i1, i2;
  int bs;           // current block size of the stage
  int gw, gh;       // blocks per row / column at this block size
  float lm;         // lambda * (float)lambda_multiplier, motion_framework.cpp:607
  const short2* O;  // field before the sweep ("old")
  short2* Y;        // field after the sweep
  size_t mv_plane;  // entries between pairs in O / Y
  uint32_t* list0;  // work lists, `wl_plane` entries between pairs
  uint32_t* list1;
  short2* nv;       // scratch list: blocks deferred to the second pass of a round (used as uint32 entries)
  uint32_t* stamp;  // de-duplication stamps, one per block
  size_t wl_plane;
  uint32_t* ctr;    // kCtrWords per pair
  uint32_t* hist = nullptr;  // optional (BBME_REG_PROFILE=1): 8 words per block size, phase times of pair 0 (regularize.cu)
};

// pad (cv::copyMakeBorder constant 0) both frames of n pairs into level 0
void launch_pad(const uint8_t* in1, const uint8_t* in2, size_t in_pitch, size_t in_plane, int w, int h, int pad_x,
                int pad_y, uint8_t* out1, uint8_t* out2, int out_pitch, size_t out_plane, int pw, int ph, int n,
                cudaStream_t s);
// cv::resize(INTER_LINEAR) by a power-of-two factor fused with the pad (main()'s quarter-pel wrapper, main_class.cpp:32-33)
struct ResizeTaps {
  int factor, shift;
  int off[8];  // tap position of destination phase p relative to d / factor
  int wt[8];   // 11-bit weight of the second tap
};
int make_resize_taps(int factor, ResizeTaps* t);  // 0 on success; factor must be 2, 4 or 8
void launch_resize_pad(const uint8_t* in1, const uint8_t* in2, size_t in_pitch, size_t in_plane, int w, int h,
                       const ResizeTaps& taps, int pad_x, int pad_y, uint8_t* out1, uint8_t* out2, int out_pitch,
                       size_t out_plane, int ph, int n, cudaStream_t s);
// main()'s strip + sub-sample + divide (main_class.cpp:58-70) from the 2x2-granular level-0 field
void launch_export_subsample(const short2* mv2, int gw2, size_t mv_plane, int pad_x, int pad_y, int factor, float* out,
                             int ow, int oh, size_t out_plane, int n, cudaStream_t s);
// cv::pyrDown to (sw/2, sh/2), both frames of n pairs
void launch_pyrdown(ImgView src1, ImgView src2, uint8_t* dst1, uint8_t* dst2, int dpitch, size_t dplane, int n,
                    cudaStream_t s);
// generic exhaustive search (any power-of-two block size); mv holds the prediction on entry
// variant: 0 = find_min_block_spiral (the reference's active search), 1 = find_min_block (raster scan, L1-distance tie-break)
void launch_search_generic(ImgView i1, ImgView i2, MvView mv, int bs, int R, int n, unsigned long long* counters,
                           cudaStream_t s, int variant = 0);
// MF::draw_MVimage: motion-compensated frame from image 2 and a block-granular field
void launch_compensate(ImgView i2, const short2* mv, int gw, size_t mv_plane, int bs, uint8_t* out, int out_pitch,
                       size_t out_plane, int n, cudaStream_t s);
// MF::copyMVs: coarse final field at 2x2 granularity -> fine prediction at fine block granularity
void launch_copy_mvs(const short2* coarse, int cgw2, size_t cplane, int cbs, MvView fine, int fbs, int n,
                     cudaStream_t s);
// MF::divide_blocks
void launch_divide(const short2* in, int gw, int gh, size_t in_plane, short2* out, size_t out_plane, int n,
                   cudaStream_t s);
// dense CV_32FC2 field from the 2x2-granular level-0 field
void launch_export(const short2* mv2, int gw2, size_t mv_plane, float* out, int pw, int ph, size_t out_plane, int n,
                   cudaStream_t s);
void launch_export_compact(const short2* mv2, int gw2, int gh2, size_t mv_plane, int16_t* out, size_t out_plane, int n,
                           cudaStream_t s);
// The whole schedule of a level in one launch: a.bs / a.gw / a.gh describe the level's INITIAL block grid, a.O holds the field
// after the search; the result ends in a.O or a.Y depending on the number of sweeps and splits (the caller tracks the
// ping-pong like the per-sweep path does).  Returns 0 or -1 (launch failure).
// first_mult: lambda_multiplier of the first sweep (1 in the schedule); single_stage: stop after the sweeps of a.bs.
int launch_reg_level(const RegArgs& a, int sweeps, float lambda0, int first_mult, int single_stage, int n, int sm_budget,
                     cudaStream_t s);

// VABSDIFF4 issue-rate micro-benchmark (kernels.cu); returns 0 on success
int measure_int_peak(int sm_count, double* absdiff_per_s, double* sm_mhz);

// ---- TMA search kernel (search_tma.cu)
struct TmaSearchPlan {
  int supported;      // 0 if this (bs, R, geometry) is not handled by the TMA kernel
  int pre;            // 1: map_win is the 4-D map over the four byte-shifted copies of image 2 (launch_shift4 first)
  int bs, R;
  int seg;            // candidate rows per thread item
  int band_rows;      // candidate rows per staged band
  int box_w, box_h;   // TMA box (bytes, rows)
  int n_box_x;        // boxes per copy along x
  int threads;
  int stages;
  size_t smem_bytes;
  CUtensorMap map_win;   // image 2 of this level: dims (w, h, n); pre: dims (w, 4 copies, h, n)
  CUtensorMap map_blk;   // image 1 of this level
};
// returns 0 on success; fills plan->supported
// img2_shift4: room for 4 * plane bytes per pair (the byte-shifted copies written by launch_shift4 before every search launch
// of a plan with pre = 1), or nullptr to stay with the funnel-shift kernels; tma_search_wants_pre says whether it would be used.
int tma_search_plan(TmaSearchPlan* plan, const uint8_t* img1, const uint8_t* img2, const uint8_t* img2_shift4, int w, int h,
                    int pitch, size_t plane, int n_planes, int bs, int R, char* err, size_t errlen);
int tma_search_wants_pre(int w, int h, int bs, int R);
// Host-only views of the planner, for tests that run without a GPU (bbme_debug_search_geometry / bbme_debug_div_magic).
struct TmaGeomInfo {
  int planned, copies, deep_ring, key64, rows_per_lane, pitch_words, stages, stage_bytes, smem_bytes, bands, segments_per_band,
      box_w, box_h, two_boxes, lanes_per_unit;
};
int tma_search_geometry(int w, int h, int bs, int R, int allow_copies, TmaGeomInfo* out);  // 0: the generic kernel's geometry
void tma_div_magic(unsigned d, uint32_t* magic, uint32_t* shift);  // x / d == umulhi(x, magic) >> shift for 0 <= x < 2^31, d >= 2
void launch_shift4(ImgView src, uint8_t* dst, int n, cudaStream_t s);
// work_ctr: one device word owned by the caller's stream (zeroed here, then the kernel's block counter); nullptr = blocks strided by CTA
int launch_search_tma(const TmaSearchPlan& plan, ImgView i1, ImgView i2, MvView mv, int n,
                       unsigned long long* counters, unsigned int* work_ctr, int sm_count, cudaStream_t s);

}  // namespace bbme
