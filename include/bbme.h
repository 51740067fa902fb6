/*
 * bbme.h -- C ABI of the B200 block-matching motion estimator (libbbme.so).
 *
 * This is the drop-in boundary for ONE path of ashish-nr/BlockBasedMotionEstimation:
 *     MF::MF + MF::calcMotionBlockMatching          (reference motion_framework.h:12-13, motion_framework.cpp:4-219)
 *     Flow::ReadFlowFile / WriteFlowFile / CalculateMSE (reference rw_flow.h:17-22, rw_flow.cpp:50-200,309-332)
 * The reference has no FFI of its own (it is one C++ program); include/motion_framework.h and
 * include/rw_flow.h re-create its C++ classes on top of the functions below, and INTEGRATION.md shows the
 * binding a maintainer of the reference would add.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 (BBME_OK) or a negative bbme_status;
 * bbme_last_error() gives the text.  A context owns one CUDA device, its streams and all device memory.
 * One context per GPU per host thread; calls on one context are not re-entrant (same as one MF object,
 * motion_framework.h:38-39).  There is no CPU fallback: without a usable CUDA device bbme_create fails.
 *
 * Index 0 of search_size[] / block_size[] is the finest pyramid level (motion_framework.cpp:67-74).
 * "search range +-R" in the reference's vocabulary is search_size = block_size + 2R (motion_framework.cpp:299).
 */
#ifndef BBME_H
#define BBME_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BBME_MAX_LEVELS 16
#define BBME_VERSION 201

typedef enum {
  BBME_OK = 0,
  BBME_E_ARG = -1,        /* null pointer, non-positive size, block size not a power of two in [2,128], ... */
  BBME_E_NOPAD = -2,      /* reference prints "Could not find any multiples of the block size..." and exits, motion_framework.cpp:21-26 */
  BBME_E_ODD_PAD = -3,    /* (padded - original) is odd: the reference pads floor(diff/2) per side and then reads out of bounds */
  BBME_E_ONE_BLOCK = -4,  /* fewer than 2 blocks on an axis at some level: the reference reads out of bounds, motion_framework.cpp:475 */
  BBME_E_NOMEM = -5,
  BBME_E_CUDA = -6,       /* CUDA runtime/driver failure, including "no device" */
  BBME_E_STATE = -7,      /* call order: estimate before plan, batch larger than planned, ... */
  BBME_E_IO = -8,         /* .flo: cannot open / short write */
  BBME_E_FORMAT = -9,     /* .flo: wrong extension, tag, size, truncated or trailing bytes */
  BBME_E_RANGE = -10      /* motion vectors would not fit int16 (image side > 16383) */
} bbme_status;

typedef struct bbme_ctx bbme_ctx;

/* What MF publishes after construction (motion_framework.h:16-19) plus the pyramid geometry. */
typedef struct {
  int width, height;                 /* original */
  int padded_width, padded_height;   /* MF::padded_width / padded_height */
  int padding_x, padding_y;          /* MF::padding_x / padding_y */
  int num_levels;
  int level_width[BBME_MAX_LEVELS];
  int level_height[BBME_MAX_LEVELS];
  int block_size[BBME_MAX_LEVELS];
  int search_size[BBME_MAX_LEVELS];
} bbme_shape;

/* Device-side timing and work counters of the last estimate call (all pairs of the call together). */
typedef struct {
  float ms_total;          /* CUDA-event time of the whole device pipeline, copies excluded */
  float ms_pyramid;        /* pad + pyrDown */
  float ms_search;         /* all levels */
  float ms_regularize;     /* all sweeps, splits and fix-up rounds */
  float ms_other;          /* MV upsample + export */
  uint32_t kernel_launches;
  uint32_t fix_rounds;     /* fix-up rounds after the first pass of every sweep, summed over sweeps and pairs */
  uint32_t fix_blocks;     /* blocks re-evaluated in those rounds */
  uint32_t search_kernel_used; /* bbme_stage_search only: 1 = generic kernel ran, 2 = TMA kernel ran */
  uint32_t search_launches;    /* search-kernel launches behind ms_search */
  uint32_t reserved;
  uint64_t search_candidates; /* in-bounds candidate positions evaluated (== oracle search_sad_calls) */
  uint64_t search_absdiffs;   /* pixels |a-b| in the search (== oracle search_absdiffs) */
} bbme_stats;

typedef struct {
  int sweeps;          /* regularisation sweeps per block size; reference hard-codes 2 (motion_framework.cpp:143,184) */
  int chunk_pairs;     /* pairs resident on the device per pipeline slot (>=1) */
  int slots;           /* pipeline slots (1..8); >1 overlaps H2D / compute / D2H of successive chunks */
  int search_kernel;   /* 0 = auto, 1 = force the generic kernel, 2 = force the TMA kernel (error if unsupported) */
  int collect_stats;   /* 1 = per-stage CUDA-event timing + work counters (adds synchronisation) */
  int keep_search_mv;  /* 1 = keep a copy of every level's field after the search (for bbme_debug_level_mv) */
  int search_variant;  /* 0 = MF::find_min_block_spiral, the search the reference runs (motion_framework.cpp:236);
                          1 = MF::find_min_block (:246-294), the raster-scan search with the L1-distance tie-break that the
                              commented line :235 would call instead (generic kernel; +-R up to 180) */
} bbme_options;

void bbme_default_options(bbme_options* o);

int bbme_version(void);
const char* bbme_status_string(int status);

/* Shape only, no device needed: the padding search of MF::MF (motion_framework.cpp:15-54). */
int bbme_plan_shape(int width, int height, int num_levels, const int* search_size, const int* block_size,
                    bbme_shape* out);

int bbme_create(bbme_ctx** ctx, int device);
void bbme_destroy(bbme_ctx* ctx);
const char* bbme_last_error(const bbme_ctx* ctx); /* ctx may be NULL: error of the last failed bbme_create */

/* Fixes geometry and parameters, allocates device memory, encodes TMA descriptors.  Replaces the parameter
 * half of MF::MF(image1, image2, search_size, block_size, num_levels) (motion_framework.h:12). */
int bbme_plan(bbme_ctx* ctx, int width, int height, int num_levels, const int* search_size,
              const int* block_size, const bbme_options* opt, bbme_shape* out);

/* One pair, host buffers: MF::MF (pad + pyramid) followed by MF::calcMotionBlockMatching.
 * im1/im2: 8-bit luma, `pitch_bytes` between rows.  flow: padded_height x padded_width x 2 float32 (u,v),
 * the CV_32FC2 field the reference returns (motion_framework.cpp:218). */
int bbme_estimate(bbme_ctx* ctx, const uint8_t* im1, const uint8_t* im2, size_t pitch_bytes, float* flow);

/* n independent pairs, host buffers (pinned buffers from bbme_host_alloc make the copies asynchronous).
 * Pairs are processed in chunks of chunk_pairs over `slots` streams. */
int bbme_estimate_batch(bbme_ctx* ctx, int n, const uint8_t* const* im1, const uint8_t* const* im2,
                        size_t pitch_bytes, float* const* flow);

/* Same as bbme_estimate_batch but returns after enqueueing the copies and kernels; the results are in `flow` after
 * the next bbme_sync.  Successive calls pipeline over the slots (the D2H of one call overlaps the H2D and kernels of
 * the next).  Host buffers must be pinned (bbme_host_alloc / cudaHostAlloc) and stay valid until bbme_sync. */
int bbme_estimate_batch_async(bbme_ctx* ctx, int n, const uint8_t* const* im1, const uint8_t* const* im2,
                              size_t pitch_bytes, float* const* flow);

/* main()'s quarter-pel wrapper around the path, on the device (main_class.cpp:32-33 and :58-70): the frames are
 * (width / factor) x (height / factor) -- width, height as given to bbme_plan, i.e. the size MF sees -- and are up-sampled
 * with the arithmetic of cv::resize(img, img, Size(), factor, factor, INTER_LINEAR) on the way into level 0; each flow
 * buffer receives (height / factor) x (width / factor) x 2 floats: the field with its padding stripped, every factor-th
 * pixel kept and the vectors divided by factor (main's subpix_MVs).  factor: 2, 4 or 8 (main uses 4).  The copies per
 * pair shrink by factor^2 (input) and by more than that (output), so this path is not bound by the host link. */
int bbme_estimate_upsampled(bbme_ctx* ctx, int n, int factor, const uint8_t* const* im1, const uint8_t* const* im2,
                            size_t pitch_bytes, float* const* flow);
int bbme_estimate_upsampled_async(bbme_ctx* ctx, int n, int factor, const uint8_t* const* im1,
                                  const uint8_t* const* im2, size_t pitch_bytes, float* const* flow);

/* A video sequence (stream frames, reuse pyramids): n_frames consecutive frames are n_frames - 1 pairs
 * (frame t, frame t + 1), i.e. what a caller looping `MF(frame[t], frame[t+1], ...)` over a clip computes.  Every frame
 * is uploaded, padded and down-sampled ONCE and serves as image 2 of pair t - 1 and as image 1 of pair t.  flow[t]
 * (t < n_frames - 1) receives the padded dense field of pair t; flow[n_frames - 1] is not read.  Results are identical
 * to bbme_estimate_batch on the same pairs. */
int bbme_estimate_sequence(bbme_ctx* ctx, int n_frames, const uint8_t* const* frames, size_t pitch_bytes,
                           float* const* flow);
int bbme_estimate_sequence_async(bbme_ctx* ctx, int n_frames, const uint8_t* const* frames, size_t pitch_bytes,
                                 float* const* flow);
/* Frames and fields in device memory: n_frames planes of plane_stride bytes, n_frames - 1 flow planes. */
int bbme_estimate_sequence_device(bbme_ctx* ctx, int n_frames, const uint8_t* d_frames, size_t pitch_bytes,
                                  size_t plane_stride, float* d_flow, size_t flow_plane_stride);

/* n <= chunk_pairs pairs already in device memory (same device as the context).  Frames are n planes of
 * `plane_stride` bytes; flow is n planes of flow_plane_stride floats.  Runs on slot 0's stream and returns
 * after enqueueing; call bbme_sync before reading.  This is the "inputs resident in HBM" entry point. */
int bbme_estimate_device(bbme_ctx* ctx, int n, const uint8_t* d_im1, const uint8_t* d_im2, size_t pitch_bytes,
                         size_t plane_stride, float* d_flow, size_t flow_plane_stride);

/* bbme_estimate_upsampled with frames and result in device memory (frames: (height/factor) rows of (width/factor)
 * bytes; flow planes of at least (height/factor) * (width/factor) * 2 floats). */
int bbme_estimate_upsampled_device(bbme_ctx* ctx, int n, int factor, const uint8_t* d_im1, const uint8_t* d_im2,
                                   size_t pitch_bytes, size_t plane_stride, float* d_flow, size_t flow_plane_stride);

/* Compact result of the same computation: the 2x2-granular int16 field (padded_height/2 x padded_width/2 x 2),
 * i.e. the information content of the dense float field (motion_framework.cpp:205-206 replicates it 2x2). */
int bbme_estimate_device_compact(bbme_ctx* ctx, int n, const uint8_t* d_im1, const uint8_t* d_im2,
                                 size_t pitch_bytes, size_t plane_stride, int16_t* d_mv, size_t mv_plane_stride);

/* Host-side expansion of a compact field into the dense CV_32FC2 layout (motion_framework.cpp:205-206,218): width2 x height2
 * int16 (u, v) entries -> (2 * height2) x (2 * width2) x 2 floats, every 2x2 pixel block one vector.  Runs on the library's
 * worker threads (the same code the host-buffer entry points use after their D2H copy); needs no GPU. */
int bbme_expand_compact(const int16_t* mv2, int width2, int height2, float* dense);

/* Both outputs of one run: the dense field (may be NULL) and the compact field (may be NULL), at least one given. */
int bbme_estimate_device_both(bbme_ctx* ctx, int n, const uint8_t* d_im1, const uint8_t* d_im2, size_t pitch_bytes,
                              size_t plane_stride, float* d_flow, size_t flow_plane_stride, int16_t* d_mv,
                              size_t mv_plane_stride);

/* Waits for all slots; with collect_stats it then folds the CUDA-event intervals and work counters of everything
 * enqueued since the previous bbme_sync / bbme_estimate_batch into the stats. */
int bbme_sync(bbme_ctx* ctx);
int bbme_get_stats(bbme_ctx* ctx, bbme_stats* out);
/* Run slot i on caller stream streams[i] (cudaStream_t handles, e.g. torch.cuda.Stream.cuda_stream) instead of
 * the context's own streams, so that the caller can bracket the work with its own events.  n must equal `slots`. */
int bbme_set_streams(bbme_ctx* ctx, int n, void* const* streams);
/* Live micro-benchmark of the VABSDIFF4.U8.ACC issue rate on this device (register-only, dependence-free chains on
 * every SM, ~1 ms): the integer roofline denominator, in |a-b| per second.  Also returns the SM clock it ran at. */
int bbme_measure_int_peak(bbme_ctx* ctx, double* absdiff_per_s, double* sm_mhz);
int bbme_get_shape(const bbme_ctx* ctx, bbme_shape* out);

/* The ceilings of the host-buffer entry points, measured on this box: pinned H2D and D2H copies of `bytes` bytes (alone and
 * both directions at once) and the worker threads' non-temporal write bandwidth into host memory.  bbme_estimate_batch moves,
 * per pair, 2*W*H bytes H2D and padded_w*padded_h/2 bytes D2H (the 2x2-granular int16 field) and writes the dense
 * 8*padded_w*padded_h-byte CV_32FC2 field (motion_framework.cpp:218) into the caller's buffer with the worker threads. */
typedef struct {
  double h2d_gbs, d2h_gbs, duplex_gbs_per_direction;
  double host_stream_write_gbs;
  int host_threads;
} bbme_host_link;
int bbme_measure_host_link(bbme_ctx* ctx, size_t bytes, bbme_host_link* out);

/* ---- the reference's calling pattern: one MF object per frame pair (main_class.cpp:45-50) ----
 * bbme_mf_open returns a context planned for (device, geometry, sweeps) with chunk_pairs = slots = 1 -- a parked one of the
 * same key if there is one, else a new one; bbme_mf_close parks it again (at most four idle contexts are kept, the oldest
 * is destroyed beyond that).  MF::MF / MF::~MF of include/bbme/dropin.hpp are these two calls, so constructing an MF per
 * pair does not pay context creation, device allocation and descriptor encoding every time.  On a planning error the
 * context is still returned for bbme_last_error; pass it to bbme_mf_close.  Thread-safe. */
int bbme_mf_open(bbme_ctx** ctx, int device, int width, int height, int num_levels, const int* search_size,
                 const int* block_size, int sweeps, bbme_shape* out);
void bbme_mf_close(bbme_ctx* ctx);
void bbme_mf_cache_clear(void);

/* ---- all GPUs of the box behind one call ----
 * A pool owns one context per device (n_devices = 0: every visible device; devices = NULL: 0..n-1).  bbme_pool_plan plans
 * them all alike; bbme_pool_estimate_batch cuts the n pairs into contiguous balanced shards, one host thread per device
 * runs its shard through bbme_estimate_batch, and every field lands in the caller's flow[i] -- pairs are independent
 * (one MF per pair), so there is no inter-GPU traffic and no collective. */
typedef struct bbme_pool bbme_pool;
int bbme_pool_create(bbme_pool** pool, int n_devices, const int* devices);
void bbme_pool_destroy(bbme_pool* pool);
int bbme_pool_device_count(const bbme_pool* pool);
const char* bbme_pool_last_error(const bbme_pool* pool);
int bbme_pool_plan(bbme_pool* pool, int width, int height, int num_levels, const int* search_size, const int* block_size,
                   const bbme_options* opt, bbme_shape* out);
int bbme_pool_estimate_batch(bbme_pool* pool, int n, const uint8_t* const* im1, const uint8_t* const* im2,
                             size_t pitch_bytes, float* const* flow);

/* ---- results of the other ranks without a communication kernel (one process per GPU, e.g. under torchrun) ----
 * The path has no data exchange between pairs; what a multi-process job may still want is every rank's fields in ONE place.
 * The owner allocates a buffer (bbme_device_alloc) and exports it (bbme_ipc_export: a 64-byte cudaIpcMemHandle_t that travels
 * through any channel, e.g. an NCCL / gloo broadcast); every other rank maps it with ITS OWN device current (bbme_ipc_open: peer
 * access over NVLink, no context on the owner's GPU) and pushes its slice with bbme_copy_async -- a copy-engine transfer that
 * takes no SM from the persistent search kernel, unlike a collective's kernels.  `device` is the calling rank's device. */
int bbme_device_alloc(int device, size_t bytes, void** p);
void bbme_device_free(int device, void* p);
int bbme_ipc_export(int device, void* p, unsigned char* handle64);
int bbme_ipc_open(int device, const unsigned char* handle64, void** p);
int bbme_ipc_close(int device, void* p);
int bbme_copy_async(int device, void* dst, const void* src, size_t bytes, void* stream);

/* Pinned host memory for asynchronous copies. */
int bbme_host_alloc(void** p, size_t bytes);
void bbme_host_free(void* p);

/* Measurement aid: with `on` != 0, bbme_estimate_batch[_async] performs its host<->device copies and the host-side expansion
 * but launches no kernel (the fields written are whatever the last real call left on the device).  Timing that call gives the
 * ceiling of the host side of the path -- link, pinned staging, worker threads, host memory -- under the same traffic mix. */
int bbme_debug_skip_compute(bbme_ctx* ctx, int on);
/* Test aid: sets the per-pair epoch of the regularisation's de-duplication stamps (a 32-bit counter that grows by a few
 * hundred per chunk for the lifetime of a plan; the kernel clears the stamps and restarts it before it can wrap). */
int bbme_debug_set_stamp_epoch(bbme_ctx* ctx, uint32_t epoch);
/* Test aids that need no GPU: how the search planner would run one pyramid level (which kernel family, ring, shared memory), and
 * the multiply-high constants the search kernel divides with (x / d == (uint64(x) * magic >> 32) >> shift for 0 <= x < 2^31). */
typedef struct bbme_search_geometry {
  int planned;            /* 0: not a TMA-kernel geometry (the generic kernel runs it); the other fields are 0 then */
  int copies;             /* 1: the kernel reads four byte-shifted copies of image 2 (no funnel shifts in the SAD loop) */
  int deep_ring;          /* 1: ring of 8 or 16 stages (few work items per unit) */
  int key64;              /* 1: 64-bit (SAD, spiral rank) keys */
  int rows_per_lane;      /* candidate rows per lane (SEG) */
  int pitch_words;        /* window row pitch class in 32-bit words */
  int stages, stage_bytes, smem_bytes;
  int bands, segments_per_band;
  int box_w, box_h;       /* TMA box of one window copy: bytes x rows */
  int two_boxes;          /* window wider than one 256-byte TMA box */
  int lanes_per_unit;
} bbme_search_geometry;
int bbme_debug_search_geometry(int level_width, int level_height, int block_size, int search_size, int allow_copies,
                               bbme_search_geometry* out);
int bbme_debug_div_magic(unsigned divisor, unsigned* magic, unsigned* shift);

/* ---- state of the last bbme_estimate* call on slot 0, for per-stage parity tests (host output buffers) ---- */
/* frame: 0 = image1, 1 = image2.  out: level_height x level_width bytes, dense. */
int bbme_debug_level_image(bbme_ctx* ctx, int pair, int frame, int level, uint8_t* out);
/* which: 0 = after the whole regularisation schedule (2x2-granular: level_height/2 x level_width/2 x 2 int16),
 *        1 = after the search (block-granular: level_height/bs x level_width/bs x 2 int16; needs keep_search_mv). */
int bbme_debug_level_mv(bbme_ctx* ctx, int pair, int level, int which, int16_t* out);

/* ---- single stages on host buffers (upload, one kernel, download): parity tests against the oracle ---- */
/* cv::pyrDown(src, Size(w/2, h/2)), motion_framework.cpp:89-90. */
int bbme_stage_pyrdown(bbme_ctx* ctx, const uint8_t* src, int w, int h, uint8_t* dst);
/* cv::resize(src, Size(), factor, factor, INTER_LINEAR), main_class.cpp:32-33.  dst: (factor*h) x (factor*w) bytes. */
int bbme_stage_resize(bbme_ctx* ctx, const uint8_t* src, int w, int h, int factor, uint8_t* dst);
/* MF::calcLevelBM (motion_framework.cpp:226-244) on one level.  mv: (h/bs) x (w/bs) x 2 int16, holds the
 * prediction on entry and the result on exit.  kernel: as bbme_options.search_kernel. */
int bbme_stage_search(bbme_ctx* ctx, const uint8_t* im1, const uint8_t* im2, int w, int h, int block_size,
                      int search_size, int16_t* mv, int kernel, bbme_stats* st);
/* MF::calcLevelBM with MF::find_min_block (motion_framework.cpp:246-294) instead of the spiral search: same arguments. */
int bbme_stage_search_raster(bbme_ctx* ctx, const uint8_t* im1, const uint8_t* im2, int w, int h, int block_size,
                             int search_size, int16_t* mv);
/* MF::draw_MVimage (motion_framework.cpp:887-905): the motion-compensated frame -- each block_size x block_size block of im2 at
 * (block position + vector) copied to the block's position in out (w x h bytes, zero where the source leaves the image).
 * mv: (h/bs) x (w/bs) x 2 int16.  After calcMotionBlockMatching the reference would call it with block_size 2 (:205,213-216). */
int bbme_stage_compensate(bbme_ctx* ctx, const uint8_t* im2, int w, int h, int block_size, const int16_t* mv, uint8_t* out);
/* One MF::regularize_MVs sweep (motion_framework.cpp:424-530) with in-place raster semantics.
 * lambda is the level's current lambda, lambda_multiplier the sweep's multiplier (motion_framework.cpp:607). */
int bbme_stage_regularize(bbme_ctx* ctx, const uint8_t* im1, const uint8_t* im2, int w, int h, int block_size,
                          float lambda, int lambda_multiplier, int16_t* mv, uint32_t* rounds_out);
/* MF::divide_blocks (motion_framework.cpp:845-862): (h/bs) x (w/bs) field -> (2h/bs) x (2w/bs). */
int bbme_stage_divide(bbme_ctx* ctx, const int16_t* mv_in, int gw, int gh, int16_t* mv_out);
/* MF::copyMVs (motion_framework.cpp:828-843): coarse 2x2-granular final field (ch/2 x cw/2) ->
 * fine prediction at fine_block_size granularity ((2ch/fbs) x (2cw/fbs)). */
int bbme_stage_copy_mvs(bbme_ctx* ctx, const int16_t* coarse_mv, int cw, int ch, int coarse_block_size,
                        int fine_block_size, int16_t* fine_pred);

/* ---- .flo codec and the endpoint-error metric (host code; rw_flow.cpp) ---- */
int bbme_flo_read_header(const char* path, int* width, int* height);
int bbme_flo_read(const char* path, float* data, int width, int height);          /* Flow::ReadFlowFile */
int bbme_flo_write(const char* path, const float* data, int width, int height);   /* Flow::WriteFlowFile */
/* Flow::CalculateMSE: average endpoint error over gt-known pixels (rw_flow.cpp:309-332). */
double bbme_flow_aee(const float* gt, const float* flow, int width, int height);
/* Flow::MotionToColor (rw_flow.cpp:202-274): Middlebury colour coding.  bgr: height x width x 3 bytes in OpenCV's
 * channel order (what main() writes to flow.png, main_class.cpp:73-75); unknown vectors (|u| or |v| > 1e9, NaN) are
 * black.  maxmotion <= 0: normalise by the largest known vector, like main() does (-1).  range5 (may be NULL) receives
 * {max radius, min u, max u, min v, max v} over the known vectors -- the line the reference prints. */
int bbme_flow_to_color(const float* flow, int width, int height, float maxmotion, uint8_t* bgr, float* range5);
/* main()'s post-processing (main_class.cpp:58-70): strip padding, keep every `factor`-th pixel, divide by factor.
 * out: (height/factor) x (width/factor) x 2 with width/height the ORIGINAL (pre-padding) size. */
int bbme_flow_strip_subsample(const float* padded_flow, const bbme_shape* shape, int factor, float* out);

#ifdef __cplusplus
}
#endif
#endif /* BBME_H */
