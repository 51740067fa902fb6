// capi.cu -- the C ABI of libbbme.so (include/bbme.h): context, plan, batched pipeline driver, stage entry points.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/bbme.h"
#include "hostpool.h"
#include "kernels.h"

using namespace bbme;

namespace {

thread_local std::string g_create_error;

struct Slot {
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  uint8_t* in1 = nullptr;  // staged host frames: chunk planes of in_pitch x height
  uint8_t* in2 = nullptr;
  uint8_t* img[2][kMaxLevels] = {};
  short2* mv_a[kMaxLevels] = {};
  short2* mv_b[kMaxLevels] = {};
  short2* mv_search[kMaxLevels] = {};
  short2* mv_final[kMaxLevels] = {};
  uint32_t* list0 = nullptr;
  uint32_t* list1 = nullptr;
  short2* nv = nullptr;
  uint32_t* stamp = nullptr;
  uint32_t* ctr = nullptr;
  unsigned long long* counters = nullptr;  // [0] candidates, [1] absdiffs
  uint32_t* hist = nullptr;                // BBME_REG_PROFILE=1: 64 words per level (RegArgs::hist)
  float* out = nullptr;               // device output of the up-sampled path (stripped, sub-sampled field)
  // host-buffer path: the compact int16 field is copied into a pinned staging buffer and expanded on the host
  cudaEvent_t done_ev = nullptr;  // cudaEventBlockingSync: bbme_sync sleeps instead of spinning (the cores belong to the expansion)
  int16_t* stage[2] = {nullptr, nullptr};
  Ticket* ticket[2] = {nullptr, nullptr};
  int stage_turn = 0;
  unsigned int* work_ctr = nullptr;  // block counter of the search launches (one word per level)
  uint8_t* sh4[kMaxLevels] = {};  // image 2 of the level in four byte phases (search kernels with pre = 1), 4 planes per pair
  TmaSearchPlan tma[kMaxLevels];
  TmaSearchPlan tma_seq[kMaxLevels];  // sequence mode: image 2 of pair i is plane i + 1 of the image-1 array
  std::vector<cudaEvent_t> ev;
  std::vector<int> ev_tag;
  int last_n = 0;
  // CUDA graphs of the per-chunk kernel sequence, keyed by everything run_chunk's launches depend on
  struct GraphEntry {
    int n, factor, seq;
    const void *in1, *in2, *flow, *compact;
    size_t in_pitch, in_plane, flow_plane, compact_plane;
    cudaGraphExec_t exec;
    uint32_t launches, search_launches;  // kernels inside the graph (for the stats)
  };
  std::vector<GraphEntry> graphs;
};

constexpr int kHistSweeps = 256;
enum { TAG_PYR = 0, TAG_SEARCH = 1, TAG_REG = 2, TAG_OTHER = 3, TAG_BEGIN = 4 };

}  // namespace

struct bbme_ctx {
  int device = 0;
  int sm_count = 0;
  std::string err;
  bool planned = false;
  bbme_shape shape{};
  bbme_options opt{};
  int pitch[kMaxLevels] = {};
  size_t plane[kMaxLevels] = {};
  size_t cap[kMaxLevels] = {};  // mv entries per pair per level (2x2 granularity)
  int in_pitch = 0;
  size_t in_plane = 0;
  size_t out_plane = 0;  // floats
  std::vector<Slot> slots;
  std::vector<void*> allocs;
  bbme_stats stats{};
  uint32_t launches = 0;
  uint32_t search_launches = 0;
  bool stats_armed = false;
  int skip_compute = 0;  // bbme_debug_skip_compute: host-buffer calls do their copies and expansions but launch no kernel
  int use_graphs = 1;   // small chunks replay a captured CUDA graph of their ~70-230 launches (BBME_GRAPHS=0 disables)
  int next_slot = 0;    // round-robin position over the slots across asynchronous calls
};

namespace {

int fail(bbme_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  else g_create_error = buf;
  return code;
}

#define CUDA_TRY(c, expr)                                                                                    \
  do {                                                                                                       \
    cudaError_t e_ = (expr);                                                                                 \
    if (e_ != cudaSuccess) return fail((c), BBME_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                                       __FILE__, __LINE__);                                                  \
  } while (0)

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
int round_up(int v, int m) { return (v + m - 1) / m * m; }

template <typename T>
int dev_alloc(bbme_ctx* c, T** p, size_t count, bool zero) {
  void* q = nullptr;
  size_t bytes = count * sizeof(T) + 256;  // slack: kernels may read a few bytes past the last row
  cudaError_t e = cudaMalloc(&q, bytes);
  if (e != cudaSuccess) return fail(c, BBME_E_NOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
  if (zero) {
    e = cudaMemset(q, 0, bytes);
    if (e != cudaSuccess) return fail(c, BBME_E_CUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
  }
  c->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return BBME_OK;
}

void release_plan(bbme_ctx* c) {
  for (Slot& s : c->slots) {
    // after a failed stream the queued host functions may never run: leak the ticket rather than wait for ever
    const bool drained = !s.stream || cudaStreamSynchronize(s.stream) == cudaSuccess;
    for (int j = 0; j < 2; ++j) {
      if (s.ticket[j] && drained) { s.ticket[j]->wait(); delete s.ticket[j]; }
      s.ticket[j] = nullptr;
      if (s.stage[j]) { cudaFreeHost(s.stage[j]); s.stage[j] = nullptr; }
    }
    for (cudaEvent_t e : s.ev) cudaEventDestroy(e);
    if (s.done_ev) { cudaEventDestroy(s.done_ev); s.done_ev = nullptr; }
    for (auto& g : s.graphs) cudaGraphExecDestroy(g.exec);
    s.graphs.clear();
    if (s.stream && s.own_stream) cudaStreamDestroy(s.stream);
  }
  c->slots.clear();
  for (void* p : c->allocs) cudaFree(p);
  c->allocs.clear();
  c->planned = false;
}

int shape_status(int w, int h, int levels, const int* ss, const int* bs, bbme_shape* out) {
  if (w <= 0 || h <= 0 || levels <= 0 || levels > BBME_MAX_LEVELS || !bs || !ss || !out) return BBME_E_ARG;
  for (int i = 0; i < levels; ++i)
    if (!is_pow2(bs[i]) || bs[i] < 2 || bs[i] > 128 || ss[i] <= 0) return BBME_E_ARG;
  // Padding search of MF::MF (motion_framework.cpp:15-46): smallest Hp >= H, Wp >= W divisible by 2^i * bs[i]
  // for every level, each axis bumped independently; the reference aborts when either axis reaches twice the
  // original, and tests that BEFORE testing divisibility (:21-26).
  long long th = h, tw = w;
  for (;;) {
    if (th == 2LL * h || tw == 2LL * w) return BBME_E_NOPAD;
    bool okh = true, okw = true;
    for (int i = 0; i < levels; ++i) {
      const long long q = (1LL << i) * bs[i];
      okh = okh && (th % q == 0);
      okw = okw && (tw % q == 0);
    }
    if (okh && okw) break;
    if (!okh) ++th;
    if (!okw) ++tw;
  }
  memset(out, 0, sizeof(*out));
  out->width = w;
  out->height = h;
  out->padded_width = (int)tw;
  out->padded_height = (int)th;
  out->padding_x = ((int)tw - w) / 2;
  out->padding_y = ((int)th - h) / 2;
  out->num_levels = levels;
  if (w + 2 * out->padding_x != out->padded_width || h + 2 * out->padding_y != out->padded_height) return BBME_E_ODD_PAD;
  int lw = out->padded_width, lh = out->padded_height;
  for (int i = 0; i < levels; ++i) {
    out->level_width[i] = lw;
    out->level_height[i] = lh;
    out->block_size[i] = bs[i];
    out->search_size[i] = ss[i];
    if (lw / bs[i] < 2 || lh / bs[i] < 2) return BBME_E_ONE_BLOCK;
    lw /= 2;
    lh /= 2;
  }
  if (out->padded_width > 16383 || out->padded_height > 16383) return BBME_E_RANGE;
  return BBME_OK;
}

int radius_of(int ss, int bs) {
  const int shift = ss - bs;
  return shift > 0 ? (shift >> 1) : 0;  // spiral covers [-R,R]^2 with R = (ss-bs)>>1 (motion_framework.cpp:299,326-411)
}

void mark(bbme_ctx* c, Slot& s, int tag) {
  if (!c->opt.collect_stats) return;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, s.stream);
  s.ev.push_back(e);
  s.ev_tag.push_back(tag);
}

void fold_events(bbme_ctx* c, Slot& s) {
  if (s.ev.size() >= 2) {
    for (size_t i = 1; i < s.ev.size(); ++i) {
      float ms = 0.f;
      if (s.ev_tag[i] == TAG_BEGIN) continue;  // gap between two chunks on this stream (copies, idle)
      if (cudaEventElapsedTime(&ms, s.ev[i - 1], s.ev[i]) != cudaSuccess) continue;
      switch (s.ev_tag[i]) {
        case TAG_PYR: c->stats.ms_pyramid += ms; break;
        case TAG_SEARCH: c->stats.ms_search += ms; break;
        case TAG_REG: c->stats.ms_regularize += ms; break;
        default: c->stats.ms_other += ms; break;
      }
      c->stats.ms_total += ms;
    }
  }
  for (cudaEvent_t e : s.ev) cudaEventDestroy(e);
  s.ev.clear();
  s.ev_tag.clear();
}

// The device pipeline for n pairs resident in slot `s` input planes (d_in1/d_in2), enqueued on s.stream.
// factor > 1: main()'s quarter-pel wrapper (main_class.cpp:32-33,58-70) -- the frames are (width / factor) x (height /
// factor) and are up-sampled on the way in; d_flow then receives the stripped, sub-sampled, divided field.
int run_chunk(bbme_ctx* c, Slot& s, int n, const uint8_t* d_in1, const uint8_t* d_in2, size_t in_pitch,
              size_t in_plane, float* d_flow, size_t flow_plane, int16_t* d_compact, size_t compact_plane,
              int factor = 1, bool seq = false) {
  const bbme_shape& sh = c->shape;
  const int L = sh.num_levels;
  cudaStream_t st = s.stream;
  s.last_n = n > s.last_n ? n : s.last_n;
  mark(c, s, TAG_BEGIN);
  // ---- MF::MF: pad + Gaussian pyramid (motion_framework.cpp:57-106)
  // Sequence mode (n pairs from n + 1 consecutive frames at d_in1): every frame is padded and down-sampled once, into
  // plane f of the image-1 array; pair i reads planes i and i + 1.  The two-frames-per-pair kernels do that unchanged
  // when "pair" p is handed frames 2p and 2p + 1 (pair stride = two planes); with an even n the last launch slot works
  // on a spare plane that nobody reads.
  const int npf = seq ? (n + 2) / 2 : n;  // launch slots of the pad / pyrDown kernels
  uint8_t* img2_l0 = seq ? s.img[0][0] + c->plane[0] : s.img[1][0];
  if (factor > 1) {
    ResizeTaps taps;
    if (make_resize_taps(factor, &taps) != 0) return fail(c, BBME_E_ARG, "up-sampling factor %d (supported: 2, 4, 8)", factor);
    launch_resize_pad(d_in1, seq ? d_in1 + in_plane : d_in2, in_pitch, seq ? 2 * in_plane : in_plane, sh.width / factor,
                      sh.height / factor, taps, sh.padding_x, sh.padding_y, s.img[0][0], img2_l0, c->pitch[0],
                      seq ? 2 * c->plane[0] : c->plane[0], sh.padded_height, npf, st);
  } else {
    launch_pad(d_in1, seq ? d_in1 + in_plane : d_in2, in_pitch, seq ? 2 * in_plane : in_plane, sh.width, sh.height,
               sh.padding_x, sh.padding_y, s.img[0][0], img2_l0, c->pitch[0], seq ? 2 * c->plane[0] : c->plane[0],
               sh.padded_width, sh.padded_height, npf, st);
  }
  ++c->launches;
  for (int l = 1; l < L; ++l) {
    const size_t sp = seq ? 2 * c->plane[l - 1] : c->plane[l - 1];
    ImgView a{s.img[0][l - 1], sh.level_width[l - 1], sh.level_height[l - 1], c->pitch[l - 1], sp};
    ImgView b{seq ? s.img[0][l - 1] + c->plane[l - 1] : s.img[1][l - 1], sh.level_width[l - 1], sh.level_height[l - 1],
              c->pitch[l - 1], sp};
    launch_pyrdown(a, b, s.img[0][l], seq ? s.img[0][l] + c->plane[l] : s.img[1][l], c->pitch[l],
                   seq ? 2 * c->plane[l] : c->plane[l], npf, st);
    ++c->launches;
  }
  mark(c, s, TAG_PYR);
  // ---- MF::calcMotionBlockMatching: coarse to fine (motion_framework.cpp:113-219)
  for (int l = L - 1; l >= 0; --l) {
    const int lw = sh.level_width[l], lh = sh.level_height[l];
    const int bs0 = sh.block_size[l];
    const int R = radius_of(sh.search_size[l], bs0);
    ImgView i1{s.img[0][l], lw, lh, c->pitch[l], c->plane[l]};
    ImgView i2{seq ? s.img[0][l] + c->plane[l] : s.img[1][l], lw, lh, c->pitch[l], c->plane[l]};
    short2* cur = s.mv_a[l];
    short2* nxt = s.mv_b[l];
    int g = bs0, gw = lw / g, gh = lh / g;
    MvView field{cur, gw, gh, c->cap[l]};
    if (l == L - 1) {
      CUDA_TRY(c, cudaMemsetAsync(cur, 0, (size_t)n * c->cap[l] * sizeof(short2), st));  // level_flow starts at zero (:70,92)
    } else {
      launch_copy_mvs(s.mv_final[l + 1], sh.level_width[l + 1] / 2, c->cap[l + 1], sh.block_size[l + 1], field, g, n, st);
      ++c->launches;
    }
    const TmaSearchPlan& tplan = seq ? s.tma_seq[l] : s.tma[l];
    const bool use_tma = tplan.supported && c->opt.search_kernel != 1 && c->opt.search_variant == 0;
    if (use_tma && tplan.pre) {  // the window image in its four byte phases (sequence mode: pair i's image 2 is frame i + 1)
      launch_shift4(i2, s.sh4[l], n, st);
      ++c->launches;
    }
    mark(c, s, TAG_OTHER);
    unsigned long long* ctrs = c->opt.collect_stats ? s.counters : nullptr;
    if (use_tma) {
      if (launch_search_tma(tplan, i1, i2, field, n, ctrs, s.work_ctr + l, c->sm_count, st) != 0)
        return fail(c, BBME_E_CUDA, "level %d: no TMA search kernel for block %d, R %d (plan and launch disagree)", l, g, R);
    } else {
      launch_search_generic(i1, i2, field, g, R, n, ctrs, st, c->opt.search_variant);
    }
    ++c->launches;
    ++c->search_launches;
    mark(c, s, TAG_SEARCH);
    if (c->opt.keep_search_mv && s.mv_search[l]) {
      CUDA_TRY(c, cudaMemcpy2DAsync(s.mv_search[l], (size_t)gw * gh * sizeof(short2), cur, c->cap[l] * sizeof(short2),
                                    (size_t)gw * gh * sizeof(short2), n, cudaMemcpyDeviceToDevice, st));
    }
    // regularisation schedule (motion_framework.cpp:133-154): per block size `sweeps` sweeps with
    // lambda_multiplier 1..sweeps, then split; lambda starts at bs/2 (integer division, :73,95) and doubles.
    float lambda = (float)(bs0 / 2);
    {
      // one launch per level: a cluster of CTAs per pair walks through every sweep, fix-up round and split
      RegArgs ra;
      ra.i1 = i1; ra.i2 = i2;
      ra.bs = g; ra.gw = gw; ra.gh = gh;
      ra.lm = 0.f;
      ra.O = cur; ra.Y = nxt;
      ra.mv_plane = c->cap[l];
      ra.list0 = s.list0; ra.list1 = s.list1; ra.nv = s.nv; ra.stamp = s.stamp;
      ra.wl_plane = c->cap[0];
      ra.ctr = s.ctr;
      ra.hist = s.hist ? s.hist + 64 * l : nullptr;  // BBME_REG_PROFILE: 8 words per block size, 64 per level
      if (launch_reg_level(ra, c->opt.sweeps, lambda, 1, 0, n, c->sm_count / (int)c->slots.size(), st) != 0)
        return fail(c, BBME_E_CUDA, "regularisation kernel launch failed at level %d: %s", l, cudaGetErrorString(cudaGetLastError()));
      ++c->launches;
      int swaps = 0;
      for (int gg = g; gg > 1; gg >>= 1) swaps += c->opt.sweeps + (gg > 2 ? 1 : 0);
      if (swaps & 1) cur = nxt;
    }
    s.mv_final[l] = cur;
    mark(c, s, TAG_REG);
  }
  // ---- final dense field (motion_framework.cpp:205-206,218)
  if (d_flow && factor > 1) {
    launch_export_subsample(s.mv_final[0], sh.padded_width / 2, c->cap[0], sh.padding_x, sh.padding_y, factor, d_flow,
                            sh.width / factor, sh.height / factor, flow_plane, n, st);
    ++c->launches;
  } else if (d_flow) {
    launch_export(s.mv_final[0], sh.padded_width / 2, c->cap[0], d_flow, sh.padded_width, sh.padded_height, flow_plane, n, st);
    ++c->launches;
  }
  if (d_compact) {
    launch_export_compact(s.mv_final[0], sh.padded_width / 2, sh.padded_height / 2, c->cap[0], d_compact, compact_plane, n, st);
    ++c->launches;
  }
  mark(c, s, TAG_OTHER);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(c, BBME_E_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
  return BBME_OK;
}

// run_chunk through a CUDA graph.  A chunk is 70-230 dependent launches, most of them a few microseconds long: for small
// chunks (a single pair above all) the host's launch calls and the gaps between dependent kernels are a large part of the
// latency.  The launch sequence has no host-side decisions, so it is captured once per distinct argument set and
// replayed.  Large chunks (every kernel runs for >> a launch) and stats collection (events between stages) launch
// directly.
int run_chunk_graphed(bbme_ctx* c, Slot& s, int n, const uint8_t* d_in1, const uint8_t* d_in2, size_t in_pitch,
                      size_t in_plane, float* d_flow, size_t flow_plane, int16_t* d_compact, size_t compact_plane,
                      int factor = 1, bool seq = false) {
  if (!c->use_graphs || c->opt.collect_stats || c->opt.keep_search_mv || n > 16)
    return run_chunk(c, s, n, d_in1, d_in2, in_pitch, in_plane, d_flow, flow_plane, d_compact, compact_plane, factor, seq);
  for (auto& g : s.graphs) {
    if (g.n == n && g.factor == factor && g.seq == (int)seq && g.in1 == d_in1 && g.in2 == d_in2 && g.flow == d_flow &&
        g.compact == d_compact && g.in_pitch == in_pitch && g.in_plane == in_plane && g.flow_plane == flow_plane &&
        g.compact_plane == compact_plane) {
      CUDA_TRY(c, cudaGraphLaunch(g.exec, s.stream));
      c->launches += g.launches;
      c->search_launches += g.search_launches;
      return BBME_OK;
    }
  }
  const uint32_t launches_before = c->launches, search_before = c->search_launches;
  if (cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    return run_chunk(c, s, n, d_in1, d_in2, in_pitch, in_plane, d_flow, flow_plane, d_compact, compact_plane, factor, seq);
  }
  int rc = run_chunk(c, s, n, d_in1, d_in2, in_pitch, in_plane, d_flow, flow_plane, d_compact, compact_plane, factor, seq);
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(s.stream, &graph);
  cudaGraphExec_t exec = nullptr;
  if (rc == BBME_OK && e == cudaSuccess && graph) e = cudaGraphInstantiate(&exec, graph, 0);
  if (graph) cudaGraphDestroy(graph);
  if (rc != BBME_OK) return rc;
  if (e != cudaSuccess || !exec) {  // capture not possible here (e.g. a caller stream that is already capturing): launch directly
    cudaGetLastError();
    c->launches = launches_before;
    c->search_launches = search_before;
    return run_chunk(c, s, n, d_in1, d_in2, in_pitch, in_plane, d_flow, flow_plane, d_compact, compact_plane, factor, seq);
  }
  if (s.graphs.size() >= 8) {
    cudaGraphExecDestroy(s.graphs.front().exec);
    s.graphs.erase(s.graphs.begin());
  }
  s.graphs.push_back({n, factor, (int)seq, d_in1, d_in2, d_flow, d_compact, in_pitch, in_plane, flow_plane, compact_plane, exec,
                      c->launches - launches_before, c->search_launches - search_before});
  CUDA_TRY(c, cudaGraphLaunch(exec, s.stream));
  return BBME_OK;
}

int collect_after_sync(bbme_ctx* c) {
  if (!c->opt.collect_stats || !c->stats_armed) return BBME_OK;
  for (Slot& s : c->slots) {
    fold_events(c, s);
    if (s.last_n > 0) {
      unsigned long long h[2] = {0, 0};
      CUDA_TRY(c, cudaMemcpy(h, s.counters, sizeof(h), cudaMemcpyDeviceToHost));
      c->stats.search_candidates += h[0];
      c->stats.search_absdiffs += h[1];
      std::vector<uint32_t> ctr((size_t)s.last_n * kCtrWords);
      CUDA_TRY(c, cudaMemcpy(ctr.data(), s.ctr, ctr.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
      for (int p = 0; p < s.last_n; ++p) {
        c->stats.fix_rounds += ctr[(size_t)p * kCtrWords + CTR_ROUNDS];
        c->stats.fix_blocks += ctr[(size_t)p * kCtrWords + CTR_BLOCKS];
      }
      s.last_n = 0;
    }
  }
  for (Slot& s : c->slots) {
    if (!s.hist) continue;
    std::vector<uint32_t> hh((size_t)kHistSweeps * 64);
    CUDA_TRY(c, cudaMemcpy(hh.data(), s.hist, hh.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    CUDA_TRY(c, cudaMemset(s.hist, 0, hh.size() * sizeof(uint32_t)));
    {
      for (int l = 0; l < c->shape.num_levels; ++l)
        for (int k = 1; k < 8; ++k) {
          const uint32_t* q = &hh[(size_t)l * 64 + 8 * k];
          if (!q[0] && !q[1]) continue;
          fprintf(stderr, "regprofile level %d bs %d: classify %.1f us, first pass %.1f us (%u listed, %u deferred), %u later rounds %.1f us (%u blocks)\n",
                  l, 1 << k, q[0] * 1e-3, q[1] * 1e-3, q[4], q[6], q[3], q[2] * 1e-3, q[5]);
        }
    }
  }
  c->stats.search_launches = c->search_launches;
  c->stats_armed = false;
  return BBME_OK;
}

int sync_all(bbme_ctx* c) {
  for (Slot& s : c->slots) {
    // a blocking event instead of cudaStreamSynchronize: the calling thread sleeps, it does not spin on a core that the
    // expansion threads (and, with one process per GPU, the other ranks) can use
    if (!s.done_ev) CUDA_TRY(c, cudaEventCreateWithFlags(&s.done_ev, cudaEventBlockingSync | cudaEventDisableTiming));
    CUDA_TRY(c, cudaEventRecord(s.done_ev, s.stream));
  }
  for (Slot& s : c->slots) CUDA_TRY(c, cudaEventSynchronize(s.done_ev));
  // host functions are stream-ordered: every expansion has been queued by now; wait for the worker threads
  for (Slot& s : c->slots)
    for (int j = 0; j < 2; ++j)
      if (s.ticket[j]) s.ticket[j]->wait();
  return BBME_OK;
}

// Dense result of a chunk into the callers' host buffers (motion_framework.cpp:218: padded CV_32FC2, every 2x2 block one
// vector): D2H of the 2x2-granular int16 field the schedule ends with (1/8 of the dense bytes) into a pinned staging
// buffer, then a stream-ordered host function hands the expansion (int16 -> float, 2x2 replication) to the worker threads.
struct ExpandJob {
  const int16_t* stage;
  std::vector<float*> dst;
  int gw2, gh2;
  size_t pw, plane;  // plane: int16 per pair in the staging buffer
  Ticket* ticket;
};

void CUDART_CB expand_cb(void* p) {
  ExpandJob* j = static_cast<ExpandJob*>(p);
  HostPool& pool = HostPool::instance();
  for (size_t i = 0; i < j->dst.size(); ++i)
    pool.expand_async(j->stage + i * j->plane, j->gw2, j->gh2, j->dst[i], j->pw, j->ticket);
  j->ticket->done(1);  // the job's own reference
  delete j;
}

int enqueue_dense_result(bbme_ctx* c, Slot& s, int m, float* const* flow) {
  const size_t plane = c->cap[0] * 2;  // int16 per pair
  const int turn = s.stage_turn;
  for (int j = 0; j < 2; ++j) {  // first host-buffer call on this slot: both staging buffers at once (pinning is slow)
    if (s.stage[j]) continue;
    void* q = nullptr;
    const size_t bytes = (size_t)c->opt.chunk_pairs * plane * sizeof(int16_t);
    if (cudaHostAlloc(&q, bytes, cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      return fail(c, BBME_E_NOMEM, "cudaHostAlloc(%zu bytes) for the result staging buffer failed", bytes);
    }
    s.stage[j] = static_cast<int16_t*>(q);
    s.ticket[j] = new Ticket();
  }
  s.ticket[turn]->wait();  // the chunk that used this staging buffer two turns ago has been expanded
  CUDA_TRY(c, cudaMemcpyAsync(s.stage[turn], s.mv_final[0], (size_t)m * plane * sizeof(int16_t), cudaMemcpyDeviceToHost, s.stream));
  ExpandJob* j = new ExpandJob();
  j->stage = s.stage[turn];
  j->dst.assign(flow, flow + m);
  j->gw2 = c->shape.padded_width / 2;
  j->gh2 = c->shape.padded_height / 2;
  j->pw = (size_t)c->shape.padded_width;
  j->plane = plane;
  j->ticket = s.ticket[turn];
  j->ticket->add(1);
  cudaError_t e = cudaLaunchHostFunc(s.stream, expand_cb, j);
  if (e != cudaSuccess) {
    j->ticket->done(1);
    delete j;
    return fail(c, BBME_E_CUDA, "cudaLaunchHostFunc failed: %s", cudaGetErrorString(e));
  }
  s.stage_turn ^= 1;
  return BBME_OK;
}

// Start of an estimate call: reset the stats and (asynchronously, on each slot's stream) the device counters.
void begin_call(bbme_ctx* c) {
  memset(&c->stats, 0, sizeof(c->stats));
  c->launches = 0;
  c->search_launches = 0;
  if (!c->opt.collect_stats) return;
  for (Slot& s : c->slots) {
    for (cudaEvent_t e : s.ev) cudaEventDestroy(e);
    s.ev.clear();
    s.ev_tag.clear();
    s.last_n = 0;
    cudaMemsetAsync(s.counters, 0, 2 * sizeof(unsigned long long), s.stream);
    cudaMemset2DAsync(s.ctr + CTR_ROUNDS, kCtrWords * sizeof(uint32_t), 0, 2 * sizeof(uint32_t), c->opt.chunk_pairs, s.stream);
  }
  c->stats_armed = true;
}

}  // namespace

extern "C" {

void bbme_default_options(bbme_options* o) {
  if (!o) return;
  o->sweeps = 2;
  o->chunk_pairs = 1;
  o->slots = 1;
  o->search_kernel = 0;
  o->collect_stats = 0;
  o->keep_search_mv = 0;
  o->search_variant = 0;
}

int bbme_version(void) { return BBME_VERSION; }

const char* bbme_status_string(int s) {
  switch (s) {
    case BBME_OK: return "ok";
    case BBME_E_ARG: return "invalid argument";
    case BBME_E_NOPAD: return "Could not find any multiples of the block size that match padded image dimensions";
    case BBME_E_ODD_PAD: return "padded size minus original size is odd (unsupported: the reference reads out of bounds)";
    case BBME_E_ONE_BLOCK: return "fewer than two blocks along an axis at some level (unsupported: the reference reads out of bounds)";
    case BBME_E_NOMEM: return "out of memory";
    case BBME_E_CUDA: return "CUDA error";
    case BBME_E_STATE: return "invalid call order or size for this plan";
    case BBME_E_IO: return "file I/O error";
    case BBME_E_FORMAT: return "bad .flo file";
    case BBME_E_RANGE: return "image too large for int16 motion vectors";
    default: return "unknown status";
  }
}

int bbme_plan_shape(int width, int height, int num_levels, const int* search_size, const int* block_size,
                    bbme_shape* out) {
  return shape_status(width, height, num_levels, search_size, block_size, out);
}

int bbme_create(bbme_ctx** out, int device) {
  if (!out) return fail(nullptr, BBME_E_ARG, "bbme_create: null output pointer");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(nullptr, BBME_E_CUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= count) return fail(nullptr, BBME_E_ARG, "device %d out of range (0..%d)", device, count - 1);
  cudaDeviceProp prop;
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    return fail(nullptr, BBME_E_CUDA, "cannot select device %d: %s", device, cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, BBME_E_CUDA, "device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major,
                prop.minor);
  bbme_ctx* c = new bbme_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  if (const char* g = getenv("BBME_GRAPHS")) c->use_graphs = atoi(g) != 0;
  bbme_default_options(&c->opt);
  *out = c;
  return BBME_OK;
}

void bbme_destroy(bbme_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  release_plan(c);
  delete c;
}

const char* bbme_last_error(const bbme_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int bbme_plan(bbme_ctx* c, int width, int height, int num_levels, const int* search_size, const int* block_size,
              const bbme_options* opt, bbme_shape* out) {
  if (!c) return BBME_E_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  release_plan(c);
  bbme_shape sh;
  int rc = shape_status(width, height, num_levels, search_size, block_size, &sh);
  if (rc != BBME_OK) return fail(c, rc, "bbme_plan(%dx%d, %d levels): %s", width, height, num_levels, bbme_status_string(rc));
  bbme_options o;
  bbme_default_options(&o);
  if (opt) o = *opt;
  if (o.sweeps < 0 || o.chunk_pairs < 1 || o.slots < 1 || o.slots > 8 || o.search_kernel < 0 || o.search_kernel > 2 ||
      o.search_variant < 0 || o.search_variant > 1 || (o.search_variant == 1 && o.search_kernel == 2))
    return fail(c, BBME_E_ARG, "bbme_plan: bad options (sweeps=%d chunk_pairs=%d slots=%d search_kernel=%d search_variant=%d)",
                o.sweeps, o.chunk_pairs, o.slots, o.search_kernel, o.search_variant);
  if (o.search_variant == 1)
    for (int l = 0; l < sh.num_levels; ++l)
      if (2 * radius_of(sh.search_size[l], sh.block_size[l]) + 1 > 362)
        return fail(c, BBME_E_ARG, "bbme_plan: search_variant 1 supports +-R up to 180 (level %d)", l);
  c->shape = sh;
  c->opt = o;
  const int L = sh.num_levels;
  for (int l = 0; l < L; ++l) {
    c->pitch[l] = round_up(sh.level_width[l] + 4, 64);
    c->plane[l] = (size_t)c->pitch[l] * sh.level_height[l];
    c->cap[l] = (size_t)(sh.level_width[l] / 2) * (sh.level_height[l] / 2);
  }
  c->in_pitch = round_up(width, 16);
  c->in_plane = (size_t)c->in_pitch * height;
  c->out_plane = (size_t)sh.padded_width * sh.padded_height * 2;
  const size_t n = (size_t)o.chunk_pairs;
  c->slots.resize(o.slots);
  for (Slot& s : c->slots) {
    CUDA_TRY(c, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    // image-1 side buffers hold two spare planes: sequence mode keeps n + 1 frames there (+1 for an odd launch slot)
    if ((rc = dev_alloc(c, &s.in1, (n + 2) * c->in_plane, false)) || (rc = dev_alloc(c, &s.in2, n * c->in_plane, false))) return rc;
    for (int l = 0; l < L; ++l) {
      for (int f = 0; f < 2; ++f)
        if ((rc = dev_alloc(c, &s.img[f][l], (f == 0 ? n + 2 : n) * c->plane[l], true))) return rc;
      if ((rc = dev_alloc(c, &s.mv_a[l], n * c->cap[l], true)) || (rc = dev_alloc(c, &s.mv_b[l], n * c->cap[l], true))) return rc;
      if (o.keep_search_mv) {
        const size_t blocks = (size_t)(sh.level_width[l] / sh.block_size[l]) * (sh.level_height[l] / sh.block_size[l]);
        if ((rc = dev_alloc(c, &s.mv_search[l], n * blocks, true))) return rc;
      }
      s.mv_final[l] = s.mv_a[l];
    }
    if ((rc = dev_alloc(c, &s.list0, n * c->cap[0], false)) || (rc = dev_alloc(c, &s.list1, n * c->cap[0], false)) ||
        (rc = dev_alloc(c, &s.nv, n * c->cap[0], false)) || (rc = dev_alloc(c, &s.stamp, n * c->cap[0], true)) ||
        (rc = dev_alloc(c, &s.ctr, n * kCtrWords, true)) || (rc = dev_alloc(c, &s.counters, (size_t)2, true)) ||
        (rc = dev_alloc(c, &s.out, n * (c->out_plane / 4 + 64), false)) || (rc = dev_alloc(c, &s.work_ctr, (size_t)kMaxLevels, true)))
      return rc;
    if (getenv("BBME_REG_PROFILE") && (rc = dev_alloc(c, &s.hist, (size_t)kHistSweeps * 64, true))) return rc;
    for (int l = 0; l < L; ++l) {
      memset(&s.tma[l], 0, sizeof(s.tma[l]));
      memset(&s.tma_seq[l], 0, sizeof(s.tma_seq[l]));
      if (o.search_kernel == 1) continue;
      char msg[256] = {0};
      const int R = radius_of(sh.search_size[l], sh.block_size[l]);
      if (tma_search_wants_pre(sh.level_width[l], sh.level_height[l], sh.block_size[l], R) &&
          (rc = dev_alloc(c, &s.sh4[l], n * 4 * c->plane[l], false)))
        return rc;
      int trc = tma_search_plan(&s.tma[l], s.img[0][l], s.img[1][l], s.sh4[l], sh.level_width[l], sh.level_height[l], c->pitch[l],
                                c->plane[l], o.chunk_pairs, sh.block_size[l], R, msg, sizeof(msg));
      if (trc != 0) return fail(c, BBME_E_CUDA, "TMA search plan failed at level %d: %s", l, msg);
      if (trc == 0)
        trc = tma_search_plan(&s.tma_seq[l], s.img[0][l], s.img[0][l] + c->plane[l], s.sh4[l], sh.level_width[l], sh.level_height[l],
                              c->pitch[l], c->plane[l], o.chunk_pairs, sh.block_size[l], R, msg, sizeof(msg));
      if (trc != 0) return fail(c, BBME_E_CUDA, "TMA search plan (sequence mode) failed at level %d: %s", l, msg);
      if (o.search_kernel == 2 && !s.tma[l].supported)
        return fail(c, BBME_E_ARG, "search_kernel=2 but level %d (block %d, R %d) is not covered by the TMA kernel", l,
                    sh.block_size[l], R);
    }
  }
  c->planned = true;
  if (out) *out = sh;
  return BBME_OK;
}

int bbme_get_shape(const bbme_ctx* c, bbme_shape* out) {
  if (!c || !out) return BBME_E_ARG;
  if (!c->planned) return BBME_E_STATE;
  *out = c->shape;
  return BBME_OK;
}

int bbme_sync(bbme_ctx* c) {
  if (!c) return BBME_E_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  int rc = sync_all(c);
  if (rc) return rc;
  return collect_after_sync(c);
}

int bbme_set_streams(bbme_ctx* c, int n, void* const* streams) {
  if (!c || !streams) return BBME_E_ARG;
  if (!c->planned) return fail(c, BBME_E_STATE, "bbme_set_streams before bbme_plan");
  if (n != (int)c->slots.size()) return fail(c, BBME_E_ARG, "bbme_set_streams: %d streams for %zu slots", n, c->slots.size());
  CUDA_TRY(c, cudaSetDevice(c->device));
  int rc = sync_all(c);
  if (rc) return rc;
  for (int i = 0; i < n; ++i) {
    Slot& s = c->slots[i];
    if (s.stream && s.own_stream) cudaStreamDestroy(s.stream);
    s.stream = reinterpret_cast<cudaStream_t>(streams[i]);
    s.own_stream = false;
  }
  return BBME_OK;
}

int bbme_measure_int_peak(bbme_ctx* c, double* absdiff_per_s, double* sm_mhz) {
  if (!c || !absdiff_per_s) return BBME_E_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  double best = 0.0, mhz = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    double v = 0.0, f = 0.0;
    if (measure_int_peak(c->sm_count, &v, &f) != 0) return fail(c, BBME_E_CUDA, "int-peak micro-benchmark failed");
    if (v > best) { best = v; mhz = f; }
  }
  *absdiff_per_s = best;
  if (sm_mhz) *sm_mhz = mhz;
  return BBME_OK;
}

int bbme_measure_host_link(bbme_ctx* c, size_t bytes, bbme_host_link* out) {
  if (!c || !out || bytes < (1u << 20)) return BBME_E_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  memset(out, 0, sizeof(*out));
  void *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
  cudaStream_t s0 = nullptr, s1 = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = BBME_OK;
  auto time_ms = [&](int mode) -> float {  // 0 = H2D, 1 = D2H, 2 = both directions at once
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaDeviceSynchronize();
      cudaEventRecord(e0, s0);
      cudaStreamWaitEvent(s1, e0, 0);
      if (mode != 1) cudaMemcpyAsync(d_in, h_in, bytes, cudaMemcpyHostToDevice, s0);
      if (mode != 0) cudaMemcpyAsync(h_out, d_out, bytes, cudaMemcpyDeviceToHost, s1);
      cudaEventRecord(e1, s1);
      cudaStreamWaitEvent(s0, e1, 0);
      cudaEventRecord(e1, s0);
      cudaStreamSynchronize(s0);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms > 0.f && ms < best) best = ms;
    }
    return best;
  };
  if (cudaHostAlloc(&h_in, bytes, cudaHostAllocDefault) != cudaSuccess || cudaHostAlloc(&h_out, bytes, cudaHostAllocDefault) != cudaSuccess ||
      cudaMalloc(&d_in, bytes) != cudaSuccess || cudaMalloc(&d_out, bytes) != cudaSuccess ||
      cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&e0) != cudaSuccess ||
      cudaEventCreate(&e1) != cudaSuccess) {
    cudaGetLastError();
    rc = fail(c, BBME_E_NOMEM, "bbme_measure_host_link: allocation of %zu-byte test buffers failed", bytes);
  } else {
    memset(h_in, 1, bytes);
    memset(h_out, 1, bytes);
    const double gb = (double)bytes / 1e9;
    out->h2d_gbs = gb / (time_ms(0) * 1e-3);
    out->d2h_gbs = gb / (time_ms(1) * 1e-3);
    out->duplex_gbs_per_direction = gb / (time_ms(2) * 1e-3);
    HostPool& pool = HostPool::instance();
    out->host_threads = pool.threads();
    out->host_stream_write_gbs = pool.measure_stream_write(h_out, bytes, 3);
    if (cudaGetLastError() != cudaSuccess) rc = fail(c, BBME_E_CUDA, "bbme_measure_host_link: CUDA error during the copies");
  }
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (s0) cudaStreamDestroy(s0);
  if (s1) cudaStreamDestroy(s1);
  if (d_in) cudaFree(d_in);
  if (d_out) cudaFree(d_out);
  if (h_in) cudaFreeHost(h_in);
  if (h_out) cudaFreeHost(h_out);
  return rc;
}

int bbme_debug_set_stamp_epoch(bbme_ctx* c, uint32_t epoch) {
  if (!c) return BBME_E_ARG;
  if (!c->planned) return BBME_E_STATE;
  CUDA_TRY(c, cudaSetDevice(c->device));
  int rc = sync_all(c);
  if (rc) return rc;
  for (Slot& s : c->slots) {
    std::vector<uint32_t> ctr((size_t)c->opt.chunk_pairs * kCtrWords, 0u);
    CUDA_TRY(c, cudaMemcpy(ctr.data(), s.ctr, ctr.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (int p = 0; p < c->opt.chunk_pairs; ++p) ctr[(size_t)p * kCtrWords + CTR_EPOCH] = epoch;
    CUDA_TRY(c, cudaMemcpy(s.ctr, ctr.data(), ctr.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  }
  return BBME_OK;
}

int bbme_debug_search_geometry(int level_width, int level_height, int block_size, int search_size, int allow_copies,
                               bbme_search_geometry* out) {
  if (!out || level_width < 1 || level_height < 1 || !is_pow2(block_size) || search_size < block_size) return BBME_E_ARG;
  static_assert(sizeof(bbme_search_geometry) == sizeof(TmaGeomInfo), "bbme_search_geometry mirrors TmaGeomInfo field by field");
  TmaGeomInfo gi;
  tma_search_geometry(level_width, level_height, block_size, radius_of(search_size, block_size), allow_copies, &gi);
  memcpy(out, &gi, sizeof(gi));
  return BBME_OK;
}

int bbme_debug_div_magic(unsigned divisor, unsigned* magic, unsigned* shift) {
  if (divisor < 2 || !magic || !shift) return BBME_E_ARG;
  uint32_t m, s;
  tma_div_magic(divisor, &m, &s);
  *magic = m;
  *shift = s;
  return BBME_OK;
}

int bbme_debug_skip_compute(bbme_ctx* c, int on) {
  if (!c) return BBME_E_ARG;
  c->skip_compute = on != 0;
  return BBME_OK;
}

int bbme_get_stats(bbme_ctx* c, bbme_stats* out) {
  if (!c || !out) return BBME_E_ARG;
  *out = c->stats;
  out->kernel_launches = c->launches;
  return BBME_OK;
}

static int estimate_batch_async_impl(bbme_ctx* c, int n, int factor, const uint8_t* const* im1, const uint8_t* const* im2,
                                     size_t pitch, float* const* flow) {
  if (!c) return BBME_E_ARG;
  if (!c->planned) return fail(c, BBME_E_STATE, "bbme_estimate_batch before bbme_plan");
  if (factor < 1 || c->shape.width % factor || c->shape.height % factor)
    return fail(c, BBME_E_ARG, "up-sampling factor %d does not divide the planned size %dx%d", factor, c->shape.width, c->shape.height);
  const int w = c->shape.width / factor, h = c->shape.height / factor;  // size of the frames handed in
  if (n <= 0 || !im1 || !im2 || !flow || pitch < (size_t)w) return fail(c, BBME_E_ARG, "bbme_estimate_batch: bad arguments");
  for (int i = 0; i < n; ++i)
    if (!im1[i] || !im2[i] || !flow[i]) return fail(c, BBME_E_ARG, "bbme_estimate_batch: null buffer for pair %d", i);
  CUDA_TRY(c, cudaSetDevice(c->device));
  begin_call(c);
  const int chunk = c->opt.chunk_pairs;
  const size_t in_pitch = factor > 1 ? (size_t)round_up(w, 16) : (size_t)c->in_pitch;
  const size_t in_plane = factor > 1 ? in_pitch * h : c->in_plane;
  const size_t out_plane = factor > 1 ? (size_t)w * h * 2 : c->out_plane;  // floats per pair
  const size_t flow_bytes = out_plane * sizeof(float);
  int ci = c->next_slot;
  for (int start = 0; start < n; start += chunk, ++ci) {
    Slot& s = c->slots[ci % c->slots.size()];
    const int m = (n - start < chunk) ? (n - start) : chunk;
    // frames that follow each other in host memory (an array of frames: plane i + 1 starts where plane i ends) go up as
    // one copy per image instead of one per pair
    auto upload = [&](uint8_t* dst, const uint8_t* const* src) -> cudaError_t {
      bool contiguous = in_plane == in_pitch * (size_t)h;
      for (int i = 1; i < m && contiguous; ++i) contiguous = src[i] == src[i - 1] + pitch * (size_t)h;
      if (contiguous) return cudaMemcpy2DAsync(dst, in_pitch, src[0], pitch, w, (size_t)h * m, cudaMemcpyHostToDevice, s.stream);
      for (int i = 0; i < m; ++i) {
        cudaError_t e = cudaMemcpy2DAsync(dst + (size_t)i * in_plane, in_pitch, src[i], pitch, w, h, cudaMemcpyHostToDevice, s.stream);
        if (e != cudaSuccess) return e;
      }
      return cudaSuccess;
    };
    CUDA_TRY(c, upload(s.in1, im1 + start));
    CUDA_TRY(c, upload(s.in2, im2 + start));
    if (factor > 1) {
      int rc = run_chunk_graphed(c, s, m, s.in1, s.in2, in_pitch, in_plane, s.out, out_plane, nullptr, 0, factor);
      if (rc) return rc;
      for (int i = 0; i < m; ++i)
        CUDA_TRY(c, cudaMemcpyAsync(flow[start + i], s.out + (size_t)i * out_plane, flow_bytes, cudaMemcpyDeviceToHost, s.stream));
    } else {
      int rc = c->skip_compute ? BBME_OK : run_chunk_graphed(c, s, m, s.in1, s.in2, in_pitch, in_plane, nullptr, 0, nullptr, 0, 1);
      if (rc || (rc = enqueue_dense_result(c, s, m, flow + start))) return rc;
    }
  }
  c->next_slot = ci % (int)c->slots.size();
  return BBME_OK;
}

int bbme_estimate_batch_async(bbme_ctx* c, int n, const uint8_t* const* im1, const uint8_t* const* im2, size_t pitch,
                              float* const* flow) {
  return estimate_batch_async_impl(c, n, 1, im1, im2, pitch, flow);
}

int bbme_estimate_upsampled_async(bbme_ctx* c, int n, int factor, const uint8_t* const* im1, const uint8_t* const* im2,
                                  size_t pitch, float* const* flow) {
  if (factor < 2) return c ? fail(c, BBME_E_ARG, "bbme_estimate_upsampled: factor must be 2, 4 or 8") : BBME_E_ARG;
  return estimate_batch_async_impl(c, n, factor, im1, im2, pitch, flow);
}

int bbme_estimate_upsampled(bbme_ctx* c, int n, int factor, const uint8_t* const* im1, const uint8_t* const* im2,
                            size_t pitch, float* const* flow) {
  int rc = bbme_estimate_upsampled_async(c, n, factor, im1, im2, pitch, flow);
  if (rc) return rc;
  rc = sync_all(c);
  if (rc) return rc;
  return collect_after_sync(c);
}

// A video sequence: n_frames frames are n_frames - 1 pairs (t, t + 1).  Chunks of chunk_pairs pairs take chunk_pairs + 1
// frames; consecutive chunks share one frame (uploaded and down-sampled again, 1 / chunk_pairs of the work).
int bbme_estimate_sequence_async(bbme_ctx* c, int n_frames, const uint8_t* const* frames, size_t pitch, float* const* flow) {
  if (!c) return BBME_E_ARG;
  if (!c->planned) return fail(c, BBME_E_STATE, "bbme_estimate_sequence before bbme_plan");
  if (n_frames < 2 || !frames || !flow || pitch < (size_t)c->shape.width) return fail(c, BBME_E_ARG, "bbme_estimate_sequence: bad arguments");
  for (int i = 0; i < n_frames; ++i)
    if (!frames[i] || (i + 1 < n_frames && !flow[i])) return fail(c, BBME_E_ARG, "bbme_estimate_sequence: null buffer at frame %d", i);
  CUDA_TRY(c, cudaSetDevice(c->device));
  begin_call(c);
  const int chunk = c->opt.chunk_pairs, n = n_frames - 1;
  int ci = c->next_slot;
  for (int start = 0; start < n; start += chunk, ++ci) {
    Slot& s = c->slots[ci % c->slots.size()];
    const int m = (n - start < chunk) ? (n - start) : chunk;
    for (int i = 0; i <= m; ++i)
      CUDA_TRY(c, cudaMemcpy2DAsync(s.in1 + (size_t)i * c->in_plane, c->in_pitch, frames[start + i], pitch, c->shape.width,
                                    c->shape.height, cudaMemcpyHostToDevice, s.stream));
    int rc = run_chunk_graphed(c, s, m, s.in1, nullptr, c->in_pitch, c->in_plane, nullptr, 0, nullptr, 0, 1, true);
    if (rc || (rc = enqueue_dense_result(c, s, m, flow + start))) return rc;
  }
  c->next_slot = ci % (int)c->slots.size();
  return BBME_OK;
}

int bbme_estimate_sequence(bbme_ctx* c, int n_frames, const uint8_t* const* frames, size_t pitch, float* const* flow) {
  int rc = bbme_estimate_sequence_async(c, n_frames, frames, pitch, flow);
  if (rc) return rc;
  rc = sync_all(c);
  if (rc) return rc;
  return collect_after_sync(c);
}

int bbme_estimate_sequence_device(bbme_ctx* c, int n_frames, const uint8_t* d_frames, size_t pitch, size_t plane,
                                  float* d_flow, size_t flow_plane) {
  if (!c) return BBME_E_ARG;
  if (!c->planned) return fail(c, BBME_E_STATE, "bbme_estimate_sequence_device before bbme_plan");
  if (n_frames < 2 || !d_frames || !d_flow || pitch < (size_t)c->shape.width || plane < pitch * (size_t)c->shape.height ||
      flow_plane < c->out_plane)
    return fail(c, BBME_E_ARG, "bbme_estimate_sequence_device: bad arguments");
  CUDA_TRY(c, cudaSetDevice(c->device));
  begin_call(c);
  const int chunk = c->opt.chunk_pairs, n = n_frames - 1;
  int ci = 0;
  for (int start = 0; start < n; start += chunk, ++ci) {
    Slot& s = c->slots[ci % c->slots.size()];
    const int m = (n - start < chunk) ? (n - start) : chunk;
    // an even m makes the pad kernel's last launch slot read one frame past frame start + m: stage through the slot
    // buffer when that frame does not exist (end of the sequence)
    const uint8_t* src = d_frames + (size_t)start * plane;
    size_t sp = pitch, spl = plane;
    if ((m & 1) == 0 && start + m + 1 >= n_frames) {
      CUDA_TRY(c, cudaMemcpy2DAsync(s.in1, c->in_pitch, src, pitch, c->shape.width, (size_t)c->shape.height, cudaMemcpyDeviceToDevice, s.stream));
      for (int i = 1; i <= m; ++i)
        CUDA_TRY(c, cudaMemcpy2DAsync(s.in1 + (size_t)i * c->in_plane, c->in_pitch, src + (size_t)i * plane, pitch, c->shape.width,
                                      (size_t)c->shape.height, cudaMemcpyDeviceToDevice, s.stream));
      src = s.in1; sp = c->in_pitch; spl = c->in_plane;
    }
    int rc = run_chunk_graphed(c, s, m, src, nullptr, sp, spl, d_flow + (size_t)start * flow_plane, flow_plane, nullptr, 0, 1, true);
    if (rc) return rc;
  }
  return BBME_OK;
}

int bbme_estimate_batch(bbme_ctx* c, int n, const uint8_t* const* im1, const uint8_t* const* im2, size_t pitch,
                        float* const* flow) {
  int rc = bbme_estimate_batch_async(c, n, im1, im2, pitch, flow);
  if (rc) return rc;
  rc = sync_all(c);
  if (rc) return rc;
  return collect_after_sync(c);
}

int bbme_estimate(bbme_ctx* c, const uint8_t* im1, const uint8_t* im2, size_t pitch, float* flow) {
  return bbme_estimate_batch(c, 1, &im1, &im2, pitch, &flow);
}

static int estimate_device_impl(bbme_ctx* c, int n, const uint8_t* d1, const uint8_t* d2, size_t pitch, size_t plane,
                                float* d_flow, size_t flow_plane, int16_t* d_mv, size_t mv_plane, int factor = 1) {
  if (!c) return BBME_E_ARG;
  if (!c->planned) return fail(c, BBME_E_STATE, "bbme_estimate_device before bbme_plan");
  if (factor < 1 || c->shape.width % factor || c->shape.height % factor)
    return fail(c, BBME_E_ARG, "up-sampling factor %d does not divide the planned size %dx%d", factor, c->shape.width, c->shape.height);
  const size_t fw = (size_t)(c->shape.width / factor), fh = (size_t)(c->shape.height / factor);
  if (n <= 0 || !d1 || !d2 || pitch < fw || plane < pitch * fh)
    return fail(c, BBME_E_ARG, "bbme_estimate_device: bad arguments");
  if (d_flow && flow_plane < (factor > 1 ? fw * fh * 2 : c->out_plane))
    return fail(c, BBME_E_ARG, "bbme_estimate_device: flow plane stride too small");
  if (d_mv && mv_plane < c->cap[0] * 2) return fail(c, BBME_E_ARG, "bbme_estimate_device: mv plane stride too small");
  CUDA_TRY(c, cudaSetDevice(c->device));
  begin_call(c);
  const int chunk = c->opt.chunk_pairs;
  int ci = 0;
  for (int start = 0; start < n; start += chunk, ++ci) {
    Slot& s = c->slots[ci % c->slots.size()];
    const int m = (n - start < chunk) ? (n - start) : chunk;
    int rc = run_chunk_graphed(c, s, m, d1 + (size_t)start * plane, d2 + (size_t)start * plane, pitch, plane,
                       d_flow ? d_flow + (size_t)start * flow_plane : nullptr, flow_plane,
                       d_mv ? d_mv + (size_t)start * mv_plane : nullptr, mv_plane, factor);
    if (rc) return rc;
  }
  return BBME_OK;
}

int bbme_estimate_device(bbme_ctx* c, int n, const uint8_t* d1, const uint8_t* d2, size_t pitch, size_t plane,
                         float* d_flow, size_t flow_plane) {
  if (!d_flow) return c ? fail(c, BBME_E_ARG, "bbme_estimate_device: null flow") : BBME_E_ARG;
  return estimate_device_impl(c, n, d1, d2, pitch, plane, d_flow, flow_plane, nullptr, 0);
}

int bbme_estimate_upsampled_device(bbme_ctx* c, int n, int factor, const uint8_t* d1, const uint8_t* d2, size_t pitch,
                                   size_t plane, float* d_flow, size_t flow_plane) {
  if (!d_flow || factor < 2) return c ? fail(c, BBME_E_ARG, "bbme_estimate_upsampled_device: null flow or factor < 2") : BBME_E_ARG;
  return estimate_device_impl(c, n, d1, d2, pitch, plane, d_flow, flow_plane, nullptr, 0, factor);
}

int bbme_estimate_device_compact(bbme_ctx* c, int n, const uint8_t* d1, const uint8_t* d2, size_t pitch, size_t plane,
                                 int16_t* d_mv, size_t mv_plane) {
  if (!d_mv) return c ? fail(c, BBME_E_ARG, "bbme_estimate_device_compact: null mv") : BBME_E_ARG;
  return estimate_device_impl(c, n, d1, d2, pitch, plane, nullptr, 0, d_mv, mv_plane);
}

int bbme_estimate_device_both(bbme_ctx* c, int n, const uint8_t* d1, const uint8_t* d2, size_t pitch, size_t plane,
                              float* d_flow, size_t flow_plane, int16_t* d_mv, size_t mv_plane) {
  if (!d_flow && !d_mv) return c ? fail(c, BBME_E_ARG, "bbme_estimate_device_both: no output buffer") : BBME_E_ARG;
  return estimate_device_impl(c, n, d1, d2, pitch, plane, d_flow, flow_plane, d_mv, mv_plane);
}

int bbme_expand_compact(const int16_t* mv2, int width2, int height2, float* dense) {
  if (!mv2 || !dense || width2 <= 0 || height2 <= 0) return BBME_E_ARG;
  Ticket t;
  HostPool::instance().expand_async(mv2, width2, height2, dense, (size_t)width2 * 2, &t);
  t.wait();
  return BBME_OK;
}

// ---- peer memory between the one-process-per-GPU ranks of a box (results gathered without a communication kernel) ----
int bbme_device_alloc(int device, size_t bytes, void** p) {
  if (!p || bytes == 0) return BBME_E_ARG;
  *p = nullptr;
  if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(p, bytes) != cudaSuccess) {
    cudaGetLastError();
    return BBME_E_NOMEM;
  }
  return BBME_OK;
}

void bbme_device_free(int device, void* p) {
  if (p && cudaSetDevice(device) == cudaSuccess) cudaFree(p);
}

int bbme_ipc_export(int device, void* p, unsigned char* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (!p || !handle64) return BBME_E_ARG;
  cudaIpcMemHandle_t h;
  if (cudaSetDevice(device) != cudaSuccess || cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
    cudaGetLastError();
    return BBME_E_CUDA;
  }
  memcpy(handle64, &h, 64);
  return BBME_OK;
}

int bbme_ipc_open(int device, const unsigned char* handle64, void** p) {
  if (!p || !handle64) return BBME_E_ARG;
  *p = nullptr;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  // opened with THIS rank's device current: the exporter's memory is mapped with peer access from here, and this process
  // never creates a context on the exporter's GPU (a second context there would time-slice with the exporter's kernels)
  if (cudaSetDevice(device) != cudaSuccess || cudaIpcOpenMemHandle(p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
    cudaGetLastError();
    return BBME_E_CUDA;
  }
  return BBME_OK;
}

int bbme_ipc_close(int device, void* p) {
  if (!p) return BBME_E_ARG;
  if (cudaSetDevice(device) != cudaSuccess || cudaIpcCloseMemHandle(p) != cudaSuccess) {
    cudaGetLastError();
    return BBME_E_CUDA;
  }
  return BBME_OK;
}

int bbme_copy_async(int device, void* dst, const void* src, size_t bytes, void* stream) {
  if (!dst || !src) return BBME_E_ARG;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, reinterpret_cast<cudaStream_t>(stream)) != cudaSuccess) {
    cudaGetLastError();
    return BBME_E_CUDA;
  }
  return BBME_OK;
}

int bbme_host_alloc(void** p, size_t bytes) {
  if (!p) return BBME_E_ARG;
  return cudaHostAlloc(p, bytes, cudaHostAllocDefault) == cudaSuccess ? BBME_OK : BBME_E_NOMEM;
}

void bbme_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------------------------------------ debug state

int bbme_debug_level_image(bbme_ctx* c, int pair, int frame, int level, uint8_t* out) {
  if (!c || !out) return BBME_E_ARG;
  if (!c->planned) return BBME_E_STATE;
  if (pair < 0 || pair >= c->opt.chunk_pairs || frame < 0 || frame > 1 || level < 0 || level >= c->shape.num_levels)
    return fail(c, BBME_E_ARG, "bbme_debug_level_image: index out of range");
  CUDA_TRY(c, cudaSetDevice(c->device));
  Slot& s = c->slots[0];
  CUDA_TRY(c, cudaStreamSynchronize(s.stream));
  const int w = c->shape.level_width[level], h = c->shape.level_height[level];
  CUDA_TRY(c, cudaMemcpy2D(out, w, s.img[frame][level] + (size_t)pair * c->plane[level], c->pitch[level], w, h,
                           cudaMemcpyDeviceToHost));
  return BBME_OK;
}

int bbme_debug_level_mv(bbme_ctx* c, int pair, int level, int which, int16_t* out) {
  if (!c || !out) return BBME_E_ARG;
  if (!c->planned) return BBME_E_STATE;
  if (pair < 0 || pair >= c->opt.chunk_pairs || level < 0 || level >= c->shape.num_levels || which < 0 || which > 1)
    return fail(c, BBME_E_ARG, "bbme_debug_level_mv: index out of range");
  CUDA_TRY(c, cudaSetDevice(c->device));
  Slot& s = c->slots[0];
  CUDA_TRY(c, cudaStreamSynchronize(s.stream));
  const int w = c->shape.level_width[level], h = c->shape.level_height[level];
  if (which == 0) {
    CUDA_TRY(c, cudaMemcpy(out, s.mv_final[level] + (size_t)pair * c->cap[level], c->cap[level] * sizeof(short2),
                           cudaMemcpyDeviceToHost));
  } else {
    if (!s.mv_search[level]) return fail(c, BBME_E_STATE, "plan was made without keep_search_mv");
    const int bs = c->shape.block_size[level];
    const size_t blocks = (size_t)(w / bs) * (h / bs);
    CUDA_TRY(c, cudaMemcpy(out, s.mv_search[level] + (size_t)pair * blocks, blocks * sizeof(short2), cudaMemcpyDeviceToHost));
  }
  return BBME_OK;
}

// ------------------------------------------------------------------------------------------------ single stages

}  // extern "C"

namespace {
struct Scratch {
  std::vector<void*> p;
  ~Scratch() {
    for (void* q : p) cudaFree(q);
  }
  template <typename T>
  T* get(size_t count, bool zero) {
    void* q = nullptr;
    if (cudaMalloc(&q, count * sizeof(T) + 256) != cudaSuccess) return nullptr;
    if (zero) cudaMemset(q, 0, count * sizeof(T) + 256);
    p.push_back(q);
    return reinterpret_cast<T*>(q);
  }
};

int upload_image(bbme_ctx* c, Scratch& sc, const uint8_t* host, int w, int h, int* pitch_out, uint8_t** dev) {
  const int pitch = round_up(w + 4, 64);
  uint8_t* d = sc.get<uint8_t>((size_t)pitch * h, true);
  if (!d) return fail(c, BBME_E_NOMEM, "stage: cudaMalloc failed");
  CUDA_TRY(c, cudaMemcpy2D(d, pitch, host, w, w, h, cudaMemcpyHostToDevice));
  *pitch_out = pitch;
  *dev = d;
  return BBME_OK;
}
}  // namespace

extern "C" {

int bbme_stage_pyrdown(bbme_ctx* c, const uint8_t* src, int w, int h, uint8_t* dst) {
  if (!c || !src || !dst || w < 2 || h < 2) return BBME_E_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  Scratch sc;
  int sp = 0, rc;
  uint8_t* ds = nullptr;
  if ((rc = upload_image(c, sc, src, w, h, &sp, &ds))) return rc;
  const int dw = w / 2, dh = h / 2, dp = round_up(dw + 4, 64);
  uint8_t* dd = sc.get<uint8_t>((size_t)dp * dh * 2, true);
  if (!dd) return fail(c, BBME_E_NOMEM, "stage: cudaMalloc failed");
  ImgView v{ds, w, h, sp, (size_t)sp * h};
  launch_pyrdown(v, v, dd, dd + (size_t)dp * dh, dp, 0, 1, 0);
  CUDA_TRY(c, cudaDeviceSynchronize());
  CUDA_TRY(c, cudaMemcpy2D(dst, dw, dd, dp, dw, dh, cudaMemcpyDeviceToHost));
  return BBME_OK;
}

int bbme_stage_resize(bbme_ctx* c, const uint8_t* src, int w, int h, int factor, uint8_t* dst) {
  if (!c || !src || !dst || w < 1 || h < 1) return BBME_E_ARG;
  ResizeTaps taps;
  if (make_resize_taps(factor, &taps) != 0) return fail(c, BBME_E_ARG, "bbme_stage_resize: factor %d (supported: 2, 4, 8)", factor);
  CUDA_TRY(c, cudaSetDevice(c->device));
  Scratch sc;
  const int sp = round_up(w, 16);
  uint8_t* ds = sc.get<uint8_t>((size_t)sp * h, true);
  const int dw = w * factor, dh = h * factor, dp = round_up(dw + 4, 64);
  uint8_t* dd = sc.get<uint8_t>((size_t)dp * dh * 2, true);
  if (!ds || !dd) return fail(c, BBME_E_NOMEM, "stage: cudaMalloc failed");
  CUDA_TRY(c, cudaMemcpy2D(ds, sp, src, w, w, h, cudaMemcpyHostToDevice));
  launch_resize_pad(ds, ds, sp, 0, w, h, taps, 0, 0, dd, dd + (size_t)dp * dh, dp, 0, dh, 1, 0);
  CUDA_TRY(c, cudaDeviceSynchronize());
  CUDA_TRY(c, cudaMemcpy2D(dst, dw, dd, dp, dw, dh, cudaMemcpyDeviceToHost));
  return BBME_OK;
}

int bbme_stage_search(bbme_ctx* c, const uint8_t* im1, const uint8_t* im2, int w, int h, int bs, int ss, int16_t* mv,
                      int kernel, bbme_stats* st) {
  if (!c || !im1 || !im2 || !mv || !is_pow2(bs) || bs < 2 || w % bs || h % bs) return BBME_E_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  Scratch sc;
  int pitch = 0, rc;
  uint8_t *d1 = nullptr, *d2 = nullptr;
  if ((rc = upload_image(c, sc, im1, w, h, &pitch, &d1)) || (rc = upload_image(c, sc, im2, w, h, &pitch, &d2))) return rc;
  const int gw = w / bs, gh = h / bs, R = radius_of(ss, bs);
  short2* dmv = sc.get<short2>((size_t)gw * gh, false);
  unsigned long long* ctr = sc.get<unsigned long long>(2, true);
  if (!dmv || !ctr) return fail(c, BBME_E_NOMEM, "stage: cudaMalloc failed");
  CUDA_TRY(c, cudaMemcpy(dmv, mv, (size_t)gw * gh * sizeof(short2), cudaMemcpyHostToDevice));
  ImgView i1{d1, w, h, pitch, (size_t)pitch * h}, i2{d2, w, h, pitch, (size_t)pitch * h};
  MvView f{dmv, gw, gh, (size_t)gw * gh};
  TmaSearchPlan plan;
  memset(&plan, 0, sizeof(plan));
  if (kernel != 1) {
    char msg[256] = {0};
    uint8_t* d2s = nullptr;  // the byte-shifted copies, where the planned path would use them too
    if (tma_search_wants_pre(w, h, bs, R) && !(d2s = sc.get<uint8_t>((size_t)pitch * h * 4, false)))
      return fail(c, BBME_E_NOMEM, "stage: cudaMalloc failed");
    if (tma_search_plan(&plan, d1, d2, d2s, w, h, pitch, (size_t)pitch * h, 1, bs, R, msg, sizeof(msg)) != 0)
      return fail(c, BBME_E_CUDA, "stage_search: %s", msg);
    if (plan.supported && plan.pre) launch_shift4(i2, d2s, 1, 0);
    if (kernel == 2 && !plan.supported) return fail(c, BBME_E_ARG, "stage_search: (block %d, R %d) not covered by the TMA kernel", bs, R);
  }
  unsigned int* wc = sc.get<unsigned int>(1, true);
  if (!wc) return fail(c, BBME_E_NOMEM, "stage: cudaMalloc failed");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, 0);
  if (plan.supported) {
    if (launch_search_tma(plan, i1, i2, f, 1, ctr, wc, c->sm_count, 0) != 0) return fail(c, BBME_E_CUDA, "stage_search: TMA kernel launch failed");
  } else {
    launch_search_generic(i1, i2, f, bs, R, 1, ctr, 0);
  }
  cudaEventRecord(e1, 0);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (e != cudaSuccess) return fail(c, BBME_E_CUDA, "stage_search kernel failed: %s", cudaGetErrorString(e));
  CUDA_TRY(c, cudaMemcpy(mv, dmv, (size_t)gw * gh * sizeof(short2), cudaMemcpyDeviceToHost));
  if (st) {
    memset(st, 0, sizeof(*st));
    unsigned long long hc[2];
    CUDA_TRY(c, cudaMemcpy(hc, ctr, sizeof(hc), cudaMemcpyDeviceToHost));
    st->ms_search = ms;
    st->ms_total = ms;
    st->kernel_launches = 1;
    st->search_candidates = hc[0];
    st->search_absdiffs = hc[1];
    st->search_kernel_used = plan.supported ? 2u : 1u;
    st->search_launches = 1;
  }
  return BBME_OK;
}

int bbme_stage_search_raster(bbme_ctx* c, const uint8_t* im1, const uint8_t* im2, int w, int h, int bs, int ss, int16_t* mv) {
  if (!c || !im1 || !im2 || !mv || !is_pow2(bs) || bs < 2 || w % bs || h % bs) return BBME_E_ARG;
  const int R = radius_of(ss, bs);
  if (2 * R + 1 > 362) return fail(c, BBME_E_ARG, "stage_search_raster: +-R up to 180");
  CUDA_TRY(c, cudaSetDevice(c->device));
  Scratch sc;
  int pitch = 0, rc;
  uint8_t *d1 = nullptr, *d2 = nullptr;
  if ((rc = upload_image(c, sc, im1, w, h, &pitch, &d1)) || (rc = upload_image(c, sc, im2, w, h, &pitch, &d2))) return rc;
  const int gw = w / bs, gh = h / bs;
  short2* dmv = sc.get<short2>((size_t)gw * gh, false);
  if (!dmv) return fail(c, BBME_E_NOMEM, "stage: cudaMalloc failed");
  CUDA_TRY(c, cudaMemcpy(dmv, mv, (size_t)gw * gh * sizeof(short2), cudaMemcpyHostToDevice));
  ImgView i1{d1, w, h, pitch, (size_t)pitch * h}, i2{d2, w, h, pitch, (size_t)pitch * h};
  MvView f{dmv, gw, gh, (size_t)gw * gh};
  launch_search_generic(i1, i2, f, bs, R, 1, nullptr, 0, 1);
  CUDA_TRY(c, cudaDeviceSynchronize());
  CUDA_TRY(c, cudaMemcpy(mv, dmv, (size_t)gw * gh * sizeof(short2), cudaMemcpyDeviceToHost));
  return BBME_OK;
}

int bbme_stage_compensate(bbme_ctx* c, const uint8_t* im2, int w, int h, int bs, const int16_t* mv, uint8_t* out) {
  if (!c || !im2 || !mv || !out || !is_pow2(bs) || bs < 2 || w % bs || h % bs) return BBME_E_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  Scratch sc;
  int pitch = 0, rc;
  uint8_t* d2 = nullptr;
  if ((rc = upload_image(c, sc, im2, w, h, &pitch, &d2))) return rc;
  const int gw = w / bs, gh = h / bs;
  short2* dmv = sc.get<short2>((size_t)gw * gh, false);
  uint8_t* dout = sc.get<uint8_t>((size_t)pitch * h, true);
  if (!dmv || !dout) return fail(c, BBME_E_NOMEM, "stage: cudaMalloc failed");
  CUDA_TRY(c, cudaMemcpy(dmv, mv, (size_t)gw * gh * sizeof(short2), cudaMemcpyHostToDevice));
  ImgView i2{d2, w, h, pitch, (size_t)pitch * h};
  launch_compensate(i2, dmv, gw, (size_t)gw * gh, bs, dout, pitch, (size_t)pitch * h, 1, 0);
  CUDA_TRY(c, cudaDeviceSynchronize());
  CUDA_TRY(c, cudaMemcpy2D(out, w, dout, pitch, w, h, cudaMemcpyDeviceToHost));
  return BBME_OK;
}

int bbme_stage_regularize(bbme_ctx* c, const uint8_t* im1, const uint8_t* im2, int w, int h, int bs, float lambda,
                          int mult, int16_t* mv, uint32_t* rounds_out) {
  if (!c || !im1 || !im2 || !mv || !is_pow2(bs) || bs < 2 || w % bs || h % bs || w / bs < 2 || h / bs < 2) return BBME_E_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  Scratch sc;
  int pitch = 0, rc;
  uint8_t *d1 = nullptr, *d2 = nullptr;
  if ((rc = upload_image(c, sc, im1, w, h, &pitch, &d1)) || (rc = upload_image(c, sc, im2, w, h, &pitch, &d2))) return rc;
  const int gw = w / bs, gh = h / bs;
  const size_t nb = (size_t)gw * gh;
  short2* O = sc.get<short2>(nb, false);
  short2* Y = sc.get<short2>(nb, false);
  uint32_t* l0 = sc.get<uint32_t>(nb, false);
  uint32_t* l1 = sc.get<uint32_t>(nb, false);
  short2* nv = sc.get<short2>(nb, false);
  uint32_t* stamp = sc.get<uint32_t>(nb, true);
  uint32_t* ctr = sc.get<uint32_t>(kCtrWords, true);
  if (!O || !Y || !l0 || !l1 || !nv || !stamp || !ctr) return fail(c, BBME_E_NOMEM, "stage: cudaMalloc failed");
  CUDA_TRY(c, cudaMemcpy(O, mv, nb * sizeof(short2), cudaMemcpyHostToDevice));
  RegArgs ra;
  ra.i1 = ImgView{d1, w, h, pitch, (size_t)pitch * h};
  ra.i2 = ImgView{d2, w, h, pitch, (size_t)pitch * h};
  ra.bs = bs; ra.gw = gw; ra.gh = gh;
  ra.lm = lambda * (float)mult;
  ra.O = O; ra.Y = Y; ra.mv_plane = nb;
  ra.list0 = l0; ra.list1 = l1; ra.nv = nv; ra.stamp = stamp; ra.wl_plane = nb; ra.ctr = ctr;
  ra.lm = 0.f;
  if (launch_reg_level(ra, 1, lambda, mult, 1, 1, c->sm_count, 0) != 0)
    return fail(c, BBME_E_CUDA, "stage_regularize launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return fail(c, BBME_E_CUDA, "stage_regularize kernel failed: %s", cudaGetErrorString(e));
  CUDA_TRY(c, cudaMemcpy(mv, Y, nb * sizeof(short2), cudaMemcpyDeviceToHost));
  if (rounds_out) {
    uint32_t hc[kCtrWords];
    CUDA_TRY(c, cudaMemcpy(hc, ctr, sizeof(hc), cudaMemcpyDeviceToHost));
    *rounds_out = hc[CTR_ROUNDS];
  }
  return BBME_OK;
}

int bbme_stage_divide(bbme_ctx* c, const int16_t* mv_in, int gw, int gh, int16_t* mv_out) {
  if (!c || !mv_in || !mv_out || gw < 1 || gh < 1) return BBME_E_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  Scratch sc;
  const size_t nb = (size_t)gw * gh;
  short2* a = sc.get<short2>(nb, false);
  short2* b = sc.get<short2>(nb * 4, false);
  if (!a || !b) return fail(c, BBME_E_NOMEM, "stage: cudaMalloc failed");
  CUDA_TRY(c, cudaMemcpy(a, mv_in, nb * sizeof(short2), cudaMemcpyHostToDevice));
  launch_divide(a, gw, gh, nb, b, nb * 4, 1, 0);
  CUDA_TRY(c, cudaDeviceSynchronize());
  CUDA_TRY(c, cudaMemcpy(mv_out, b, nb * 4 * sizeof(short2), cudaMemcpyDeviceToHost));
  return BBME_OK;
}

int bbme_stage_copy_mvs(bbme_ctx* c, const int16_t* coarse_mv, int cw, int ch, int cbs, int fbs, int16_t* fine_pred) {
  if (!c || !coarse_mv || !fine_pred || !is_pow2(cbs) || !is_pow2(fbs) || cbs < 2 || fbs < 2 || cw % cbs || ch % cbs ||
      (2 * cw) % fbs || (2 * ch) % fbs)
    return BBME_E_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  Scratch sc;
  const size_t nc = (size_t)(cw / 2) * (ch / 2);
  const int fgw = 2 * cw / fbs, fgh = 2 * ch / fbs;
  short2* a = sc.get<short2>(nc, false);
  short2* b = sc.get<short2>((size_t)fgw * fgh, false);
  if (!a || !b) return fail(c, BBME_E_NOMEM, "stage: cudaMalloc failed");
  CUDA_TRY(c, cudaMemcpy(a, coarse_mv, nc * sizeof(short2), cudaMemcpyHostToDevice));
  MvView f{b, fgw, fgh, (size_t)fgw * fgh};
  launch_copy_mvs(a, cw / 2, nc, cbs, f, fbs, 1, 0);
  CUDA_TRY(c, cudaDeviceSynchronize());
  CUDA_TRY(c, cudaMemcpy(fine_pred, b, (size_t)fgw * fgh * sizeof(short2), cudaMemcpyDeviceToHost));
  return BBME_OK;
}

}  // extern "C"
