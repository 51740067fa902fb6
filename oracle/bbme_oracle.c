/*
 * bbme_oracle.c -- CPU oracle for the block-matching hot path (TEST INFRASTRUCTURE ONLY).
 *
 * A plain-C restatement of what the reference computes on the path
 *   MF::MF -> MF::calcMotionBlockMatching  (motion_framework.cpp:4-219)
 * with the reference's loop order kept literal where the order is observable: the spiral walk
 * (strict '<' => earliest visited minimum wins), the in-place raster regularisation sweep, and the
 * un-fused float32 energy.  Build with -O2 -ffp-contract=off (GCC would otherwise contract
 * a + b*c into an FMA; MSVC /fp:precise did not).
 *
 * Arithmetic that lives in OpenCV (not vendored by the reference; 2.4.9 / 3.0.0 per its .props files)
 * is restated here from the published algorithm:
 *   cv::copyMakeBorder(BORDER_CONSTANT,0)  -> orc_pad_image
 *   cv::pyrDown(8-bit)                     -> orc_pyrdown: separable [1 4 6 4 1], BORDER_REFLECT_101,
 *                                             (sum + 128) >> 8
 *   cv::norm(a, b, NORM_L1) on CV_8UC1     -> exact integer sum of |a-b|
 *   cv::resize(INTER_LINEAR, 8-bit)        -> orc_resize_linear (main()'s x4 up-sampling, main_class.cpp:32-33):
 *                                             11-bit weights, int32 horizontal pass, two-shift vertical pass
 *
 * Parity pinning (what this oracle has been checked against):
 *   - oracle/_ref: the reference's own motion_framework.cpp / rw_flow.cpp compiled where they lie,
 *     against oracle/cvshim (a minimal own implementation of the cv:: subset they use); dense fields
 *     are compared bit-for-bit in tests/test_oracle.py and frozen in tests/golden/mf_reference.npz.
 *   - cv2 4.13.0 (this container's Python wheel) for pyrDown / copyMakeBorder / norm, frozen in
 *     tests/golden/pyrdown_*.npz by tests/golden/make_golden.py.
 *   - cv2 4.13.0 for resize(INTER_LINEAR) at factors 2 / 4 / 8 (tests/golden/resize_cv2.npz by
 *     tests/golden/make_resize_golden.py, plus a live cv2 comparison in tests/test_oracle.py).
 *   - the 8 Middlebury gt-flow .flo files for the .flo codec (byte-identical round trip).
 * The reference ships no golden motion fields of its own (it has no tests).
 *
 * Deliberately not restated: fast_array (motion_framework.cpp:77-78,286,414,594-602) -- a memo whose
 * hit condition is "identical position and block size", so it cannot change a result; oracle/_ref
 * keeps it, and the two agree.
 */
#include "bbme_oracle.h"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------------------ shape (motion_framework.cpp:15-54) */

int orc_plan_shape(int w, int h, int levels, const int* block_size, orc_shape* out) {
  if (w <= 0 || h <= 0 || levels <= 0 || levels > ORC_MAX_LEVELS || !block_size || !out) return ORC_E_ARG;
  for (int i = 0; i < levels; ++i)
    if (block_size[i] <= 0) return ORC_E_ARG;
  /* The reference searches in doubles with fmod/pow; every operand is an integer below 2^53, so the
   * search is exact.  It bumps height and width independently and aborts when either reaches twice
   * the original -- tested BEFORE the divisibility test (:21-26). */
  double th = (double)h, tw = (double)w;
  for (;;) {
    if (th == 2.0 * h || tw == 2.0 * w) return ORC_E_NOPAD;
    double rem_h = 0.0, rem_w = 0.0;
    for (int i = 0; i < levels; ++i) {
      double q = pow(2.0, (double)i) * (double)block_size[i];
      rem_h += fmod(th, q);
      rem_w += fmod(tw, q);
    }
    if (rem_h == 0.0 && rem_w == 0.0) break;
    if (rem_h != 0.0) th += 1.0;
    if (rem_w != 0.0) tw += 1.0;
  }
  memset(out, 0, sizeof(*out));
  out->padded_h = (int)th;
  out->padded_w = (int)tw;
  out->pad_x = ((int)tw - w) / 2;
  out->pad_y = ((int)th - h) / 2;
  out->num_levels = levels;
  /* the padded image really is orig + 2*pad (:57-61); with an odd difference it is one short and the
   * block grid no longer tiles it -> reference reads out of bounds. */
  if (w + 2 * out->pad_x != out->padded_w || h + 2 * out->pad_y != out->padded_h) return ORC_E_ODD_PAD;
  int lw = out->padded_w, lh = out->padded_h;
  for (int i = 0; i < levels; ++i) {
    out->level_w[i] = lw;
    out->level_h[i] = lh;
    if (lw / block_size[i] < 2 || lh / block_size[i] < 2) return ORC_E_ONE_BLOCK;
    lw /= 2;
    lh /= 2;
  }
  return ORC_OK;
}

/* ------------------------------------------------------------------ pad + pyramid */

void orc_pad_image(const uint8_t* src, int w, int h, size_t pitch, int pad_x, int pad_y, uint8_t* dst) {
  int pw = w + 2 * pad_x, ph = h + 2 * pad_y;
  memset(dst, 0, (size_t)pw * (size_t)ph);
  for (int y = 0; y < h; ++y) memcpy(dst + (size_t)(y + pad_y) * pw + pad_x, src + (size_t)y * pitch, (size_t)w);
}

static inline int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * len - 2 - p;
  }
  return p;
}

void orc_pyrdown(const uint8_t* src, int sw, int sh, uint8_t* dst) {
  static const int k5[5] = {1, 4, 6, 4, 1};
  int dw = sw / 2, dh = sh / 2;
  int* row = (int*)malloc(sizeof(int) * (size_t)dw * 5);
  /* horizontal pass for the 5 source rows feeding one output row, then the vertical pass; the 2-D sum
   * is exact in int and rounded once, so the split does not matter. */
  for (int y = 0; y < dh; ++y) {
    for (int j = 0; j < 5; ++j) {
      const uint8_t* s = src + (size_t)reflect101(2 * y + j - 2, sh) * sw;
      int* r = row + (size_t)j * dw;
      for (int x = 0; x < dw; ++x) {
        int acc = 0;
        for (int i = 0; i < 5; ++i) acc += k5[i] * (int)s[reflect101(2 * x + i - 2, sw)];
        r[x] = acc;
      }
    }
    for (int x = 0; x < dw; ++x) {
      int acc = 0;
      for (int j = 0; j < 5; ++j) acc += k5[j] * row[(size_t)j * dw + x];
      dst[(size_t)y * dw + x] = (uint8_t)((acc + 128) >> 8);
    }
  }
  free(row);
}

/* ------------------------------------------------------------------ SAD = (int)cv::norm(roi1, roi2, NORM_L1) */

static inline int sad_block(const uint8_t* a, const uint8_t* b, int pitch, int bs) {
  int s = 0;
  for (int r = 0; r < bs; ++r) {
    const uint8_t* pa = a + (size_t)r * pitch;
    const uint8_t* pb = b + (size_t)r * pitch;
    for (int c = 0; c < bs; ++c) {
      int d = (int)pa[c] - (int)pb[c];
      s += d < 0 ? -d : d;
    }
  }
  return s;
}

/* ------------------------------------------------------------------ spiral (motion_framework.cpp:296-422) */

int orc_spiral_walk(int shift, int* dxdy, int cap_pairs) {
  int n = 0, l = 0, k = 0, m, t;
#define EMIT()                         \
  do {                                 \
    if (n < cap_pairs) {               \
      dxdy[2 * n] = l;                 \
      dxdy[2 * n + 1] = k;             \
    }                                  \
    ++n;                               \
  } while (0)
  EMIT();
  for (m = 1; m < shift; m += 2) {
    for (t = 0; t < m; ++t) { l += 1; EMIT(); }
    for (t = 0; t < m; ++t) { k += 1; EMIT(); }
    for (t = 0; t < m + 1; ++t) { l -= 1; EMIT(); }
    for (t = 0; t < m + 1; ++t) { k -= 1; EMIT(); }
  }
  for (t = 0; t < m - 1; ++t) { l += 1; EMIT(); }
#undef EMIT
  return n;
}

int orc_spiral_rank(int dx, int dy) {
  int ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy;
  int r = ax > ay ? ax : ay;
  if (r == 0) return 0;
  int base = (2 * r - 1) * (2 * r - 1);
  if (dx == r && dy > -r) return base + dy + r - 1;      /* right column, going down */
  if (dy == r) return base + 2 * r + (r - 1 - dx);       /* bottom row, going left */
  if (dx == -r) return base + 4 * r + (r - 1 - dy);      /* left column, going up */
  return base + 6 * r + (dx + r - 1);                    /* top row, going right */
}

typedef struct { int x, y; } pos2i;

static pos2i spiral_search(const uint8_t* im1, const uint8_t* im2, int w, int h, int bs, int ss, int y1, int x1,
                           int y2, int x2, orc_stats* st, int level) {
  pos2i best;
  int shift = ss - bs;
  if (x2 < 0 || y2 < 0 || x2 + bs > w || y2 + bs > h) { /* :304-310 -> MV becomes 0 */
    best.x = x1;
    best.y = y1;
    return best;
  }
  const uint8_t* a = im1 + (size_t)y1 * w + x1;
  uint64_t calls = 1;
  int min_x = x2, min_y = y2;
  int sad_min = sad_block(a, im2 + (size_t)y2 * w + x2, w, bs);
  int l = x2, k = y2, m, t;
#define VISIT()                                                      \
  do {                                                               \
    if (l < 0 || k < 0 || l + bs > w || k + bs > h) break;           \
    int s_ = sad_block(a, im2 + (size_t)k * w + l, w, bs);           \
    ++calls;                                                         \
    if (s_ < sad_min) { sad_min = s_; min_x = l; min_y = k; }        \
  } while (0)
  for (m = 1; m < shift; m += 2) {
    for (t = 0; t < m; ++t) { l += 1; VISIT(); }
    for (t = 0; t < m; ++t) { k += 1; VISIT(); }
    for (t = 0; t < m + 1; ++t) { l -= 1; VISIT(); }
    for (t = 0; t < m + 1; ++t) { k -= 1; VISIT(); }
  }
  for (t = 0; t < m - 1; ++t) { l += 1; VISIT(); }
#undef VISIT
  if (st) {
    st->search_sad_calls += calls;
    st->search_absdiffs += calls * (uint64_t)(bs * bs);
    if (level >= 0 && level < ORC_MAX_LEVELS) st->level_search_absdiffs[level] += calls * (uint64_t)(bs * bs);
  }
  best.x = min_x;
  best.y = min_y;
  return best;
}

void orc_search_level(const uint8_t* im1, const uint8_t* im2, int w, int h, int bs, int ss, float* flow,
                      orc_stats* st, int level) {
  for (int i = 0; i < h; i += bs) {
    for (int j = 0; j < w; j += bs) {
      float* f = flow + ((size_t)i * w + j) * 2;
      int x2 = j + (int)f[0];
      int y2 = i + (int)f[1];
      pos2i r = spiral_search(im1, im2, w, h, bs, ss, i, j, y2, x2, st, level);
      f[0] = (float)r.x - (float)j;
      f[1] = (float)r.y - (float)i;
    }
  }
}

/* MF::find_min_block (motion_framework.cpp:246-294): the raster-scan full search that calcLevelBM's commented line :235 would
 * call instead of the spiral search.  The window is clamped to the image (:260,262; no centre test), the scan is row-major,
 * a strictly smaller SAD wins (:271), and at equal SAD the candidate with the smaller L1 distance to the block's position in
 * image 1 (:278).  With an empty window the predicted position is returned unchanged (:251-252). */
static pos2i raster_search(const uint8_t* im1, const uint8_t* im2, int w, int h, int bs, int ss, int y1, int x1, int y2,
                           int x2) {
  int start_pos = (ss - bs) >> 1;
  int sad_min = INT_MAX, l1_dist = INT_MAX;
  pos2i best;
  best.x = x2;
  best.y = y2;
  int k0 = y2 - start_pos > 0 ? y2 - start_pos : 0, k1 = h - bs + 1 < y2 + start_pos + 1 ? h - bs + 1 : y2 + start_pos + 1;
  int l0 = x2 - start_pos > 0 ? x2 - start_pos : 0, l1 = w - bs + 1 < x2 + start_pos + 1 ? w - bs + 1 : x2 + start_pos + 1;
  for (int k = k0; k < k1; ++k) {
    for (int l = l0; l < l1; ++l) {
      int sad = sad_block(im1 + (size_t)y1 * w + x1, im2 + (size_t)k * w + l, w, bs);
      int dist = abs(x1 - l) + abs(y1 - k);
      if (sad < sad_min) {
        sad_min = sad; best.x = l; best.y = k; l1_dist = dist;
      } else if (sad == sad_min && dist < l1_dist) {
        best.x = l; best.y = k; l1_dist = dist;
      }
    }
  }
  return best;
}

void orc_search_level_raster(const uint8_t* im1, const uint8_t* im2, int w, int h, int bs, int ss, float* flow) {
  for (int i = 0; i < h; i += bs) {
    for (int j = 0; j < w; j += bs) { /* calcLevelBM, :229-239, with :235 instead of :236 */
      float* f = flow + ((size_t)i * w + j) * 2;
      int x2 = j + (int)f[0];
      int y2 = i + (int)f[1];
      pos2i r = raster_search(im1, im2, w, h, bs, ss, i, j, y2, x2);
      f[0] = (float)r.x - (float)j;
      f[1] = (float)r.y - (float)i;
    }
  }
}

/* MF::draw_MVimage (motion_framework.cpp:887-905): the motion-compensated frame -- every block of image 2 at (block position +
 * vector) copied to the block's position; blocks whose source leaves the image are skipped (out keeps the caller's bytes). */
void orc_compensate(const uint8_t* im2, int w, int h, int bs, const float* flow, uint8_t* out) {
  for (int i = 0; i < h; i += bs) {
    for (int j = 0; j < w; j += bs) {
      int x2 = j + (int)flow[((size_t)i * w + j) * 2];
      int y2 = i + (int)flow[((size_t)i * w + j) * 2 + 1];
      if (x2 < 0 || x2 > w - bs || y2 < 0 || y2 > h - bs) continue;
      for (int r = 0; r < bs; ++r) memcpy(out + (size_t)(i + r) * w + j, im2 + (size_t)(y2 + r) * w + x2, (size_t)bs);
    }
  }
}

/* ------------------------------------------------------------------ regularisation (motion_framework.cpp:424-662) */

typedef struct { float u, v; } mv2f;

static inline mv2f at(const float* flow, int w, int y, int x) {
  mv2f r;
  r.u = flow[((size_t)y * w + x) * 2];
  r.v = flow[((size_t)y * w + x) * 2 + 1];
  return r;
}

static float smoothness(int cur, const mv2f* c, int n) { /* :623-644 */
  float cost = 0.0f;
  for (int i = 0; i < n; ++i) cost += fabsf(c[i].u - c[cur].u) + fabsf(c[i].v - c[cur].v);
  return cost;
}

static void pick_candidate(const uint8_t* im1, const uint8_t* im2, int w, int h, int bs, float lambda, int mult,
                           int x1, int y1, const mv2f* c, int n, float* flow, orc_stats* st, int level) {
  float energy[9];
  for (int i = 0; i < n; ++i) {
    float px = (float)x1 + c[i].u, py = (float)y1 + c[i].v; /* pos2 = pos1 + candidates[i] */
    int ix = (int)px, iy = (int)py;
    if (ix < 0 || ix > w - bs || iy < 0 || iy > h - bs) { /* :578-582 */
      energy[i] = FLT_MAX;
      continue;
    }
    int sad = sad_block(im1 + (size_t)y1 * w + x1, im2 + (size_t)iy * w + ix, w, bs);
    if (st) {
      st->reg_sad_calls += 1;
      st->reg_absdiffs += (uint64_t)(bs * bs);
      if (level >= 0 && level < ORC_MAX_LEVELS) st->level_reg_absdiffs[level] += (uint64_t)(bs * bs);
    }
    float s = smoothness(i, c, n);
    float lm = lambda * (float)mult; /* :607, left to right */
    float prod = lm * s;
    energy[i] = (float)sad + prod;
  }
  int min_pos = 0; /* :646-662 */
  float min_val = energy[0];
  for (int i = 1; i < n; ++i)
    if (energy[i] < min_val) { min_val = energy[i]; min_pos = i; }
  flow[((size_t)y1 * w + x1) * 2] = c[min_pos].u; /* in place, :616 */
  flow[((size_t)y1 * w + x1) * 2 + 1] = c[min_pos].v;
}

void orc_regularize_sweep(const uint8_t* im1, const uint8_t* im2, int w, int h, int bs, float lambda, int mult,
                          float* flow, orc_stats* st, int level) {
  mv2f c[9];
  for (int i = 0; i < h; i += bs) {
    for (int j = 0; j < w; j += bs) {
      int n = 0;
      int up = i - bs >= 0, dn = i + bs < h, lf = j - bs >= 0, rt = j + bs < w;
      /* the reference's nine-way if/else chain, :438-522, branch for branch */
      if (up && lf && rt && dn) {
        c[n++] = at(flow, w, i, j);
        c[n++] = at(flow, w, i, j - bs);
        c[n++] = at(flow, w, i, j + bs);
        c[n++] = at(flow, w, i + bs, j + bs);
        c[n++] = at(flow, w, i - bs, j - bs);
        c[n++] = at(flow, w, i - bs, j + bs);
        c[n++] = at(flow, w, i - bs, j);
        c[n++] = at(flow, w, i + bs, j);
        c[n++] = at(flow, w, i + bs, j - bs);
      } else if (lf && rt && i == 0) { /* top row */
        c[n++] = at(flow, w, i, j);
        c[n++] = at(flow, w, i, j - bs);
        c[n++] = at(flow, w, i, j + bs);
        c[n++] = at(flow, w, i + bs, j + bs);
        c[n++] = at(flow, w, i + bs, j);
        c[n++] = at(flow, w, i + bs, j - bs);
      } else if (lf && rt && i == h - bs) { /* bottom row */
        c[n++] = at(flow, w, i, j);
        c[n++] = at(flow, w, i, j - bs);
        c[n++] = at(flow, w, i, j + bs);
        c[n++] = at(flow, w, i - bs, j - bs);
        c[n++] = at(flow, w, i - bs, j + bs);
        c[n++] = at(flow, w, i - bs, j);
      } else if (j == 0 && up && dn) { /* left column */
        c[n++] = at(flow, w, i, j);
        c[n++] = at(flow, w, i, j + bs);
        c[n++] = at(flow, w, i + bs, j + bs);
        c[n++] = at(flow, w, i - bs, j + bs);
        c[n++] = at(flow, w, i - bs, j);
        c[n++] = at(flow, w, i + bs, j);
      } else if (j == w - bs && up && dn) { /* right column */
        c[n++] = at(flow, w, i, j);
        c[n++] = at(flow, w, i, j - bs);
        c[n++] = at(flow, w, i - bs, j - bs);
        c[n++] = at(flow, w, i - bs, j);
        c[n++] = at(flow, w, i + bs, j);
        c[n++] = at(flow, w, i + bs, j - bs);
      } else if (i == 0 && j == 0) { /* top-left */
        c[n++] = at(flow, w, i, j);
        c[n++] = at(flow, w, i, j + bs);
        c[n++] = at(flow, w, i + bs, j + bs);
        c[n++] = at(flow, w, i + bs, j);
      } else if (i == 0) { /* top-right */
        c[n++] = at(flow, w, i, j);
        c[n++] = at(flow, w, i, j - bs);
        c[n++] = at(flow, w, i + bs, j);
        c[n++] = at(flow, w, i + bs, j - bs);
      } else if (j == 0) { /* bottom-left */
        c[n++] = at(flow, w, i, j);
        c[n++] = at(flow, w, i, j + bs);
        c[n++] = at(flow, w, i - bs, j + bs);
        c[n++] = at(flow, w, i - bs, j);
      } else { /* bottom-right */
        c[n++] = at(flow, w, i, j);
        c[n++] = at(flow, w, i, j - bs);
        c[n++] = at(flow, w, i - bs, j - bs);
        c[n++] = at(flow, w, i - bs, j);
      }
      pick_candidate(im1, im2, w, h, bs, lambda, mult, j, i, c, n, flow, st, level);
    }
  }
}

void orc_divide_blocks(int w, int h, int bs, float* flow) {
  int hb = bs >> 1;
  for (int i = 0; i < h; i += bs)
    for (int j = 0; j < w; j += bs) {
      mv2f m = at(flow, w, i, j);
      float* p;
      p = flow + ((size_t)(i + hb) * w + j) * 2; p[0] = m.u; p[1] = m.v;
      p = flow + ((size_t)i * w + j + hb) * 2; p[0] = m.u; p[1] = m.v;
      p = flow + ((size_t)(i + hb) * w + j + hb) * 2; p[0] = m.u; p[1] = m.v;
    }
}

static void fill_block(float* flow, int w, int i, int j, int bs, mv2f m) { /* :803-813 */
  for (int k = i; k < i + bs; ++k)
    for (int l = j; l < j + bs; ++l) {
      flow[((size_t)k * w + l) * 2] = m.u;
      flow[((size_t)k * w + l) * 2 + 1] = m.v;
    }
}

void orc_copy_mvs(const float* coarse, int cw, int ch, int cbs, float* fine) {
  int fw = 2 * cw;
  for (int i = 0; i < ch; i += cbs)
    for (int j = 0; j < cw; j += cbs) {
      mv2f m = at(coarse, cw, i, j);
      m.u *= 2.0f;
      m.v *= 2.0f;
      fill_block(fine, fw, i << 1, j << 1, cbs << 1, m);
    }
}

void orc_copy_to_all_pixels(int w, int h, int bs, float* flow) {
  for (int i = 0; i < h; i += bs)
    for (int j = 0; j < w; j += bs) fill_block(flow, w, i, j, bs, at(flow, w, i, j));
}

/* ------------------------------------------------------------------ whole pair */

static int estimate_impl(const uint8_t* im1, const uint8_t* im2, int w, int h, size_t pitch, int levels,
                         const int* search_size, const int* block_size, int sweeps, float* flow_out, orc_stats* st,
                         uint8_t* const* pyr1, uint8_t* const* pyr2, float* const* after_search,
                         float* const* after_reg, int raster_search_variant) {
  orc_shape sh;
  orc_stats local;
  if (!st) st = &local;
  memset(st, 0, sizeof(*st));
  if (!im1 || !im2 || !search_size || !flow_out || sweeps < 0) return ORC_E_ARG;
  int rc = orc_plan_shape(w, h, levels, block_size, &sh);
  if (rc != ORC_OK) return rc;

  double t0 = now_s();
  uint8_t* i1[ORC_MAX_LEVELS] = {0};
  uint8_t* i2[ORC_MAX_LEVELS] = {0};
  float* fl[ORC_MAX_LEVELS] = {0};
  rc = ORC_OK;
  for (int l = 0; l < levels; ++l) {
    size_t px = (size_t)sh.level_w[l] * sh.level_h[l];
    i1[l] = (uint8_t*)malloc(px);
    i2[l] = (uint8_t*)malloc(px);
    fl[l] = (float*)calloc(px * 2, sizeof(float)); /* cv::Mat::zeros, :70,92 */
    if (!i1[l] || !i2[l] || !fl[l]) rc = ORC_E_NOMEM;
  }
  if (rc == ORC_OK) {
    orc_pad_image(im1, w, h, pitch, sh.pad_x, sh.pad_y, i1[0]);
    orc_pad_image(im2, w, h, pitch, sh.pad_x, sh.pad_y, i2[0]);
    for (int l = 1; l < levels; ++l) {
      orc_pyrdown(i1[l - 1], sh.level_w[l - 1], sh.level_h[l - 1], i1[l]);
      orc_pyrdown(i2[l - 1], sh.level_w[l - 1], sh.level_h[l - 1], i2[l]);
    }
    st->t_ctor_s = now_s() - t0;
    for (int l = 0; l < levels; ++l) {
      size_t px = (size_t)sh.level_w[l] * sh.level_h[l];
      if (pyr1 && pyr1[l]) memcpy(pyr1[l], i1[l], px);
      if (pyr2 && pyr2[l]) memcpy(pyr2[l], i2[l], px);
    }

    double t1 = now_s();
    for (int l = levels - 1; l >= 0; --l) { /* :115 */
      int lw = sh.level_w[l], lh = sh.level_h[l];
      size_t fbytes = (size_t)lw * lh * 2 * sizeof(float);
      if (l != levels - 1) orc_copy_mvs(fl[l + 1], sh.level_w[l + 1], sh.level_h[l + 1], block_size[l + 1], fl[l]);
      if (raster_search_variant) orc_search_level_raster(i1[l], i2[l], lw, lh, block_size[l], search_size[l], fl[l]);
      else orc_search_level(i1[l], i2[l], lw, lh, block_size[l], search_size[l], fl[l], st, l);
      if (after_search && after_search[l]) memcpy(after_search[l], fl[l], fbytes);
      int bs = block_size[l];
      float lambda = (float)(block_size[l] / 2); /* integer division, :73,95 */
      while (bs > 1) {                           /* :141-152 */
        for (int s = 0; s < sweeps; ++s) orc_regularize_sweep(i1[l], i2[l], lw, lh, bs, lambda, s + 1, fl[l], st, l);
        orc_divide_blocks(lw, lh, bs, fl[l]);
        bs >>= 1;
        lambda = lambda * 2;
      }
      if (after_reg && after_reg[l]) memcpy(after_reg[l], fl[l], fbytes);
    }
    orc_copy_to_all_pixels(sh.level_w[0], sh.level_h[0], 2, fl[0]); /* :205-206 */
    st->t_run_s = now_s() - t1;
    memcpy(flow_out, fl[0], (size_t)sh.padded_w * sh.padded_h * 2 * sizeof(float));
  }
  for (int l = 0; l < levels; ++l) {
    free(i1[l]);
    free(i2[l]);
    free(fl[l]);
  }
  return rc;
}

int orc_estimate_debug(const uint8_t* im1, const uint8_t* im2, int w, int h, size_t pitch, int levels,
                       const int* search_size, const int* block_size, int sweeps, float* flow_out, orc_stats* st,
                       uint8_t* const* pyr1, uint8_t* const* pyr2, float* const* after_search,
                       float* const* after_reg) {
  return estimate_impl(im1, im2, w, h, pitch, levels, search_size, block_size, sweeps, flow_out, st, pyr1, pyr2, after_search,
                       after_reg, 0);
}

/* The whole path with find_min_block (:246-294) in calcLevelBM, i.e. with the comment markers of :235 and :236 swapped. */
int orc_estimate_raster(const uint8_t* im1, const uint8_t* im2, int w, int h, size_t pitch, int levels,
                        const int* search_size, const int* block_size, int sweeps, float* flow_out) {
  return estimate_impl(im1, im2, w, h, pitch, levels, search_size, block_size, sweeps, flow_out, NULL, NULL, NULL, NULL, NULL, 1);
}

int orc_estimate(const uint8_t* im1, const uint8_t* im2, int w, int h, size_t pitch, int levels,
                 const int* search_size, const int* block_size, int sweeps, float* flow_out, orc_stats* st) {
  return orc_estimate_debug(im1, im2, w, h, pitch, levels, search_size, block_size, sweeps, flow_out, st, NULL, NULL,
                            NULL, NULL);
}

typedef struct {
  int n, next, w, h, levels, sweeps, rc;
  size_t pitch;
  const uint8_t* const* im1;
  const uint8_t* const* im2;
  const int* ss;
  const int* bs;
  float* const* out;
  pthread_mutex_t mu;
  orc_stats sum;
} many_job;

static void* many_worker(void* arg) {
  many_job* j = (many_job*)arg;
  for (;;) {
    pthread_mutex_lock(&j->mu);
    int i = j->next++;
    pthread_mutex_unlock(&j->mu);
    if (i >= j->n) break;
    orc_stats st;
    int rc = orc_estimate(j->im1[i], j->im2[i], j->w, j->h, j->pitch, j->levels, j->ss, j->bs, j->sweeps, j->out[i], &st);
    pthread_mutex_lock(&j->mu);
    if (rc != ORC_OK) j->rc = rc;
    j->sum.t_ctor_s += st.t_ctor_s;
    j->sum.t_run_s += st.t_run_s;
    j->sum.search_sad_calls += st.search_sad_calls;
    j->sum.search_absdiffs += st.search_absdiffs;
    j->sum.reg_sad_calls += st.reg_sad_calls;
    j->sum.reg_absdiffs += st.reg_absdiffs;
    for (int l = 0; l < ORC_MAX_LEVELS; ++l) {
      j->sum.level_search_absdiffs[l] += st.level_search_absdiffs[l];
      j->sum.level_reg_absdiffs[l] += st.level_reg_absdiffs[l];
    }
    pthread_mutex_unlock(&j->mu);
  }
  return NULL;
}

int orc_estimate_many(int n, const uint8_t* const* im1, const uint8_t* const* im2, int w, int h, size_t pitch,
                      int levels, const int* search_size, const int* block_size, int sweeps, float* const* flow_out,
                      int threads, orc_stats* st_sum) {
  if (n <= 0 || threads <= 0) return ORC_E_ARG;
  many_job j;
  memset(&j, 0, sizeof(j));
  j.n = n; j.w = w; j.h = h; j.pitch = pitch; j.levels = levels; j.sweeps = sweeps;
  j.im1 = im1; j.im2 = im2; j.ss = search_size; j.bs = block_size; j.out = flow_out;
  pthread_mutex_init(&j.mu, NULL);
  if (threads > n) threads = n;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
  for (int t = 0; t < threads; ++t) pthread_create(&th[t], NULL, many_worker, &j);
  for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
  free(th);
  pthread_mutex_destroy(&j.mu);
  if (st_sum) *st_sum = j.sum;
  return j.rc;
}

/* ------------------------------------------------------------------ .flo codec + AEE (rw_flow.cpp) */

static const float kFloTag = 202021.25f; /* "PIEH", rw_flow.cpp:25-26 */

/* ------------------------------------------------------------------------------------------------ main()'s wrapper
 * cv::resize(INTER_LINEAR), 8-bit, integer up-sampling factor (main_class.cpp:32-33).  Restated from OpenCV's
 * published algorithm (see the header); every step is integer except the tap positions, which OpenCV computes as
 * float((d + 0.5) * scale - 0.5) with scale = 1 / factor in double. */
static void resize_taps(int d, int factor, int n_src, int clamp_weight, int* s0, int* s1, int* w0, int* w1) {
  const double scale = 1.0 / (double)factor;
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (clamp_weight) { /* x direction: out-of-range taps move inside and lose their weight */
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= n_src - 1) { s = n_src - 1; f = 0.f; }
  }
  *w0 = (int)lrintf((1.f - f) * 2048.f);
  *w1 = (int)lrintf(f * 2048.f);
  int a = s, b = s + 1;
  if (a < 0) a = 0;
  if (a > n_src - 1) a = n_src - 1;
  if (b < 0) b = 0;
  if (b > n_src - 1) b = n_src - 1;
  *s0 = a;
  *s1 = b;
}

void orc_resize_linear(const uint8_t* src, int w, int h, int factor, uint8_t* dst) {
  const int dw = w * factor, dh = h * factor;
  int* xs0 = (int*)malloc(sizeof(int) * 4 * (size_t)dw);
  int *xs1 = xs0 + dw, *xw0 = xs1 + dw, *xw1 = xw0 + dw;
  for (int x = 0; x < dw; ++x) resize_taps(x, factor, w, 1, &xs0[x], &xs1[x], &xw0[x], &xw1[x]);
  for (int y = 0; y < dh; ++y) {
    int y0, y1, b0, b1;
    resize_taps(y, factor, h, 0, &y0, &y1, &b0, &b1);
    const uint8_t* r0 = src + (size_t)y0 * w;
    const uint8_t* r1 = src + (size_t)y1 * w;
    for (int x = 0; x < dw; ++x) {
      const int h0 = r0[xs0[x]] * xw0[x] + r0[xs1[x]] * xw1[x];
      const int h1 = r1[xs0[x]] * xw0[x] + r1[xs1[x]] * xw1[x];
      int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
      if (v < 0) v = 0;
      if (v > 255) v = 255;
      dst[(size_t)y * dw + x] = (uint8_t)v;
    }
  }
  free(xs0);
}

void orc_strip_subsample(const float* padded_flow, int padded_w, int padded_h, int pad_x, int pad_y, int factor,
                         float* out) {
  const int ow = (padded_w - 2 * pad_x) / factor;
  for (int i = pad_y; i < padded_h - pad_y; i += factor) {      /* main_class.cpp:62-70 */
    for (int j = pad_x; j < padded_w - pad_x; j += factor) {
      const float* p = padded_flow + ((size_t)i * padded_w + j) * 2;
      float* q = out + ((size_t)((i - pad_y) / factor) * ow + (j - pad_x) / factor) * 2;
      q[0] = p[0] / (float)factor;
      q[1] = p[1] / (float)factor;
    }
  }
}


static int flo_open(const char* path, FILE** fp, int* w, int* h) {
  if (!path) return ORC_E_ARG;
  const char* dot = strrchr(path, '.');
  if (!dot || strcmp(dot, ".flo") != 0) return ORC_E_FORMAT; /* :58-63 */
  FILE* f = fopen(path, "rb");
  if (!f) return ORC_E_IO;
  float tag;
  int32_t ww, hh;
  if (fread(&tag, 4, 1, f) != 1 || fread(&ww, 4, 1, f) != 1 || fread(&hh, 4, 1, f) != 1) { fclose(f); return ORC_E_FORMAT; }
  if (tag != kFloTag || ww < 1 || ww > 99999 || hh < 1 || hh > 99999) { fclose(f); return ORC_E_FORMAT; } /* :82-98 */
  *w = ww;
  *h = hh;
  *fp = f;
  return ORC_OK;
}

int orc_flo_read_header(const char* path, int* w, int* h) {
  FILE* f;
  int rc = flo_open(path, &f, w, h);
  if (rc == ORC_OK) fclose(f);
  return rc;
}

int orc_flo_read(const char* path, float* data, int w, int h) {
  FILE* f;
  int fw, fh;
  int rc = flo_open(path, &f, &fw, &fh);
  if (rc != ORC_OK) return rc;
  if (fw != w || fh != h) { fclose(f); return ORC_E_ARG; }
  size_t n = (size_t)w * h * 2;
  if (fread(data, sizeof(float), n, f) != n) { fclose(f); return ORC_E_FORMAT; } /* "file is too short" */
  if (fgetc(f) != EOF) { fclose(f); return ORC_E_FORMAT; }                         /* "file is too long", :129-133 */
  fclose(f);
  return ORC_OK;
}

int orc_flo_write(const char* path, const float* data, int w, int h) {
  if (!path) return ORC_E_ARG;
  const char* dot = strrchr(path, '.');
  if (!dot || strcmp(dot, ".flo") != 0) return ORC_E_FORMAT;
  FILE* f = fopen(path, "wb");
  if (!f) return ORC_E_IO;
  int32_t ww = w, hh = h;
  size_t n = (size_t)w * h * 2;
  int ok = fwrite("PIEH", 1, 4, f) == 4 && fwrite(&ww, 4, 1, f) == 1 && fwrite(&hh, 4, 1, f) == 1 &&
           fwrite(data, sizeof(float), n, f) == n;
  fclose(f);
  return ok ? ORC_OK : ORC_E_IO;
}

static int unknown_flow(float u, float v) { /* rw_flow.cpp:39-43 */
  return fabs((double)u) > 1e9 || fabs((double)v) > 1e9 || isnan(u) || isnan(v);
}

double orc_aee(const float* gt, const float* flow, int w, int h) { /* rw_flow.cpp:309-332 */
  int count = 0;
  double error = 0.0;
  for (size_t i = 0; i < (size_t)w * h; ++i) {
    float gu = gt[2 * i], gv = gt[2 * i + 1];
    if (unknown_flow(gu, gv)) continue;
    ++count;
    /* float differences, products and sum; the unqualified sqrt on a float picks the C++ float
     * overload (MSVC <math.h> and libstdc++ alike); the accumulator is double (:312,327) */
    float du = gu - flow[2 * i], dv = gv - flow[2 * i + 1];
    float s = du * du + dv * dv;
    error += (double)sqrtf(s);
  }
  return error / count;
}
