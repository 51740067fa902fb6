// flo.cpp -- Middlebury .flo codec, endpoint-error metric and main()'s field post-processing (host code).
//
// Behaviour follows Flow::ReadFlowFile / WriteFlowFile / CalculateMSE (reference rw_flow.cpp:50-136, 139-200,
// 309-332) and main_class.cpp:58-70, with status codes instead of print-and-exit and whole-buffer I/O instead
// of one fread/fwrite per float.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/bbme.h"

namespace {
const float kTagFloat = 202021.25f;  // the bytes "PIEH" read as a little-endian float (rw_flow.cpp:25-26)

bool has_flo_extension(const char* path) {
  const char* dot = path ? strrchr(path, '.') : nullptr;
  return dot && strcmp(dot, ".flo") == 0;
}

int open_and_check(const char* path, FILE** out, int* w, int* h) {
  if (!path || !w || !h) return BBME_E_ARG;
  if (!has_flo_extension(path)) return BBME_E_FORMAT;  // "extension .flo expected"
  FILE* f = fopen(path, "rb");
  if (!f) return BBME_E_IO;
  float tag = 0.f;
  int32_t dims[2] = {0, 0};
  if (fread(&tag, sizeof(tag), 1, f) != 1 || fread(dims, sizeof(int32_t), 2, f) != 2) {
    fclose(f);
    return BBME_E_FORMAT;  // "problem reading file"
  }
  if (tag != kTagFloat || dims[0] < 1 || dims[0] > 99999 || dims[1] < 1 || dims[1] > 99999) {
    fclose(f);
    return BBME_E_FORMAT;  // wrong tag / illegal width / illegal height
  }
  *w = dims[0];
  *h = dims[1];
  *out = f;
  return BBME_OK;
}

bool unknown_flow(float u, float v) {  // rw_flow.cpp:39-43
  return fabs((double)u) > 1e9 || fabs((double)v) > 1e9 || isnan(u) || isnan(v);
}
}  // namespace

extern "C" {

int bbme_flo_read_header(const char* path, int* width, int* height) {
  FILE* f = nullptr;
  int rc = open_and_check(path, &f, width, height);
  if (rc == BBME_OK) fclose(f);
  return rc;
}

int bbme_flo_read(const char* path, float* data, int width, int height) {
  if (!data) return BBME_E_ARG;
  FILE* f = nullptr;
  int w = 0, h = 0;
  int rc = open_and_check(path, &f, &w, &h);
  if (rc != BBME_OK) return rc;
  if (w != width || h != height) {
    fclose(f);
    return BBME_E_ARG;
  }
  const size_t count = (size_t)w * h * 2;
  const bool short_file = fread(data, sizeof(float), count, f) != count;  // "file is too short"
  const bool long_file = !short_file && fgetc(f) != EOF;                  // "file is too long"
  fclose(f);
  return (short_file || long_file) ? BBME_E_FORMAT : BBME_OK;
}

int bbme_flo_write(const char* path, const float* data, int width, int height) {
  if (!path || !data || width < 1 || height < 1) return BBME_E_ARG;
  if (!has_flo_extension(path)) return BBME_E_FORMAT;
  FILE* f = fopen(path, "wb");
  if (!f) return BBME_E_IO;
  const int32_t dims[2] = {width, height};  // width first, then height (rw_flow.cpp:169-170)
  const size_t count = (size_t)width * height * 2;
  const bool ok = fwrite("PIEH", 1, 4, f) == 4 && fwrite(dims, sizeof(int32_t), 2, f) == 2 &&
                  fwrite(data, sizeof(float), count, f) == count;
  const bool closed = fclose(f) == 0;
  return (ok && closed) ? BBME_OK : BBME_E_IO;
}

double bbme_flow_aee(const float* gt, const float* flow, int width, int height) {
  if (!gt || !flow || width < 1 || height < 1) return NAN;
  long long known = 0;
  double total = 0.0;
  const size_t px = (size_t)width * height;
  for (size_t i = 0; i < px; ++i) {
    const float gu = gt[2 * i], gv = gt[2 * i + 1];
    if (unknown_flow(gu, gv)) continue;
    ++known;
    const float du = gu - flow[2 * i], dv = gv - flow[2 * i + 1];
    const float sq = du * du + dv * dv;  // float arithmetic, float sqrt, double accumulator (rw_flow.cpp:312,327)
    total += (double)sqrtf(sq);
  }
  return total / (double)known;
}

int bbme_flow_strip_subsample(const float* padded, const bbme_shape* sh, int factor, float* out) {
  if (!padded || !sh || !out || factor < 1) return BBME_E_ARG;
  // main_class.cpp:58-70: i from pad_y while i < padded_h - pad_y step factor; MV / factor
  const int pw = sh->padded_width, ph = sh->padded_height, px = sh->padding_x, py = sh->padding_y;
  const int ow = sh->width / factor;
  const float div = (float)factor;
  for (int i = py; i < ph - py; i += factor)
    for (int j = px; j < pw - px; j += factor) {
      const float* s = padded + ((size_t)i * pw + j) * 2;
      float* d = out + ((size_t)((i - py) / factor) * ow + (j - px) / factor) * 2;
      d[0] = s[0] / div;
      d[1] = s[1] / div;
    }
  return BBME_OK;
}

}  // extern "C"

// Flow::MotionToColor + computeColor + makecolorwheel (rw_flow.cpp:202-300): Middlebury colour coding of a flow field.
// Host code like the .flo codec (the reference writes flow.png from it, main_class.cpp:73-75).
//
// PROVENANCE.  This block is a deliberate restatement of a published algorithm, not independent design: the reference's
// routine is itself a copy of D. Scharstein's Middlebury `colorcode.cpp` (vendored in the reference under
// middlebury/flow-code/), and a flow.png that matches the reference byte for byte leaves no freedom in the colour-wheel table
// (segment lengths RY YG GC CB BM MR = 15, 6, 4, 11, 13, 6), in the operation order, or in which operations are float and which
// double: float sqrt / atan2 / division by the reference's FLOAT pi (rw_flow.cpp:36), double only in `col *= .75` and
// `255.0 * col`.  The variable names for the running extrema and the wheel interpolation follow the published routine for that
// reason.  The reference is GPLv3; this file is test-pinned against the reference's own function compiled from the reference
// tree (tests/golden/flow_color_ref.npz), which is how the byte-exactness claim is checked.  The .flo codec and the AEE above
// are independently written (whole-buffer I/O, status codes).
namespace {
struct ColorWheel {
  int n = 0;
  int c[60][3];  // MAXCOLS, rw_flow.h:5
  void set(int r, int g, int b, int k) { c[k][0] = r; c[k][1] = g; c[k][2] = b; }
  ColorWheel() {  // makecolorwheel, rw_flow.cpp:276-300
    const int RY = 15, YG = 6, GC = 4, CB = 11, BM = 13, MR = 6;
    int k = 0;
    for (int i = 0; i < RY; i++) set(255, 255 * i / RY, 0, k++);
    for (int i = 0; i < YG; i++) set(255 - 255 * i / YG, 255, 0, k++);
    for (int i = 0; i < GC; i++) set(0, 255, 255 * i / GC, k++);
    for (int i = 0; i < CB; i++) set(0, 255 - 255 * i / CB, 255, k++);
    for (int i = 0; i < BM; i++) set(255 * i / BM, 0, 255, k++);
    for (int i = 0; i < MR; i++) set(255, 0, 255 - 255 * i / MR, k++);
    n = k;
  }
};

}  // namespace

extern "C" int bbme_flow_to_color(const float* flow, int width, int height, float maxmotion, uint8_t* bgr, float* range5) {
  if (!flow || !bgr || width < 1 || height < 1) return BBME_E_ARG;
  static const ColorWheel wheel;
  const float kPi = 3.14159265358979323846f;  // the reference redefines M_PI as a float literal (rw_flow.cpp:36)
  float maxx = -999, maxy = -999, minx = 999, miny = 999, maxrad = -1;  // rw_flow.cpp:205-207
  const size_t n = (size_t)width * height;
  for (size_t i = 0; i < n; ++i) {
    const float fx = flow[2 * i], fy = flow[2 * i + 1];
    if (unknown_flow(fx, fy)) continue;
    maxx = maxx > fx ? maxx : fx;
    maxy = maxy > fy ? maxy : fy;
    minx = minx < fx ? minx : fx;
    miny = miny < fy ? miny : fy;
    const float rad = sqrtf(fx * fx + fy * fy);
    maxrad = maxrad > rad ? maxrad : rad;
  }
  if (range5) { range5[0] = maxrad; range5[1] = minx; range5[2] = maxx; range5[3] = miny; range5[4] = maxy; }
  if (maxmotion > 0) maxrad = maxmotion;  // :225-226
  if (maxrad == 0) maxrad = 1;            // :228-229
  for (size_t i = 0; i < n; ++i) {
    uint8_t* pix = bgr + 3 * i;
    const float u = flow[2 * i], v = flow[2 * i + 1];
    if (unknown_flow(u, v)) { pix[0] = pix[1] = pix[2] = 0; continue; }
    const float fx = u / maxrad, fy = v / maxrad;  // :245
    const float rad = sqrtf(fx * fx + fy * fy);
    const float a = atan2f(-fy, -fx) / kPi;
    const float fk = (a + 1.0f) / 2.0f * (wheel.n - 1);
    const int k0 = (int)fk;
    const int k1 = (k0 + 1) % wheel.n;
    const float f = fk - k0;
    for (int b = 0; b < 3; b++) {
      const float col0 = wheel.c[k0][b] / 255.0f;
      const float col1 = wheel.c[k1][b] / 255.0f;
      float col = (1 - f) * col0 + f * col1;
      if (rad <= 1) col = 1 - rad * (1 - col);  // increase saturation with radius
      else col *= .75;                          // out of range
      pix[2 - b] = (uint8_t)(int)(255.0 * col);
    }
  }
  return BBME_OK;
}
