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
