// ref_glue.cpp -- C entry points around the reference's own classes (compiled from /root/reference, unmodified).
// TEST INFRASTRUCTURE ONLY: built into oracle/_ref/libbbme_ref.so by oracle/Makefile.
#include <time.h>

// The standard and cv:: headers first, with their own access specifiers; then the reference's class with its private members
// opened up, so that its UNCALLED functions (find_min_block :246-294, draw_MVimage :887-905 -- the call sites at :235 and
// :213-216 are commented out) can be run as they stand.  motion_framework.cpp itself is compiled unmodified.
#include "opencv_headers.h"
#include "standard_headers.h"
#include <string.h>
#include <vector>
#define private public
#include "motion_framework.h"
#undef private
#include "rw_flow.h"

static double now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

extern "C" {

// MF(image1, image2, search_size, block_size, num_levels) + calcMotionBlockMatching(), as main_class.cpp:45-50 does.
// dims4 = {padded_width, padded_height, padding_x, padding_y}.  The caller pre-checks that the shape is paddable
// (otherwise the reference calls getchar() and exit(1), motion_framework.cpp:21-26).
int ref_mf_run(const unsigned char* im1, const unsigned char* im2, int w, int h, const int* search_size,
               const int* block_size, int num_levels, float* flow_out, int* dims4, double* t_ctor, double* t_run) {
  cv::Mat a(h, w, CV_8UC1, (void*)im1, (size_t)w), b(h, w, CV_8UC1, (void*)im2, (size_t)w);
  double t0 = now_s();
  MF mf(a, b, search_size, block_size, num_levels);
  double t1 = now_s();
  cv::Mat flow = mf.calcMotionBlockMatching();
  double t2 = now_s();
  if (t_ctor) *t_ctor = t1 - t0;
  if (t_run) *t_run = t2 - t1;
  dims4[0] = mf.padded_width; dims4[1] = mf.padded_height; dims4[2] = mf.padding_x; dims4[3] = mf.padding_y;
  if (flow.cols != mf.padded_width || flow.rows != mf.padded_height) return -1;
  for (int i = 0; i < flow.rows; ++i) memcpy(flow_out + (size_t)i * flow.cols * 2, flow.data + (size_t)i * flow.step, (size_t)flow.cols * 8);
  return 0;
}

// MF::calcLevelBM with the reference's own find_min_block (:246-294) in place of the spiral search (the swap the commented line
// :235 describes), on ONE level: flow holds the predictions at block corners on entry and the vectors on exit (w, h multiples
// of the block size with at least two blocks per axis, so that MF::MF pads nothing).
int ref_find_min_block_level(const unsigned char* im1, const unsigned char* im2, int w, int h, int block_size, int search_size,
                             float* flow) {
  cv::Mat a(h, w, CV_8UC1, (void*)im1, (size_t)w), b(h, w, CV_8UC1, (void*)im2, (size_t)w);
  const int ss[1] = {search_size}, bs[1] = {block_size};
  MF mf(a, b, ss, bs, 1);
  if (mf.padded_width != w || mf.padded_height != h) return -1;
  mf.curr_level = 0;
  for (int i = 0; i < h; i += block_size) {
    for (int j = 0; j < w; j += block_size) {
      float* f = flow + ((size_t)i * w + j) * 2;
      const int x2 = j + (int)f[0], y2 = i + (int)f[1];
      BlockPosition r = mf.find_min_block(i, j, y2, x2);
      f[0] = (float)r.pos_x - j;
      f[1] = (float)r.pos_y - i;
    }
  }
  return 0;
}

// The reference's own draw_MVimage (:887-905) on one level: out is zero-filled first.
int ref_draw_mvimage(const unsigned char* im1, const unsigned char* im2, int w, int h, int block_size, const float* flow,
                     unsigned char* out) {
  cv::Mat a(h, w, CV_8UC1, (void*)im1, (size_t)w), b(h, w, CV_8UC1, (void*)im2, (size_t)w);
  const int ss[1] = {block_size + 2}, bs[1] = {block_size};
  MF mf(a, b, ss, bs, 1);
  if (mf.padded_width != w || mf.padded_height != h) return -1;
  mf.curr_level = 0;
  cv::Mat& lf = mf.level_data[0].level_flow;
  for (int i = 0; i < h; ++i) memcpy(lf.data + (size_t)i * lf.step, flow + (size_t)i * w * 2, (size_t)w * 8);
  cv::Mat img = cv::Mat::zeros(h, w, CV_8UC1);
  mf.draw_MVimage(img);
  for (int i = 0; i < h; ++i) memcpy(out + (size_t)i * w, img.data + (size_t)i * img.step, (size_t)w);
  return 0;
}

int ref_flow_read(const char* path, float* out, int* w, int* h, int cap_floats) {
  Flow f;
  cv::Mat img;
  f.ReadFlowFile(img, path);
  *w = img.cols; *h = img.rows;
  if ((long long)img.cols * img.rows * 2 > cap_floats) return -1;
  for (int i = 0; i < img.rows; ++i) memcpy(out + (size_t)i * img.cols * 2, img.data + (size_t)i * img.step, (size_t)img.cols * 8);
  return 0;
}

int ref_flow_write(const char* path, const float* data, int w, int h) {
  Flow f;
  cv::Mat img(h, w, CV_32FC2, (void*)data, (size_t)w * 8);
  f.WriteFlowFile(img, path);
  return 0;
}

double ref_flow_mse(const float* gt, const float* flow, int w, int h) {
  Flow f;
  cv::Mat a(h, w, CV_32FC2, (void*)gt, (size_t)w * 8), b(h, w, CV_32FC2, (void*)flow, (size_t)w * 8);
  return f.CalculateMSE(a, b);
}

// Flow::MotionToColor (rw_flow.cpp:202-249) of the reference itself; out: h x w x 3 bytes.
int ref_flow_color(const float* flow, int w, int h, float maxmotion, unsigned char* out) {
  Flow f;
  cv::Mat in(h, w, CV_32FC2, (void*)flow, (size_t)w * 8), img;
  f.MotionToColor(in, img, maxmotion);
  if (img.cols != w || img.rows != h) return -1;
  for (int i = 0; i < h; ++i) memcpy(out + (size_t)i * w * 3, img.data + (size_t)i * img.step, (size_t)w * 3);
  return 0;
}

}  // extern "C"
