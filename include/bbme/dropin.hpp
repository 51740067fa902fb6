// dropin.hpp -- the reference's C++ entry points re-created on top of the C ABI (bbme.h).
//
//   class MF             <- motion_framework.h:9-54   (constructor :12, calcMotionBlockMatching :13, public ints :16-19)
//   class PyramidLevel   <- pyramid_level.h:7-16
//   class BlockPosition  <- block_position.h:4-9
//   class Flow           <- rw_flow.h:9-38            (ReadFlowFile, WriteFlowFile, MotionToColor, CalculateMSE)
//
// Header-only: an application that used the reference's classes includes "motion_framework.h" / "rw_flow.h" from this
// include/ directory instead of the reference's, links libbbme.so, and keeps its source unchanged.  What differs:
//   * the work happens on a B200.  The constructor copies the two frames (the reference's constructor copies them too,
//     with copyMakeBorder, motion_framework.cpp:60-61, so a caller may reuse or release its Mats afterwards) and takes a
//     planned context for this geometry from the library's cache (bbme_mf_open); calcMotionBlockMatching() uploads the
//     frames and runs pad, pyramid, search and regularisation; the destructor parks the context for the next MF of the
//     same geometry (main builds one MF per pair, main_class.cpp:45);
//   * the GPU is device 0 unless the BBME_DEVICE environment variable or MF::set_device() names another one;
//   * errors: where the reference prints and calls getchar()/exit(1) (motion_framework.cpp:21-26, rw_flow.cpp), these
//     classes print the same message to std::cout and call exit(1) (no getchar); define BBME_DROPIN_THROW to get a
//     std::runtime_error instead;
//   * MF takes one extra, defaulted constructor argument (regularisation sweeps, the reference's hard-coded 2).
#ifndef BBME_DROPIN_HPP
#define BBME_DROPIN_HPP

#if defined(BBME_USE_OPENCV) || defined(OPENCV_CORE_HPP) || defined(__OPENCV_CORE_HPP__)
#include <opencv2/core/core.hpp>
#else
#include "cvmat_min.hpp"
#endif

#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../bbme.h"

namespace bbme_dropin {
inline void fatal(const std::string& msg) {
#ifdef BBME_DROPIN_THROW
  throw std::runtime_error(msg);
#else
  std::cout << msg << std::endl;
  std::exit(1);
#endif
}
}  // namespace bbme_dropin

class BlockPosition {  // block_position.h:4-9
 public:
  int pos_x;
  int pos_y;
};

class PyramidLevel {  // pyramid_level.h:7-16
 public:
  cv::Mat level_flow;  // filled for level 0 after calcMotionBlockMatching(); coarser levels stay on the device
  int block_size;
  int search_size;
  float lambda;
  cv::Mat image1;
  cv::Mat image2;
};

class MF {
 public:
  MF(cv::Mat& image1, cv::Mat& image2, const int search_size[], const int block_size[], const int num_levels,
     const int sweeps = 2)
      : padded_height(0), padded_width(0), padding_x(0), padding_y(0), ctx_(nullptr), image1_(image1), image2_(image2) {
    if (num_levels <= 0) bbme_dropin::fatal("MF: num_levels must be > 0");                       // assert, motion_framework.cpp:7
    if (image1.rows != image2.rows || image1.cols != image2.cols) bbme_dropin::fatal("MF: image sizes differ");  // :8
    if (image1.type() != CV_8UC1 || image2.type() != CV_8UC1) bbme_dropin::fatal("MF: images must be CV_8UC1");
    image1_ = image1.clone();  // dense copies: the caller's Mats may change or go away before the estimation runs
    image2_ = image2.clone();
    bbme_shape shape;
    const int rc = bbme_mf_open(&ctx_, device(), image1.cols, image1.rows, num_levels, search_size, block_size, sweeps, &shape);
    if (rc == BBME_E_NOPAD) fail("Could not find any multiples of the block size that match padded image dimensions");
    if (rc != BBME_OK) fail(std::string("MF: ") + (ctx_ ? bbme_last_error(ctx_) : bbme_last_error(nullptr)));
    padded_height = shape.padded_height;
    padded_width = shape.padded_width;
    padding_x = shape.padding_x;
    padding_y = shape.padding_y;
    level_data.resize(static_cast<size_t>(num_levels));
    for (int l = 0; l < num_levels; ++l) {
      level_data[l].block_size = block_size[l];
      level_data[l].search_size = search_size[l];
      level_data[l].lambda = static_cast<float>(block_size[l] / 2);  // motion_framework.cpp:73,95
    }
  }

  // Perform block matching for the whole hierarchy/pyramid.  Returns the padded CV_32FC2 field
  // (padded_height x padded_width), which stays valid after the MF object is gone (motion_framework.cpp:218).
  cv::Mat calcMotionBlockMatching() {
    cv::Mat flow(padded_height, padded_width, CV_32FC2);
    const int rc = bbme_estimate(ctx_, image1_.data, image2_.data, image1_.step, reinterpret_cast<float*>(flow.data));
    if (rc != BBME_OK) fail(std::string("MF::calcMotionBlockMatching: ") + bbme_last_error(ctx_));
    level_data[0].level_flow = flow;
    return flow;
  }

  ~MF() {
    if (ctx_) bbme_mf_close(ctx_);
  }

  // Which GPU MF objects constructed from now on use (default: BBME_DEVICE or 0).
  static void set_device(int d) { device_ref() = d; }
  static int device() {
    int d = device_ref();
    if (d < 0) {
      const char* e = std::getenv("BBME_DEVICE");
      d = e ? std::atoi(e) : 0;
    }
    return d;
  }

  int padded_height;
  int padded_width;
  int padding_x;
  int padding_y;

  std::vector<PyramidLevel> level_data;  // private in the reference; exposed read-only here for inspection

 private:
  MF(const MF&);
  MF& operator=(const MF&);
  static int& device_ref() {
    static int d = -1;
    return d;
  }
  void fail(const std::string& msg) {
    const std::string m = msg;  // the text may live in the context
    if (ctx_) bbme_mf_close(ctx_);
    ctx_ = nullptr;
    bbme_dropin::fatal(m);
  }
  bbme_ctx* ctx_;
  cv::Mat image1_, image2_;
};

class Flow {
 public:
  // read a flow file into 2-band image (rw_flow.cpp:50-136)
  void ReadFlowFile(cv::Mat& img, const char* filename) {
    if (filename == nullptr) bbme_dropin::fatal("ReadFlowFile: empty filename");
    int w = 0, h = 0;
    int rc = bbme_flo_read_header(filename, &w, &h);
    if (rc == BBME_OK) {
      img = cv::Mat(h, w, CV_32FC2);
      rc = bbme_flo_read(filename, reinterpret_cast<float*>(img.data), w, h);
    }
    if (rc == BBME_E_IO) bbme_dropin::fatal("ReadFlowFile: could not open file");
    if (rc != BBME_OK) bbme_dropin::fatal("ReadFlowFile: bad .flo file (extension, tag, size or length)");
  }
  // write a 2-band image into flow file (rw_flow.cpp:139-200)
  void WriteFlowFile(cv::Mat img, const char* filename) {
    if (filename == nullptr) bbme_dropin::fatal("WriteFlowFile: empty filename");
    cv::Mat dense = (img.step == static_cast<size_t>(img.cols) * 8) ? img : img.clone();
    const int rc = bbme_flo_write(filename, reinterpret_cast<const float*>(dense.data), dense.cols, dense.rows);
    if (rc == BBME_E_FORMAT) bbme_dropin::fatal("WriteFlowFile: filename should have extension '.flo'");
    if (rc != BBME_OK) bbme_dropin::fatal("WriteFlowFile: could not open file or problem writing data");
  }
  // Color code motion vectors for easier visualization (rw_flow.cpp:202-249); prints the reference's range line
  void MotionToColor(cv::Mat& input_img, cv::Mat& output_img, float maxmotion) {
    cv::Mat dense = (input_img.step == static_cast<size_t>(input_img.cols) * 8) ? input_img : input_img.clone();
    output_img = cv::Mat::zeros(dense.rows, dense.cols, CV_8UC3);
    float r[5];
    const int rc = bbme_flow_to_color(reinterpret_cast<const float*>(dense.data), dense.cols, dense.rows, maxmotion,
                                      output_img.data, r);
    if (rc != BBME_OK) bbme_dropin::fatal("MotionToColor: bad arguments");
    std::printf("max motion: %.4f  motion range: u = %.3f .. %.3f;  v = %.3f .. %.3f\n", r[0], r[1], r[2], r[3], r[4]);
  }
  // average endpoint error against the ground truth, unknown pixels skipped (rw_flow.cpp:309-332)
  double CalculateMSE(cv::Mat& gtruth, cv::Mat& flow) {
    cv::Mat a = (gtruth.step == static_cast<size_t>(gtruth.cols) * 8) ? gtruth : gtruth.clone();
    cv::Mat b = (flow.step == static_cast<size_t>(flow.cols) * 8) ? flow : flow.clone();
    return bbme_flow_aee(reinterpret_cast<const float*>(a.data), reinterpret_cast<const float*>(b.data), a.cols, a.rows);
  }
};

#endif
