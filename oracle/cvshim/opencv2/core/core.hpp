// cvshim: a minimal, self-written stand-in for the subset of OpenCV's C++ API that the reference's
// motion_framework.cpp / rw_flow.cpp / parallel.h use.  TEST INFRASTRUCTURE ONLY: it exists so that the
// reference's own sources can be compiled in this container (which has no OpenCV C++ headers or libraries)
// into oracle/_ref, to pin the oracle against the reference's real control flow.
//
// Numeric functions follow OpenCV's documented 8-bit behaviour and are themselves pinned against the
// container's cv2 4.13.0 wheel by tests/test_oracle.py (pyrDown: 5x5 [1 4 6 4 1]^2 / 256 with rounding,
// BORDER_REFLECT_101; copyMakeBorder: constant border; norm(NORM_L1): exact sum of |a-b|).
// Display / file output functions are no-ops.
#ifndef CVSHIM_CORE_HPP
#define CVSHIM_CORE_HPP

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <queue>
#include <string>
#include <vector>

#define CV_CN_SHIFT 3
#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC4 CV_MAKETYPE(CV_32F, 4)
#define CV_32SC4 CV_MAKETYPE(CV_32S, 4)

namespace cv {

enum { NORM_L1 = 2 };
enum { BORDER_CONSTANT = 0, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };

template <typename T, int N>
struct Vec {
  T val[N];
  Vec() { for (int i = 0; i < N; ++i) val[i] = T(); }
  Vec(T a) { for (int i = 0; i < N; ++i) val[i] = T(); val[0] = a; }
  template <typename A, typename B>
  Vec(A a, B b) { for (int i = 0; i < N; ++i) val[i] = T(); val[0] = (T)a; val[1] = (T)b; }
  template <typename A, typename B, typename C>
  Vec(A a, B b, C c) { for (int i = 0; i < N; ++i) val[i] = T(); val[0] = (T)a; val[1] = (T)b; val[2] = (T)c; }
  template <typename A, typename B, typename C, typename D>
  Vec(A a, B b, C c, D d) { val[0] = (T)a; val[1] = (T)b; val[2] = (T)c; val[3] = (T)d; }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
  Vec mul(const Vec& o) const { Vec r; for (int i = 0; i < N; ++i) r.val[i] = val[i] * o.val[i]; return r; }
  Vec operator+(const Vec& o) const { Vec r; for (int i = 0; i < N; ++i) r.val[i] = val[i] + o.val[i]; return r; }
  Vec operator-(const Vec& o) const { Vec r; for (int i = 0; i < N; ++i) r.val[i] = val[i] - o.val[i]; return r; }
};
typedef Vec<float, 2> Vec2f;
typedef Vec<float, 4> Vec4f;
typedef Vec<int, 4> Vec4i;
typedef Vec<unsigned char, 3> Vec3b;

struct Scalar {
  double val[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
};
struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
  bool operator==(const Size& o) const { return width == o.width && height == o.height; }
  bool operator!=(const Size& o) const { return !(*this == o); }
};
struct Point {
  int x, y;
  Point() : x(0), y(0) {}
  Point(int x_, int y_) : x(x_), y(y_) {}
};
struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {}
};
struct Range {
  int start, end;
  Range() : start(0), end(0) {}
  Range(int s, int e) : start(s), end(e) {}
};

class Mat {
 public:
  int rows, cols;
  unsigned char* data;
  size_t step;

  Mat() : rows(0), cols(0), data(0), step(0), type_(0) {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, const Scalar& s) { create(r, c, type); fill(s); }
  // header over caller memory (not owned)
  Mat(int r, int c, int type, void* ext, size_t ext_step) : rows(r), cols(c), data((unsigned char*)ext), step(ext_step), type_(type) {}

  static Mat zeros(int r, int c, int type) { Mat m(r, c, type); memset(m.data, 0, m.step * (size_t)r); return m; }

  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    step = (size_t)c * elemSize();
    size_t bytes = step * (size_t)r;
    buf_ = std::shared_ptr<unsigned char>(new unsigned char[bytes ? bytes : 1], std::default_delete<unsigned char[]>());
    data = buf_.get();
  }
  int type() const { return type_; }
  int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
  size_t elemSize1() const { int d = type_ & 7; return d == CV_8U ? 1 : 4; }
  size_t elemSize() const { return elemSize1() * (size_t)channels(); }
  Size size() const { return Size(cols, rows); }
  bool empty() const { return data == 0 || rows == 0 || cols == 0; }

  template <typename T> T& at(int i, int j) { return *reinterpret_cast<T*>(data + (size_t)i * step + (size_t)j * sizeof(T)); }
  template <typename T> const T& at(int i, int j) const { return *reinterpret_cast<const T*>(data + (size_t)i * step + (size_t)j * sizeof(T)); }

  Mat operator()(const Rect& r) const {  // region of interest sharing storage
    Mat m;
    m.rows = r.height; m.cols = r.width; m.type_ = type_; m.step = step; m.buf_ = buf_;
    m.data = data + (size_t)r.y * step + (size_t)r.x * elemSize();
    return m;
  }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int i = 0; i < rows; ++i) memcpy(m.data + (size_t)i * m.step, data + (size_t)i * step, (size_t)cols * elemSize());
    return m;
  }
  void copyTo(Mat dst) const {
    if (dst.rows != rows || dst.cols != cols || dst.type_ != type_) return;  // ROI destination of the same size only
    for (int i = 0; i < rows; ++i) memcpy(dst.data + (size_t)i * dst.step, data + (size_t)i * step, (size_t)cols * elemSize());
  }

 private:
  void fill(const Scalar& s) {
    const int cn = channels(), d = type_ & 7;
    for (int i = 0; i < rows; ++i)
      for (int j = 0; j < cols; ++j)
        for (int k = 0; k < cn; ++k) {
          unsigned char* p = data + (size_t)i * step + ((size_t)j * cn + k) * elemSize1();
          if (d == CV_8U) *p = (unsigned char)s.val[k];
          else if (d == CV_32S) *reinterpret_cast<int*>(p) = (int)s.val[k];
          else *reinterpret_cast<float*>(p) = (float)s.val[k];
        }
  }
  int type_;
  std::shared_ptr<unsigned char> buf_;
};

// sum of |a-b| over two equally sized CV_8UC1 views
inline double norm(const Mat& a, const Mat& b, int /*NORM_L1*/) {
  long long s = 0;
  for (int i = 0; i < a.rows; ++i) {
    const unsigned char* pa = a.data + (size_t)i * a.step;
    const unsigned char* pb = b.data + (size_t)i * b.step;
    int r = 0;
    for (int j = 0; j < a.cols; ++j) r += abs((int)pa[j] - (int)pb[j]);
    s += r;
  }
  return (double)s;
}

inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int /*BORDER_CONSTANT*/,
                           const Scalar& v = Scalar()) {
  Mat out(src.rows + top + bottom, src.cols + left + right, src.type());
  memset(out.data, (int)v.val[0], out.step * (size_t)out.rows);
  for (int i = 0; i < src.rows; ++i) memcpy(out.data + (size_t)(i + top) * out.step + (size_t)left, src.data + (size_t)i * src.step, (size_t)src.cols);
  dst = out;
}

inline int shimBorder101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
  return p;
}

// 8-bit pyrDown: direct 5x5 binomial stencil on even source centres, one rounding at the end
inline void pyrDown(const Mat& src, Mat& dst, const Size& dsize = Size(), int /*borderType*/ = BORDER_DEFAULT) {
  const int dw = dsize.width > 0 ? dsize.width : (src.cols + 1) / 2;
  const int dh = dsize.height > 0 ? dsize.height : (src.rows + 1) / 2;
  static const int k[5] = {1, 4, 6, 4, 1};
  Mat out(dh, dw, src.type());
  for (int y = 0; y < dh; ++y)
    for (int x = 0; x < dw; ++x) {
      int acc = 0;
      for (int j = -2; j <= 2; ++j) {
        const unsigned char* row = src.data + (size_t)shimBorder101(2 * y + j, src.rows) * src.step;
        for (int i = -2; i <= 2; ++i) acc += k[j + 2] * k[i + 2] * (int)row[shimBorder101(2 * x + i, src.cols)];
      }
      out.data[(size_t)y * out.step + x] = (unsigned char)((acc + 128) >> 8);
    }
  dst = out;
}

class ParallelLoopBody {
 public:
  virtual ~ParallelLoopBody() {}
  virtual void operator()(const Range& range) const = 0;
};
inline void parallel_for_(const Range& r, const ParallelLoopBody& body, double = -1.) { body(r); }

// drawing / GUI / image files: not part of the hot path
inline void line(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0) {}
inline bool imwrite(const std::string&, const Mat&) { return true; }
inline void imshow(const std::string&, const Mat&) {}
inline void namedWindow(const std::string&, int = 1) {}
inline int waitKey(int = 0) { return -1; }

}  // namespace cv

// MSVC <stdlib.h> macros used by rw_flow.cpp:214-219.  Defined last: every standard header the reference
// includes afterwards is already include-guarded above, so libstdc++'s own `__max` identifiers are not touched.
#ifndef __max
#define __max(a, b) (((a) > (b)) ? (a) : (b))
#define __min(a, b) (((a) < (b)) ? (a) : (b))
#endif
#endif
