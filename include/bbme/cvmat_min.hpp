// cvmat_min.hpp -- the few cv:: types the drop-in classes need, for builds WITHOUT OpenCV.
//
// The reference's public signatures take and return cv::Mat (motion_framework.h:12-13, rw_flow.h:17-22).  When the
// host application has OpenCV, include <opencv2/core.hpp> BEFORE the drop-in headers (or define BBME_USE_OPENCV) and
// the real cv::Mat is used.  Otherwise this header provides a small reference-counted cv::Mat with the members the
// drop-in path touches: rows, cols, data, step, type(), at<T>(), clone(), zeros(), ROI via operator()(Rect).
#ifndef BBME_CVMAT_MIN_HPP
#define BBME_CVMAT_MIN_HPP

#include <cstddef>
#include <cstring>
#include <memory>

#ifndef CV_8UC1
#define BBME_CV_DEPTH_MASK 7
#define CV_8U 0
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#endif

namespace cv {

template <typename T, int N>
struct Vec {
  T val[N];
  Vec() { for (int i = 0; i < N; ++i) val[i] = T(); }
  Vec(T a, T b) { static_assert(N >= 2, "Vec(a,b)"); for (int i = 0; i < N; ++i) val[i] = T(); val[0] = a; val[1] = b; }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
};
typedef Vec<float, 2> Vec2f;

struct Rect {
  int x, y, width, height;
  Rect(int x_ = 0, int y_ = 0, int w_ = 0, int h_ = 0) : x(x_), y(y_), width(w_), height(h_) {}
};

class Mat {
 public:
  int rows, cols;
  unsigned char* data;
  std::size_t step;

  Mat() : rows(0), cols(0), data(nullptr), step(0), type_(0) {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* external, std::size_t external_step)
      : rows(r), cols(c), data(static_cast<unsigned char*>(external)), step(external_step), type_(type) {}

  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    step = static_cast<std::size_t>(c) * elemSize();
    const std::size_t bytes = step * static_cast<std::size_t>(r);
    owner_.reset(new unsigned char[bytes ? bytes : 1], std::default_delete<unsigned char[]>());
    data = owner_.get();
  }
  static Mat zeros(int r, int c, int type) {
    Mat m(r, c, type);
    std::memset(m.data, 0, m.step * static_cast<std::size_t>(r));
    return m;
  }
  int type() const { return type_; }
  int channels() const { return (type_ >> 3) + 1; }
  std::size_t elemSize() const { return ((type_ & 7) == CV_8U ? 1u : 4u) * static_cast<std::size_t>(channels()); }
  bool empty() const { return !data || rows == 0 || cols == 0; }
  template <typename T> T& at(int i, int j) { return *reinterpret_cast<T*>(data + static_cast<std::size_t>(i) * step + static_cast<std::size_t>(j) * sizeof(T)); }
  template <typename T> const T& at(int i, int j) const { return *reinterpret_cast<const T*>(data + static_cast<std::size_t>(i) * step + static_cast<std::size_t>(j) * sizeof(T)); }
  Mat operator()(const Rect& r) const {
    Mat m;
    m.rows = r.height; m.cols = r.width; m.type_ = type_; m.step = step; m.owner_ = owner_;
    m.data = data + static_cast<std::size_t>(r.y) * step + static_cast<std::size_t>(r.x) * elemSize();
    return m;
  }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int i = 0; i < rows; ++i) std::memcpy(m.data + static_cast<std::size_t>(i) * m.step, data + static_cast<std::size_t>(i) * step, static_cast<std::size_t>(cols) * elemSize());
    return m;
  }

 private:
  int type_;
  std::shared_ptr<unsigned char> owner_;
};

}  // namespace cv
#endif
