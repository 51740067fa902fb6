// opencv_io_stubs.hpp -- TEST ONLY: declarations of the three OpenCV I/O / imgproc functions main_class.cpp calls besides the
// classes on the path (imread, resize, imwrite), so that the reference's unmodified main() can be syntax-checked against
// include/ without an OpenCV installation (tests/test_capi_host.py::test_reference_main_compiles_against_dropin_headers).
#pragma once
#include <string>
#include <opencv2/core/core.hpp>
namespace cv {
enum { INTER_LINEAR = 1 };
Mat imread(const std::string& filename, int flags);
bool imwrite(const std::string& filename, const Mat& img);
void resize(const Mat& src, Mat& dst, Size dsize, double fx, double fy, int interpolation);
}  // namespace cv
