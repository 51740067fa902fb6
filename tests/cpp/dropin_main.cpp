// dropin_main.cpp -- a caller written against the REFERENCE's class API (the shape of main_class.cpp:19-82), compiled
// against include/ + libbbme.so.  Reads two raw 8-bit frames, runs MF, strips the padding like main() does and writes
// the field with Flow::WriteFlowFile.  tests/test_gpu_dropin_cpp.py builds it, runs it and compares with the oracle.
//   usage: dropin_main <w> <h> <frame1.raw> <frame2.raw> <out.flo> <levels> <search_size...> <block_size...>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "motion_framework.h"
#include "rw_flow.h"

int main(int argc, char** argv) {
  if (argc < 8) return 2;
  const int w = atoi(argv[1]), h = atoi(argv[2]);
  const int num_levels = atoi(argv[6]);
  if (argc != 7 + 2 * num_levels) return 2;
  std::vector<int> search_size(num_levels), block_size(num_levels);
  for (int i = 0; i < num_levels; ++i) {
    search_size[i] = atoi(argv[7 + i]);
    block_size[i] = atoi(argv[7 + num_levels + i]);
  }
  cv::Mat image1(h, w, CV_8UC1), image2(h, w, CV_8UC1);
  FILE* f = fopen(argv[3], "rb");
  if (!f || fread(image1.data, 1, (size_t)w * h, f) != (size_t)w * h) return 3;
  fclose(f);
  f = fopen(argv[4], "rb");
  if (!f || fread(image2.data, 1, (size_t)w * h, f) != (size_t)w * h) return 3;
  fclose(f);

  MF motion_pair(image1, image2, search_size.data(), block_size.data(), num_levels);
  cv::Mat flow_res = motion_pair.calcMotionBlockMatching();

  // main_class.cpp:58-70 with an interpolation factor of 1: strip the padding
  const int pad_y = motion_pair.padding_y, pad_x = motion_pair.padding_x;
  cv::Mat mvs(h, w, CV_32FC2);
  for (int i = pad_y; i < motion_pair.padded_height - pad_y; ++i)
    for (int j = pad_x; j < motion_pair.padded_width - pad_x; ++j)
      mvs.at<cv::Vec2f>(i - pad_y, j - pad_x) = flow_res.at<cv::Vec2f>(i, j);

  Flow file;
  // main_class.cpp:73-75: colour-code the field (main() hands the image to cv::imwrite)
  cv::Mat flow_img;
  file.MotionToColor(mvs, flow_img, -1);
  if (flow_img.rows != h || flow_img.cols != w) return 5;
  file.WriteFlowFile(mvs, argv[5]);
  cv::Mat back;
  file.ReadFlowFile(back, argv[5]);
  double err = file.CalculateMSE(back, mvs);
  printf("padded %dx%d pad (%d,%d) self-AEE %.6f\n", motion_pair.padded_width, motion_pair.padded_height, pad_x, pad_y, err);
  if (err != 0.0) return 4;

  // The reference's caller builds one MF per frame pair (main_class.cpp:45-50) and its constructor COPIES the frames
  // (copyMakeBorder, motion_framework.cpp:60-61): a video loop may overwrite its buffers right after construction.
  double ms = 0.0;
  const int reps = 5;
  for (int r = 0; r < reps; ++r) {
    cv::Mat a = image1.clone(), b = image2.clone();
    const auto t0 = std::chrono::steady_clock::now();
    {
      MF again(a, b, search_size.data(), block_size.data(), num_levels);
      memset(a.data, 0, (size_t)w * h);  // the caller reuses its buffers
      memset(b.data, 255, (size_t)w * h);
      cv::Mat f2 = again.calcMotionBlockMatching();
      if (memcmp(f2.data, flow_res.data, (size_t)flow_res.rows * flow_res.cols * 8) != 0) return 6;
    }
    ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  printf("mf_ms_per_pair %.3f (constructor + calcMotionBlockMatching + destructor, %d objects)\n", ms / reps, reps);
  return 0;
}
