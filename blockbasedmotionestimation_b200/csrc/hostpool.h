// hostpool.h -- host-side worker threads of libbbme.so.
//
// The reference returns a dense padded CV_32FC2 field (motion_framework.cpp:218) in which every 2x2 pixel block shares
// one integer-valued vector (motion_framework.cpp:205-206).  Shipping that field over the host link costs 8 bytes per
// pixel; the device's own result is the 2x2-granular int16 field, 1/8 of the bytes.  The host-buffer entry points
// therefore copy the compact field to a pinned staging buffer and these threads expand it into the caller's dense
// buffer (int16 -> float, 2x2 replication, non-temporal stores).  Same bytes in the caller's buffer, 8x less link traffic.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <mutex>

namespace bbme {

// Completion counter of a group of tasks (one staging buffer, one call ...).
struct Ticket {
  std::atomic<int> pending{0};
  std::mutex mu;
  std::condition_variable cv;
  void add(int n) { pending.fetch_add(n, std::memory_order_acq_rel); }
  void done(int n);
  void wait();
};

// compact rows [y0, y1) of a gw2-wide short2 field -> dense rows [2*y0, 2*y1) of a pw-pixel-wide float2 field
void expand_rows(const int16_t* src, int gw2, int y0, int y1, float* dst, size_t pw);

class HostPool {
 public:
  static HostPool& instance();  // process-wide, created on first use (BBME_HOST_THREADS overrides the thread count)
  int threads() const { return nthreads_; }
  // Expands one pair's compact field (gh2 rows of gw2 short2) into dst (2*gh2 rows of pw float2) in row bands;
  // `ticket` is incremented per band before this returns and decremented as the bands finish.
  void expand_async(const int16_t* src, int gw2, int gh2, float* dst, size_t pw, Ticket* ticket);
  // Fills `bytes` of dst with non-temporal stores on all threads, returns GB/s (the host-side ceiling of the expansion).
  double measure_stream_write(void* dst, size_t bytes, int reps);

 private:
  HostPool();
  ~HostPool();
  struct Impl;
  Impl* impl_;
  int nthreads_;
};

}  // namespace bbme
