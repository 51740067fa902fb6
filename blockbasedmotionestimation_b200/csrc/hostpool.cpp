// hostpool.cpp -- see hostpool.h.
#include "hostpool.h"

#include <immintrin.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <deque>
#include <thread>
#include <vector>

namespace bbme {

void Ticket::done(int n) {
  // decrement and notify under the lock: a waiter that sees zero may destroy the ticket as soon as it owns the lock again
  std::lock_guard<std::mutex> lk(mu);
  if (pending.fetch_sub(n, std::memory_order_acq_rel) == n) cv.notify_all();
}

void Ticket::wait() {
  std::unique_lock<std::mutex> lk(mu);
  cv.wait(lk, [&] { return pending.load(std::memory_order_acquire) == 0; });
}

// ---- the expansion: entry (u, v) int16 -> floats (u, v, u, v) in two consecutive pixel rows ----------------------

static void expand_rows_scalar(const int16_t* src, int gw2, int y0, int y1, float* dst, size_t pw) {
  for (int y2 = y0; y2 < y1; ++y2) {
    const int16_t* s = src + (size_t)y2 * gw2 * 2;
    float* o0 = dst + (size_t)(2 * y2) * pw * 2;
    float* o1 = o0 + pw * 2;
    for (int x = 0; x < gw2; ++x) {
      const float u = (float)s[2 * x], v = (float)s[2 * x + 1];
      o0[4 * x] = u; o0[4 * x + 1] = v; o0[4 * x + 2] = u; o0[4 * x + 3] = v;
      o1[4 * x] = u; o1[4 * x + 1] = v; o1[4 * x + 2] = u; o1[4 * x + 3] = v;
    }
  }
}

__attribute__((target("avx2"))) static void expand_rows_avx2(const int16_t* src, int gw2, int y0, int y1, float* dst,
                                                              size_t pw) {
  const __m256i lo_idx = _mm256_setr_epi32(0, 1, 0, 1, 2, 3, 2, 3);
  const __m256i hi_idx = _mm256_setr_epi32(4, 5, 4, 5, 6, 7, 6, 7);
  // non-temporal stores need 32-byte aligned addresses: rows are pw * 8 bytes apart
  const bool nt = ((reinterpret_cast<uintptr_t>(dst) & 31) == 0) && ((pw & 3) == 0);
  for (int y2 = y0; y2 < y1; ++y2) {
    const int16_t* s = src + (size_t)y2 * gw2 * 2;
    float* o0 = dst + (size_t)(2 * y2) * pw * 2;
    float* o1 = o0 + pw * 2;
    int x = 0;
    for (; x + 4 <= gw2; x += 4) {
      const __m128i v16 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 2 * x));
      const __m256 f = _mm256_cvtepi32_ps(_mm256_cvtepi16_epi32(v16));
      const __m256 lo = _mm256_permutevar8x32_ps(f, lo_idx), hi = _mm256_permutevar8x32_ps(f, hi_idx);
      if (nt) {
        _mm256_stream_ps(o0 + 4 * x, lo); _mm256_stream_ps(o0 + 4 * x + 8, hi);
        _mm256_stream_ps(o1 + 4 * x, lo); _mm256_stream_ps(o1 + 4 * x + 8, hi);
      } else {
        _mm256_storeu_ps(o0 + 4 * x, lo); _mm256_storeu_ps(o0 + 4 * x + 8, hi);
        _mm256_storeu_ps(o1 + 4 * x, lo); _mm256_storeu_ps(o1 + 4 * x + 8, hi);
      }
    }
    for (; x < gw2; ++x) {
      const float u = (float)s[2 * x], v = (float)s[2 * x + 1];
      o0[4 * x] = u; o0[4 * x + 1] = v; o0[4 * x + 2] = u; o0[4 * x + 3] = v;
      o1[4 * x] = u; o1[4 * x + 1] = v; o1[4 * x + 2] = u; o1[4 * x + 3] = v;
    }
  }
  if (nt) _mm_sfence();
}

void expand_rows(const int16_t* src, int gw2, int y0, int y1, float* dst, size_t pw) {
  static const bool have_avx2 = __builtin_cpu_supports("avx2");
  if (have_avx2) expand_rows_avx2(src, gw2, y0, y1, dst, pw);
  else expand_rows_scalar(src, gw2, y0, y1, dst, pw);
}

__attribute__((target("avx2"))) static void fill_nt_avx2(float* p, size_t n_floats, float val) {
  const __m256 v = _mm256_set1_ps(val);
  for (size_t i = 0; i + 8 <= n_floats; i += 8) _mm256_stream_ps(p + i, v);
  _mm_sfence();
}

// ---- the pool --------------------------------------------------------------------------------------------------

struct Task {
  int kind;  // 0 = expand, 1 = fill
  const int16_t* src;
  int gw2, y0, y1;
  float* dst;
  size_t pw;
  Ticket* ticket;
};

struct HostPool::Impl {
  std::mutex mu;
  std::condition_variable cv;
  std::deque<Task> q;
  std::vector<std::thread> th;
  bool stop = false;

  void run() {
    for (;;) {
      Task t;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return stop || !q.empty(); });
        if (q.empty()) return;
        t = q.front();
        q.pop_front();
      }
      if (t.kind == 0) {
        expand_rows(t.src, t.gw2, t.y0, t.y1, t.dst, t.pw);
      } else {
        const size_t n = t.pw;
        if (__builtin_cpu_supports("avx2") && (reinterpret_cast<uintptr_t>(t.dst) & 31) == 0) fill_nt_avx2(t.dst, n, 1.0f);
        else for (size_t i = 0; i < n; ++i) t.dst[i] = 1.0f;
      }
      t.ticket->done(1);
    }
  }
};

HostPool::HostPool() : impl_(new Impl()), nthreads_(0) {
  int n = 0;
  if (const char* e = getenv("BBME_HOST_THREADS")) n = atoi(e);
  if (n <= 0) {
    // one process per GPU under torchrun: share the host cores between the local ranks
    int local = 1;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) local = atoi(e) > 0 ? atoi(e) : 1;
    const int hw = (int)std::thread::hardware_concurrency();
    n = (hw > 0 ? hw : 8) / local;
    if (n > 16) n = 16;
    if (n < 2) n = 2;
  }
  nthreads_ = n;
  for (int i = 0; i < n; ++i) impl_->th.emplace_back([this] { impl_->run(); });
}

HostPool::~HostPool() {
  {
    std::lock_guard<std::mutex> lk(impl_->mu);
    impl_->stop = true;
  }
  impl_->cv.notify_all();
  for (auto& t : impl_->th) t.join();
  delete impl_;
}

HostPool& HostPool::instance() {
  static HostPool* pool = new HostPool();  // intentionally never destroyed: worker threads may outlive static destructors
  return *pool;
}

void HostPool::expand_async(const int16_t* src, int gw2, int gh2, float* dst, size_t pw, Ticket* ticket) {
  const int band = 32;  // compact rows per task: 64 dense rows, 0.5-1 MB of output at 1080p-4K
  const int nb = (gh2 + band - 1) / band;
  ticket->add(nb);
  {
    std::lock_guard<std::mutex> lk(impl_->mu);
    for (int b = 0; b < nb; ++b) {
      Task t{0, src, gw2, b * band, (b + 1) * band < gh2 ? (b + 1) * band : gh2, dst, pw, ticket};
      impl_->q.push_back(t);
    }
  }
  impl_->cv.notify_all();
}

double HostPool::measure_stream_write(void* dst, size_t bytes, int reps) {
  const size_t n_floats = bytes / 4;
  const size_t per = ((n_floats / (size_t)(nthreads_ * 4)) / 8) * 8;  // four tasks per thread
  if (per == 0) return 0.0;
  double best = 0.0;
  for (int r = 0; r < reps; ++r) {
    Ticket tk;
    const auto t0 = std::chrono::steady_clock::now();
    size_t off = 0;
    int cnt = 0;
    {
      std::lock_guard<std::mutex> lk(impl_->mu);
      while (off + per <= n_floats) {
        Task t{1, nullptr, 0, 0, 0, static_cast<float*>(dst) + off, per, &tk};
        tk.add(1);
        impl_->q.push_back(t);
        off += per;
        ++cnt;
      }
    }
    impl_->cv.notify_all();
    tk.wait();
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const double gbs = (double)off * 4.0 / s / 1e9;
    if (gbs > best) best = gbs;
  }
  return best;
}

}  // namespace bbme
