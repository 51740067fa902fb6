// multi.cpp -- what sits between the reference's calling pattern and the planned contexts of capi.cu:
//
//  * bbme_mf_open / bbme_mf_close: the reference's caller builds one MF object per frame pair (main_class.cpp:45-50).  A
//    context + plan is streams, ~40 MB per pair of device memory and TMA descriptors; paying that per object would dwarf the
//    1-2 ms the estimation takes.  Closed contexts are parked in a small geometry-keyed cache and handed to the next
//    MF of the same geometry.
//  * bbme_pool_*: one planned context per visible GPU behind one call.  Frame pairs are independent (one MF per pair), so
//    a batch is cut into contiguous shards, one host thread per GPU runs its shard through bbme_estimate_batch, and the
//    fields land in the caller's host array -- no inter-GPU traffic at all, no NCCL needed when results go to host memory.
//
// Everything here goes through the public C ABI of include/bbme.h (no access to the context's internals).
#include <string.h>

#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bbme.h"

namespace {

struct CacheKey {
  int device, w, h, levels, sweeps;
  int ss[BBME_MAX_LEVELS], bs[BBME_MAX_LEVELS];
  bool operator==(const CacheKey& o) const {
    if (device != o.device || w != o.w || h != o.h || levels != o.levels || sweeps != o.sweeps) return false;
    for (int i = 0; i < levels; ++i)
      if (ss[i] != o.ss[i] || bs[i] != o.bs[i]) return false;
    return true;
  }
};

struct CacheEntry {
  CacheKey key;
  bbme_ctx* ctx;
  bbme_shape shape;
  bool idle;
};

std::mutex g_cache_mu;
std::vector<CacheEntry> g_cache;  // idle and handed-out contexts (the latter to find the key again at close)
constexpr size_t kMaxIdle = 4;

}  // namespace

struct bbme_pool {
  std::vector<int> devices;
  std::vector<bbme_ctx*> ctx;
  std::string err;
  bool planned = false;
};

extern "C" {

int bbme_mf_open(bbme_ctx** out, int device, int width, int height, int num_levels, const int* search_size,
                 const int* block_size, int sweeps, bbme_shape* shape) {
  if (!out || !search_size || !block_size || num_levels <= 0 || num_levels > BBME_MAX_LEVELS) return BBME_E_ARG;
  *out = nullptr;
  CacheKey key;
  memset(&key, 0, sizeof(key));
  key.device = device; key.w = width; key.h = height; key.levels = num_levels; key.sweeps = sweeps;
  for (int i = 0; i < num_levels; ++i) { key.ss[i] = search_size[i]; key.bs[i] = block_size[i]; }
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (CacheEntry& e : g_cache) {
      if (e.idle && e.key == key) {
        e.idle = false;
        *out = e.ctx;
        if (shape) *shape = e.shape;
        return BBME_OK;
      }
    }
  }
  bbme_ctx* c = nullptr;
  int rc = bbme_create(&c, device);
  if (rc != BBME_OK) return rc;
  bbme_options opt;
  bbme_default_options(&opt);
  opt.sweeps = sweeps;
  CacheEntry e;
  e.key = key;
  e.ctx = c;
  e.idle = false;
  rc = bbme_plan(c, width, height, num_levels, search_size, block_size, &opt, &e.shape);
  if (rc != BBME_OK) {
    *out = c;  // the caller reads bbme_last_error(ctx) and then calls bbme_mf_close, which destroys an unplanned context
    return rc;
  }
  if (shape) *shape = e.shape;
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_cache.push_back(e);
  }
  *out = c;
  return BBME_OK;
}

void bbme_mf_close(bbme_ctx* ctx) {
  if (!ctx) return;
  bbme_ctx* victim = nullptr;
  bool known = false;
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    size_t idle = 0;
    for (CacheEntry& e : g_cache) {
      if (e.ctx == ctx) { e.idle = true; known = true; }
      if (e.idle) ++idle;
    }
    if (known && idle > kMaxIdle) {  // park at most kMaxIdle contexts: drop the oldest idle one
      for (size_t i = 0; i < g_cache.size(); ++i) {
        if (g_cache[i].idle && g_cache[i].ctx != ctx) {
          victim = g_cache[i].ctx;
          g_cache.erase(g_cache.begin() + (long)i);
          break;
        }
      }
    }
  }
  if (!known) victim = ctx;  // failed plan: never entered the cache
  if (victim) bbme_destroy(victim);
}

void bbme_mf_cache_clear(void) {
  std::vector<bbme_ctx*> victims;
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (size_t i = 0; i < g_cache.size();) {
      if (g_cache[i].idle) {
        victims.push_back(g_cache[i].ctx);
        g_cache.erase(g_cache.begin() + (long)i);
      } else {
        ++i;
      }
    }
  }
  for (bbme_ctx* c : victims) bbme_destroy(c);
}

// ---------------------------------------------------------------------------------------------------------- pool

int bbme_pool_create(bbme_pool** out, int n_devices, const int* devices) {
  if (!out || n_devices < 0) return BBME_E_ARG;
  *out = nullptr;
  bbme_pool* p = new bbme_pool();
  if (n_devices == 0) {
    // all visible devices: keep creating contexts until the device index runs out
    for (int d = 0; d < 64; ++d) {
      bbme_ctx* c = nullptr;
      if (bbme_create(&c, d) != BBME_OK) break;
      p->devices.push_back(d);
      p->ctx.push_back(c);
    }
    if (p->ctx.empty()) {
      delete p;
      return BBME_E_CUDA;  // bbme_last_error(NULL) has the reason ("no usable CUDA device ...")
    }
  } else {
    for (int i = 0; i < n_devices; ++i) {
      bbme_ctx* c = nullptr;
      const int d = devices ? devices[i] : i;
      const int rc = bbme_create(&c, d);
      if (rc != BBME_OK) {
        for (bbme_ctx* q : p->ctx) bbme_destroy(q);
        delete p;
        return rc;
      }
      p->devices.push_back(d);
      p->ctx.push_back(c);
    }
  }
  *out = p;
  return BBME_OK;
}

void bbme_pool_destroy(bbme_pool* p) {
  if (!p) return;
  for (bbme_ctx* c : p->ctx) bbme_destroy(c);
  delete p;
}

int bbme_pool_device_count(const bbme_pool* p) { return p ? (int)p->ctx.size() : 0; }

const char* bbme_pool_last_error(const bbme_pool* p) { return p ? p->err.c_str() : ""; }

int bbme_pool_plan(bbme_pool* p, int width, int height, int num_levels, const int* search_size, const int* block_size,
                   const bbme_options* opt, bbme_shape* out) {
  if (!p) return BBME_E_ARG;
  p->planned = false;
  for (size_t i = 0; i < p->ctx.size(); ++i) {
    const int rc = bbme_plan(p->ctx[i], width, height, num_levels, search_size, block_size, opt, i == 0 ? out : nullptr);
    if (rc != BBME_OK) {
      p->err = "device " + std::to_string(p->devices[i]) + ": " + bbme_last_error(p->ctx[i]);
      return rc;
    }
  }
  p->planned = true;
  return BBME_OK;
}

int bbme_pool_estimate_batch(bbme_pool* p, int n, const uint8_t* const* im1, const uint8_t* const* im2, size_t pitch_bytes,
                             float* const* flow) {
  if (!p || n <= 0 || !im1 || !im2 || !flow) return BBME_E_ARG;
  if (!p->planned) {
    p->err = "bbme_pool_estimate_batch before bbme_pool_plan";
    return BBME_E_STATE;
  }
  const int g = (int)p->ctx.size();
  std::vector<int> rc((size_t)g, BBME_OK);
  std::vector<std::thread> th;
  // contiguous, balanced shards: device i takes pairs [start_i, start_i + cnt_i)
  int start = 0;
  for (int i = 0; i < g; ++i) {
    const int cnt = n / g + (i < n % g ? 1 : 0);
    if (cnt > 0) {
      bbme_ctx* c = p->ctx[(size_t)i];
      int* r = &rc[(size_t)i];
      const int s0 = start;
      th.emplace_back([=] { *r = bbme_estimate_batch(c, cnt, im1 + s0, im2 + s0, pitch_bytes, flow + s0); });
    }
    start += cnt;
  }
  for (std::thread& t : th) t.join();
  for (int i = 0; i < g; ++i) {
    if (rc[(size_t)i] != BBME_OK) {
      p->err = "device " + std::to_string(p->devices[(size_t)i]) + ": " + bbme_last_error(p->ctx[(size_t)i]);
      return rc[(size_t)i];
    }
  }
  return BBME_OK;
}

}  // extern "C"
