// common.cuh -- shared device helpers and launch-argument structs (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bbme {

constexpr int kMaxLevels = 16;

// One pyramid level of one frame set: `n` planes (one per pair of the chunk) of pitch x h bytes.
struct ImgView {
  const uint8_t* p;
  int w, h, pitch;
  size_t plane;  // bytes between pairs
};

// Block-granular motion field: gw x gh entries of short2 (x = u, y = v), `plane` entries between pairs.
struct MvView {
  short2* p;
  int gw, gh;
  size_t plane;
};

// 4 packed |a-b| + c : one VABSDIFF4.U8.ACC on sm_100a (64 lanes/clk/SM, measured in bench_micro/int_peak.cu).
__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// Visit rank of displacement (dx right, dy down) in the reference's spiral walk
// (motion_framework.cpp:326-411): ring r = max(|dx|,|dy|), rings walked right-column-down,
// bottom-row-left, left-column-up, top-row-right.  Ties in SAD go to the smaller rank (strict '<', :339).
__device__ __forceinline__ uint32_t spiral_rank(int dx, int dy) {
  int ax = abs(dx), ay = abs(dy);
  int r = max(ax, ay);
  if (r == 0) return 0u;
  int base = (2 * r - 1) * (2 * r - 1);
  int v;
  if (dx == r && dy > -r) v = base + dy + r - 1;
  else if (dy == r) v = base + 2 * r + (r - 1 - dx);
  else if (dx == -r) v = base + 4 * r + (r - 1 - dy);
  else v = base + 6 * r + (dx + r - 1);
  return (uint32_t)v;
}

// SAD of a bs x bs block: `a` is 4-byte aligned when bs >= 4 (2-byte aligned for bs == 2), `b` has any
// alignment.  Reads up to 3 bytes past the end of each `b` row (inside the row pitch / buffer slack).
__device__ __forceinline__ uint32_t sad_block_unaligned(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                        int pitch, int bs) {
  uint32_t acc = 0;
  if (bs >= 4) {
    const int words = bs >> 2;
    for (int r = 0; r < bs; ++r) {
      const uint32_t* aw = reinterpret_cast<const uint32_t*>(a + (size_t)r * pitch);
      const uint8_t* bb = b + (size_t)r * pitch;
      const uintptr_t ab = reinterpret_cast<uintptr_t>(bb);
      const uint32_t* bw = reinterpret_cast<const uint32_t*>(ab & ~(uintptr_t)3);
      const uint32_t sh = (uint32_t)(ab & 3) * 8u;
      uint32_t w0 = __ldg(bw);
      for (int k = 0; k < words; ++k) {
        uint32_t w1 = __ldg(bw + k + 1);
        acc = sad4(__ldg(aw + k), __funnelshift_r(w0, w1, sh), acc);
        w0 = w1;
      }
    }
  } else {  // bs == 2
    uint32_t a0 = *reinterpret_cast<const uint16_t*>(a);
    uint32_t a1 = *reinterpret_cast<const uint16_t*>(a + pitch);
    uint32_t bv = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[pitch] << 16) | ((uint32_t)b[pitch + 1] << 24);
    acc = sad4(a0 | (a1 << 16), bv, 0u);
  }
  return acc;
}

__device__ __forceinline__ uint32_t pack_mv(short2 v) {
  return (uint32_t)(uint16_t)v.x | ((uint32_t)(uint16_t)v.y << 16);
}

}  // namespace bbme
