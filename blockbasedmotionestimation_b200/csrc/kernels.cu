// kernels.cu -- pyramid build, generic block search, MV plumbing and the exact regularisation sweep.
//
// Reference semantics (cited as file:line of /root/reference) are restated in DESIGN.md; nothing here is a
// translation of the reference's loops: the data layout is pitched uint8 planes + block-granular short2
// fields, and the in-place raster sweep is reproduced by a Jacobi pass followed by fixed-point rounds.
#include "kernels.h"

#include <cooperative_groups.h>
#include <float.h>
#include <math.h>
#include <stdlib.h>

namespace bbme {

// ============================================================================================ pad
// cv::copyMakeBorder(BORDER_CONSTANT, 0) of both frames (motion_framework.cpp:60-61).
// One thread writes 16 output bytes; the zero border and the zero pitch tail are written too.
__global__ void __launch_bounds__(256) k_pad(const uint8_t* __restrict__ in1, const uint8_t* __restrict__ in2,
                                             size_t in_pitch, size_t in_plane, int w, int h, int pad_x, int pad_y,
                                             uint8_t* __restrict__ out1, uint8_t* __restrict__ out2, int out_pitch,
                                             size_t out_plane, int ph) {
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
  const int y = blockIdx.y;
  const int pair = blockIdx.z >> 1;
  const int frame = blockIdx.z & 1;
  if (x0 >= out_pitch || y >= ph) return;
  const uint8_t* in = (frame ? in2 : in1) + (size_t)pair * in_plane;
  uint8_t* out = (frame ? out2 : out1) + (size_t)pair * out_plane + (size_t)y * out_pitch + x0;
  const int sy = y - pad_y;
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (sy >= 0 && sy < h) {
    const int sx = x0 - pad_x;
    const uint8_t* src = in + (size_t)sy * in_pitch + sx;
    if (sx >= 0 && sx + 16 <= w && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
      v = __ldg(reinterpret_cast<const uint4*>(src));
    } else if (sx + 16 > 0 && sx < w) {
      uint32_t wd[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        int xx = sx + i;
        uint32_t b = (xx >= 0 && xx < w) ? (uint32_t)__ldg(src + i) : 0u;
        wd[i >> 2] |= b << ((i & 3) * 8);
      }
      v = make_uint4(wd[0], wd[1], wd[2], wd[3]);
    }
  }
  *reinterpret_cast<uint4*>(out) = v;
}

void launch_pad(const uint8_t* in1, const uint8_t* in2, size_t in_pitch, size_t in_plane, int w, int h, int pad_x,
                int pad_y, uint8_t* out1, uint8_t* out2, int out_pitch, size_t out_plane, int pw, int ph, int n,
                cudaStream_t s) {
  (void)pw;
  dim3 block(128);
  dim3 grid((out_pitch / 16 + block.x - 1) / block.x, ph, 2 * n);
  k_pad<<<grid, block, 0, s>>>(in1, in2, in_pitch, in_plane, w, h, pad_x, pad_y, out1, out2, out_pitch, out_plane, ph);
}

// ============================================================================================ resize + pad
// main()'s quarter-pel wrapper (main_class.cpp:32-33): cv::resize(img, img, Size(), f, f, INTER_LINEAR) on 8-bit frames,
// fused with the copyMakeBorder of MF::MF (motion_framework.cpp:60-61): the up-sampled frame is never stored unpadded.
// OpenCV's fixed-point algorithm (pinned against cv2 by tests/golden/resize_cv2.npz): 11-bit tap weights,
// horizontal pass in int32, vertical pass (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2; x taps that
// fall outside move inside and lose their weight, y taps are clamped by row index.  For the power-of-two factors
// supported here the tap position of destination index d = q * f + p is exactly q + off[p] with weight wt[p].
__global__ void __launch_bounds__(128) k_resize_pad(const uint8_t* __restrict__ in1, const uint8_t* __restrict__ in2,
                                                    size_t in_pitch, size_t in_plane, int w, int h, ResizeTaps taps,
                                                    int pad_x, int pad_y, uint8_t* __restrict__ out1,
                                                    uint8_t* __restrict__ out2, int out_pitch, size_t out_plane, int ph) {
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
  const int y = blockIdx.y;
  const int pair = blockIdx.z >> 1;
  const int frame = blockIdx.z & 1;
  if (x0 >= out_pitch || y >= ph) return;
  const uint8_t* in = (frame ? in2 : in1) + (size_t)pair * in_plane;
  uint8_t* out = (frame ? out2 : out1) + (size_t)pair * out_plane + (size_t)y * out_pitch + x0;
  const int f = taps.factor, sh = taps.shift;
  const int dy = y - pad_y;
  uint32_t wd[4] = {0u, 0u, 0u, 0u};
  if (dy >= 0 && dy < h * f) {
    const int py = dy & (f - 1);
    const int sy = (dy >> sh) + taps.off[py];
    const int b1 = taps.wt[py], b0 = 2048 - b1;
    const uint8_t* r0 = in + (size_t)min(max(sy, 0), h - 1) * in_pitch;
    const uint8_t* r1 = in + (size_t)min(max(sy + 1, 0), h - 1) * in_pitch;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int dx = x0 + i - pad_x;
      if (dx < 0 || dx >= w * f) continue;
      const int px = dx & (f - 1);
      int sx = (dx >> sh) + taps.off[px];
      int a1 = taps.wt[px];
      if (sx < 0) { sx = 0; a1 = 0; }
      if (sx >= w - 1) { sx = w - 1; a1 = 0; }
      const int a0 = 2048 - a1;
      const int sx1 = min(sx + 1, w - 1);
      const int h0 = (int)__ldg(r0 + sx) * a0 + (int)__ldg(r0 + sx1) * a1;
      const int h1 = (int)__ldg(r1 + sx) * a0 + (int)__ldg(r1 + sx1) * a1;
      const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
      wd[i >> 2] |= (uint32_t)v << ((i & 3) * 8);
    }
  }
  *reinterpret_cast<uint4*>(out) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
}

// Same result when pad_x is a multiple of the factor (the usual case): the phase of output byte i of a 16-byte strip is
// then i % F at compile time, so the 16 / F + 2 source pixels a strip touches are loaded once per source row into
// registers (12 byte loads instead of 64 for F = 4).  Clamped loads reproduce the "tap moves inside and loses its weight"
// rule exactly: a clamped pair of taps reads the same pixel twice, and the weights sum to 2048.
template <int F>
__global__ void __launch_bounds__(128) k_resize_pad_aligned(const uint8_t* __restrict__ in1, const uint8_t* __restrict__ in2,
                                                            size_t in_pitch, size_t in_plane, int w, int h, ResizeTaps taps,
                                                            int pad_x, int pad_y, uint8_t* __restrict__ out1,
                                                            uint8_t* __restrict__ out2, int out_pitch, size_t out_plane,
                                                            int ph) {
  constexpr int SH = F == 2 ? 1 : (F == 4 ? 2 : 3);
  constexpr int NV = 16 / F + 2;
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
  const int y = blockIdx.y;
  const int pair = blockIdx.z >> 1;
  const int frame = blockIdx.z & 1;
  if (x0 >= out_pitch || y >= ph) return;
  const uint8_t* in = (frame ? in2 : in1) + (size_t)pair * in_plane;
  uint8_t* out = (frame ? out2 : out1) + (size_t)pair * out_plane + (size_t)y * out_pitch + x0;
  const int dy = y - pad_y;
  const int dx0 = x0 - pad_x;  // multiple of F
  uint32_t wd[4] = {0u, 0u, 0u, 0u};
  if (dy >= 0 && dy < h * F && dx0 + 16 > 0 && dx0 < w * F) {
    const int py = dy & (F - 1);
    const int sy = (dy >> SH) + taps.off[py];
    const int b1 = taps.wt[py], b0 = 2048 - b1;
    const uint8_t* r0 = in + (size_t)min(max(sy, 0), h - 1) * in_pitch;
    const uint8_t* r1 = in + (size_t)min(max(sy + 1, 0), h - 1) * in_pitch;
    const int q0 = dx0 >> SH;  // arithmetic shift: strips that start in the left padding have negative q0
    int v0[NV], v1[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int idx = min(max(q0 - 1 + k, 0), w - 1);
      v0[k] = (int)__ldg(r0 + idx);
      v1[k] = (int)__ldg(r1 + idx);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int p = i % F;
      const int j = i / F + (p < F / 2 ? 0 : 1);  // = i / F + off[p] + 1 with off[p] = -1 for the first half of the phases
      const int a1 = taps.wt[p], a0 = 2048 - a1;
      const int h0 = v0[j] * a0 + v0[j + 1] * a1;
      const int h1 = v1[j] * a0 + v1[j + 1] * a1;
      const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
      const int dx = dx0 + i;
      if (dx >= 0 && dx < w * F) wd[i >> 2] |= (uint32_t)v << ((i & 3) * 8);
    }
  }
  *reinterpret_cast<uint4*>(out) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
}

int make_resize_taps(int factor, ResizeTaps* t) {
  if (!(factor == 2 || factor == 4 || factor == 8)) return -1;
  t->factor = factor;
  t->shift = factor == 2 ? 1 : (factor == 4 ? 2 : 3);
  for (int p = 0; p < 8; ++p) { t->off[p] = 0; t->wt[p] = 0; }
  const double scale = 1.0 / (double)factor;
  for (int p = 0; p < factor; ++p) {
    // OpenCV: fx = (float)((dx + 0.5) * scale_x - 0.5); sx = cvFloor(fx); fx -= sx; weight = cvRound(fx * 2048)
    float fx = (float)(((double)p + 0.5) * scale - 0.5);
    const int sx = (int)floorf(fx);
    fx -= (float)sx;
    t->off[p] = sx;
    t->wt[p] = (int)lrintf(fx * 2048.f);
  }
  return 0;
}

void launch_resize_pad(const uint8_t* in1, const uint8_t* in2, size_t in_pitch, size_t in_plane, int w, int h,
                       const ResizeTaps& taps, int pad_x, int pad_y, uint8_t* out1, uint8_t* out2, int out_pitch,
                       size_t out_plane, int ph, int n, cudaStream_t s) {
  dim3 block(128);
  dim3 grid((out_pitch / 16 + block.x - 1) / block.x, ph, 2 * n);
  const bool aligned = pad_x % taps.factor == 0;
  if (aligned && taps.factor == 4)
    k_resize_pad_aligned<4><<<grid, block, 0, s>>>(in1, in2, in_pitch, in_plane, w, h, taps, pad_x, pad_y, out1, out2, out_pitch, out_plane, ph);
  else if (aligned && taps.factor == 2)
    k_resize_pad_aligned<2><<<grid, block, 0, s>>>(in1, in2, in_pitch, in_plane, w, h, taps, pad_x, pad_y, out1, out2, out_pitch, out_plane, ph);
  else if (aligned && taps.factor == 8)
    k_resize_pad_aligned<8><<<grid, block, 0, s>>>(in1, in2, in_pitch, in_plane, w, h, taps, pad_x, pad_y, out1, out2, out_pitch, out_plane, ph);
  else
    k_resize_pad<<<grid, block, 0, s>>>(in1, in2, in_pitch, in_plane, w, h, taps, pad_x, pad_y, out1, out2, out_pitch,
                                        out_plane, ph);
}

// ============================================================================================ pyrDown
// cv::pyrDown(src, dst, Size(cols/2, rows/2)) for 8-bit (motion_framework.cpp:89-90): 5x5 separable
// [1 4 6 4 1], BORDER_REFLECT_101, (sum + 128) >> 8.
// One thread owns one aligned 16-byte source strip [16t, 16t+16) -> 8 output pixels; per source row it issues one
// coalesced 128-bit load and fetches the 2-byte left / 1-byte right halo from the neighbouring lanes with two
// shuffles (lanes at a warp or image edge read the halo bytes directly, with BORDER_REFLECT_101).  HBM/L2-bound.
__device__ __forceinline__ int reflect101(int p, int len) {
  if (p < 0) p = -p;
  if (p >= len) p = 2 * len - 2 - p;
  return p;
}

__global__ void __launch_bounds__(128) k_pyrdown(ImgView s1, ImgView s2, uint8_t* __restrict__ d1,
                                                 uint8_t* __restrict__ d2, int dw, int dh, int dpitch, size_t dplane) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;  // strip index: source bytes 16t..16t+15, outputs 8t..8t+7
  const int lane = threadIdx.x & 31;
  const int y = blockIdx.y;
  const int pair = blockIdx.z >> 1;
  const int frame = blockIdx.z & 1;
  const ImgView sv = frame ? s2 : s1;
  const uint8_t* src = sv.p + (size_t)pair * sv.plane;
  const int sw = sv.w, sh = sv.h;
  const bool live = 8 * t < dw;            // whole warps stay alive for the shuffles
  const bool full = 16 * t + 16 <= sv.pitch;
  // Output i of the strip is the 5x5 window centred on source pixel 16t + 2i.  The taps stay packed: a window row is
  // four bytes of one (half-word shifted) source word times (1,4,6,4) plus one byte of the next word, i.e. two
  // byte-dot-products (IDP.4A) that accumulate straight into acc[i]; the vertical weight of the row is folded into the
  // dot-product constants (6 * 6 = 36 fits a byte).  100 instructions per strip instead of ~300 with unpacked pixels --
  // the kernel was bound by the ALU pipe, not by HBM.
  uint32_t acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0u;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const uint32_t wj = (j == 0 || j == 4) ? 1u : ((j == 2) ? 6u : 4u);
    const uint32_t k4 = wj * 0x04060401u;  // bytes (b0..b3) x (1,4,6,4)
    const uint32_t k_b2 = wj << 16;        // byte 2 x 1
    const uint32_t k_b0 = wj;              // byte 0 x 1
    const uint8_t* row = src + (size_t)reflect101(2 * y + j - 2, sh) * sv.pitch;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (live && full) v = __ldg(reinterpret_cast<const uint4*>(row + 16 * t));
    // halo: bytes 16t-2, 16t-1 (high half of the left neighbour's last word) and 16t+16 (right neighbour's first byte)
    uint32_t L = __shfl_up_sync(0xffffffffu, v.w, 1);
    uint32_t R = __shfl_down_sync(0xffffffffu, v.x, 1);
    if (live) {
      if (lane == 0 || t == 0)
        L = ((uint32_t)__ldg(row + reflect101(16 * t - 2, sw)) << 16) | ((uint32_t)__ldg(row + reflect101(16 * t - 1, sw)) << 24);
      if (lane == 31 || 16 * t + 16 >= sw) R = (uint32_t)__ldg(row + reflect101(16 * t + 16, sw));
      uint32_t w[6] = {L, v.x, v.y, v.z, v.w, R};
      if (16 * t + 16 > sw || !full) {  // ragged right edge: pixels past the image width are reflected, not read from the pitch tail
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          if (16 * t + q >= sw || !full) {
            const uint32_t px = (uint32_t)__ldg(row + reflect101(16 * t + q, sw));
            w[1 + (q >> 2)] = (w[1 + (q >> 2)] & ~(0xffu << ((q & 3) * 8))) | (px << ((q & 3) * 8));
          }
        }
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        // even output 2m: pixels 16t+4m-2 .. 16t+4m+2 = (w[m].b2, w[m].b3, w[m+1].b0, w[m+1].b1), w[m+1].b2
        const uint32_t x = __funnelshift_r(w[m], w[m + 1], 16);
        acc[2 * m] = __dp4a(x, k4, acc[2 * m]);
        acc[2 * m] = __dp4a(w[m + 1], k_b2, acc[2 * m]);
        // odd output 2m+1: pixels 16t+4m .. 16t+4m+4 = w[m+1].b0..b3, w[m+2].b0
        acc[2 * m + 1] = __dp4a(w[m + 1], k4, acc[2 * m + 1]);
        acc[2 * m + 1] = __dp4a(w[m + 2], k_b0, acc[2 * m + 1]);
      }
    }
  }
  if (!live || y >= dh) return;
  uint8_t* dst = (frame ? d2 : d1) + (size_t)pair * dplane + (size_t)y * dpitch + 8 * t;
  uint32_t lo = 0, hi = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    lo |= ((acc[i] + 128u) >> 8) << (8 * i);
    hi |= ((acc[4 + i] + 128u) >> 8) << (8 * i);
  }
  if (8 * t + 8 <= dw) {
    *reinterpret_cast<uint2*>(dst) = make_uint2(lo, hi);
  } else {
    for (int i = 0; 8 * t + i < dw; ++i) dst[i] = (uint8_t)((i < 4 ? lo >> (8 * i) : hi >> (8 * (i - 4))) & 0xffu);
  }
}

void launch_pyrdown(ImgView src1, ImgView src2, uint8_t* dst1, uint8_t* dst2, int dpitch, size_t dplane, int n,
                    cudaStream_t s) {
  const int dw = src1.w / 2, dh = src1.h / 2;
  dim3 block(128);
  dim3 grid(((dw + 7) / 8 + block.x - 1) / block.x, dh, 2 * n);
  k_pyrdown<<<grid, block, 0, s>>>(src1, src2, dst1, dst2, dw, dh, dpitch, dplane);
}

// ============================================================================================ generic search
// MF::calcLevelBM + find_min_block_spiral (motion_framework.cpp:226-244, 296-422) for any power-of-two block
// size: one CTA per block, threads stride over the (2R+1)^2 displacements, argmin on the key (SAD, spiral rank).
// Bring-up / fallback path (block sizes the TMA kernel does not cover) and the in-library cross-check of it.
// variant 1 = MF::find_min_block (motion_framework.cpp:246-294, the raster-scan search the commented line :235 would call): no
// centre test -- the window is clamped to the image (:260,262) and an empty window keeps the prediction (:251-252) -- and ties
// go to the smaller L1 distance from the block's position in image 1, then to the earlier position in row-major order
// (:271-283): key = SAD << 32 | L1 << 17 | row-major index (needs 2R + 1 <= 362, L1 < 2^15).
__global__ void __launch_bounds__(128) k_search_generic(ImgView i1, ImgView i2, MvView mv, int bs, int R,
                                                        unsigned long long* __restrict__ counters, int variant) {
  const int pair = blockIdx.y;
  const int bx = blockIdx.x % mv.gw, by = blockIdx.x / mv.gw;
  const int x = bx * bs, y = by * bs;
  const int w = i1.w, h = i1.h, pitch = i1.pitch;
  short2* slot = mv.p + (size_t)pair * mv.plane + (size_t)by * mv.gw + bx;
  const short2 pred = *slot;
  const int x2 = x + pred.x, y2 = y + pred.y;
  if (variant == 0 && (x2 < 0 || y2 < 0 || x2 + bs > w || y2 + bs > h)) {  // :304-310 -> MV 0, no search
    if (threadIdx.x == 0) *slot = make_short2(0, 0);
    return;
  }
  const uint8_t* a = i1.p + (size_t)pair * i1.plane + (size_t)y * pitch + x;
  const uint8_t* b0 = i2.p + (size_t)pair * i2.plane;
  const int n = 2 * R + 1;
  unsigned long long best = ~0ull;
  for (int c = threadIdx.x; c < n * n; c += blockDim.x) {
    const int dx = c % n - R, dy = c / n - R;
    const int px = x2 + dx, py = y2 + dy;
    if (px < 0 || py < 0 || px + bs > w || py + bs > h) continue;  // skipped, walk continues (:335-336)
    const uint32_t sad = sad_block_unaligned(a, b0 + (size_t)py * pitch + px, pitch, bs);
    const uint32_t rank = variant == 0 ? spiral_rank(dx, dy) : (((uint32_t)(abs(pred.x + dx) + abs(pred.y + dy)) << 17) | (uint32_t)c);
    const unsigned long long key = ((unsigned long long)sad << 32) | rank;
    best = key < best ? key : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other < best ? other : best;
  }
  __shared__ unsigned long long s_best[4];
  if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) best = s_best[i] < best ? s_best[i] : best;
    // decode the rank back into (dx, dy) by evaluating the same closed form over the ring
    const uint32_t rank = (uint32_t)best;
    int dx = 0, dy = 0;
    if (variant != 0) {
      if (best != ~0ull) {  // else: empty window, the prediction stays
        const int c = (int)(rank & 0x1ffffu);
        dx = c % n - R;
        dy = c / n - R;
      }
    } else if (rank != 0) {
      int r = 1;
      while ((uint32_t)((2 * r + 1) * (2 * r + 1)) <= rank) ++r;
      const int base = (2 * r - 1) * (2 * r - 1);
      const int o = (int)rank - base;
      if (o < 2 * r) { dx = r; dy = o - r + 1; }
      else if (o < 4 * r) { dy = r; dx = r - 1 - (o - 2 * r); }
      else if (o < 6 * r) { dx = -r; dy = r - 1 - (o - 4 * r); }
      else { dy = -r; dx = (o - 6 * r) - r + 1; }
    }
    *slot = make_short2((short)(pred.x + dx), (short)(pred.y + dy));
    if (counters) {
      const int nx = max(min(R, w - bs - x2) - max(-R, -x2) + 1, 0);
      const int ny = max(min(R, h - bs - y2) - max(-R, -y2) + 1, 0);
      atomicAdd(&counters[0], (unsigned long long)(nx * ny));
      atomicAdd(&counters[1], (unsigned long long)(nx * ny) * (unsigned long long)(bs * bs));
    }
  }
}

void launch_search_generic(ImgView i1, ImgView i2, MvView mv, int bs, int R, int n, unsigned long long* counters,
                           cudaStream_t s, int variant) {
  dim3 grid(mv.gw * mv.gh, n);
  k_search_generic<<<grid, 128, 0, s>>>(i1, i2, mv, bs, R, counters, variant);
}

// MF::draw_MVimage (motion_framework.cpp:887-905): the motion-compensated frame.  One thread per 4 output bytes of a block row
// (2x2 blocks: per block row); blocks whose source leaves the image keep the bytes already in `out`.
__global__ void __launch_bounds__(256) k_compensate(ImgView i2, const short2* __restrict__ mv, int gw, size_t mv_plane, int bs,
                                                    uint8_t* __restrict__ out, int out_pitch, size_t out_plane) {
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 2;  // two pixels per thread (block sizes are even)
  const int y = blockIdx.y;
  const int pair = blockIdx.z;
  if (x >= i2.w || y >= i2.h) return;
  const int bx = x / bs, by = y / bs;
  const short2 m = __ldg(&mv[(size_t)pair * mv_plane + (size_t)by * gw + bx]);
  const int x2 = bx * bs + m.x, y2 = by * bs + m.y;
  if (x2 < 0 || x2 > i2.w - bs || y2 < 0 || y2 > i2.h - bs) return;  // :897-898
  const uint8_t* src = i2.p + (size_t)pair * i2.plane + (size_t)(y2 + (y - by * bs)) * i2.pitch + x2 + (x - bx * bs);
  uint8_t* dst = out + (size_t)pair * out_plane + (size_t)y * out_pitch + x;
  dst[0] = __ldg(src);
  dst[1] = __ldg(src + 1);
}

void launch_compensate(ImgView i2, const short2* mv, int gw, size_t mv_plane, int bs, uint8_t* out, int out_pitch,
                       size_t out_plane, int n, cudaStream_t s) {
  dim3 grid((i2.w / 2 + 255) / 256, i2.h, n);
  k_compensate<<<grid, 256, 0, s>>>(i2, mv, gw, mv_plane, bs, out, out_pitch, out_plane);
}

// ============================================================================================ MV plumbing
// MF::copyMVs + fill_block_MV (motion_framework.cpp:828-843, 803-813): only the MV at each coarse block's
// top-left pixel (on the coarse level's INITIAL block grid) is propagated, doubled, over a 2bs x 2bs region.
__global__ void __launch_bounds__(256) k_copy_mvs(const short2* __restrict__ coarse, int cgw2, size_t cplane, int cbs,
                                                  MvView fine, int fbs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int pair = blockIdx.y;
  if (i >= fine.gw * fine.gh) return;
  const int bx = i % fine.gw, by = i / fine.gw;
  const int cx = (bx * fbs) / (2 * cbs), cy = (by * fbs) / (2 * cbs);  // coarse block (initial grid)
  const int half = cbs >> 1;                                          // its corner in the 2x2-granular field
  const short2 c = coarse[(size_t)pair * cplane + (size_t)(cy * half) * cgw2 + cx * half];
  fine.p[(size_t)pair * fine.plane + i] = make_short2((short)(2 * c.x), (short)(2 * c.y));
}

void launch_copy_mvs(const short2* coarse, int cgw2, size_t cplane, int cbs, MvView fine, int fbs, int n,
                     cudaStream_t s) {
  dim3 grid((fine.gw * fine.gh + 255) / 256, n);
  k_copy_mvs<<<grid, 256, 0, s>>>(coarse, cgw2, cplane, cbs, fine, fbs);
}

// MF::divide_blocks (motion_framework.cpp:845-862): every block hands its MV to its four quadrants.
__global__ void __launch_bounds__(256) k_divide(const short2* __restrict__ in, int gw, int gh, size_t in_plane,
                                                short2* __restrict__ out, size_t out_plane) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int pair = blockIdx.y;
  const int ow = 2 * gw;
  if (i >= ow * 2 * gh) return;
  const int x = i % ow, y = i / ow;
  out[(size_t)pair * out_plane + i] = in[(size_t)pair * in_plane + (size_t)(y >> 1) * gw + (x >> 1)];
}

// Same, two input entries per thread: one 8-byte load, two 16-byte stores (the rows 2y and 2y + 1 of the finer grid are
// identical).  Needs an even grid width and 16-byte aligned planes.
__global__ void __launch_bounds__(256) k_divide2(const short2* __restrict__ in, int gw, int gh, size_t in_plane,
                                                 short2* __restrict__ out, size_t out_plane) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // pair of input entries
  const int pair = blockIdx.y;
  const int hw = gw >> 1;
  if (i >= hw * gh) return;
  const int x2 = i % hw, y = i / hw;
  const uint2 v = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint32_t*>(in + (size_t)pair * in_plane) + (size_t)y * gw + 2 * x2);
  const uint4 o = make_uint4(v.x, v.x, v.y, v.y);
  uint32_t* dst = reinterpret_cast<uint32_t*>(out + (size_t)pair * out_plane) + (size_t)(2 * y) * (2 * gw) + 4 * x2;
  *reinterpret_cast<uint4*>(dst) = o;
  *reinterpret_cast<uint4*>(dst + 2 * gw) = o;
}

void launch_divide(const short2* in, int gw, int gh, size_t in_plane, short2* out, size_t out_plane, int n,
                   cudaStream_t s) {
  const bool vec = (gw & 1) == 0 && (in_plane & 1) == 0 && (out_plane & 3) == 0 &&
                   (reinterpret_cast<uintptr_t>(in) & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (vec) {
    dim3 grid((gw / 2 * gh + 255) / 256, n);
    k_divide2<<<grid, 256, 0, s>>>(in, gw, gh, in_plane, out, out_plane);
  } else {
    dim3 grid((4 * gw * gh + 255) / 256, n);
    k_divide<<<grid, 256, 0, s>>>(in, gw, gh, in_plane, out, out_plane);
  }
}

// Final dense field (motion_framework.cpp:205-206, 815-826, 218): CV_32FC2, every 2x2 block shares one MV.
// One thread = one entry of the 2x2-granular field = one 16-byte streaming store into each of the two pixel rows it
// covers; a CTA walks over (pair, entry row) pairs, so the grid is a few waves instead of a million tiny CTAs.
// HBM-write-bound (8 B / pixel).
__global__ void __launch_bounds__(256) k_export(const short2* __restrict__ mv2, int gw2, int gh2, size_t mv_plane,
                                                float* __restrict__ out, int pw, size_t out_plane, int n) {
  const bool aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (out_plane & 3) == 0 && (pw & 1) == 0;
  for (int row = blockIdx.x; row < n * gh2; row += gridDim.x) {
    const int pair = row / gh2, y2 = row - pair * gh2;
    const short2* src = mv2 + (size_t)pair * mv_plane + (size_t)y2 * gw2;
    float* o0 = out + (size_t)pair * out_plane + (size_t)(2 * y2) * pw * 2;
    float* o1 = o0 + (size_t)pw * 2;
    for (int x2 = threadIdx.x; x2 < gw2; x2 += blockDim.x) {
      const short2 m = __ldg(&src[x2]);
      const float u = (float)m.x, v = (float)m.y;
      if (aligned) {
        __stcs(reinterpret_cast<float4*>(o0 + 4 * x2), make_float4(u, v, u, v));
        __stcs(reinterpret_cast<float4*>(o1 + 4 * x2), make_float4(u, v, u, v));
      } else {
        float* a = o0 + 4 * x2;
        float* b = o1 + 4 * x2;
        a[0] = u; a[1] = v; a[2] = u; a[3] = v;
        b[0] = u; b[1] = v; b[2] = u; b[3] = v;
      }
    }
  }
}

void launch_export(const short2* mv2, int gw2, size_t mv_plane, float* out, int pw, int ph, size_t out_plane, int n,
                   cudaStream_t s) {
  const int gh2 = ph / 2;
  long long rows = (long long)n * gh2;
  int grid = (int)(rows < 148 * 16 ? rows : 148 * 16);
  k_export<<<grid, 256, 0, s>>>(mv2, gw2, gh2, mv_plane, out, pw, out_plane, n);
}

__global__ void __launch_bounds__(256) k_export_compact(const short2* __restrict__ mv2, int count, size_t mv_plane,
                                                        int16_t* __restrict__ out, size_t out_plane) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int pair = blockIdx.y;
  if (i >= count) return;
  reinterpret_cast<short2*>(out + (size_t)pair * out_plane)[i] = mv2[(size_t)pair * mv_plane + i];
}

void launch_export_compact(const short2* mv2, int gw2, int gh2, size_t mv_plane, int16_t* out, size_t out_plane, int n,
                           cudaStream_t s) {
  dim3 grid((gw2 * gh2 + 255) / 256, n);
  k_export_compact<<<grid, 256, 0, s>>>(mv2, gw2 * gh2, mv_plane, out, out_plane);
}

// main()'s post-processing (main_class.cpp:58-70) on the device: strip the padding, keep every factor-th pixel, divide
// the vectors by the factor -> (height / factor) x (width / factor) x 2 floats.  The dense padded field is never built.
__global__ void __launch_bounds__(256) k_export_subsample(const short2* __restrict__ mv2, int gw2, size_t mv_plane,
                                                          int pad_x, int pad_y, int factor, float* __restrict__ out,
                                                          int ow, int oh, size_t out_plane) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int pair = blockIdx.z;
  if (x >= ow || y >= oh) return;
  const int sx = pad_x + x * factor, sy = pad_y + y * factor;  // pixel (i, j) of the padded field, :62-68
  const short2 m = __ldg(&mv2[(size_t)pair * mv_plane + (size_t)(sy >> 1) * gw2 + (sx >> 1)]);
  const float inv = (float)factor;
  reinterpret_cast<float2*>(out + (size_t)pair * out_plane)[(size_t)y * ow + x] =
      make_float2(__fdiv_rn((float)m.x, inv), __fdiv_rn((float)m.y, inv));
}

void launch_export_subsample(const short2* mv2, int gw2, size_t mv_plane, int pad_x, int pad_y, int factor, float* out,
                             int ow, int oh, size_t out_plane, int n, cudaStream_t s) {
  dim3 grid((ow + 255) / 256, oh, n);
  k_export_subsample<<<grid, 256, 0, s>>>(mv2, gw2, mv_plane, pad_x, pad_y, factor, out, ow, oh, out_plane);
}

// ============================================================================================ regularisation
// MF::regularize_MVs / find_min_candidate / calculate_smoothness / min_energy_candidate
// (motion_framework.cpp:424-662).  Candidate slots in the reference's order [C, L, R, DR, UL, UR, U, D, DL];
// neighbours outside the grid are dropped (the reference's nine-way if/else chain, :438-522, is exactly that).
// In the reference's in-place raster sweep L, UL, U, UR have already been updated ("pred" neighbours, read
// from P) and C, R, DR, D, DL have not (read from O).
//
// Energy (:607) is float32 and un-fused: (float)SAD + ((lambda * (float)mult) * S); S is a sum of
// integer-valued floats (< 2^24, exact), so it is accumulated in int and converted once.
//
// Every evaluator is branch-free over the nine slots: all nine candidate windows are addressed first and their
// loads issued together (missing neighbours and out-of-image candidates read the block's own position and are masked
// out of the argmin), so one evaluation costs one memory round trip instead of nine dependent ones -- these kernels
// are latency-bound (the candidate windows of a 128-pair chunk do not fit the L2).
//
// Small blocks (2x2, 4x4: 95 % of all block evaluations) are evaluated by one thread; the only branch is
// warp-uniform (every lane's nine candidates identical -> nothing can change).
template <int BS>  // 2 or 4
__device__ __forceinline__ short2 reg_eval_small(const RegArgs& a, int pair, const short2* O,
                                                 const short2* P, int bx, int by, bool live) {
  const int gw = a.gw, gh = a.gh;
  const int idx = by * gw + bx;
  const bool up = by > 0, dn = by < gh - 1, lf = bx > 0, rt = bx < gw - 1;
  const short2 c0 = O[idx];
  short2 c[9];
  c[0] = c0;
  c[1] = lf ? P[idx - 1] : c0;
  c[2] = rt ? O[idx + 1] : c0;
  c[3] = (dn && rt) ? O[idx + gw + 1] : c0;
  c[4] = (up && lf) ? P[idx - gw - 1] : c0;
  c[5] = (up && rt) ? P[idx - gw + 1] : c0;
  c[6] = up ? P[idx - gw] : c0;
  c[7] = dn ? O[idx + gw] : c0;
  c[8] = (dn && lf) ? O[idx + gw - 1] : c0;
  const uint32_t k0 = pack_mv(c0);
  bool all_same = true;
#pragma unroll
  for (int i = 1; i < 9; ++i) all_same = all_same && (pack_mv(c[i]) == k0);
  // warp-uniform early out: identical candidates have identical energies, index 0 wins (:653-659)
  if (__all_sync(__activemask(), all_same || !live)) return c0;

  const uint32_t vmask = 1u | (lf ? 2u : 0u) | (rt ? 4u : 0u) | ((dn && rt) ? 8u : 0u) | ((up && lf) ? 16u : 0u) |
                         ((up && rt) ? 32u : 0u) | (up ? 64u : 0u) | (dn ? 128u : 0u) | ((dn && lf) ? 256u : 0u);
  const float n_missing = (float)(9 - __popc(vmask));
  int cx[9], cy[9];
  float fx[9], fy[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    cx[i] = c[i].x; cy[i] = c[i].y;
    fx[i] = (float)cx[i]; fy[i] = (float)cy[i];
  }
  // S_i over the gathered candidates (:637-641), accumulated in float like the reference (integer-valued, exact):
  // |a - b| + acc is FADD + FADD-with-|.|-modifier on the FMA pipe.  Sum over all nine slots, then remove the
  // missing slots' share (each holds C, i.e. contributes d(i, 0)).
  float S[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) S[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
#pragma unroll
    for (int k = i + 1; k < 9; ++k) {
      const float d = __fadd_rn(fabsf(__fsub_rn(fx[i], fx[k])), fabsf(__fsub_rn(fy[i], fy[k])));
      S[i] = __fadd_rn(S[i], d);
      S[k] = __fadd_rn(S[k], d);
    }
  }
  {
    // d(i, 0) for the correction; S[0] needs none (d(0,0) = 0)
#pragma unroll
    for (int i = 1; i < 9; ++i) {
      const float d0 = __fadd_rn(fabsf(__fsub_rn(fx[i], fx[0])), fabsf(__fsub_rn(fy[i], fy[0])));
      S[i] = __fmaf_rn(-n_missing, d0, S[i]);
    }
  }

  const int x = bx * BS, y = by * BS;
  const int w = a.i1.w, h = a.i1.h, pitch = a.i1.pitch;
  const uint8_t* blk = a.i1.p + (size_t)pair * a.i1.plane + (size_t)y * pitch + x;
  const uint8_t* ref = a.i2.p + (size_t)pair * a.i2.plane;
  uint32_t A[BS == 2 ? 1 : 4];
  if (BS == 2) {
    A[0] = (uint32_t)*reinterpret_cast<const uint16_t*>(blk) | ((uint32_t)*reinterpret_cast<const uint16_t*>(blk + pitch) << 16);
  } else {
#pragma unroll
    for (int r = 0; r < (BS == 2 ? 1 : 4); ++r) A[r] = *reinterpret_cast<const uint32_t*>(blk + (size_t)r * pitch);
  }
  // A listed block has 2-4 DISTINCT vectors among its nine candidates (2.3 on average: it sits on the border between
  // two or three motion layers), and the SAD depends on the vector only: the window of slot i is loaded only if no
  // earlier slot holds the same vector ("need"), the others copy the SAD.  All needed windows are still addressed first
  // and loaded together (predicated loads), so an evaluation stays one memory round trip.
  uint32_t pk[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) pk[i] = pack_mv(c[i]);
  bool inb[9], need[9];
  const uint8_t* bp[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const int px = x + cx[i], py = y + cy[i];
    inb[i] = (unsigned)px <= (unsigned)(w - BS) && (unsigned)py <= (unsigned)(h - BS);  // :578
    bool dup = false;
#pragma unroll
    for (int j = 0; j < i; ++j) dup = dup || pk[j] == pk[i];
    need[i] = inb[i] && !dup;
    bp[i] = ref + (size_t)(need[i] ? py : y) * pitch + (need[i] ? px : x);
  }
  uint32_t sadv[9];
  if (BS == 2) {
    // a 2x2 window row is two bytes at any alignment: the aligned word that holds the first byte, plus the next word
    // only when the row starts at byte 3
    uint32_t w0[9][2], w1[9][2], sh[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const uintptr_t ab = reinterpret_cast<uintptr_t>(bp[i]);
      sh[i] = (uint32_t)(ab & 3);
      const uint32_t* q = reinterpret_cast<const uint32_t*>(ab & ~(uintptr_t)3);
      const bool two = need[i] && sh[i] == 3u;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const uint32_t* qr = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(q) + (size_t)r * pitch);
        w0[i][r] = need[i] ? qr[0] : 0u;
        w1[i][r] = two ? qr[1] : 0u;
      }
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const uint32_t r0 = __funnelshift_r(w0[i][0], w1[i][0], sh[i] * 8u) & 0xffffu;
      const uint32_t r1 = __funnelshift_r(w0[i][1], w1[i][1], sh[i] * 8u) & 0xffffu;
      sadv[i] = sad4(A[0], r0 | (r1 << 16), 0u);
    }
  } else {
    uint32_t w0[9][4], w1[9][4], sh[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const uintptr_t ab = reinterpret_cast<uintptr_t>(bp[i]);
      sh[i] = (uint32_t)(ab & 3);
      const uint32_t* q = reinterpret_cast<const uint32_t*>(ab & ~(uintptr_t)3);
      const bool two = need[i] && sh[i] != 0u;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint32_t* qr = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(q) + (size_t)r * pitch);
        w0[i][r] = need[i] ? qr[0] : 0u;
        w1[i][r] = two ? qr[1] : 0u;
      }
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      uint32_t sad = 0;
#pragma unroll
      for (int r = 0; r < 4; ++r) sad = sad4(A[r % (BS == 2 ? 1 : 4)], __funnelshift_r(w0[i][r], w1[i][r], sh[i] * 8u), sad);
      sadv[i] = sad;
    }
  }
  // duplicates copy the SAD of an earlier slot with the same vector (every earlier copy already holds it)
#pragma unroll
  for (int i = 1; i < 9; ++i) {
#pragma unroll
    for (int j = 0; j < i; ++j) sadv[i] = (pk[j] == pk[i]) ? sadv[j] : sadv[i];
  }
  float best = 0.f;
  int best_i = 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const float e = inb[i] ? __fadd_rn(__uint2float_rn(sadv[i]), __fmul_rn(a.lm, S[i])) : FLT_MAX;
    if (i == 0) {
      best = e;
    } else {
      const bool take = ((vmask >> i) & 1u) && (e < best);
      best = take ? e : best;
      best_i = take ? i : best_i;
    }
  }
  short2 r = c0;
#pragma unroll
  for (int i = 1; i < 9; ++i) r = (best_i == i) ? c[i] : r;
  return r;
}

// Blocks of 8x8 and larger are evaluated by a TEAM of adjacent lanes (8 for 8x8, 16 for 16x16, 32 above): a lane
// takes one block row (32x32 and larger: bs/32 rows, 16 bytes at a time), loads its slice of all nine candidate
// windows at once; the partial SADs are reduce-scattered inside the team so that lane i holds candidate i's SAD, computes
// that candidate's smoothness and energy, and a shuffle argmin leaves every lane of the team with the same winner.
template <int TEAM>
__device__ __forceinline__ short2 reg_eval_team(const RegArgs& a, int pair, const short2* O,
                                                const short2* P, int bx, int by, int tl, uint32_t team_mask) {
  const int gw = a.gw, gh = a.gh, bs = a.bs;
  const int idx = by * gw + bx;
  const bool up = by > 0, dn = by < gh - 1, lf = bx > 0, rt = bx < gw - 1;
  const short2 c0 = O[idx];
  short2 c[9];
  c[0] = c0;
  c[1] = lf ? P[idx - 1] : c0;
  c[2] = rt ? O[idx + 1] : c0;
  c[3] = (dn && rt) ? O[idx + gw + 1] : c0;
  c[4] = (up && lf) ? P[idx - gw - 1] : c0;
  c[5] = (up && rt) ? P[idx - gw + 1] : c0;
  c[6] = up ? P[idx - gw] : c0;
  c[7] = dn ? O[idx + gw] : c0;
  c[8] = (dn && lf) ? O[idx + gw - 1] : c0;
  const uint32_t vmask = 1u | (lf ? 2u : 0u) | (rt ? 4u : 0u) | ((dn && rt) ? 8u : 0u) | ((up && lf) ? 16u : 0u) |
                         ((up && rt) ? 32u : 0u) | (up ? 64u : 0u) | (dn ? 128u : 0u) | ((dn && lf) ? 256u : 0u);

  const int x = bx * bs, y = by * bs;
  const int w = a.i1.w, h = a.i1.h, pitch = a.i1.pitch;
  const uint8_t* blk = a.i1.p + (size_t)pair * a.i1.plane + (size_t)y * pitch + x;
  const uint8_t* ref = a.i2.p + (size_t)pair * a.i2.plane;
  // candidate windows: out-of-image ones (:578-582) read the block's own position and get FLT_MAX below
  // Only the first slot of every distinct vector loads its window ("need", team-uniform; a listed block has 2.3
  // distinct vectors among its nine candidates on average); the other slots copy the partial sums below.
  uint32_t inb = 0;
  const uint8_t* bp[9];
  uint32_t pk[9];
  bool need[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    pk[i] = pack_mv(c[i]);
    const int px = x + c[i].x, py = y + c[i].y;
    const bool ok = (unsigned)px <= (unsigned)(w - bs) && (unsigned)py <= (unsigned)(h - bs);
    inb |= ok ? (1u << i) : 0u;
    bool dup = false;
#pragma unroll
    for (int j = 0; j < i; ++j) dup = dup || pk[j] == pk[i];
    need[i] = ok && !dup;
    bp[i] = ref + (size_t)(need[i] ? py : y) * pitch + (need[i] ? px : x);
  }
  // Partial SADs of this lane's rows.  A window row starts at any byte: it is fetched as the two aligned 16-byte (8x8
  // blocks: 8-byte) vectors that contain it -- two requests per row instead of five (three) 32-bit ones; a team's lanes
  // read 16 different rows, i.e. 16 cache lines per request, and the L1's line rate, not its bandwidth, bounded these
  // kernels -- and the wanted words are selected by the start offset's word index before the byte shift.
  uint32_t v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0u;
  if (TEAM == 8) {
    const size_t ro = (size_t)tl * pitch;
    const uint2 A = __ldg(reinterpret_cast<const uint2*>(blk + ro));
    uint2 q0[9], q1[9];
    uint32_t off[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const uintptr_t ab = reinterpret_cast<uintptr_t>(bp[i] + ro);
      const uint2* q = reinterpret_cast<const uint2*>(ab & ~(uintptr_t)7);
      off[i] = (uint32_t)(ab & 7);
      q0[i] = need[i] ? __ldg(q) : make_uint2(0u, 0u);
      q1[i] = need[i] ? __ldg(q + 1) : make_uint2(0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const bool w1 = (off[i] & 4u) != 0u;
      const uint32_t sh = (off[i] & 3u) * 8u;
      const uint32_t a0 = w1 ? q0[i].y : q0[i].x, a1 = w1 ? q1[i].x : q0[i].y, a2 = w1 ? q1[i].y : q1[i].x;
      v[i] = sad4(A.y, __funnelshift_r(a1, a2, sh), sad4(A.x, __funnelshift_r(a0, a1, sh), 0u));
    }
  } else {
    // 16 bytes of one row per step; TEAM == 16: one step, TEAM == 32: (bs / 32) rows x (bs / 16) column chunks
    const int rows = TEAM == 16 ? 1 : bs / 32;
    const int chunks = TEAM == 16 ? 1 : bs / 16;
    for (int rr = 0; rr < rows; ++rr) {
      for (int ch = 0; ch < chunks; ++ch) {
        const size_t ro = (size_t)(rr * TEAM + tl) * pitch + ch * 16;
        const uint4 A = __ldg(reinterpret_cast<const uint4*>(blk + ro));
        uint4 q0[9], q1[9];
        uint32_t off[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const uintptr_t ab = reinterpret_cast<uintptr_t>(bp[i] + ro);
          const uint4* q = reinterpret_cast<const uint4*>(ab & ~(uintptr_t)15);
          off[i] = (uint32_t)(ab & 15);
          q0[i] = need[i] ? __ldg(q) : make_uint4(0u, 0u, 0u, 0u);
          q1[i] = need[i] ? __ldg(q + 1) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const bool s2 = (off[i] & 8u) != 0u, s1 = (off[i] & 4u) != 0u;
          const uint32_t sh = (off[i] & 3u) * 8u;
          // words 0..7 of the 32 aligned bytes; skip two words, then one
          const uint32_t t0 = s2 ? q0[i].z : q0[i].x, t1 = s2 ? q0[i].w : q0[i].y, t2 = s2 ? q1[i].x : q0[i].z,
                         t3 = s2 ? q1[i].y : q0[i].w, t4 = s2 ? q1[i].z : q1[i].x, t5 = s2 ? q1[i].w : q1[i].y;
          const uint32_t a0 = s1 ? t1 : t0, a1 = s1 ? t2 : t1, a2 = s1 ? t3 : t2, a3 = s1 ? t4 : t3, a4 = s1 ? t5 : t4;
          uint32_t sum = v[i];
          sum = sad4(A.x, __funnelshift_r(a0, a1, sh), sum);
          sum = sad4(A.y, __funnelshift_r(a1, a2, sh), sum);
          sum = sad4(A.z, __funnelshift_r(a2, a3, sh), sum);
          sum = sad4(A.w, __funnelshift_r(a3, a4, sh), sum);
          v[i] = sum;
        }
      }
    }
  }

  // slots that share a vector share the partial sums
#pragma unroll
  for (int i = 1; i < 9; ++i) {
#pragma unroll
    for (int j = 0; j < i; ++j) v[i] = (pk[j] == pk[i]) ? v[j] : v[i];
  }
  // Reduce-scatter inside the team: the nine sums live in 16 slots; at each step a lane keeps one half of its slots and
  // hands the other half to its partner, so that lane i (8x8 teams: lane i / 2) ends with the team total of candidate i
  // -- 15 (14) shuffles instead of 9 * log2(TEAM).  32-lane teams first fold their two halves together.
  if (TEAM == 32) {
#pragma unroll
    for (int i = 0; i < 9; ++i) v[i] += __shfl_xor_sync(team_mask, v[i], 16);
  }
  constexpr int W = TEAM >= 16 ? 16 : 8;  // lanes the slots are scattered over
  constexpr int NS = 16 / W;              // slots a lane ends up with
#pragma unroll
  for (int off = W / 2, n = 8; off >= 1; off >>= 1, n >>= 1) {
    const bool upper = (tl & off) != 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < n) {
        const uint32_t lo = v[j], hi = v[j + n];
        const uint32_t recv = __shfl_xor_sync(team_mask, upper ? lo : hi, off);
        v[j] = (upper ? hi : lo) + recv;
      }
    }
  }
  const int cid0 = W == 16 ? (tl & 15) : 2 * (tl & 7);  // candidate of this lane's slot 0 (the reduce-scatter's bit order)

  // Each lane finishes ITS candidate(s): smoothness S_i = sum over the gathered candidates k of |c_k.x - c_i.x| +
  // |c_k.y - c_i.y| (:637-641) in float like the reference (integer-valued, < 2^24, exact; FADD with |.| modifiers on the
  // FMA pipe); all nine slots are summed and the missing slots' share removed (each holds a copy of C, i.e. d(i, 0)).
  // Before, every lane of a team repeated all 36 pair distances.
  float fx[9], fy[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    fx[i] = (float)c[i].x;
    fy[i] = (float)c[i].y;
  }
  const float n_missing = (float)(9 - __popc(vmask));
  float best_e = FLT_MAX;
  int best_i = 15;
#pragma unroll
  for (int sidx = 0; sidx < NS; ++sidx) {
    const int i = cid0 + sidx;
    float mx = fx[0], my = fy[0];
#pragma unroll
    for (int q = 1; q < 9; ++q) {
      mx = (i == q) ? fx[q] : mx;
      my = (i == q) ? fy[q] : my;
    }
    float S = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) S = __fadd_rn(S, __fadd_rn(fabsf(__fsub_rn(mx, fx[k])), fabsf(__fsub_rn(my, fy[k]))));
    const float d0 = __fadd_rn(fabsf(__fsub_rn(mx, fx[0])), fabsf(__fsub_rn(my, fy[0])));
    S = __fmaf_rn(-n_missing, d0, S);
    const bool valid = i < 9 && ((vmask >> i) & 1u);
    const bool in_image = ((inb >> i) & 1u) != 0u;
    // (:607) un-fused, (:578-582) FLT_MAX outside the image; slots without a neighbour can never win
    const float e = (valid && in_image) ? __fadd_rn(__uint2float_rn(v[sidx]), __fmul_rn(a.lm, S)) : FLT_MAX;
    const int ii = valid ? i : 15;
    const bool take = e < best_e || (e == best_e && ii < best_i);
    best_e = take ? e : best_e;
    best_i = take ? ii : best_i;
  }
  // argmin over the team: smallest energy, ties to the smallest index == the reference's scan with strict '<' from
  // index 1 (:653-659); index 0 (C) is always present, so an all-FLT_MAX block keeps its vector
#pragma unroll
  for (int off = W / 2; off >= 1; off >>= 1) {
    const float oe = __shfl_xor_sync(team_mask, best_e, off);
    const int oi = __shfl_xor_sync(team_mask, best_i, off);
    const bool take = oe < best_e || (oe == best_e && oi < best_i);
    best_e = take ? oe : best_e;
    best_i = take ? oi : best_i;
  }
  uint32_t r = pk[0];
#pragma unroll
  for (int i = 1; i < 9; ++i) r = (best_i == i) ? pk[i] : r;
  return make_short2((short)(r & 0xffffu), (short)(r >> 16));
}

// TEAM == 1 evaluates 2x2 blocks, TEAM == 2 is the tag for "one thread per 4x4 block" (TEAMSZ below is 1 for both)
template <int TEAM>
__device__ __forceinline__ short2 reg_eval_any(const RegArgs& a, int pair, const short2* O, const short2* P,
                                               int bx, int by, int tl, uint32_t team_mask, bool live) {
  if (TEAM == 1) {
    return reg_eval_small<2>(a, pair, O, P, bx, by, live);
  } else if (TEAM == 2) {
    return reg_eval_small<4>(a, pair, O, P, bx, by, live);
  } else {
    return reg_eval_team<TEAM>(a, pair, O, P, bx, by, tl, team_mask);
  }
}

// A block that changed invalidates the evaluations of the blocks that read it as a "pred" neighbour: its right,
// lower-left, lower and lower-right neighbours.  They are appended to the next round's work list, de-duplicated by an
// epoch stamp per block.  Called by all 32 lanes of a warp together: the four stamp exchanges of a lane are issued
// back to back (independent atomics, one round trip), and the list slots of the whole warp are reserved with ONE
// atomicAdd (a per-entry atomicAdd on the pair's counter serialises in the L2).
__device__ __forceinline__ void push_dependents(bool changed, int bx, int by, int gw, int gh, uint32_t* stamp,
                                                uint32_t ep, uint32_t* list, uint32_t* count) {
  const int lane = threadIdx.x & 31;
  int d[4];
  d[0] = by * gw + bx + 1;
  d[1] = (by + 1) * gw + bx - 1;
  d[2] = (by + 1) * gw + bx;
  d[3] = (by + 1) * gw + bx + 1;
  const bool rt = bx + 1 < gw, dn = by + 1 < gh;
  const bool ex[4] = {changed && rt, changed && dn && bx > 0, changed && dn, changed && dn && rt};
  uint32_t old[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) old[j] = ex[j] ? atomicExch(&stamp[d[j]], ep) : ep;
  int k = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) k += (old[j] != ep) ? 1 : 0;
  const uint32_t any = __ballot_sync(0xffffffffu, k > 0);
  if (any == 0u) return;
  int incl = k;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  uint32_t base = 0;
  if (lane == 31) base = atomicAdd(count, (uint32_t)incl);
  base = __shfl_sync(0xffffffffu, base, 31);
  uint32_t pos = base + (uint32_t)(incl - k);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (old[j] != ep) list[pos++] = (uint32_t)d[j];
}

// Pass 1 of a sweep (a Jacobi step: every block evaluated with the OLD field in all nine slots), in two kernels:
//  k_reg_classify  copies O to Y and lists the blocks whose nine candidates are not all identical.  A block whose
//                  candidates are identical keeps its vector (all energies equal, index 0 wins, :653-659); on real
//                  fields 80-96 % of the 2x2 / 4x4 blocks are of that kind.  One thread handles four horizontally
//                  adjacent blocks: three 128-bit row loads + six halo entries instead of 36 scalar loads; the list
//                  slots of a warp are reserved with one atomicAdd, so neighbours stay neighbours in the list.
//  k_reg_eval      evaluation of the listed blocks by a fixed-size grid that strides over the list; blocks whose
//                  value changed enqueue their dependents (they may have used a stale "pred" value) for the rounds.
__global__ void __launch_bounds__(256) k_reg_classify4(RegArgs a) {
  // block = 64 x 4 threads: 64 four-block groups along a row, four rows (no integer division per thread)
  const int tx = blockIdx.x * 64 + threadIdx.x, by = blockIdx.y * 4 + threadIdx.y;
  const int pair = blockIdx.z;
  const int gw = a.gw, gh = a.gh, gw4 = gw >> 2;
  const int lane = threadIdx.x & 31;
  const bool live = tx < gw4 && by < gh;
  const uint32_t* __restrict__ O = reinterpret_cast<const uint32_t*>(a.O + (size_t)pair * a.mv_plane);
  uint32_t work = 0;
  int i0 = 0;
  if (live) {
    const int bx = tx * 4;
    i0 = by * gw + bx;
    // clamped coordinates: a clamped neighbour is the block itself or another neighbour, so the test is unchanged
    const int ru = max(by - 1, 0) * gw, rm = by * gw, rd = min(by + 1, gh - 1) * gw;
    const int cl = max(bx - 1, 0), cr = min(bx + 4, gw - 1);
    const uint4 U = *reinterpret_cast<const uint4*>(O + ru + bx);
    const uint4 M = *reinterpret_cast<const uint4*>(O + rm + bx);
    const uint4 D = *reinterpret_cast<const uint4*>(O + rd + bx);
    const uint32_t u[6] = {O[ru + cl], U.x, U.y, U.z, U.w, O[ru + cr]};
    const uint32_t m[6] = {O[rm + cl], M.x, M.y, M.z, M.w, O[rm + cr]};
    const uint32_t d[6] = {O[rd + cl], D.x, D.y, D.z, D.w, O[rd + cr]};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t k0 = m[j + 1];
      const bool same = u[j] == k0 && u[j + 1] == k0 && u[j + 2] == k0 && m[j] == k0 && m[j + 2] == k0 && d[j] == k0 &&
                        d[j + 1] == k0 && d[j + 2] == k0;
      work |= same ? 0u : (1u << j);
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(a.Y + (size_t)pair * a.mv_plane) + i0) = M;
  }
  const int k = __popc(work);
  if (__ballot_sync(0xffffffffu, k > 0) == 0u) return;
  int incl = k;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  uint32_t base = 0;
  if (lane == 31) base = atomicAdd(&a.ctr[(size_t)pair * kCtrWords + CTR_COUNT_EVAL], (uint32_t)incl);
  base = __shfl_sync(0xffffffffu, base, 31);
  uint32_t* list = a.list1 + (size_t)pair * a.wl_plane + base + (uint32_t)(incl - k);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if ((work >> j) & 1u) *list++ = (uint32_t)(i0 + j);
}

// grids whose width is not a multiple of four (only the coarsest stages of small images): one thread per block
__global__ void __launch_bounds__(256) k_reg_classify(RegArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int pair = blockIdx.y;
  const int gw = a.gw, gh = a.gh;
  const bool live = i < gw * gh;
  const uint32_t* __restrict__ O = reinterpret_cast<const uint32_t*>(a.O + (size_t)pair * a.mv_plane);
  bool work = false;
  if (live) {
    const int bx = i % gw, by = i / gw;
    const int ru = max(by - 1, 0) * gw, rm = by * gw, rd = min(by + 1, gh - 1) * gw;
    const int cl = max(bx - 1, 0), cr = min(bx + 1, gw - 1);
    const uint32_t k0 = O[i];
    const uint32_t v[8] = {O[ru + cl], O[ru + bx], O[ru + cr], O[rm + cl], O[rm + cr], O[rd + cl], O[rd + bx], O[rd + cr]};
    bool same = true;
#pragma unroll
    for (int j = 0; j < 8; ++j) same = same && v[j] == k0;
    reinterpret_cast<uint32_t*>(a.Y + (size_t)pair * a.mv_plane)[i] = k0;
    work = !same;
  }
  const uint32_t m = __ballot_sync(0xffffffffu, work);
  if (m) {
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(&a.ctr[(size_t)pair * kCtrWords + CTR_COUNT_EVAL], (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (work) a.list1[(size_t)pair * a.wl_plane + base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)i;
  }
}

template <int TEAM>
__global__ void __launch_bounds__(128, TEAM <= 2 ? 4 : 3) k_reg_eval(RegArgs a) {
  constexpr int TEAMSZ = TEAM <= 2 ? 1 : TEAM;
  constexpr int TPW = 32 / TEAMSZ;  // teams per warp
  const int pair = blockIdx.y;
  uint32_t* ctr = a.ctr + (size_t)pair * kCtrWords;
  const uint32_t cnt = ctr[CTR_COUNT_EVAL];
  if (cnt == 0) return;
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t team = gtid / TEAMSZ, nteams = gridDim.x * blockDim.x / TEAMSZ;
  const int tl = (int)(gtid % TEAMSZ);
  const uint32_t team_mask = TEAMSZ >= 32 ? 0xffffffffu : (((1u << TEAMSZ) - 1u) << ((threadIdx.x & 31) / TEAMSZ * TEAMSZ));
  const short2* O = a.O + (size_t)pair * a.mv_plane;
  short2* Y = a.Y + (size_t)pair * a.mv_plane;
  const uint32_t* lc = a.list1 + (size_t)pair * a.wl_plane;
  uint32_t* stamp = a.stamp + (size_t)pair * a.wl_plane;
  uint32_t* ln = a.list0 + (size_t)pair * a.wl_plane;
  const uint32_t ep = ctr[CTR_EPOCH] + 1u;
  const uint32_t limit = (cnt + TPW - 1) / TPW * TPW;  // whole warps iterate together
  for (uint32_t e = team; e < limit; e += nteams) {
    const bool live = e < cnt;
    const int i = (int)lc[live ? e : cnt - 1];
    const int bx = i % a.gw, by = i / a.gw;
    const short2 nv = reg_eval_any<TEAM>(a, pair, O, O, bx, by, tl, team_mask, live);
    const bool lead = tl == 0 && live;
    const bool changed = lead && pack_mv(nv) != pack_mv(O[i]);
    if (changed) Y[i] = nv;  // k_reg_classify copied O to Y
    push_dependents(changed, bx, by, a.gw, a.gh, stamp, ep, ln, &ctr[CTR_COUNT0]);
  }
}

// Passes 2..: rounds on the active set until nothing changes.  The update map is triangular in raster order (a
// block depends on earlier blocks' NEW values and later blocks' OLD values only), so the fixed point is unique and
// equals the reference's in-place raster sweep.  Rounds update Y in place ("chaotic" iteration): an evaluation may
// read a neighbour before or after that neighbour's update of the same round; whenever a block changes, its
// dependents are (re-)enqueued for the next round, so a block that read a stale value is always evaluated again
// after the next barrier / kernel boundary.  Only the path to the fixed point varies, not the result.
//
// Round r reads list[r & 1] (count in counter r % 3), appends to list[(r + 1) & 1] (counter (r + 1) % 3, stamp
// epoch + 2 + r) and clears counter (r + 2) % 3 for the round after.  With few pairs in flight the first rounds (the
// big ones) run as grid-wide kernels over all pairs (k_reg_round); the tail, where rounds are short and
// latency-bound, runs as one CTA per pair that loops until its list is empty (k_reg_fix).
__device__ __forceinline__ int ctr_index(int k) { return k == 2 ? CTR_COUNT2 : k; }

template <int TEAM>
__global__ void __launch_bounds__(256, TEAM <= 2 ? 2 : 1) k_reg_round(RegArgs a, int r) {
  constexpr int TEAMSZ = TEAM <= 2 ? 1 : TEAM;
  constexpr int TPW = 32 / TEAMSZ;
  const int pair = blockIdx.y;
  uint32_t* ctr = a.ctr + (size_t)pair * kCtrWords;
  const uint32_t cnt = ctr[ctr_index(r % 3)];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ctr[ctr_index((r + 2) % 3)] = 0;
    if (cnt) { ctr[CTR_ROUNDS] += 1; ctr[CTR_BLOCKS] += cnt; }
  }
  if (cnt == 0) return;
  const short2* O = a.O + (size_t)pair * a.mv_plane;
  short2* Y = a.Y + (size_t)pair * a.mv_plane;
  uint32_t* stamp = a.stamp + (size_t)pair * a.wl_plane;
  const uint32_t* lc = ((r & 1) ? a.list1 : a.list0) + (size_t)pair * a.wl_plane;
  uint32_t* ln = ((r & 1) ? a.list0 : a.list1) + (size_t)pair * a.wl_plane;
  uint32_t* next_count = &ctr[ctr_index((r + 1) % 3)];
  const uint32_t ep = ctr[CTR_EPOCH] + 2u + (uint32_t)r;
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t team = gtid / TEAMSZ, tl = gtid % TEAMSZ, nteams = gridDim.x * blockDim.x / TEAMSZ;
  const uint32_t team_mask = TEAMSZ >= 32 ? 0xffffffffu : (((1u << TEAMSZ) - 1u) << ((threadIdx.x & 31) / TEAMSZ * TEAMSZ));
  const uint32_t limit = (cnt + TPW - 1) / TPW * TPW;
  for (uint32_t e = team; e < limit; e += nteams) {
    const bool live = e < cnt;
    const int b = (int)lc[live ? e : cnt - 1];
    const int bx = b % a.gw, by = b / a.gw;
    const short2 nv = reg_eval_any<TEAM>(a, pair, O, Y, bx, by, (int)tl, team_mask, live);
    const bool changed = tl == 0 && live && pack_mv(nv) != pack_mv(Y[b]);
    if (changed) Y[b] = nv;
    push_dependents(changed, bx, by, a.gw, a.gh, stamp, ep, ln, next_count);
  }
}

template <int TEAM>
__global__ void __launch_bounds__(512) k_reg_fix(RegArgs a, int r0) {
  constexpr int TEAMSZ = TEAM <= 2 ? 1 : TEAM;
  constexpr int TPW = 32 / TEAMSZ;
  const int pair = blockIdx.x;
  const short2* O = a.O + (size_t)pair * a.mv_plane;
  short2* Y = a.Y + (size_t)pair * a.mv_plane;
  uint32_t* ctr = a.ctr + (size_t)pair * kCtrWords;
  uint32_t* stamp = a.stamp + (size_t)pair * a.wl_plane;
  uint32_t* lists[2] = {a.list0 + (size_t)pair * a.wl_plane, a.list1 + (size_t)pair * a.wl_plane};
  __shared__ uint32_t s_next[2];
  uint32_t cnt = ctr[ctr_index(r0 % 3)];
  uint32_t ep = ctr[CTR_EPOCH] + 1u + (uint32_t)r0;  // appends of round r use ep + 1 = epoch + 2 + r
  const uint32_t team = threadIdx.x / TEAMSZ, tl = threadIdx.x % TEAMSZ, nteams = blockDim.x / TEAMSZ;
  const uint32_t team_mask = TEAMSZ >= 32 ? 0xffffffffu : (((1u << TEAMSZ) - 1u) << ((threadIdx.x & 31) / TEAMSZ * TEAMSZ));
  uint32_t rounds = 0, blocks = 0;
  int cur = r0 & 1;
  if (threadIdx.x == 0) { s_next[0] = 0; s_next[1] = 0; }
  __syncthreads();
  while (cnt > 0) {
    const uint32_t* lc = lists[cur];
    uint32_t* ln = lists[cur ^ 1];
    uint32_t* next_count = &s_next[cur ^ 1];
    const uint32_t limit = (cnt + TPW - 1) / TPW * TPW;  // whole warps iterate together
    for (uint32_t e = team; e < limit; e += nteams) {
      const bool live = e < cnt;
      const int b = (int)lc[live ? e : cnt - 1];
      const int bx = b % a.gw, by = b / a.gw;
      const short2 nv = reg_eval_any<TEAM>(a, pair, O, Y, bx, by, (int)tl, team_mask, live);
      const bool changed = tl == 0 && live && pack_mv(nv) != pack_mv(Y[b]);
      if (changed) Y[b] = nv;
      push_dependents(changed, bx, by, a.gw, a.gh, stamp, ep + 1u, ln, next_count);
    }
    __syncthreads();
    if (a.hist && threadIdx.x == 0) atomicAdd(&a.hist[2 + min(rounds + (uint32_t)r0, 59u)], cnt);
    blocks += cnt;
    cnt = *next_count;
    if (threadIdx.x == 0) s_next[cur] = 0;  // becomes the append counter of the round after next
    cur ^= 1;
    ++ep;
    ++rounds;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (a.hist) {
      atomicAdd(&a.hist[0], ctr[CTR_COUNT_EVAL]);
      atomicMax(&a.hist[62], rounds + (uint32_t)r0);
      atomicAdd(&a.hist[63], rounds);
    }
    ctr[CTR_COUNT0] = 0;
    ctr[CTR_COUNT1] = 0;
    ctr[CTR_COUNT2] = 0;
    ctr[CTR_COUNT_EVAL] = 0;
    ctr[CTR_EPOCH] = ep;
    ctr[CTR_ROUNDS] += rounds;
    ctr[CTR_BLOCKS] += blocks;
    ctr[CTR_TAIL_BLOCKS] += blocks;
  }
}

// evaluator variant per block size: 1 = one thread per 2x2 block, 2 = one thread per 4x4 block, else lanes per block
static int team_for(int bs) { return bs >= 32 ? 32 : (bs >= 8 ? bs : (bs == 4 ? 2 : 1)); }

void launch_reg_full(const RegArgs& a, int n, cudaStream_t s) {
  const int team = team_for(a.bs);
  const int lanes = team <= 2 ? 1 : team;
  const size_t nb = (size_t)a.gw * a.gh;
  if ((a.gw & 3) == 0 && (a.mv_plane & 3) == 0 && a.gh <= 4 * 65535)
    k_reg_classify4<<<dim3((unsigned)((a.gw / 4 + 63) / 64), (unsigned)((a.gh + 3) / 4), n), dim3(64, 4), 0, s>>>(a);
  else k_reg_classify<<<dim3((unsigned)((nb + 255) / 256), n), 256, 0, s>>>(a);
  // a fixed-size grid strides over each pair's list (its length is only known on the device): enough CTAs to fill the
  // chip a few times over, never more than the list could need
  size_t want = (nb * lanes + 127) / 128;
  const size_t cap = (size_t)(148 * 16 + n - 1) / n;
  unsigned gx = (unsigned)(want < cap ? want : cap);
  if (gx < 1) gx = 1;
  dim3 grid(gx, n);
  switch (team) {
    case 32: k_reg_eval<32><<<grid, 128, 0, s>>>(a); break;
    case 16: k_reg_eval<16><<<grid, 128, 0, s>>>(a); break;
    case 8: k_reg_eval<8><<<grid, 128, 0, s>>>(a); break;
    case 2: k_reg_eval<2><<<grid, 128, 0, s>>>(a); break;
    default: k_reg_eval<1><<<grid, 128, 0, s>>>(a); break;
  }
}

void launch_reg_round(const RegArgs& a, int r, int n, cudaStream_t s) {
  const int team = team_for(a.bs);
  const int lanes = team <= 2 ? 1 : team;
  // enough CTAs per pair to spread a few-percent active set of this grid over the chip, at most 16
  size_t want = ((size_t)a.gw * a.gh * lanes / 16 + 255) / 256;
  unsigned bx = (unsigned)(want < 1 ? 1 : (want > 16 ? 16 : want));
  dim3 grid(bx, n);
  switch (team) {
    case 32: k_reg_round<32><<<grid, 256, 0, s>>>(a, r); break;
    case 16: k_reg_round<16><<<grid, 256, 0, s>>>(a, r); break;
    case 8: k_reg_round<8><<<grid, 256, 0, s>>>(a, r); break;
    case 2: k_reg_round<2><<<grid, 256, 0, s>>>(a, r); break;
    default: k_reg_round<1><<<grid, 256, 0, s>>>(a, r); break;
  }
}

void launch_reg_fix(const RegArgs& a, int r0, int n, cudaStream_t s) {
  switch (team_for(a.bs)) {
    case 32: k_reg_fix<32><<<n, 512, 0, s>>>(a, r0); break;
    case 16: k_reg_fix<16><<<n, 512, 0, s>>>(a, r0); break;
    case 8: k_reg_fix<8><<<n, 512, 0, s>>>(a, r0); break;
    case 2: k_reg_fix<2><<<n, 512, 0, s>>>(a, r0); break;
    default: k_reg_fix<1><<<n, 512, 0, s>>>(a, r0); break;
  }
}

// ============================================================================================ lean evaluators
// What the evaluators above cost is instructions, not bytes: 1300-1600 per lane and evaluation (nine-slot select chains,
// 36 pair distances, every lane of a team repeating all of it) at 16 warps per SM.  A listed block, however, sits on the
// border between two or three motion layers: its nine candidates hold 2.0-2.6 DISTINCT vectors on average (4 at most
// stages' worst).  The evaluators below work on the distinct vectors.  Same arithmetic per candidate as the reference
// (motion_framework.cpp:578-582,605-607,637-641,653-659), hence the same field; only who computes what changes.

__device__ __forceinline__ int mv_x(uint32_t pk) { return (int)(short)(pk & 0xffffu); }
__device__ __forceinline__ int mv_y(uint32_t pk) { return (int)(short)(pk >> 16); }

// SAD of a 2x2 / 4x4 block of image 1 (rows in A) against the window at `b` (any alignment) of image 2
template <int BS>
__device__ __forceinline__ uint32_t sad_small(const uint32_t* A, const uint8_t* b, int pitch) {
  const uintptr_t ab = reinterpret_cast<uintptr_t>(b);
  const uint32_t sh = (uint32_t)(ab & 3);
  const uint8_t* q = reinterpret_cast<const uint8_t*>(ab & ~(uintptr_t)3);
  if (BS == 2) {
    uint32_t w0[2], w1[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const uint32_t* qr = reinterpret_cast<const uint32_t*>(q + (size_t)r * pitch);
      w0[r] = __ldg(qr);
      w1[r] = sh == 3u ? __ldg(qr + 1) : 0u;
    }
    const uint32_t r0 = __funnelshift_r(w0[0], w1[0], sh * 8u) & 0xffffu;
    const uint32_t r1 = __funnelshift_r(w0[1], w1[1], sh * 8u) & 0xffffu;
    return sad4(A[0], r0 | (r1 << 16), 0u);
  } else {
    uint32_t w0[4], w1[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint32_t* qr = reinterpret_cast<const uint32_t*>(q + (size_t)r * pitch);
      w0[r] = __ldg(qr);
      w1[r] = sh != 0u ? __ldg(qr + 1) : 0u;
    }
    uint32_t sad = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) sad = sad4(A[r % (BS == 2 ? 1 : 4)], __funnelshift_r(w0[r], w1[r], sh * 8u), sad);
    return sad;
  }
}

// The nine candidate vectors of block (bx, by), packed: A0 = the block's own (slot 0), pk[0..7] = slots 1..8
// [L, R, DR, UL, UR, U, D, DL]; a missing neighbour holds A0 (it drops out of every count below).  Split from the evaluation so
// that the caller can issue these loads one iteration ahead (the evaluation's only other memory round trip is the windows).
__device__ __forceinline__ void small_gather(const RegArgs& a, const short2* O, const short2* P, int bx, int by, uint32_t& A0,
                                             uint32_t (&pk)[8]) {
  const int gw = a.gw, gh = a.gh;
  const int idx = by * gw + bx;
  const bool up = by > 0, dn = by < gh - 1, lf = bx > 0, rt = bx < gw - 1;
  const uint32_t* Ou = reinterpret_cast<const uint32_t*>(O);
  const uint32_t* Pu = reinterpret_cast<const uint32_t*>(P);
  A0 = Ou[idx];
  pk[0] = Pu[lf ? idx - 1 : idx];
  pk[1] = Ou[rt ? idx + 1 : idx];
  pk[2] = Ou[(dn && rt) ? idx + gw + 1 : idx];
  pk[3] = Pu[(up && lf) ? idx - gw - 1 : idx];
  pk[4] = Pu[(up && rt) ? idx - gw + 1 : idx];
  pk[5] = Pu[up ? idx - gw : idx];
  pk[6] = Ou[dn ? idx + gw : idx];
  pk[7] = Ou[(dn && lf) ? idx + gw - 1 : idx];
}

// SAD of one row of W bytes (8 or 16) of the block (a: aligned) against the window row at b (any alignment), added to acc.  The
// window row is fetched as the two aligned vectors that contain it; the wanted words are selected by the start offset.
template <int W>
__device__ __forceinline__ uint32_t row_sad(const uint8_t* a, const uint8_t* b, uint32_t acc) {
  const uintptr_t ab = reinterpret_cast<uintptr_t>(b);
  if (W == 8) {
    const uint2 A = __ldg(reinterpret_cast<const uint2*>(a));
    const uint2* q = reinterpret_cast<const uint2*>(ab & ~(uintptr_t)7);
    const uint32_t off = (uint32_t)(ab & 7);
    const uint2 q0 = __ldg(q), q1 = __ldg(q + 1);
    const bool w1 = (off & 4u) != 0u;
    const uint32_t sh = (off & 3u) * 8u;
    const uint32_t a0 = w1 ? q0.y : q0.x, a1 = w1 ? q1.x : q0.y, a2 = w1 ? q1.y : q1.x;
    return sad4(A.y, __funnelshift_r(a1, a2, sh), sad4(A.x, __funnelshift_r(a0, a1, sh), acc));
  } else {
    const uint4 A = __ldg(reinterpret_cast<const uint4*>(a));
    const uint4* q = reinterpret_cast<const uint4*>(ab & ~(uintptr_t)15);
    const uint32_t off = (uint32_t)(ab & 15);
    const uint4 q0 = __ldg(q), q1 = __ldg(q + 1);
    const bool s2 = (off & 8u) != 0u, s1 = (off & 4u) != 0u;
    const uint32_t sh = (off & 3u) * 8u;
    const uint32_t t0 = s2 ? q0.z : q0.x, t1 = s2 ? q0.w : q0.y, t2 = s2 ? q1.x : q0.z, t3 = s2 ? q1.y : q0.w,
                   t4 = s2 ? q1.z : q1.x, t5 = s2 ? q1.w : q1.y;
    const uint32_t a0 = s1 ? t1 : t0, a1 = s1 ? t2 : t1, a2 = s1 ? t3 : t2, a3 = s1 ? t4 : t3, a4 = s1 ? t5 : t4;
    acc = sad4(A.x, __funnelshift_r(a0, a1, sh), acc);
    acc = sad4(A.y, __funnelshift_r(a1, a2, sh), acc);
    acc = sad4(A.z, __funnelshift_r(a2, a3, sh), acc);
    return sad4(A.w, __funnelshift_r(a3, a4, sh), acc);
  }
}

// SAD of the whole block at blk (image 1) against the window at win (image 2), one thread
template <int BSK>
__device__ __forceinline__ uint32_t block_sad_thread(const uint8_t* blk, const uint8_t* win, int pitch, int bs) {
  if (BSK <= 4) {
    uint32_t Ab[BSK == 2 ? 1 : 4];
    if (BSK == 2) {
      Ab[0] = (uint32_t)__ldg(reinterpret_cast<const uint16_t*>(blk)) | ((uint32_t)__ldg(reinterpret_cast<const uint16_t*>(blk + pitch)) << 16);
    } else {
#pragma unroll
      for (int r = 0; r < (BSK == 2 ? 1 : 4); ++r) Ab[r] = __ldg(reinterpret_cast<const uint32_t*>(blk + (size_t)r * pitch));
    }
    return sad_small<BSK <= 2 ? 2 : 4>(Ab, win, pitch);
  } else if (BSK == 8) {
    uint32_t acc = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) acc = row_sad<8>(blk + (size_t)r * pitch, win + (size_t)r * pitch, acc);
    return acc;
  } else {
    uint32_t acc = 0;
    const int chunks = BSK == 16 ? 1 : bs / 16;
    const int rows = BSK == 16 ? 16 : bs;
#pragma unroll 4
    for (int r = 0; r < rows; ++r)
      for (int ch = 0; ch < chunks; ++ch) acc = row_sad<16>(blk + (size_t)r * pitch + 16 * ch, win + (size_t)r * pitch + 16 * ch, acc);
    return acc;
  }
}

// One thread per block, any block size, for blocks whose nine candidates hold at most THREE distinct vectors u0 (the block's
// own, slot 0), u1, u2 in order of first appearance -- a listed block sits on the border between two or three motion layers, so
// this is nearly all of them.  With multiplicities m_k over the slots that have a neighbour: S_j = sum_k m_k * d(u_j, u_k)
// (:637-641; integer-valued, exact in float like the reference's running sum), E_j = (float)SAD_j + (lambda * mult) * S_j (:607,
// un-fused), FLT_MAX outside the image (:578-582); the smallest energy wins and ties go to the earlier first appearance, which is
// the reference's scan with strict '<' (:653-659) because slots with the same vector have the same energy.  Returns false if a
// fourth distinct vector shows up and any_count is false (the caller defers the block to a second pass that allows any count, so
// that the warps of the first pass stay converged); *out = the new vector otherwise.
template <int BSK>
__device__ __forceinline__ bool reg_eval_thread(const RegArgs& a, int pair, int bx, int by, uint32_t A0, const uint32_t (&pkin)[8],
                                                uint32_t* out, bool any_count, uint32_t* s_u) {
  const int gw = a.gw, gh = a.gh, bs = BSK >= 32 ? a.bs : BSK;
  const bool up = by > 0, dn = by < gh - 1, lf = bx > 0, rt = bx < gw - 1;
  const bool has[8] = {lf, rt, dn && rt, up && lf, up && rt, up, dn, dn && lf};  // slots 1..8: L, R, DR, UL, UR, U, D, DL
  uint32_t u1 = A0, u2 = A0;
  int n = 1, m0 = 1, m1 = 0, m2 = 0;
  bool more = false;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t v = pkin[i];
    if (has[i]) {  // a missing neighbour is not a candidate (its slot was read from the block's own index)
      if (v == A0) {
        ++m0;
      } else if (n >= 2 && v == u1) {
        ++m1;
      } else if (n >= 3 && v == u2) {
        ++m2;
      } else if (n == 1) {
        u1 = v; m1 = 1; n = 2;
      } else if (n == 2) {
        u2 = v; m2 = 1; n = 3;
      } else {
        more = true;
      }
    }
  }
  *out = A0;
  if (n == 1) return true;  // all candidates identical: index 0 wins
  const int x = bx * bs, y = by * bs;
  const int w = a.i1.w, h = a.i1.h, pitch = a.i1.pitch;
  const uint8_t* blk = a.i1.p + (size_t)pair * a.i1.plane + (size_t)y * pitch + x;
  const uint8_t* ref = a.i2.p + (size_t)pair * a.i2.plane;
  if (more) {
    if (!any_count) return false;
    // Four or more distinct vectors (rare outside the first sweep of a level): the same computation with the distinct vectors in
    // a per-thread column of shared memory (s_u[j * blockDim.x]) and their multiplicities packed four bits each.
    unsigned long long M = 1ull;
    int nd = 1;
    s_u[0] = A0;
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
      if (!has[i]) continue;
      const uint32_t v = pkin[i];
      int j = 0;
      while (j < nd && s_u[j * blockDim.x] != v) ++j;
      if (j == nd) { s_u[nd * blockDim.x] = v; ++nd; }
      M += 1ull << (4 * j);
    }
    float best = FLT_MAX;
    uint32_t r = A0;
#pragma unroll 1
    for (int j = 0; j < nd; ++j) {
      const uint32_t uj = s_u[j * blockDim.x];
      const int xj = mv_x(uj), yj = mv_y(uj);
      int S = 0;
      for (int k = 0; k < nd; ++k) {
        const uint32_t uk = s_u[k * blockDim.x];
        S += (int)((M >> (4 * k)) & 15ull) * (abs(xj - mv_x(uk)) + abs(yj - mv_y(uk)));
      }
      float E = FLT_MAX;
      if ((unsigned)(x + xj) <= (unsigned)(w - bs) && (unsigned)(y + yj) <= (unsigned)(h - bs))  // :578
        E = __fadd_rn(__uint2float_rn(block_sad_thread<BSK>(blk, ref + (size_t)(y + yj) * pitch + (x + xj), pitch, bs)),
                      __fmul_rn(a.lm, (float)S));
      if (j == 0 || E < best) { best = E; r = uj; }  // first appearance order, strict '<' (:653-659)
    }
    *out = r;
    return true;
  }
  const int x0 = mv_x(A0), y0 = mv_y(A0), x1 = mv_x(u1), y1 = mv_y(u1), x2 = mv_x(u2), y2 = mv_y(u2);
  const int d01 = abs(x0 - x1) + abs(y0 - y1), d02 = abs(x0 - x2) + abs(y0 - y2), d12 = abs(x1 - x2) + abs(y1 - y2);
  const float S0 = (float)(m1 * d01 + m2 * d02), S1 = (float)(m0 * d01 + m2 * d12), S2 = (float)(m0 * d02 + m1 * d12);
  const bool in0 = (unsigned)(x + x0) <= (unsigned)(w - bs) && (unsigned)(y + y0) <= (unsigned)(h - bs);  // :578
  const bool in1 = (unsigned)(x + x1) <= (unsigned)(w - bs) && (unsigned)(y + y1) <= (unsigned)(h - bs);
  const bool in2 = n == 3 && (unsigned)(x + x2) <= (unsigned)(w - bs) && (unsigned)(y + y2) <= (unsigned)(h - bs);
  float E0 = FLT_MAX, E1 = FLT_MAX, E2 = FLT_MAX;
  if (BSK <= 4) {
    // tiny windows: branch-free, out-of-image candidates read the block's own position
    const uint32_t sad0 = block_sad_thread<BSK>(blk, ref + (size_t)(in0 ? y + y0 : y) * pitch + (in0 ? x + x0 : x), pitch, bs);
    const uint32_t sad1 = block_sad_thread<BSK>(blk, ref + (size_t)(in1 ? y + y1 : y) * pitch + (in1 ? x + x1 : x), pitch, bs);
    E0 = in0 ? __fadd_rn(__uint2float_rn(sad0), __fmul_rn(a.lm, S0)) : FLT_MAX;
    E1 = in1 ? __fadd_rn(__uint2float_rn(sad1), __fmul_rn(a.lm, S1)) : FLT_MAX;
    if (n == 3) {
      const uint32_t sad2 = block_sad_thread<BSK>(blk, ref + (size_t)(in2 ? y + y2 : y) * pitch + (in2 ? x + x2 : x), pitch, bs);
      E2 = in2 ? __fadd_rn(__uint2float_rn(sad2), __fmul_rn(a.lm, S2)) : FLT_MAX;
    }
  } else {
    if (in0) E0 = __fadd_rn(__uint2float_rn(block_sad_thread<BSK>(blk, ref + (size_t)(y + y0) * pitch + (x + x0), pitch, bs)), __fmul_rn(a.lm, S0));
    if (in1) E1 = __fadd_rn(__uint2float_rn(block_sad_thread<BSK>(blk, ref + (size_t)(y + y1) * pitch + (x + x1), pitch, bs)), __fmul_rn(a.lm, S1));
    if (in2) E2 = __fadd_rn(__uint2float_rn(block_sad_thread<BSK>(blk, ref + (size_t)(y + y2) * pitch + (x + x2), pitch, bs)), __fmul_rn(a.lm, S2));
  }
  uint32_t r = A0;
  float best = E0;
  if (E1 < best) { best = E1; r = u1; }
  if (n == 3 && E2 < best) r = u2;
  *out = r;
  return true;
}

// partial SAD of this lane's share of one candidate window.  BSK == 8: row `row` (8 bytes) of an 8x8 block; BSK == 16: row
// `row` of a 16x16 block; BSK == 32: rows row, row + 32, ... of a block of 32 or more, in 16-byte chunks.  Windows start at
// any byte: a row is fetched as the two aligned vectors that contain it and the wanted words are selected by the start
// offset before the byte shift (two requests per row instead of three or five 32-bit ones).
template <int BSK>
__device__ __forceinline__ uint32_t team_partial_sad(const uint8_t* blk, const uint8_t* win, int pitch, int bs, int row) {
  uint32_t sum = 0;
  if (BSK == 8) {
    const size_t ro = (size_t)row * pitch;
    const uint2 A = __ldg(reinterpret_cast<const uint2*>(blk + ro));
    const uintptr_t ab = reinterpret_cast<uintptr_t>(win + ro);
    const uint2* q = reinterpret_cast<const uint2*>(ab & ~(uintptr_t)7);
    const uint32_t off = (uint32_t)(ab & 7);
    const uint2 q0 = __ldg(q), q1 = __ldg(q + 1);
    const bool w1 = (off & 4u) != 0u;
    const uint32_t sh = (off & 3u) * 8u;
    const uint32_t a0 = w1 ? q0.y : q0.x, a1 = w1 ? q1.x : q0.y, a2 = w1 ? q1.y : q1.x;
    sum = sad4(A.y, __funnelshift_r(a1, a2, sh), sad4(A.x, __funnelshift_r(a0, a1, sh), 0u));
  } else {
    const int rows = BSK == 16 ? 1 : bs / 32;
    const int chunks = BSK == 16 ? 1 : bs / 16;
    for (int rr = 0; rr < rows; ++rr) {
      for (int ch = 0; ch < chunks; ++ch) {
        const size_t ro = (size_t)(rr * 32 + row) * pitch + ch * 16;
        const uint4 A = __ldg(reinterpret_cast<const uint4*>(blk + ro));
        const uintptr_t ab = reinterpret_cast<uintptr_t>(win + ro);
        const uint4* q = reinterpret_cast<const uint4*>(ab & ~(uintptr_t)15);
        const uint32_t off = (uint32_t)(ab & 15);
        const uint4 q0 = __ldg(q), q1 = __ldg(q + 1);
        const bool s2 = (off & 8u) != 0u, s1 = (off & 4u) != 0u;
        const uint32_t sh = (off & 3u) * 8u;
        const uint32_t t0 = s2 ? q0.z : q0.x, t1 = s2 ? q0.w : q0.y, t2 = s2 ? q1.x : q0.z, t3 = s2 ? q1.y : q0.w,
                       t4 = s2 ? q1.z : q1.x, t5 = s2 ? q1.w : q1.y;
        const uint32_t a0 = s1 ? t1 : t0, a1 = s1 ? t2 : t1, a2 = s1 ? t3 : t2, a3 = s1 ? t4 : t3, a4 = s1 ? t5 : t4;
        sum = sad4(A.x, __funnelshift_r(a0, a1, sh), sum);
        sum = sad4(A.y, __funnelshift_r(a1, a2, sh), sum);
        sum = sad4(A.z, __funnelshift_r(a2, a3, sh), sum);
        sum = sad4(A.w, __funnelshift_r(a3, a4, sh), sum);
      }
    }
  }
  return sum;
}

// Blocks of 8x8 and larger, a team of adjacent lanes per block (16 lanes for 8x8 and 16x16 blocks, 32 above; whole warps call
// this together).  Lane s < 9 of a team OWNS candidate slot s: it loads that one vector, finds out whether an earlier slot
// holds the same one (match.any), sums its smoothness over the other lanes' vectors and ends with its energy.  The windows of
// the DISTINCT in-image vectors (a team-uniform list) are summed row-wise by all lanes, several per memory round trip
// (16x16 and up: three windows, one row each; 8x8: four windows, each half of the team takes two); a shuffle argmin over
// (energy, slot) gives every lane the winner.  ~300 instructions per lane instead of ~1600.
// Slot tl of block (bx, by): [C, L, R, DR, UL, UR, U, D, DL] (:441-449); L, UL, UR, U come from the new field.  Returns the
// slot's vector (an invalid slot returns the entry at the block's own index; the evaluator replaces it by C's).
__device__ __forceinline__ uint32_t team_slot_load(const RegArgs& a, const short2* O, const short2* P, int bx, int by, int tl,
                                                   bool& valid) {
  const int gw = a.gw, gh = a.gh;
  const int s = tl < 9 ? tl : 0;
  const int ddx = (s == 2 || s == 3 || s == 5) ? 1 : ((s == 1 || s == 4 || s == 8) ? -1 : 0);
  const int ddy = (s == 3 || s == 7 || s == 8) ? 1 : ((s == 4 || s == 5 || s == 6) ? -1 : 0);
  const bool from_new = s == 1 || s == 4 || s == 5 || s == 6;
  const int nx = bx + ddx, ny = by + ddy;
  valid = tl < 9 && nx >= 0 && nx < gw && ny >= 0 && ny < gh;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(from_new ? P : O);
  const int idx = by * gw + bx;
  return src[valid ? idx + ddy * gw + ddx : idx];
}

template <int BSK>  // 8, 16, or 32 (= 32 and larger)
__device__ __forceinline__ uint32_t reg_eval_team_lean(const RegArgs& a, int pair, int bx, int by, int tl, bool live, uint32_t my,
                                                       bool valid) {
  constexpr int TEAMSZ = BSK >= 32 ? 32 : 16;
  constexpr uint32_t FULL = 0xffffffffu;
  const int bs = a.bs;
  const int lane = threadIdx.x & 31;
  const int base = lane - tl;  // first lane of this team inside the warp
  const uint32_t c0 = __shfl_sync(FULL, my, base);  // slot 0 is always valid and reads O
  if (!valid) my = c0;
  // first slot of the team that holds my vector (invalid slots hold C's, i.e. slot 0's)
  const uint32_t same = __match_any_sync(FULL, my) & (0x1ffu << base);
  const int first = __ffs(same) - 1 - base;
  const int x = bx * bs, y = by * bs;
  const int w = a.i1.w, h = a.i1.h, pitch = a.i1.pitch;
  const int mx = mv_x(my), myy = mv_y(my);
  const bool inb = (unsigned)(x + mx) <= (unsigned)(w - bs) && (unsigned)(y + myy) <= (unsigned)(h - bs);  // :578
  const uint32_t valid_mask = (__ballot_sync(FULL, valid) >> base) & 0x1ffu;
  uint32_t need_mask = (__ballot_sync(FULL, live && valid && inb && first == tl) >> base) & 0x1ffu;
  // smoothness of my slot over all gathered candidates (:637-641): integer-valued, exact in float like the running sum
  int S = 0;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const uint32_t vk = __shfl_sync(FULL, my, base + k);
    const int dk = abs(mx - mv_x(vk)) + abs(myy - mv_y(vk));
    S += ((valid_mask >> k) & 1u) ? dk : 0;
  }
  const uint8_t* blk = a.i1.p + (size_t)pair * a.i1.plane + (size_t)y * pitch + x;
  const uint8_t* ref = a.i2.p + (size_t)pair * a.i2.plane;
  uint32_t mysad = 0;
  if (BSK == 8) {
    // four windows per round trip: half hf of the team sums windows j[hf] and j[2 + hf], one row per lane
    const int hf = tl >> 3, row = tl & 7;
    const int batches = __reduce_max_sync(FULL, (unsigned)((__popc(need_mask) + 3) / 4));
    for (int it = 0; it < batches; ++it) {
      int j[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        j[q] = need_mask ? __ffs(need_mask) - 1 : -1;
        need_mask &= need_mask - 1u;
      }
      uint32_t part[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int jj = hf ? j[2 * q + 1] : j[2 * q];
        const uint32_t vj = __shfl_sync(FULL, my, base + (jj >= 0 ? jj : 0));
        part[q] = 0;
        if (jj >= 0) part[q] = team_partial_sad<8>(blk, ref + (size_t)(y + mv_y(vj)) * pitch + (x + mv_x(vj)), pitch, bs, row);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int o = 4; o >= 1; o >>= 1) part[q] += __shfl_xor_sync(FULL, part[q], o);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {  // window j[c] was summed by half (c & 1) as its part[c >> 1]
        const uint32_t tot = __shfl_sync(FULL, part[c >> 1], base + 8 * (c & 1));
        mysad = (j[c] >= 0 && first == j[c]) ? tot : mysad;
      }
    }
  } else {
    const int batches = __reduce_max_sync(FULL, (unsigned)((__popc(need_mask) + 2) / 3));
    for (int it = 0; it < batches; ++it) {
      int j[3];
      uint32_t part[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        j[q] = need_mask ? __ffs(need_mask) - 1 : -1;
        need_mask &= need_mask - 1u;
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const uint32_t vj = __shfl_sync(FULL, my, base + (j[q] >= 0 ? j[q] : 0));
        part[q] = 0;
        if (j[q] >= 0) part[q] = team_partial_sad<BSK>(blk, ref + (size_t)(y + mv_y(vj)) * pitch + (x + mv_x(vj)), pitch, bs, tl);
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) {
#pragma unroll
        for (int o = TEAMSZ / 2; o >= 1; o >>= 1) part[q] += __shfl_xor_sync(FULL, part[q], o);
        mysad = (j[q] >= 0 && first == j[q]) ? part[q] : mysad;
      }
    }
  }
  // (:607) un-fused; (:578-582) FLT_MAX outside the image; slots without a neighbour can never win
  float e = (valid && inb) ? __fadd_rn(__uint2float_rn(mysad), __fmul_rn(a.lm, (float)S)) : FLT_MAX;
  int bi = valid ? tl : 15;
  // argmin over the team: smallest energy, ties to the smallest slot == the reference's scan with strict '<' (:653-659);
  // slot 0 (C) is always present, so an all-FLT_MAX block keeps its vector
#pragma unroll
  for (int o = TEAMSZ / 2; o >= 1; o >>= 1) {
    const float oe = __shfl_xor_sync(FULL, e, o);
    const int oi = __shfl_xor_sync(FULL, bi, o);
    const bool take = oe < e || (oe == e && oi < bi);
    e = take ? oe : e;
    bi = take ? oi : bi;
  }
  return __shfl_sync(FULL, my, base + (bi < 9 ? bi : 0));
}

// ============================================================================================ fused level schedule
// The whole regularisation schedule of one pyramid level (motion_framework.cpp:133-154: for every block size from the
// level's initial one down to 2, `sweeps` sweeps with lambda_multiplier 1..sweeps, then divide_blocks, lambda *= 2) in ONE
// launch: a cluster of CS CTAs owns a frame pair and walks through classify -> evaluation rounds -> next sweep -> split with
// cluster barriers only.  What the per-sweep launches of round 1 lost is gone: ~30 dependent launches per level (each at
// least a few microseconds, i.e. most of a single pair's latency), and a kernel boundary per phase at which every pair of a
// chunk waited for the slowest one (now a pair's cluster runs ahead on its own; pairs only meet at the end of the level).
// Large chunks run CS = 1 (one CTA per pair, plain __syncthreads, counters in shared memory), small chunks spread a pair
// over up to 8 SMs (barrier.cluster, counters in the DSMEM of rank 0; the barrier also invalidates the L1, so fields
// written by a sibling CTA are re-read from the L2).
//
// A sweep is the same fixed-point iteration as before, with one change: the first pass over the listed blocks already reads
// its "pred" neighbours from the NEW field (chaotic iteration from the start; a block that read a stale value is
// re-enqueued by the neighbour that changed, so the unique fixed point -- the reference's in-place raster result -- is
// reached whatever the interleaving), which shortens the tail because the list is in raster order.
namespace cg = cooperative_groups;

#ifndef BBME_LEVEL_THREADS
#define BBME_LEVEL_THREADS 512  // threads per CTA of the level kernel (one CTA per SM); 1024 (64 registers) measured slower: spills
#endif

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct LevelCtx {
  uint32_t* cnt;     // three rotating list counters + two alternating counters of the deferred list (shared memory of rank 0)
  int rank, cs;
  uint32_t gtid, gthreads;
};

template <bool MULTI>
__device__ __forceinline__ void level_sync() {
  if (MULTI) cg::this_cluster().sync();
  else __syncthreads();
}

// classify: copy O -> Y, list the blocks whose nine gathered candidates are not all identical (see k_reg_classify4)
__device__ __forceinline__ void level_classify(const RegArgs& a, int pair, const LevelCtx& lc, uint32_t* list) {
  const int gw = a.gw, gh = a.gh;
  // no __restrict__ / read-only loads on the fields: the two buffers swap roles from sweep to sweep inside one launch
  const uint32_t* O = reinterpret_cast<const uint32_t*>(a.O + (size_t)pair * a.mv_plane);
  uint32_t* Y = reinterpret_cast<uint32_t*>(a.Y + (size_t)pair * a.mv_plane);
  const int lane = threadIdx.x & 31;
  if ((gw & 3) == 0 && (a.mv_plane & 3) == 0) {
    // One thread = a strip of 4 x 4 blocks: six 128-bit row loads (rows by0 - 1 .. by0 + 4, clamped: a clamped neighbour is the
    // block itself or another neighbour, so the test is unchanged) issued together, left / right halo entries from the
    // neighbouring lanes by shuffle (explicit loads only at warp edges), four 128-bit stores.
    const int gw4 = gw >> 2;
    const uint32_t strips = (uint32_t)gw4 * (uint32_t)((gh + 3) >> 2);
    const uint32_t limit = (strips + 31u) / 32u * 32u;
    for (uint32_t t = lc.gtid; t < limit; t += lc.gthreads) {
      const bool live = t < strips;
      const int st = live ? (int)(t / gw4) : 0, cg = live ? (int)(t - (uint32_t)st * gw4) : 0;
      const int bx = cg * 4, by0 = st * 4;
      uint4 R[6];
      int ry[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        ry[i] = min(max(by0 - 1 + i, 0), gh - 1) * gw;
        R[i] = live ? *reinterpret_cast<const uint4*>(O + ry[i] + bx) : make_uint4(0u, 0u, 0u, 0u);
      }
      uint32_t Lh[6], Rh[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        Lh[i] = __shfl_up_sync(0xffffffffu, R[i].w, 1);
        Rh[i] = __shfl_down_sync(0xffffffffu, R[i].x, 1);
        if (cg == 0) Lh[i] = R[i].x;               // clamped: the block itself
        else if (lane == 0 && live) Lh[i] = O[ry[i] + bx - 1];
        if (cg == gw4 - 1) Rh[i] = R[i].w;
        else if (lane == 31 && live) Rh[i] = O[ry[i] + bx + 4];
      }
      uint32_t work = 0;  // bit 4 * r + j: block (bx + j, by0 + r) has candidates that differ
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (live && by0 + r < gh) {
          const uint32_t u[6] = {Lh[r], R[r].x, R[r].y, R[r].z, R[r].w, Rh[r]};
          const uint32_t m[6] = {Lh[r + 1], R[r + 1].x, R[r + 1].y, R[r + 1].z, R[r + 1].w, Rh[r + 1]};
          const uint32_t d[6] = {Lh[r + 2], R[r + 2].x, R[r + 2].y, R[r + 2].z, R[r + 2].w, Rh[r + 2]};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t k0 = m[j + 1];
            const bool same = u[j] == k0 && u[j + 1] == k0 && u[j + 2] == k0 && m[j] == k0 && m[j + 2] == k0 && d[j] == k0 &&
                              d[j + 1] == k0 && d[j + 2] == k0;
            work |= same ? 0u : (1u << (4 * r + j));
          }
          *reinterpret_cast<uint4*>(Y + (size_t)(by0 + r) * gw + bx) = R[r + 1];
        }
      }
      const int k = __popc(work);
      if (__ballot_sync(0xffffffffu, k > 0) == 0u) continue;
      int incl = k;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      uint32_t base = 0;
      if (lane == 31) base = atomicAdd(&lc.cnt[0], (uint32_t)incl);
      base = __shfl_sync(0xffffffffu, base, 31);
      uint32_t* dst = list + base + (uint32_t)(incl - k);
#pragma unroll
      for (int q = 0; q < 16; ++q)
        if ((work >> q) & 1u) *dst++ = (uint32_t)((by0 + (q >> 2)) * gw + bx + (q & 3));
    }
  } else {
    const uint32_t nb = (uint32_t)gw * gh;
    const uint32_t limit = (nb + 31u) / 32u * 32u;
    for (uint32_t t = lc.gtid; t < limit; t += lc.gthreads) {
      bool work = false;
      if (t < nb) {
        const int by = (int)(t / gw), bx = (int)(t - (uint32_t)by * gw);
        const int ru = max(by - 1, 0) * gw, rm = by * gw, rd = min(by + 1, gh - 1) * gw;
        const int cl = max(bx - 1, 0), cr = min(bx + 1, gw - 1);
        const uint32_t k0 = O[t];
        const uint32_t v[8] = {O[ru + cl], O[ru + bx], O[ru + cr], O[rm + cl], O[rm + cr], O[rd + cl], O[rd + bx], O[rd + cr]};
        bool same = true;
#pragma unroll
        for (int j = 0; j < 8; ++j) same = same && v[j] == k0;
        Y[t] = k0;
        work = !same;
      }
      const uint32_t m = __ballot_sync(0xffffffffu, work);
      if (m) {
        const int leader = __ffs(m) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&lc.cnt[0], (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (work) list[base + __popc(m & ((1u << lane) - 1u))] = t;
      }
    }
  }
}

// One sweep at the current block size: classify, then rounds until the work list is empty.  Every listed block is evaluated
// by one thread (reg_eval_thread).  In a large round (the sweep's first pass) the few blocks with four or more distinct candidate
// vectors are deferred to a second pass of the same round, so that the warps of the first pass stay converged on the register-only
// path; a small round (the fix-up tail, where a thread has at most a couple of blocks) evaluates them in line and saves the barrier.
template <int BSK, bool MULTI>
__device__ __forceinline__ void level_sweep(const RegArgs& a, int pair, const LevelCtx& lc, uint32_t& ep, uint32_t& rounds,
                                            uint32_t& blocks, uint32_t* s_u) {
  const short2* O = a.O + (size_t)pair * a.mv_plane;
  short2* Y = a.Y + (size_t)pair * a.mv_plane;
  uint32_t* Yu = reinterpret_cast<uint32_t*>(Y);
  uint32_t* stamp = a.stamp + (size_t)pair * a.wl_plane;
  uint32_t* lists[2] = {a.list0 + (size_t)pair * a.wl_plane, a.list1 + (size_t)pair * a.wl_plane};
  uint32_t* dlist = reinterpret_cast<uint32_t*>(a.nv) + (size_t)pair * a.wl_plane;  // blocks deferred to the second pass
  const int lane = threadIdx.x & 31;
  if (lc.gtid == 0) {
#pragma unroll
    for (int i = 0; i < 5; ++i) lc.cnt[i] = 0;
  }
  level_sync<MULTI>();
  // BBME_REG_PROFILE: per (level, block size) wall time of pair 0's phases in ns, words [0] classify [1] first pass [2] later
  // rounds [3] rounds [4] listed blocks [5] blocks of later rounds [6] deferred blocks
  uint32_t* prof = (a.hist && pair == 0 && lc.gtid == 0) ? a.hist + 8 * (31 - __clz(BSK)) : nullptr;
  unsigned long long t0 = 0;
  if (prof) t0 = globaltimer_ns();
  level_classify(a, pair, lc, lists[0]);
  level_sync<MULTI>();
  if (prof) { const unsigned long long t1 = globaltimer_ns(); prof[0] += (uint32_t)(t1 - t0); t0 = t1; prof[4] += lc.cnt[0]; }
  for (int r = 0;; ++r) {
    // round r reads list[r & 1] (counter r % 3), appends to list[(r + 1) & 1] (counter (r + 1) % 3) and clears counter
    // (r + 2) % 3, which was last read before the barrier that precedes this round; the deferred list's counter alternates
    // between words 3 and 4 for the same reason
    const uint32_t cnt = *reinterpret_cast<volatile uint32_t*>(&lc.cnt[r % 3]);
    if (cnt == 0) break;
    if (lc.gtid == 0) {
      lc.cnt[(r + 2) % 3] = 0;
      lc.cnt[3 + ((r + 1) & 1)] = 0;
    }
    const uint32_t* lcur = lists[r & 1];
    uint32_t* lnext = lists[(r + 1) & 1];
    uint32_t* next_count = &lc.cnt[(r + 1) % 3];
    uint32_t* dcount = &lc.cnt[3 + (r & 1)];
    const bool in_line = cnt <= 2u * lc.gthreads;  // the same for every thread of the cluster
    ++ep;
    if (BSK >= 8 && in_line) {
      // A small round of large blocks is a latency problem, not a throughput problem: a team of lanes per block (one window
      // row per lane, all distinct candidates in one or two round trips) instead of one thread walking through 16 rows of
      // every candidate.
      constexpr int TEAMSZ = BSK >= 32 ? 32 : 16;
      constexpr int TPW = 32 / TEAMSZ;
      const uint32_t team = lc.gtid / TEAMSZ, tl = lc.gtid % TEAMSZ, nteams = lc.gthreads / TEAMSZ;
      const uint32_t limit = (cnt + TPW - 1) / TPW * TPW;
      for (uint32_t e = team; e < limit; e += nteams) {
        const bool live = e < cnt;
        const int b = (int)lcur[live ? e : cnt - 1];
        const int bx = b % a.gw, by = b / a.gw;
        bool valid = false;
        const uint32_t my = team_slot_load(a, O, Y, bx, by, (int)tl, valid);
        const uint32_t nv = reg_eval_team_lean<BSK <= 4 ? 8 : BSK>(a, pair, bx, by, (int)tl, live, my, valid);
        const bool changed = tl == 0 && live && nv != Yu[b];
        if (changed) Yu[b] = nv;
        push_dependents(changed, bx, by, a.gw, a.gh, stamp, ep, lnext, next_count);
      }
    } else {
      // Software-pipelined: the list entry is loaded two iterations ahead and the nine vectors one iteration ahead, so that an
      // evaluation waits for ONE memory round trip (its windows) instead of three dependent ones.  Reading the vectors early
      // is safe: a block that read a neighbour's old value is re-enqueued by that neighbour's push, whenever the read happened
      // (chaotic iteration).
      // (Tried: a thread taking runs of four consecutive list entries and patching the next entry's "L" vector with the value
      // just computed -- an in-place sweep along the row inside a run.  It saves a quarter of the later rounds' blocks but
      // not one round, and the strided list reads cost more than that: K = 1.)
      constexpr uint32_t K = 1u;
      const uint32_t runs = (cnt + K - 1u) / K;
      const uint32_t run_limit = (runs + 31u) / 32u * 32u;  // whole warps iterate together
      const uint32_t G = lc.gthreads;
      uint32_t run = lc.gtid, k = 0;
      auto advance = [&](uint32_t& rr, uint32_t& kk) { if (++kk == K) { kk = 0; rr += G; } };
      auto load_entry = [&](uint32_t rr, uint32_t kk) -> int {
        const uint32_t e = rr * K + kk;
        return rr < run_limit ? (int)lcur[e < cnt ? e : cnt - 1] : 0;
      };
      uint32_t r1 = run, k1 = k;   // position of b1
      uint32_t r2 = run, k2 = k;   // position of b2
      advance(r2, k2);
      int b1 = load_entry(r1, k1);
      int b2 = load_entry(r2, k2);
      uint32_t A1 = 0, pk1[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) pk1[i] = 0;
      if (r1 < run_limit) small_gather(a, O, Y, b1 % a.gw, b1 / a.gw, A1, pk1);
      while (r1 < run_limit) {
        const bool live = r1 * K + k1 < cnt;
        const int b = b1;
        const uint32_t A0 = A1;
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = pk1[i];
        // shift the pipeline: b1 <- b2, b2 <- the entry after it
        b1 = b2; r1 = r2; k1 = k2;
        advance(r2, k2);
        b2 = load_entry(r2, k2);
        if (r1 < run_limit) small_gather(a, O, Y, b1 % a.gw, b1 / a.gw, A1, pk1);
        const int bx = b % a.gw, by = b / a.gw;
        uint32_t nv = 0;
        const bool done = !live || reg_eval_thread<BSK>(a, pair, bx, by, A0, pk, &nv, in_line, s_u);
        if (!in_line) {
          const uint32_t dm = __ballot_sync(0xffffffffu, !done);
          if (dm) {
            uint32_t dbase = 0;
            const int leader = __ffs(dm) - 1;
            if (lane == leader) dbase = atomicAdd(dcount, (uint32_t)__popc(dm));
            dbase = __shfl_sync(0xffffffffu, dbase, leader);
            if (!done) dlist[dbase + __popc(dm & ((1u << lane) - 1u))] = (uint32_t)b;
          }
        }
        const bool changed = live && done && nv != Yu[b];
        if (changed) Yu[b] = nv;
        if (live && done && b1 == b + 1 && bx + 1 < a.gw) pk1[0] = nv;  // the next block's left neighbour is this block
        push_dependents(changed, bx, by, a.gw, a.gh, stamp, ep, lnext, next_count);
      }
    }
    if (!in_line) {
      // second pass of a large round: the blocks with four or more distinct candidate vectors
      if (MULTI) __threadfence();
      level_sync<MULTI>();
      const uint32_t dcnt = *reinterpret_cast<volatile uint32_t*>(dcount);
      if (prof) prof[6] += dcnt;
      const uint32_t dlimit = (dcnt + 31u) / 32u * 32u;
      for (uint32_t e = lc.gtid; e < dlimit; e += lc.gthreads) {
        const bool live = e < dcnt;
        const int b = (int)dlist[live ? e : dcnt - 1];
        const int bx = b % a.gw, by = b / a.gw;
        uint32_t A0, pk[8], nv = 0;
        small_gather(a, O, Y, bx, by, A0, pk);
        if (live) reg_eval_thread<BSK>(a, pair, bx, by, A0, pk, &nv, true, s_u);
        const bool changed = live && nv != Yu[b];
        if (changed) Yu[b] = nv;
        push_dependents(changed, bx, by, a.gw, a.gh, stamp, ep, lnext, next_count);
      }
    }
    if (MULTI) __threadfence();
    level_sync<MULTI>();
    if (prof) {
      const unsigned long long t1 = globaltimer_ns();
      prof[r == 0 ? 1 : 2] += (uint32_t)(t1 - t0);
      t0 = t1;
      if (r > 0) { prof[3] += 1; prof[5] += cnt; }
    }
    if (r > 0) {  // the first pass is the sweep itself; later rounds are the fix-up
      rounds += 1;
      blocks += cnt;
    }
  }
}

constexpr int kLevelThreads = BBME_LEVEL_THREADS;

template <bool MULTI>
__global__ void __launch_bounds__(kLevelThreads, 1) k_reg_level(RegArgs a, int sweeps, float lambda0, int first_mult, int single_stage) {
  __shared__ uint32_t s_cnt[8];
  __shared__ uint32_t s_ucol[9 * kLevelThreads];  // per-thread columns of distinct candidate vectors (reg_eval_thread)
  uint32_t* s_u = s_ucol + threadIdx.x;
  LevelCtx lc;
  int pair;
  if (MULTI) {
    cg::cluster_group cl = cg::this_cluster();
    lc.cs = (int)cl.num_blocks();
    lc.rank = (int)cl.block_rank();
    lc.cnt = cl.map_shared_rank(s_cnt, 0);
    pair = blockIdx.x / lc.cs;
  } else {
    lc.cs = 1;
    lc.rank = 0;
    lc.cnt = s_cnt;
    pair = blockIdx.x;
  }
  lc.gtid = (uint32_t)lc.rank * blockDim.x + threadIdx.x;
  lc.gthreads = (uint32_t)lc.cs * blockDim.x;
  uint32_t* ctr = a.ctr + (size_t)pair * kCtrWords;
  uint32_t ep = ctr[CTR_EPOCH];
  if (ep > 0xf0000000u) {  // the de-duplication stamps must stay below every epoch still to come: restart before a wrap
    uint32_t* stamp = a.stamp + (size_t)pair * a.wl_plane;
    for (size_t i = lc.gtid; i < a.wl_plane; i += lc.gthreads) stamp[i] = 0u;
    ep = 0;
    if (MULTI) __threadfence();
  }
  level_sync<MULTI>();
  uint32_t rounds = 0, blocks = 0;
  float lambda = lambda0;
  for (int g = a.bs; g > 1; g >>= 1) {
    for (int sw = first_mult; sw < first_mult + sweeps; ++sw) {
      a.lm = lambda * (float)sw;  // lambda * (float)lambda_multiplier, motion_framework.cpp:607
      switch (g >= 32 ? 32 : g) {
        case 32: level_sweep<32, MULTI>(a, pair, lc, ep, rounds, blocks, s_u); break;
        case 16: level_sweep<16, MULTI>(a, pair, lc, ep, rounds, blocks, s_u); break;
        case 8: level_sweep<8, MULTI>(a, pair, lc, ep, rounds, blocks, s_u); break;
        case 4: level_sweep<4, MULTI>(a, pair, lc, ep, rounds, blocks, s_u); break;
        default: level_sweep<2, MULTI>(a, pair, lc, ep, rounds, blocks, s_u); break;
      }
      const short2* t = a.O; a.O = a.Y; a.Y = const_cast<short2*>(t);
    }
    if (single_stage) break;
    if (g > 2) {
      // MF::divide_blocks (motion_framework.cpp:845-862): a.O (gw x gh) -> a.Y (2gw x 2gh)
      const uint32_t* in = reinterpret_cast<const uint32_t*>(a.O + (size_t)pair * a.mv_plane);
      uint32_t* out = reinterpret_cast<uint32_t*>(a.Y + (size_t)pair * a.mv_plane);
      const int gw = a.gw, gh = a.gh;
      if ((gw & 1) == 0 && (a.mv_plane & 3) == 0) {
        const int hw = gw >> 1;
        for (uint32_t i = lc.gtid; i < (uint32_t)hw * gh; i += lc.gthreads) {
          const int y = (int)(i / hw), x2 = (int)(i - (uint32_t)y * hw);
          const uint2 v = *reinterpret_cast<const uint2*>(in + (size_t)y * gw + 2 * x2);
          const uint4 o = make_uint4(v.x, v.x, v.y, v.y);
          uint32_t* dst = out + (size_t)(2 * y) * (2 * gw) + 4 * x2;
          *reinterpret_cast<uint4*>(dst) = o;
          *reinterpret_cast<uint4*>(dst + 2 * gw) = o;
        }
      } else {
        const int ow = 2 * gw;
        for (uint32_t i = lc.gtid; i < (uint32_t)ow * 2 * gh; i += lc.gthreads) {
          const int y = (int)(i / ow), x = (int)(i - (uint32_t)y * ow);
          out[i] = in[(size_t)(y >> 1) * gw + (x >> 1)];
        }
      }
      const short2* t = a.O; a.O = a.Y; a.Y = const_cast<short2*>(t);
      a.gw *= 2;
      a.gh *= 2;
      if (MULTI) __threadfence();
      level_sync<MULTI>();
    }
    a.bs = g >> 1;
    lambda = lambda * 2;
  }
  if (lc.gtid == 0) {
    ctr[CTR_EPOCH] = ep;
    ctr[CTR_ROUNDS] += rounds;
    ctr[CTR_BLOCKS] += blocks;
  }
  // rank 0 owns the counters its siblings read through DSMEM: nobody leaves before everybody has read the final zero
  if (MULTI) level_sync<MULTI>();
}

int launch_reg_level(const RegArgs& a, int sweeps, float lambda0, int first_mult, int single_stage, int n, int sm_budget,
                     cudaStream_t s) {
  // cluster size: spread a pair over several SMs while the chunk leaves SMs of its budget idle (the budget is the GPU divided by
  // the pipeline slots: chunks of other slots run beside this one)
  int cs = 1;
  while (cs < 8 && 2 * cs * n <= sm_budget) cs *= 2;
  if (const char* e = getenv("BBME_REG_CLUSTER")) {
    const int v = atoi(e);
    if (v == 1 || v == 2 || v == 4 || v == 8) cs = v;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n * cs));
  cfg.blockDim = dim3(kLevelThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e;
  if (cs == 1) {
    k_reg_level<false><<<n, kLevelThreads, 0, s>>>(a, sweeps, lambda0, first_mult, single_stage);
    e = cudaGetLastError();
  } else {
    e = cudaLaunchKernelEx(&cfg, k_reg_level<true>, a, sweeps, lambda0, first_mult, single_stage);
  }
  return e == cudaSuccess ? 0 : -1;
}

// ============================================================================================ integer peak
// Register-only, dependence-free VABSDIFF4.U8.ACC chains on every SM (same measurement as bench_micro/int_peak.cu).
__global__ void __launch_bounds__(1024) k_int_peak(uint32_t* out, uint32_t seed, long long* cyc) {
  constexpr int CH = 16, ITERS = 2048;
  uint32_t a[CH], b[CH], acc[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    a[i] = (threadIdx.x + 1) * 0x01010101u * (i + 1) + seed;
    b[i] = a[i] ^ 0x5a5a5a5au;
    acc[i] = i;
  }
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a[i]), "r"(b[i]));
  }
  const long long t1 = clock64();
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) r ^= acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int measure_int_peak(int sm_count, double* absdiff_per_s, double* sm_mhz) {
  uint32_t* d_out = nullptr;
  long long* d_cyc = nullptr;
  if (cudaMalloc(&d_out, sizeof(uint32_t) * sm_count * 1024) != cudaSuccess) return -1;
  if (cudaMalloc(&d_cyc, sizeof(long long) * sm_count) != cudaSuccess) { cudaFree(d_out); return -1; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_int_peak<<<sm_count, 1024>>>(d_out, 1u, d_cyc);
  cudaEventRecord(e0);
  k_int_peak<<<sm_count, 1024>>>(d_out, 2u, d_cyc);
  cudaEventRecord(e1);
  int rc = cudaDeviceSynchronize() == cudaSuccess ? 0 : -1;
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  long long cyc = 0;
  cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  cudaFree(d_cyc);
  if (rc == 0 && ms > 0.f) {
    const double lane_ops = (double)sm_count * 1024.0 * 2048.0 * 16.0;
    *absdiff_per_s = lane_ops * 4.0 / (ms * 1e-3);
    *sm_mhz = (double)cyc / (ms * 1e3);
  }
  return rc;
}

}  // namespace bbme
