// kernels.cu -- pyramid build, generic block search, MV plumbing, dense export (the regularisation is regularize.cu).
//
// Reference semantics (cited as file:line of /root/reference) are restated in DESIGN.md; nothing here is a
// translation of the reference's loops: the data layout is pitched uint8 planes + block-granular short2
// fields, and the in-place raster sweep is reproduced by a Jacobi pass followed by fixed-point rounds.
#include "kernels.h"

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
