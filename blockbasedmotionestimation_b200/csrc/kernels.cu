// kernels.cu -- pyramid build, generic block search, MV plumbing and the exact regularisation sweep.
//
// Reference semantics (cited as file:line of /root/reference) are restated in DESIGN.md; nothing here is a
// translation of the reference's loops: the data layout is pitched uint8 planes + block-granular short2
// fields, and the in-place raster sweep is reproduced by a Jacobi pass followed by fixed-point rounds.
#include "kernels.h"

#include <float.h>

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
  int acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const int wj = (j == 0 || j == 4) ? 1 : ((j == 2) ? 6 : 4);
    const uint8_t* row = src + (size_t)reflect101(2 * y + j - 2, sh) * sv.pitch;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (live && full) v = __ldg(reinterpret_cast<const uint4*>(row + 16 * t));
    // halo: bytes 16t-2, 16t-1 (high half of the left neighbour's last word) and 16t+16 (right neighbour's first byte)
    uint32_t left = __shfl_up_sync(0xffffffffu, v.w, 1);
    uint32_t right = __shfl_down_sync(0xffffffffu, v.x, 1);
    int p[19];  // source pixels 16t-2 .. 16t+16
    if (live) {
      if (lane == 0 || t == 0) {
        p[0] = (int)__ldg(row + reflect101(16 * t - 2, sw));
        p[1] = (int)__ldg(row + reflect101(16 * t - 1, sw));
      } else {
        p[0] = (int)((left >> 16) & 0xffu);
        p[1] = (int)(left >> 24);
      }
      if (lane == 31 || 16 * t + 16 >= sw) p[18] = (int)__ldg(row + reflect101(16 * t + 16, sw));
      else p[18] = (int)(right & 0xffu);
      const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 16; ++q) p[2 + q] = (int)((wd[q >> 2] >> ((q & 3) * 8)) & 0xffu);
      if (16 * t + 16 > sw || !full) {  // ragged right edge: pixels past the image width are reflected, not read from the pitch tail
#pragma unroll
        for (int q = 0; q < 16; ++q)
          if (16 * t + q >= sw || !full) p[2 + q] = (int)__ldg(row + reflect101(16 * t + q, sw));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int hsum = p[2 * i] + 4 * p[2 * i + 1] + 6 * p[2 * i + 2] + 4 * p[2 * i + 3] + p[2 * i + 4];
        acc[i] += wj * hsum;
      }
    }
  }
  if (!live || y >= dh) return;
  uint8_t* dst = (frame ? d2 : d1) + (size_t)pair * dplane + (size_t)y * dpitch + 8 * t;
  uint32_t lo = 0, hi = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    lo |= (uint32_t)((acc[i] + 128) >> 8) << (8 * i);
    hi |= (uint32_t)((acc[4 + i] + 128) >> 8) << (8 * i);
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
__global__ void __launch_bounds__(128) k_search_generic(ImgView i1, ImgView i2, MvView mv, int bs, int R,
                                                        unsigned long long* __restrict__ counters) {
  const int pair = blockIdx.y;
  const int bx = blockIdx.x % mv.gw, by = blockIdx.x / mv.gw;
  const int x = bx * bs, y = by * bs;
  const int w = i1.w, h = i1.h, pitch = i1.pitch;
  short2* slot = mv.p + (size_t)pair * mv.plane + (size_t)by * mv.gw + bx;
  const short2 pred = *slot;
  const int x2 = x + pred.x, y2 = y + pred.y;
  if (x2 < 0 || y2 < 0 || x2 + bs > w || y2 + bs > h) {  // :304-310 -> MV 0, no search
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
    const unsigned long long key = ((unsigned long long)sad << 32) | spiral_rank(dx, dy);
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
    if (rank != 0) {
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
      const int nx = min(R, w - bs - x2) - max(-R, -x2) + 1;
      const int ny = min(R, h - bs - y2) - max(-R, -y2) + 1;
      atomicAdd(&counters[0], (unsigned long long)(nx * ny));
      atomicAdd(&counters[1], (unsigned long long)(nx * ny) * (unsigned long long)(bs * bs));
    }
  }
}

void launch_search_generic(ImgView i1, ImgView i2, MvView mv, int bs, int R, int n, unsigned long long* counters,
                           cudaStream_t s) {
  dim3 grid(mv.gw * mv.gh, n);
  k_search_generic<<<grid, 128, 0, s>>>(i1, i2, mv, bs, R, counters);
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

void launch_divide(const short2* in, int gw, int gh, size_t in_plane, short2* out, size_t out_plane, int n,
                   cudaStream_t s) {
  dim3 grid((4 * gw * gh + 255) / 256, n);
  k_divide<<<grid, 256, 0, s>>>(in, gw, gh, in_plane, out, out_plane);
}

// Final dense field (motion_framework.cpp:205-206, 815-826, 218): CV_32FC2, every 2x2 block shares one MV.
// One thread = two horizontally adjacent pixels = one 16-byte store; HBM-write-bound (8 B / pixel).
__global__ void __launch_bounds__(256) k_export(const short2* __restrict__ mv2, int gw2, size_t mv_plane,
                                                float* __restrict__ out, int pw, int ph, size_t out_plane) {
  const int x2 = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int pair = blockIdx.z;
  if (x2 >= gw2 || y >= ph) return;
  const short2 m = __ldg(&mv2[(size_t)pair * mv_plane + (size_t)(y >> 1) * gw2 + x2]);
  const float u = (float)m.x, v = (float)m.y;
  float* o = out + (size_t)pair * out_plane + ((size_t)y * pw + 2 * x2) * 2;
  if ((reinterpret_cast<uintptr_t>(o) & 15) == 0) {
    __stcs(reinterpret_cast<float4*>(o), make_float4(u, v, u, v));
  } else {
    o[0] = u; o[1] = v; o[2] = u; o[3] = v;
  }
}

void launch_export(const short2* mv2, int gw2, size_t mv_plane, float* out, int pw, int ph, size_t out_plane, int n,
                   cudaStream_t s) {
  dim3 grid((gw2 + 127) / 128, ph, n);
  k_export<<<grid, 128, 0, s>>>(mv2, gw2, mv_plane, out, pw, ph, out_plane);
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
// Small blocks (2x2, 4x4: 95 % of all block evaluations) are evaluated by one thread, branch-free: all nine slots
// are always computed (missing neighbours hold a copy of C and are masked out of the argmin; their contribution to
// the smoothness sums is removed arithmetically), so a warp never diverges; the only branch is warp-uniform (every
// lane's nine candidates identical -> nothing can change).
template <int BS>  // 2 or 4
__device__ __forceinline__ short2 reg_eval_small(const RegArgs& a, int pair, const short2* __restrict__ O,
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
  float best = 0.f;
  int best_i = 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const int px = x + cx[i], py = y + cy[i];
    const bool inb = (unsigned)px <= (unsigned)(w - BS) && (unsigned)py <= (unsigned)(h - BS);  // :578
    const uint8_t* b = ref + (size_t)(inb ? py : y) * pitch + (inb ? px : x);  // out-of-image candidates read the block's own position
    uint32_t sad;
    if (BS == 2) {
      const uint32_t bv = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[pitch] << 16) | ((uint32_t)b[pitch + 1] << 24);
      sad = sad4(A[0], bv, 0u);
    } else {
      sad = 0;
      const uintptr_t ab = reinterpret_cast<uintptr_t>(b);
      const uint32_t sh = (uint32_t)(ab & 3) * 8u;
      const uint32_t* bw = reinterpret_cast<const uint32_t*>(ab & ~(uintptr_t)3);
#pragma unroll
      for (int r = 0; r < (BS == 2 ? 1 : 4); ++r) {
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(bw) + (size_t)r * pitch);
        sad = sad4(A[r], __funnelshift_r(rw[0], rw[1], sh), sad);
      }
    }
    const float e = inb ? __fadd_rn(__uint2float_rn(sad), __fmul_rn(a.lm, S[i])) : FLT_MAX;
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

//
// A block is evaluated by a TEAM of adjacent lanes (1 for block sizes 2 and 4, min(bs, 32) above): every lane takes
// bs / TEAM rows of each candidate's SAD and the partial sums are combined with xor-shuffles inside the team, so
// the latency of one evaluation no longer grows with the block area (the fix-up rounds are latency-bound).
template <int TEAM>
__device__ __forceinline__ short2 reg_eval(const RegArgs& a, int pair, const short2* __restrict__ O,
                                           const short2* P, int bx, int by, int tl, uint32_t team_mask) {
  const int gw = a.gw, gh = a.gh, bs = a.bs;
  const int idx = by * gw + bx;
  const bool up = by > 0, dn = by < gh - 1, lf = bx > 0, rt = bx < gw - 1;
  short2 c[9];
  uint32_t mask = 1u;
  c[0] = O[idx];
#pragma unroll
  for (int i = 1; i < 9; ++i) c[i] = c[0];
  if (lf) { c[1] = P[idx - 1]; mask |= 1u << 1; }
  if (rt) { c[2] = O[idx + 1]; mask |= 1u << 2; }
  if (dn && rt) { c[3] = O[idx + gw + 1]; mask |= 1u << 3; }
  if (up && lf) { c[4] = P[idx - gw - 1]; mask |= 1u << 4; }
  if (up && rt) { c[5] = P[idx - gw + 1]; mask |= 1u << 5; }
  if (up) { c[6] = P[idx - gw]; mask |= 1u << 6; }
  if (dn) { c[7] = O[idx + gw]; mask |= 1u << 7; }
  if (dn && lf) { c[8] = O[idx + gw - 1]; mask |= 1u << 8; }

  // all candidates identical -> every energy is equal -> index 0 wins (:653-659)
  const uint32_t k0 = pack_mv(c[0]);
  bool all_same = true;
#pragma unroll
  for (int i = 1; i < 9; ++i) all_same = all_same && (pack_mv(c[i]) == k0);  // dropped slots hold c[0]
  if (all_same) return c[0];

  // smoothness: S_i = sum over gathered candidates k of |c_k.x - c_i.x| + |c_k.y - c_i.y|  (:637-641)
  int S[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) S[i] = 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
#pragma unroll
    for (int k = i + 1; k < 9; ++k) {
      const int d = abs((int)c[i].x - (int)c[k].x) + abs((int)c[i].y - (int)c[k].y);
      const bool both = ((mask >> i) & (mask >> k) & 1u) != 0u;
      S[i] += both ? d : 0;
      S[k] += both ? d : 0;
    }
  }

  const int x = bx * bs, y = by * bs;
  const int w = a.i1.w, h = a.i1.h, pitch = a.i1.pitch;
  const uint8_t* blk = a.i1.p + (size_t)pair * a.i1.plane + (size_t)y * pitch + x;
  const uint8_t* ref = a.i2.p + (size_t)pair * a.i2.plane;
  const int rows = TEAM > 1 ? bs / TEAM : bs;  // rows of the block this lane sums
  const int r0 = TEAM > 1 ? tl * rows : 0;
  float best = 0.f;
  int best_i = 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    if (!((mask >> i) & 1u)) continue;
    // a candidate equal to an earlier gathered one has the identical energy and cannot win by strict '<'
    bool dup = false;
#pragma unroll
    for (int j = 0; j < i; ++j) dup = dup || (((mask >> j) & 1u) && pack_mv(c[j]) == pack_mv(c[i]));
    if (dup) continue;
    const int px = x + c[i].x, py = y + c[i].y;
    float e;
    if (px < 0 || px > w - bs || py < 0 || py > h - bs) {
      e = FLT_MAX;  // :578-582
    } else {
      uint32_t sad;
      if (TEAM == 1) {
        sad = sad_block_unaligned(blk, ref + (size_t)py * pitch + px, pitch, bs);
      } else {
        sad = 0;
        const int words = bs >> 2;
        for (int r = r0; r < r0 + rows; ++r) {
          const uint32_t* aw = reinterpret_cast<const uint32_t*>(blk + (size_t)r * pitch);
          const uintptr_t ab = reinterpret_cast<uintptr_t>(ref + (size_t)(py + r) * pitch + px);
          const uint32_t* bw = reinterpret_cast<const uint32_t*>(ab & ~(uintptr_t)3);
          const uint32_t sh = (uint32_t)(ab & 3) * 8u;
          uint32_t w0 = __ldg(bw);
          for (int k = 0; k < words; ++k) {
            const uint32_t w1 = __ldg(bw + k + 1);
            sad = sad4(__ldg(aw + k), __funnelshift_r(w0, w1, sh), sad);
            w0 = w1;
          }
        }
#pragma unroll
        for (int off = TEAM / 2; off > 0; off >>= 1) sad += __shfl_xor_sync(team_mask, sad, off);
      }
      e = __fadd_rn(__uint2float_rn(sad), __fmul_rn(a.lm, __int2float_rn(S[i])));
    }
    if (i == 0) { best = e; best_i = 0; }
    else if (e < best) { best = e; best_i = i; }
  }
  short2 r = c[0];
#pragma unroll
  for (int i = 1; i < 9; ++i) r = (best_i == i) ? c[i] : r;
  return r;
}

// Blocks whose "pred" neighbour is block (bx,by): its right, lower-left, lower and lower-right neighbours.
template <typename Push>
__device__ __forceinline__ void for_each_dependent(int bx, int by, int gw, int gh, Push push) {
  if (bx + 1 < gw) push(by * gw + bx + 1);
  if (by + 1 < gh) {
    if (bx > 0) push((by + 1) * gw + bx - 1);
    push((by + 1) * gw + bx);
    if (bx + 1 < gw) push((by + 1) * gw + bx + 1);
  }
}

// TEAM == 1 evaluates 2x2 blocks, TEAM == 2 is the tag for "one thread per 4x4 block" (TEAMSZ below is 1 for both)
template <int TEAM>
__device__ __forceinline__ short2 reg_eval_any(const RegArgs& a, int pair, const short2* __restrict__ O, const short2* P,
                                               int bx, int by, int tl, uint32_t team_mask, bool live) {
  if (TEAM == 1) {
    return reg_eval_small<2>(a, pair, O, P, bx, by, live);
  } else if (TEAM == 2) {
    return reg_eval_small<4>(a, pair, O, P, bx, by, live);
  } else {
    return reg_eval<TEAM>(a, pair, O, P, bx, by, tl, team_mask);
  }
}

// Pass 1 of a sweep (a Jacobi step: every block evaluated with the OLD field in all nine slots), in two kernels:
//  k_reg_classify  one thread per block, ~20 registers, full occupancy: blocks whose nine candidates are identical
//                  keep their vector (all energies equal, index 0 wins, :653-659) and are done; the others are
//                  compacted into the evaluation list (warp-aggregated append, so neighbours stay neighbours).
//                  On real fields 80-96 % of the 2x2 / 4x4 blocks are of the first kind.
//  k_reg_eval      dense evaluation of the listed blocks; blocks whose value changed enqueue their dependents
//                  (they may have used a stale "pred" value) for the fix-up rounds.
__global__ void __launch_bounds__(256) k_reg_classify(RegArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int pair = blockIdx.y;
  const int gw = a.gw, gh = a.gh;
  const bool live = i < gw * gh;
  const short2* __restrict__ O = a.O + (size_t)pair * a.mv_plane;
  bool work = false;
  if (live) {
    const int bx = i % gw, by = i / gw;
    const bool up = by > 0, dn = by < gh - 1, lf = bx > 0, rt = bx < gw - 1;
    const short2 c0 = O[i];
    const uint32_t k0 = pack_mv(c0);
    bool same = true;
    if (lf) same = same && pack_mv(O[i - 1]) == k0;
    if (rt) same = same && pack_mv(O[i + 1]) == k0;
    if (up) {
      same = same && pack_mv(O[i - gw]) == k0;
      if (lf) same = same && pack_mv(O[i - gw - 1]) == k0;
      if (rt) same = same && pack_mv(O[i - gw + 1]) == k0;
    }
    if (dn) {
      same = same && pack_mv(O[i + gw]) == k0;
      if (lf) same = same && pack_mv(O[i + gw - 1]) == k0;
      if (rt) same = same && pack_mv(O[i + gw + 1]) == k0;
    }
    if (same) a.Y[(size_t)pair * a.mv_plane + i] = c0;
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
__global__ void __launch_bounds__(128, TEAM <= 2 ? 4 : 1) k_reg_eval(RegArgs a) {
  constexpr int TEAMSZ = TEAM <= 2 ? 1 : TEAM;
  const int pair = blockIdx.y;
  uint32_t* ctr = a.ctr + (size_t)pair * kCtrWords;
  const uint32_t cnt = ctr[CTR_COUNT_EVAL];
  const uint32_t first = (uint32_t)(blockIdx.x * blockDim.x) / TEAMSZ;
  if (first >= cnt) return;  // the grid is sized for "every block listed"
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t e = t / TEAMSZ;
  const int tl = (int)(t % TEAMSZ);
  const bool live = e < cnt;
  if (!live) {
    if (TEAMSZ > 1) return;  // whole teams leave together
    e = cnt - 1;             // one thread per block: keep the warp converged
  }
  const uint32_t team_mask = TEAMSZ >= 32 ? 0xffffffffu : (((1u << TEAMSZ) - 1u) << ((threadIdx.x & 31) / TEAMSZ * TEAMSZ));
  const short2* O = a.O + (size_t)pair * a.mv_plane;
  short2* Y = a.Y + (size_t)pair * a.mv_plane;
  const int i = (int)a.list1[(size_t)pair * a.wl_plane + e];
  const int bx = i % a.gw, by = i / a.gw;
  const short2 nv = reg_eval_any<TEAM>(a, pair, O, O, bx, by, tl, team_mask, live);
  if (tl != 0 || !live) return;
  Y[i] = nv;
  if (pack_mv(nv) != pack_mv(O[i])) {
    const uint32_t ep = ctr[CTR_EPOCH] + 1u;
    uint32_t* stamp = a.stamp + (size_t)pair * a.wl_plane;
    uint32_t* list = a.list0 + (size_t)pair * a.wl_plane;
    for_each_dependent(bx, by, a.gw, a.gh, [&](int d) {
      if (atomicExch(&stamp[d], ep) != ep) list[atomicAdd(&ctr[CTR_COUNT0], 1u)] = (uint32_t)d;
    });
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
// epoch + 2 + r) and clears counter (r + 2) % 3 for the round after.  The first rounds are the big ones and run as
// grid-wide kernels over all pairs (k_reg_round); the tail, where rounds are short and latency-bound, runs as one
// CTA per pair that loops until its list is empty (k_reg_fix).
__device__ __forceinline__ int ctr_index(int k) { return k == 2 ? CTR_COUNT2 : k; }

template <int TEAM>
__global__ void __launch_bounds__(256, TEAM <= 2 ? 2 : 1) k_reg_round(RegArgs a, int r) {
  constexpr int TEAMSZ = TEAM <= 2 ? 1 : TEAM;
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
  const uint32_t limit = TEAMSZ == 1 ? ((cnt + 31u) & ~31u) : cnt;
  for (uint32_t e = team; e < limit; e += nteams) {
    const bool live = e < cnt;
    const int b = (int)lc[live ? e : cnt - 1];
    const int bx = b % a.gw, by = b / a.gw;
    const short2 nv = reg_eval_any<TEAM>(a, pair, O, Y, bx, by, (int)tl, team_mask, live);
    if (tl == 0 && live && pack_mv(nv) != pack_mv(Y[b])) {
      Y[b] = nv;
      for_each_dependent(bx, by, a.gw, a.gh, [&](int d) {
        if (atomicExch(&stamp[d], ep) != ep) ln[atomicAdd(next_count, 1u)] = (uint32_t)d;
      });
    }
  }
}

template <int TEAM>
__global__ void __launch_bounds__(TEAM <= 2 ? 512 : 1024) k_reg_fix(RegArgs a, int r0) {
  constexpr int TEAMSZ = TEAM <= 2 ? 1 : TEAM;
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
    // one thread per block: whole warps iterate together (the evaluator's early-out is warp-uniform)
    const uint32_t limit = TEAMSZ == 1 ? ((cnt + 31u) & ~31u) : cnt;
    for (uint32_t e = team; e < limit; e += nteams) {
      const bool live = e < cnt;
      const int b = (int)lc[live ? e : cnt - 1];
      const int bx = b % a.gw, by = b / a.gw;
      const short2 nv = reg_eval_any<TEAM>(a, pair, O, Y, bx, by, (int)tl, team_mask, live);
      if (tl == 0 && live && pack_mv(nv) != pack_mv(Y[b])) {
        Y[b] = nv;
        for_each_dependent(bx, by, a.gw, a.gh, [&](int d) {
          if (atomicExch(&stamp[d], ep + 1u) != ep + 1u) ln[atomicAdd(next_count, 1u)] = (uint32_t)d;
        });
      }
    }
    __syncthreads();
    blocks += cnt;
    cnt = *next_count;
    if (threadIdx.x == 0) s_next[cur] = 0;  // becomes the append counter of the round after next
    cur ^= 1;
    ++ep;
    ++rounds;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
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
  k_reg_classify<<<dim3((unsigned)((nb + 255) / 256), n), 256, 0, s>>>(a);
  dim3 grid((unsigned)((nb * lanes + 127) / 128), n);
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
    case 32: k_reg_fix<32><<<n, 1024, 0, s>>>(a, r0); break;
    case 16: k_reg_fix<16><<<n, 1024, 0, s>>>(a, r0); break;
    case 8: k_reg_fix<8><<<n, 1024, 0, s>>>(a, r0); break;
    case 2: k_reg_fix<2><<<n, 512, 0, s>>>(a, r0); break;
    default: k_reg_fix<1><<<n, 512, 0, s>>>(a, r0); break;
  }
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
