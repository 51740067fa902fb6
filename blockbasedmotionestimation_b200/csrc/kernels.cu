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
__device__ __forceinline__ short2 reg_eval(const RegArgs& a, int pair, const short2* __restrict__ O,
                                           const short2* P, int bx, int by) {
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
      const uint32_t sad = sad_block_unaligned(blk, ref + (size_t)py * pitch + px, pitch, bs);
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

// Pass 1 of a sweep: every block evaluated with the OLD field for all nine slots (a Jacobi step).  Blocks
// whose value changed enqueue their dependents: those may have used a stale "pred" value.
__global__ void __launch_bounds__(256) k_reg_full(RegArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int pair = blockIdx.y;
  if (i >= a.gw * a.gh) return;
  const int bx = i % a.gw, by = i / a.gw;
  const short2* O = a.O + (size_t)pair * a.mv_plane;
  short2* Y = a.Y + (size_t)pair * a.mv_plane;
  uint32_t* ctr = a.ctr + (size_t)pair * kCtrWords;
  const short2 nv = reg_eval(a, pair, O, O, bx, by);
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

// Passes 2..: one CTA per pair iterates Jacobi rounds on the active set until nothing changes.  The update
// map is triangular in raster order (a block depends on earlier blocks' NEW values and later blocks' OLD
// values only), so the fixed point is unique and equals the reference's in-place raster sweep.
__global__ void __launch_bounds__(1024) k_reg_fix(RegArgs a) {
  const int pair = blockIdx.x;
  const short2* O = a.O + (size_t)pair * a.mv_plane;
  short2* Y = a.Y + (size_t)pair * a.mv_plane;
  uint32_t* ctr = a.ctr + (size_t)pair * kCtrWords;
  uint32_t* stamp = a.stamp + (size_t)pair * a.wl_plane;
  uint32_t* lists[2] = {a.list0 + (size_t)pair * a.wl_plane, a.list1 + (size_t)pair * a.wl_plane};
  short2* nvb = a.nv + (size_t)pair * a.wl_plane;
  __shared__ uint32_t s_next;
  uint32_t cnt = ctr[CTR_COUNT0];
  if (cnt == 0) return;  // the Jacobi pass was already the raster result
  uint32_t ep = ctr[CTR_EPOCH] + 1u;
  uint32_t rounds = 0, blocks = 0;
  int cur = 0;
  while (cnt > 0) {
    if (threadIdx.x == 0) s_next = 0;
    __syncthreads();
    const uint32_t* lc = lists[cur];
    uint32_t* ln = lists[cur ^ 1];
    for (uint32_t e = threadIdx.x; e < cnt; e += blockDim.x) {
      const int b = (int)lc[e];
      nvb[e] = reg_eval(a, pair, O, Y, b % a.gw, b / a.gw);
    }
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < cnt; e += blockDim.x) {
      const int b = (int)lc[e];
      const short2 nv = nvb[e];
      if (pack_mv(nv) != pack_mv(Y[b])) {
        Y[b] = nv;
        for_each_dependent(b % a.gw, b / a.gw, a.gw, a.gh, [&](int d) {
          if (atomicExch(&stamp[d], ep + 1u) != ep + 1u) ln[atomicAdd(&s_next, 1u)] = (uint32_t)d;
        });
      }
    }
    __syncthreads();
    blocks += cnt;
    cnt = s_next;
    cur ^= 1;
    ++ep;
    ++rounds;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    ctr[CTR_COUNT0] = 0;
    ctr[CTR_COUNT1] = 0;
    ctr[CTR_EPOCH] = ep;
    ctr[CTR_ROUNDS] += rounds;
    ctr[CTR_BLOCKS] += blocks;
  }
}

void launch_reg_full(const RegArgs& a, int n, cudaStream_t s) {
  dim3 grid((a.gw * a.gh + 127) / 128, n);
  k_reg_full<<<grid, 128, 0, s>>>(a);
}

void launch_reg_fix(const RegArgs& a, int n, cudaStream_t s) { k_reg_fix<<<n, 1024, 0, s>>>(a); }


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
