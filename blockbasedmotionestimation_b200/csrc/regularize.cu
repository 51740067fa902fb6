// regularize.cu -- the exact regularisation schedule of one pyramid level in one launch.
//
// Reference semantics (file:line of /root/reference): MF::regularize_MVs / find_min_candidate / calculate_smoothness /
// min_energy_candidate (motion_framework.cpp:424-662) inside the schedule of :133-154 (per block size `sweeps` sweeps with
// lambda_multiplier 1..sweeps, then divide_blocks :845-862, lambda *= 2).  Candidate slots in the reference's order
// [C, L, R, DR, UL, UR, U, D, DL]; neighbours outside the grid are dropped (the reference's nine-way if/else chain, :438-522,
// is exactly that).  The reference sweeps IN PLACE in raster order: L, UL, U, UR have already been updated ("pred"
// neighbours, read from the new field P) and C, R, DR, D, DL have not (read from the old field O).  That result is the
// unique fixed point of the update map (triangular in raster order) and is computed here by chaotic iteration: evaluate,
// and whenever a block changes re-enqueue the blocks that read it as a pred neighbour.
//
// Energy (:607) is float32 and un-fused: (float)SAD + ((lambda * (float)mult) * S); S is a sum of integer-valued floats
// (< 2^24, exact), so it is accumulated in int and converted once; out-of-image candidates get FLT_MAX (:578-582); the
// first index with a strictly smaller energy wins (:653-659).
#include "kernels.h"

#include <cooperative_groups.h>
#include <float.h>
#include <math.h>
#include <stdlib.h>

namespace bbme {


// A block that changed invalidates the evaluations of the blocks that read it as a "pred" neighbour: its right,
// lower-left, lower and lower-right neighbours.  They are appended to the next round's work list, de-duplicated by an
// epoch stamp per block.  Called by all 32 lanes of a warp together: the four stamp exchanges of a lane are issued
// back to back (independent atomics, one round trip), and the list slots of the whole warp are reserved with ONE
// atomicAdd (a per-entry atomicAdd on the pair's counter serialises in the L2).
// Split in two so that the stamp exchanges of one block are in flight while the next block is evaluated: push_issue sends the
// atomics, push_commit (an iteration later) looks at what they returned and appends.
struct PushState {
  uint32_t old[4];  // what the stamps held; == ep: nothing to append for that dependent
  uint32_t pb;      // packed entry (by << 16 | bx) of the block that changed: its dependents are pb + 1, pb + 0x10000 - 1, ...
};

__device__ __forceinline__ void push_issue(PushState& ps, bool changed, int bx, int by, int gw, int gh, uint32_t* stamp,
                                           uint32_t ep) {
  const bool rt = bx + 1 < gw, dn = by + 1 < gh;
  const int b = by * gw + bx;
  const int d[4] = {b + 1, b + gw - 1, b + gw, b + gw + 1};
  const bool ex[4] = {changed && rt, changed && dn && bx > 0, changed && dn, changed && dn && rt};
#pragma unroll
  for (int j = 0; j < 4; ++j) ps.old[j] = ex[j] ? atomicExch(&stamp[d[j]], ep) : ep;
  ps.pb = ((uint32_t)by << 16) | (uint32_t)bx;
}

// List entries are packed block coordinates (by << 16 | bx; grids are at most 8192 wide): no integer division anywhere.
__device__ __forceinline__ void push_commit(const PushState& ps, uint32_t ep, uint32_t* list, uint32_t* count) {
  const int lane = threadIdx.x & 31;
  int k = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) k += (ps.old[j] != ep) ? 1 : 0;
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
  const uint32_t d[4] = {ps.pb + 1u, ps.pb + 0x10000u - 1u, ps.pb + 0x10000u, ps.pb + 0x10000u + 1u};
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (ps.old[j] != ep) list[pos++] = d[j];
}

__device__ __forceinline__ void push_dependents(bool changed, int bx, int by, int gw, int gh, uint32_t* stamp,
                                                uint32_t ep, uint32_t* list, uint32_t* count) {
  PushState ps;
  push_issue(ps, changed, bx, by, gw, gh, stamp, ep);
  push_commit(ps, ep, list, count);
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
                                             uint32_t (&pk)[8], uint32_t& cur) {
  const int gw = a.gw, gh = a.gh;
  const int idx = by * gw + bx;
  const bool up = by > 0, dn = by < gh - 1, lf = bx > 0, rt = bx < gw - 1;
  const uint32_t* Ou = reinterpret_cast<const uint32_t*>(O);
  const uint32_t* Pu = reinterpret_cast<const uint32_t*>(P);
  A0 = Ou[idx];
  cur = Pu[idx];  // the block's value in the new field: only its own evaluation writes it, and a round lists a block once
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
#pragma unroll
    for (int i = 0; i < 8; ++i) {  // unrolled: pkin / has stay in registers (a rolled loop would index them in local memory)
      if (has[i]) {
        const uint32_t v = pkin[i];
        int j = 0;
        while (j < nd && s_u[j * blockDim.x] != v) ++j;
        if (j == nd) { s_u[nd * blockDim.x] = v; ++nd; }
        M += 1ull << (4 * j);
      }
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
#define BBME_LEVEL_THREADS 640  // threads per CTA of the level kernel (one CTA per SM, 102 registers): 512 and 768 measured 1-5 % slower, 1024 (64 registers, spills) 20 %
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
    // (Tried: issuing the next strip's six loads before this strip is examined -- slower, the registers it takes cost more than
    // the doubled bytes in flight bring.)
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
        if ((work >> q) & 1u) *dst++ = ((uint32_t)(by0 + (q >> 2)) << 16) | (uint32_t)(bx + (q & 3));
    }
  } else {
    const uint32_t nb = (uint32_t)gw * gh;
    const uint32_t limit = (nb + 31u) / 32u * 32u;
    for (uint32_t t = lc.gtid; t < limit; t += lc.gthreads) {
      bool work = false;
      uint32_t packed = 0;
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
        packed = ((uint32_t)by << 16) | (uint32_t)bx;
      }
      const uint32_t m = __ballot_sync(0xffffffffu, work);
      if (m) {
        const int leader = __ffs(m) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&lc.cnt[0], (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (work) list[base + __popc(m & ((1u << lane) - 1u))] = packed;
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
        const uint32_t pb = lcur[live ? e : cnt - 1];
        const int bx = (int)(pb & 0xffffu), by = (int)(pb >> 16);
        const int b = by * a.gw + bx;
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
      uint32_t A1 = 0, pk1[8], cur1 = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) pk1[i] = 0;
      if (r1 < run_limit) small_gather(a, O, Y, b1 & 0xffff, b1 >> 16, A1, pk1, cur1);
      PushState ps;  // the previous iteration's stamp exchanges, committed after this iteration's evaluation
      ps.pb = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) ps.old[j] = ep;
      while (r1 < run_limit) {
        const bool live = r1 * K + k1 < cnt;
        const int b = b1;
        const uint32_t A0 = A1, cur = cur1;
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = pk1[i];
        // shift the pipeline: b1 <- b2, b2 <- the entry after it
        b1 = b2; r1 = r2; k1 = k2;
        advance(r2, k2);
        b2 = load_entry(r2, k2);
        if (r1 < run_limit) small_gather(a, O, Y, b1 & 0xffff, b1 >> 16, A1, pk1, cur1);
        const int bx = b & 0xffff, by = b >> 16;  // list entries are packed coordinates
        const int idx = by * a.gw + bx;
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
        const bool changed = live && done && nv != cur;
        if (changed) Yu[idx] = nv;
        if (live && done && b1 == b + 1 && bx + 1 < a.gw) pk1[0] = nv;  // the next block's left neighbour is this block
        push_commit(ps, ep, lnext, next_count);                     // the previous block's dependents
        push_issue(ps, changed, bx, by, a.gw, a.gh, stamp, ep);       // this block's stamp exchanges fly during the next evaluation
      }
      push_commit(ps, ep, lnext, next_count);
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
        const uint32_t pb = dlist[live ? e : dcnt - 1];
        const int bx = (int)(pb & 0xffffu), by = (int)(pb >> 16);
        const int b = by * a.gw + bx;
        uint32_t A0, pk[8], cur, nv = 0;
        small_gather(a, O, Y, bx, by, A0, pk, cur);
        if (live) reg_eval_thread<BSK>(a, pair, bx, by, A0, pk, &nv, true, s_u);
        const bool changed = live && nv != cur;
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

}  // namespace bbme
