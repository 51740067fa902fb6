// search_tma.cu -- exhaustive SAD block search for sm_100a: TMA-staged windows, VABSDIFF4 inner loop.
//
// What it computes: MF::calcLevelBM + find_min_block_spiral (reference motion_framework.cpp:226-244,296-422)
// for every block of every pair of a chunk: all (2R+1)^2 displacements around the predicted position whose
// block lies inside the image, minimum SAD, ties to the earliest position of the reference's spiral walk.
//
// How (B200-first, see DESIGN.md "search kernel"):
//  * VABSDIFF4.U8.ACC issues at 64 lanes/clk/SM and shares its pipe with SHF/PRMT/LOP3 (bench_micro/int_peak.cu); IMAD.HI and
//    IMAD.WIDE cost the same slot (bench_micro/shift_pipe.cu).  The inner loop is therefore SADs and shared loads only.
//  * Byte alignment: displacement dx shifts the window by single bytes, the SADs work on aligned words, and TMA tile loads need a
//    16-byte-aligned innermost coordinate (an unaligned one raises an illegal-instruction fault -- scripts/tma_probe.cu).
//    PRE instantiations: image 2 exists in four copies shifted by 0..3 bytes (k_shift4, one HBM pass per level); one 4-D tile
//    load stages a window as [row][byte phase][pitch] and a lane reads the phase of its column -- no shift instruction at all.
//    The other instantiations (64-bit keys, pitch classes whose phases would collide in the banks, windows too large for four
//    copies) stage the window once from the aligned origin and funnel-shift TWW+1 words of a row into TWW (13 % of the pipe).
//  * One warp is the TMA producer (a ring of 5 / 8 / 16 stages, mbarrier full/empty); it takes blocks from a grid-wide counter,
//    two ahead, and sleeps between polls of a full ring.  Sixteen consumer warps pull 32-lane work items from a shared counter
//    (balances the four SM sub-partitions without CTA barriers).
//  * A lane owns one displacement column dx and SEG consecutive dy: the 16x16 (or 8x8) block tile lives in 64
//    (16) registers, SEG accumulators in registers, every window word loaded once per lane feeds up to 16 SADs.
//  * The 32 lanes of a work item take consecutive dx, i.e. consecutive bytes: 8-9 distinct consecutive words per
//    shared load (PRE: per byte phase, in disjoint bank groups).
//  * Block results are reduced with a 32-bit key (SAD << KS | spiral rank); the ranks come from a table built once
//    per CTA in shared memory, so the per-candidate epilogue is one 16-bit shared load, one IMAD and half a
//    three-input min; warps reduce with redux.sync.min and one shared atomicMin per work item.
//  * Large search ranges (K64 instantiations): when (2R+1)^2 does not fit the 32-bit key, the key is 64 bit
//    (SAD << 32 | rank) and ranks are computed arithmetically; when the window is wider than one 256-byte TMA box it
//    is staged as two boxes that overlap by 16 bytes and every lane picks the box that holds its five words.
#include "kernels.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace bbme {

// CTA shape: one producer warp + kConsumerWarps consumer warps over a ring of kStages stages, kMinCtas CTAs per SM
// (register cap = 64 K / (kMinCtas * kThreads)).  The macros exist for tuning experiments: one CTA of
// 1 + 16 warps over a five-stage ring measured 71.6 % of the integer peak on config 2, two CTAs of 1 + 8 warps over
// three stages each 68.8 %; 18-19 consumer warps or 6-8 stages change nothing, 20 warps spill.
#ifndef BBME_STAGES
#define BBME_STAGES 5
#endif
#ifndef BBME_CW
#define BBME_CW 16
#endif
#ifndef BBME_MINB
#define BBME_MINB 1
#endif
constexpr int kStages = BBME_STAGES;   // minimum ring depth: the stage budget (and with it the band split) is sized for this many
constexpr int kMaxStages = 16;         // the ring takes as many stages as fit the CTA's shared memory, up to this (TmaSearchArgs::stages)
constexpr int kConsumerWarps = BBME_CW;
constexpr int kMinCtas = BBME_MINB;
constexpr int kThreads = 32 * (1 + kConsumerWarps);

struct TmaSearchArgs {
  int w, h;            // level size
  int gw, gh;          // blocks per row / column
  int n_pairs;
  int R, n;            // n = 2R+1
  int segs_total;      // ceil(n / SEG)
  int segs_per_band;
  int nbands;
  int band_rows;       // segs_per_band * SEG
  int wi_max;          // 32-lane work items per unit (band)
  int pww;             // window row pitch in words
  int win_bytes;       // bytes reserved for the window inside a stage (multiple of 128)
  int box_bytes;       // bytes of the window box
  int blk_bytes;       // bytes of the block box
  int stage_bytes;
  int rank_off;        // byte offset of the spiral-rank table (uint16, (n + SEG) rows of 4 * pww) in dynamic shared memory
  int box1_word;       // 0: one window box; else the word column where the second (overlapping) box starts
  int box1_off_words;  // word offset of the second box inside a stage
  int stage_shift;     // log2(stages) when stages is a power of two, 0 when stages == kStages
  uint32_t n_magic, n_shift;    // x / n == umulhi(x, n_magic) >> n_shift for 0 <= x < 2^31 (magic 0: divide)
  uint32_t iu_magic, iu_shift;  // the same for the lanes per unit
  unsigned int* work_ctr;       // zeroed before the launch: the next block to hand out (nullptr: blocks strided by CTA index)
  int stages;          // ring depth of this launch (kStages .. kMaxStages): small units (32x32 blocks with +-16: three work items
                       // per unit) need a deep ring to keep sixteen consumer warps fed, large windows only fit a shallow one
  short2* mv;
  size_t mv_plane;
  unsigned long long* counters;
};

struct StageMeta {
  int x2, y2;      // predicted position of the block in image 2
  int off;         // byte offset of the window's first column inside the staged (16-byte aligned) box
  int predx, predy;
  int valid;       // 0: prediction leaves the image -> MV 0, nothing staged (motion_framework.cpp:304-310)
  int band;
  int bslot;
  int gblk;        // global block index (pair * blocks + block)
  int unit;        // the CTA-local unit staged here (consumers check it: see wait_unit)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Waits are watched: a barrier that does not complete within ~2^24 polls (seconds; a healthy wait takes microseconds)
// traps, so a protocol error surfaces as a CUDA error instead of a hung process.  Build with -DBBME_DEBUG_HANG to have
// the stuck wait print where it is (the printf costs the kernel a stack frame, hence not in the normal build).
__device__ __forceinline__ void mbar_stuck(int who, int k, uint32_t parity) {
#ifdef BBME_DEBUG_HANG
  printf("bbme search kernel: barrier wait stuck (cta %d warp %d who %d unit %d parity %u)\n", (int)blockIdx.x,
         (int)(threadIdx.x >> 5), who, k, parity);
#else
  (void)who; (void)k; (void)parity;
#endif
  __trap();
}
// SLEEP_NS > 0: back off between polls.  The producer's ring is full nearly all the time; polled back to back its wait loop
// was 5 % of the kernel's executed instructions and took issue slots and ALU-pipe slots (VIADD, ISETP) from the consumer warps
// of its sub-partition (profiles/r02_search_l0_ncu.txt).
template <int SLEEP_NS = 0>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int who = 0, int k = 0) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  uint32_t polls = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!ok) {
      if (++polls > (1u << (SLEEP_NS ? 22 : 24))) mbar_stuck(who, k, parity);
      if (SLEEP_NS) __nanosleep(SLEEP_NS);
    }
  }
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {  // one poll
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// inverse of spiral_rank: rank -> (dx, dy)
__device__ __forceinline__ void spiral_unrank(uint32_t rank, int& dx, int& dy) {
  dx = 0;
  dy = 0;
  if (rank == 0) return;
  int r = 1;
  while ((uint32_t)((2 * r + 1) * (2 * r + 1)) <= rank) ++r;
  const int o = (int)rank - (2 * r - 1) * (2 * r - 1);
  if (o < 2 * r) { dx = r; dy = o - r + 1; }
  else if (o < 4 * r) { dy = r; dx = r - 1 - (o - 2 * r); }
  else if (o < 6 * r) { dx = -r; dy = r - 1 - (o - 4 * r); }
  else { dy = -r; dx = (o - 6 * r) - r + 1; }
}

template <int BS, int SEG, int PWW, bool K64, bool DEEP, bool PRE>
__global__ void __launch_bounds__(kThreads, kMinCtas)
k_search_tma(const __grid_constant__ CUtensorMap map_win, const __grid_constant__ CUtensorMap map_blk,
             const TmaSearchArgs a) {
  constexpr int TW = BS >= 16 ? 16 : BS;   // tile width / height held in registers
  constexpr int TWW = TW / 4;              // tile words per row
  constexpr int QN = BS / TW;              // tiles per block side
  constexpr int AP = BS >= 16 ? BS : 16;   // staged block row pitch (TMA inner extent is >= 16 bytes)
  // PRE: the window image exists in four copies shifted by 0..3 bytes (k_shift4 below), staged by ONE 4-D tile load as
  // [row][copy][PWW words]; a lane reads the copy of its column's byte phase and needs no funnel shift.  The byte alignment
  // is 13 % of the ALU-pipe instructions of an item and only SHF/PRMT (ALU pipe) or IMAD.HI/IMAD.WIDE (which cost the same
  // slot: bench_micro/shift_pipe.cu) can do it in the loop.  Copy stride PWW = 24 or 40 words puts the four byte phases of
  // eight consecutive words into 32 different banks.
  constexpr int RP = PRE ? 4 * PWW : PWW;  // words between consecutive window rows in a stage

  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int SMAX = DEEP ? kMaxStages : kStages, BMAX = SMAX + 1;
  __shared__ __align__(8) uint64_t s_full[SMAX];
  __shared__ __align__(8) uint64_t s_empty[SMAX];
  __shared__ StageMeta s_meta[SMAX];
  __shared__ uint32_t s_sdone[SMAX];
  __shared__ uint32_t s_bkey[BMAX];
  __shared__ unsigned long long s_bkey64[BMAX];
  __shared__ uint32_t s_bdone[BMAX];
  __shared__ uint32_t s_bbusy[BMAX];
  // Ring depth: kStages (compile time) or, in the DEEP instantiations, 8 or 16 stages (shift and mask).  Sixteen consumer warps
  // want ~16 work items ready; where a unit holds only a few items (32x32 blocks with +-16: three) five stages starve them
  // (58 % of the integer peak, long-scoreboard stalls), sixteen do not (73 %).  Large-window geometries keep the five stages
  // they were tuned with -- and the compile-time ring arithmetic: a run-time depth in every instantiation cost config 2 1 %.
  const int NS = DEEP ? a.stages : kStages, NB = NS + 1, nshift = a.stage_shift;
  auto ring_slot = [&](int k) -> int { return DEEP ? (k & (NS - 1)) : (k % kStages); };
  auto ring_turn = [&](int k) -> int { return DEEP ? (k >> nshift) : (k / kStages); };
  __shared__ uint32_t s_next;
  __shared__ int s_end;  // number of units this CTA stages, published by the producer after the last one (INT_MAX before)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblocks = a.gw * a.gh;
  const int total_blocks = nblocks * a.n_pairs;
  const int G = gridDim.x, cta = blockIdx.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < SMAX; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 1);
      s_sdone[i] = 0;
      s_meta[i].unit = -1;  // shared memory keeps the previous CTA's values: a stale unit number must not match
    }
    for (int i = 0; i < BMAX; ++i) {
      s_bkey[i] = 0xffffffffu;
      s_bkey64[i] = ~0ull;
      s_bdone[i] = 0;
      s_bbusy[i] = 0;
    }
    s_next = 0;
    s_end = 0x7fffffff;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  // spiral visit rank of every displacement, [dy + R][dx + R] with a compile-time row pitch of NP >= n entries (the epilogue's
  // thirteen rank loads then use immediate offsets instead of a chain of address adds); rows and columns past n hold 0xffff
  constexpr int KS = BS >= 32 ? 14 : 16;  // key = SAD << KS | rank; SAD < 2^(32-KS), rank < 2^KS (checked on the host)
  constexpr int NP = 4 * PWW;             // n <= 4 * PWW - BS - 15 (the window of n + BS - 1 + 15 bytes fits the pitch)
  uint16_t* s_rank = reinterpret_cast<uint16_t*>(smem + a.rank_off);
  if (!K64) {
    for (int i = threadIdx.x; i < (a.n + SEG) * NP; i += blockDim.x) {
      const int ry = i / NP, rx = i - ry * NP;
      s_rank[i] = (ry < a.n && rx < a.n) ? (uint16_t)spiral_rank(rx - a.R, ry - a.R) : (uint16_t)0xffffu;
    }
  }
  __syncthreads();

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      // Blocks are handed out by a grid-wide counter: a CTA that starts late -- its SM was still running another stream's
      // kernel, the normal case with several chunks in flight -- takes fewer blocks instead of holding the launch open for a
      // fixed share (128 pairs as 4 x 32 on four streams: 4 360 -> 4 420 pairs/s).
      const bool dyn = a.work_ctr != nullptr;
      // two blocks ahead for the counter, one ahead for the block's predicted vector: both latencies (an L2 round trip each)
      // stay off the staging path -- with 32x32 blocks and +-16 a block is consumed in ~3 us
      auto fetch = [&](int prev) -> int { return dyn ? (int)atomicAdd(a.work_ctr, 1u) : prev + G; };
      auto load_pred = [&](int g) -> short2 {
        if (g >= total_blocks) return make_short2(0, 0);
        const int pr = g / nblocks;
        return a.mv[(size_t)pr * a.mv_plane + (g - pr * nblocks)];
      };
      int g0 = dyn ? fetch(0) : cta;
      int g1 = fetch(g0);
      short2 p0 = load_pred(g0);
      short2 pred = make_short2(0, 0);
      int gblk = 0, lb = -1, band = a.nbands - 1;
      for (int k = 0;; ++k) {
        if (++band == a.nbands) {  // next block
          band = 0;
          ++lb;
          gblk = g0;
          pred = p0;
          if (gblk >= total_blocks) {
            *reinterpret_cast<volatile int*>(&s_end) = k;
            break;
          }
          g0 = g1;
          g1 = fetch(g1);
          p0 = load_pred(g0);
        }
        const int stage = ring_slot(k);
        if (k >= NS) mbar_wait<400>(&s_empty[stage], (uint32_t)(ring_turn(k) - 1) & 1u, 1, k);
        if (band == 0) {
          // units complete out of order, so the block that used this key slot NB blocks ago may still be in
          // flight (all its units are already staged, so it will finish without this producer)
          volatile uint32_t* busy = &s_bbusy[lb % NB];
          for (uint32_t polls = 0; *busy != 0u; __nanosleep(32))
            if (++polls > (1u << 24)) mbar_stuck(2, k, 0);
          *busy = 1u;
        }
        const int pair = gblk / nblocks, b = gblk - pair * nblocks;
        const int by = b / a.gw, bx = b - by * a.gw;
        const int x2 = bx * BS + pred.x, y2 = by * BS + pred.y;
        const int valid = !(x2 < 0 || y2 < 0 || x2 + BS > a.w || y2 + BS > a.h);
        const int wx = x2 - a.R;
        const int wx_al = wx & ~15;  // floor to a multiple of 16 (two's complement, also for negative wx)
        StageMeta m;
        m.x2 = x2; m.y2 = y2; m.off = wx - wx_al; m.predx = pred.x; m.predy = pred.y;
        m.valid = valid; m.band = band; m.bslot = lb % NB; m.gblk = gblk; m.unit = k;
        s_meta[stage] = m;
        if (valid) {
          uint8_t* st = smem + (size_t)stage * a.stage_bytes;
          mbar_arrive_expect_tx(&s_full[stage], (uint32_t)((PRE ? 4 : a.box1_word ? 2 : 1) * a.box_bytes + a.blk_bytes));
          const int wy = y2 - a.R + band * a.band_rows;
          if (PRE) tma_load_4d(st, &map_win, &s_full[stage], wx_al, 0, wy, pair);
          else tma_load_3d(st, &map_win, &s_full[stage], wx_al, wy, pair);
          if (!PRE && a.box1_word) tma_load_3d(st + (size_t)a.box1_off_words * 4, &map_win, &s_full[stage], wx_al + 4 * a.box1_word, wy, pair);
          tma_load_3d(st + a.win_bytes, &map_blk, &s_full[stage], (bx * BS) & ~15, by * BS, pair);
        } else {
          mbar_arrive(&s_full[stage]);
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers
  // One flat lane space per CTA: unit k owns lanes [k * IU, (k + 1) * IU), IU = n * segs_per_band (one lane = one
  // displacement column x SEG rows).  A 32-lane work item is any aligned run of 32 lanes, so it may straddle two
  // consecutive units (IU >= 32): no lane idles at unit boundaries.  Completion is counted in lanes.
  const int IU = max(a.n * a.segs_per_band, 32);  // tiny search ranges: pad the unit to one full item (lanes past n * segs idle)
  auto finish_unit = [&](int stage) {  // lane 0 of the warp that completed the unit's last lane
    const StageMeta m = s_meta[stage];
    s_sdone[stage] = 0;
    const uint32_t bd = atomicAdd(&s_bdone[m.bslot], 1u);
    if (bd == (uint32_t)a.nbands - 1u) {  // last band of the block
      __threadfence_block();
      uint32_t key;  // the winner's spiral rank
      if (K64) {
        key = (uint32_t)*reinterpret_cast<volatile unsigned long long*>(&s_bkey64[m.bslot]);
        s_bkey64[m.bslot] = ~0ull;
      } else {
        key = *reinterpret_cast<volatile uint32_t*>(&s_bkey[m.bslot]) & ((1u << KS) - 1u);
        s_bkey[m.bslot] = 0xffffffffu;
      }
      s_bdone[m.bslot] = 0;
      __threadfence_block();
      *reinterpret_cast<volatile uint32_t*>(&s_bbusy[m.bslot]) = 0u;
      const int pair = m.gblk / nblocks, b = m.gblk - pair * nblocks;
      short2 out = make_short2(0, 0);
      if (m.valid) {
        int dx, dy;
        spiral_unrank(key, dx, dy);
        out = make_short2((short)(m.predx + dx), (short)(m.predy + dy));
        if (a.counters) {
          const int nx = min(a.R, a.w - BS - m.x2) - max(-a.R, -m.x2) + 1;
          const int ny = min(a.R, a.h - BS - m.y2) - max(-a.R, -m.y2) + 1;
          atomicAdd(&a.counters[0], (unsigned long long)(nx * ny));
          atomicAdd(&a.counters[1], (unsigned long long)(nx * ny) * (unsigned long long)(BS * BS));
        }
      }
      a.mv[(size_t)pair * a.mv_plane + b] = out;
    }
    __threadfence_block();
    mbar_arrive(&s_empty[stage]);
  };

  // Items are handed out in order but finish out of order, so a warp can hold an item of a unit that is one or more
  // ring turns ahead of the unit its stage holds (with small search ranges every item is a unit of its own, and eight
  // warps take eight units of a three-stage ring at once).  A parity wait alone then passes on an older phase of the
  // same parity -- on a fresh barrier even on the "phase before the first".  The unit number in the stage's metadata
  // (written by the producer before it arms the barrier, hence visible once the barrier completes; -1 at start) tells
  // the turns apart.
  // Returns false when unit k does not exist: the producer has run out of blocks and published the number of units it staged.
  auto wait_unit = [&](int k) -> bool {
    const int st = ring_slot(k);
    const uint32_t par = (uint32_t)ring_turn(k) & 1u;
    for (uint32_t polls = 0;; ++polls) {
      const bool phase = mbar_try(&s_full[st], par);  // a failed try has already waited in hardware
      if (phase && *reinterpret_cast<volatile int*>(&s_meta[st].unit) == k) return true;
      if (*reinterpret_cast<volatile int*>(&s_end) <= k) return false;
      if (phase) __nanosleep(64);  // an older turn of the stage satisfies the parity: the try returns at once, do not spin on it
      if (polls > (1u << 24)) mbar_stuck(4, k, par);
    }
  };

  for (;;) {
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(&s_next, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    const int T0 = (int)t * 32;
    const int k0 = a.iu_magic ? (int)(__umulhi((uint32_t)T0, a.iu_magic) >> a.iu_shift) : T0 / IU;  // unit of lane 0
    if (!wait_unit(k0)) break;                           // past the CTA's last unit
    const int split = (k0 + 1) * IU - T0;                // lanes [0, split) belong to k0, the rest to k0 + 1
    const int k1 = (split < 32 && wait_unit(k0 + 1)) ? k0 + 1 : k0;

    const int T = T0 + lane;
    const bool second = lane >= split;
    const int kl = second ? k1 : k0;
    const int stage = ring_slot(kl);
    const StageMeta m = s_meta[stage];
    const int q = T - kl * IU;
    const int segs_here = min(a.segs_per_band, a.segs_total - m.band * a.segs_per_band);
    const int sidx_raw = a.n_magic ? (int)(__umulhi((uint32_t)q, a.n_magic) >> a.n_shift) : q / a.n;  // q >= 0
    const bool active = (!second || k1 != k0) && m.valid && sidx_raw < segs_here;
    const int sidx = active ? sidx_raw : 0, o = active ? q - sidx_raw * a.n : 0;
    uint32_t best = 0xffffffffu, best_rank = 0xffffffffu;  // K64: best = SAD, best_rank = its spiral rank
    uint32_t eqmask = 0;   // K64: the lane's candidates that attain `best`
    int rank_dx = 0, rank_dy0 = 0;

    if (__any_sync(0xffffffffu, active)) {
      const int bo = m.off + o;                 // byte column of this lane's displacement inside the staged box
      const uint32_t sh = (uint32_t)(bo & 3) * 8u;
      const int wi = bo >> 2;
      const int cy0 = sidx * SEG;
      const uint8_t* st = smem + (size_t)stage * a.stage_bytes;
      const uint32_t* win = reinterpret_cast<const uint32_t*>(st) + cy0 * RP + (PRE ? (bo & 3) * PWW : 0);
      const uint8_t* blk = st + a.win_bytes + (BS == 8 ? ((m.x2 - m.predx) & 8) : 0);

      uint32_t acc[SEG];
#pragma unroll
      for (int c = 0; c < SEG; ++c) acc[c] = 0u;

#pragma unroll 1
      for (int qi = 0; qi < QN * QN; ++qi) {
        const int qy = qi / QN, qx = qi - qy * QN;
        uint32_t A[TW][TWW];
#pragma unroll
        for (int y = 0; y < TW; ++y) {
          const uint8_t* ar = blk + (size_t)(qy * TW + y) * AP + qx * TW;
          if (TWW == 4) {
            const uint4 v = *reinterpret_cast<const uint4*>(ar);
            A[y][0] = v.x; A[y][1] = v.y; A[y][2 % TWW] = v.z; A[y][3 % TWW] = v.w;
          } else {
            const uint2 v = *reinterpret_cast<const uint2*>(ar);
            A[y][0] = v.x; A[y][1 % TWW] = v.y;
          }
        }
        // word column of this lane's tile row start; with two boxes, take the one that holds words wq .. wq + TWW
        const int wq = wi + qx * TWW;
        const uint32_t* wb = win + qy * TW * RP + ((!PRE && a.box1_word && wq > PWW - 1 - TWW) ? a.box1_off_words + wq - a.box1_word : wq);
#pragma unroll
        for (int jr = 0; jr < SEG + TW - 1; ++jr) {
          uint32_t raw[TWW + 1], wv[TWW];
#pragma unroll
          for (int kk = 0; kk < TWW + (PRE ? 0 : 1); ++kk) raw[kk] = wb[jr * RP + kk];  // compile-time offsets: no address arithmetic
#pragma unroll
          for (int kk = 0; kk < TWW; ++kk) wv[kk] = PRE ? raw[kk] : __funnelshift_r(raw[kk], raw[kk + 1], sh);
#pragma unroll
          for (int kk = 0; kk < TWW; ++kk) {
#pragma unroll
            for (int y = 0; y < TW; ++y) {
              const int c = jr - y;
              if (c >= 0 && c < SEG) acc[c] = sad4(A[y][kk], wv[kk], acc[c]);
            }
          }
        }
      }

      // lane-local argmin on key = SAD << KS | rank.  Lanes whose SEG candidates are all in bounds (nearly all
      // of them) take the branch-free path; the others mask candidate by candidate.
      const int dx = o - a.R;
      const int px = m.x2 + dx;
      const bool xok = active && px >= 0 && px + BS <= a.w;
      const int dyf = m.band * a.band_rows + cy0 - a.R;       // dy of candidate c = 0
      const int c_lo = max(0, -(m.y2 + dyf));                 // py >= 0
      const int c_hi = min(min(SEG - 1, a.R - dyf), a.h - BS - m.y2 - dyf);  // dy <= R and py + BS <= h
      if (K64) {
        // 64-bit key (SAD, spiral rank).  The closed-form rank costs ~25 instructions, so it is NOT computed per
        // candidate: the lane keeps its minimum SAD and the set of candidates attaining it; after the warp has reduced
        // the SADs, only the lanes that hold the item's minimum rank their (usually single) candidate.  (Tried: no bit set
        // here, the lanes holding the minimum going through their accumulators again after the reduction -- the longer live
        // range of the accumulators costs more than the 26 instructions saved: config 5 64.6 -> 63.0 %.)
        if (xok && c_lo == 0 && c_hi == SEG - 1) {  // every candidate of the lane in the image (nearly all lanes): no masking
#pragma unroll
          for (int c = 0; c < SEG; ++c) best = min(best, acc[c]);
#pragma unroll
          for (int c = 0; c < SEG; ++c) eqmask |= (acc[c] == best) ? (1u << c) : 0u;
        } else {
#pragma unroll
          for (int c = 0; c < SEG; ++c) {
            const bool ok = xok && c >= c_lo && c <= c_hi;
            best = min(best, ok ? acc[c] : 0xffffffffu);
          }
#pragma unroll
          for (int c = 0; c < SEG; ++c) {
            const bool ok = xok && c >= c_lo && c <= c_hi;
            eqmask |= (ok && acc[c] == best) ? (1u << c) : 0u;
          }
        }
        rank_dx = dx;
        rank_dy0 = dyf;
      } else {
        // (Tried for 8x8 blocks, where the two key instructions and the rank load per candidate are 13 % of an item: bare SAD
        // minima per lane and keys only in the lanes that hold the item's minimum after a warp reduction -- config 3 fell
        // from 65 % to 56 % of the integer peak: the 43 accumulators stay live across the reduction and the re-scan diverges.)
        const uint16_t* rk = s_rank + (m.band * a.band_rows + cy0) * NP + o;
        if (xok && c_lo == 0 && c_hi == SEG - 1) {
#pragma unroll
          for (int c = 0; c < SEG; ++c) best = min(best, (acc[c] << KS) + (uint32_t)rk[c * NP]);
        } else if (xok) {
#pragma unroll
          for (int c = 0; c < SEG; ++c) {
            const uint32_t key = (acc[c] << KS) + (uint32_t)rk[c * NP];
            best = (c >= c_lo && c <= c_hi) ? min(best, key) : best;
          }
        }
      }
    }
    // per-unit reduction: lanes of k0, then lanes of k1
    const int bslot0 = __shfl_sync(0xffffffffu, m.bslot, 0);
    const int bslot1 = __shfl_sync(0xffffffffu, m.bslot, 31);
    uint32_t b0 = __reduce_min_sync(0xffffffffu, second ? 0xffffffffu : best);
    uint32_t b1 = __reduce_min_sync(0xffffffffu, second ? best : 0xffffffffu);
    uint32_t r0 = 0, r1 = 0;
    if (K64) {  // among the lanes holding the minimum SAD, the smallest rank
      const uint32_t mine = second ? b1 : b0;
      if (best == mine && best != 0xffffffffu) {
        for (uint32_t mm = eqmask; mm; mm &= mm - 1u)
          best_rank = min(best_rank, spiral_rank(rank_dx, rank_dy0 + __ffs(mm) - 1));
      }
      r0 = __reduce_min_sync(0xffffffffu, (!second && best == b0) ? best_rank : 0xffffffffu);
      r1 = __reduce_min_sync(0xffffffffu, (second && best == b1) ? best_rank : 0xffffffffu);
    }
    if (lane == 0) {
      if (K64) {
        if (b0 != 0xffffffffu) atomicMin(&s_bkey64[bslot0], ((unsigned long long)b0 << 32) | r0);
        if (k1 != k0 && b1 != 0xffffffffu) atomicMin(&s_bkey64[bslot1], ((unsigned long long)b1 << 32) | r1);
      } else {
        if (b0 != 0xffffffffu) atomicMin(&s_bkey[bslot0], b0);
        if (k1 != k0 && b1 != 0xffffffffu) atomicMin(&s_bkey[bslot1], b1);
      }
      __threadfence_block();
      const int n0 = min(32, split);
      const uint32_t d0 = atomicAdd(&s_sdone[ring_slot(k0)], (uint32_t)n0);
      if (d0 + (uint32_t)n0 == (uint32_t)IU) finish_unit(ring_slot(k0));
      if (k1 != k0) {
        const int n1 = 32 - n0;
        const uint32_t d1 = atomicAdd(&s_sdone[ring_slot(k1)], (uint32_t)n1);
        if (d1 + (uint32_t)n1 == (uint32_t)IU) finish_unit(ring_slot(k1));
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ host side

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static const PFN_encodeTiled fn = [] {  // initialised once, thread-safe (contexts may be planned from several host threads)
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      return reinterpret_cast<PFN_encodeTiled>(p);
    return static_cast<PFN_encodeTiled>(nullptr);
  }();
  return fn;
}

static int encode_u8_3d(CUtensorMap* map, const uint8_t* base, int w, int h, int pitch, size_t plane, int n,
                        int box_w, int box_h) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return -1;
  cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)plane};
  cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

// The four byte phases of the window image as one 4-D tensor: dims (x, copy, y, pair) over the [pair][y][copy][pitch] array
// that k_shift4 writes; box = (box_w, 4, box_h, 1) lands in shared memory as [row][copy][box_w].
static int encode_u8_4copies(CUtensorMap* map, const uint8_t* base, int w, int h, int pitch, size_t plane, int n, int box_w, int box_h) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return -1;
  cuuint64_t dims[4] = {(cuuint64_t)w, 4u, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)pitch, (cuuint64_t)pitch * 4u, (cuuint64_t)plane * 4u};
  cuuint32_t box[4] = {(cuuint32_t)box_w, 4u, (cuuint32_t)box_h, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<uint8_t*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

// dst[pair][y][c][x] = src[pair][y][x + c], c = 0..3 (bytes past the row's pitch read as 0): HBM-bound, 16 bytes in and
// 64 bytes out per thread.  Columns >= w - c of copy c only ever feed candidates that leave the image, which are masked.
__global__ void __launch_bounds__(128) k_shift4(const uint8_t* __restrict__ src, int pitch, size_t plane, int h, uint8_t* __restrict__ dst) {
  const int x = (blockIdx.x * 128 + threadIdx.x) * 16;
  if (x >= pitch) return;
  const int y = blockIdx.y, pair = blockIdx.z;
  const uint8_t* row = src + (size_t)pair * plane + (size_t)y * pitch;
  const uint4 v = *reinterpret_cast<const uint4*>(row + x);
  const uint32_t nx = x + 16 < pitch ? *reinterpret_cast<const uint32_t*>(row + x + 16) : 0u;
  uint8_t* out = dst + ((size_t)pair * h + y) * 4 * (size_t)pitch + x;
  *reinterpret_cast<uint4*>(out) = v;
#pragma unroll
  for (int c = 1; c < 4; ++c) {
    uint4 o;
    o.x = __funnelshift_r(v.x, v.y, 8 * c);
    o.y = __funnelshift_r(v.y, v.z, 8 * c);
    o.z = __funnelshift_r(v.z, v.w, 8 * c);
    o.w = __funnelshift_r(v.w, nx, 8 * c);
    *reinterpret_cast<uint4*>(out + (size_t)c * pitch) = o;
  }
}

void launch_shift4(ImgView src, uint8_t* dst, int n, cudaStream_t s) {
  dim3 grid((src.pitch / 16 + 127) / 128, src.h, n);
  k_shift4<<<grid, 128, 0, s>>>(src.p, src.pitch, src.plane, src.h, dst);
}

static int pick_seg(int bs, int R, bool pre) {
  // Candidate rows per lane.  Cost of a 32-lane work item per lane, in ALU-pipe instructions: the funnel shifts of
  // SEG + T - 1 window rows, SEG * T * T/4 SADs, ~2 per candidate for the key, and a fixed part (item fetch, tile
  // load, reductions) that weighs most on 8x8 blocks, whose items hold a quarter of the SADs of a 16x16 item.
  // 16x16 / 32x32: 13 or 11 (17 and 22 were measured slower, with and without register spills -- 22 again in round 2 over the
  // byte-shifted copies with 15 consumer warps and 119 registers: 76 % against 83 % of the integer peak); 8x8 blocks also get
  // 26 and 43: config 3 (8x8, +-64) went from 42 % to 64 % of the integer peak with 43.
  const int n = 2 * R + 1;
  const int T = bs >= 16 ? 16 : bs;
  const int cands[4] = {13, 11, 26, 43};
  const int ncand = bs == 8 ? 4 : 2;
  if (const char* e = getenv("BBME_SEARCH_SEG")) {  // tuning runs
    const int v = atoi(e);
    if (v == 13 || v == 11 || (bs == 8 && (v == 26 || v == 43))) return v;
  }
  int best = 13;
  double best_eff = -1.0;
  for (int i = 0; i < ncand; ++i) {
    const int seg = cands[i];
    const int segs = (n + seg - 1) / seg;
    const int wi = (n * segs + 31) / 32;
    const double per_lane = (seg + T - 1) * (T / 4 + 0.2) + (double)seg * T * (T / 4) + 2.0 * seg + 200.0;
    const double eff = (double)n * n * T * (T / 4) / ((double)wi * 32.0 * per_lane);
    if (eff > best_eff) { best_eff = eff; best = seg; }
  }
  return best;
}

struct TmaGeom {
  TmaSearchArgs a;
  int seg;
  int k64;
  int deep;  // the DEEP instantiation (ring of 8 or 16 stages) runs this geometry
  int pre;   // the PRE instantiation (four byte-shifted copies of the window, no funnel shifts) runs this geometry
  int box_w, box_h;
  size_t smem;
};

constexpr size_t kSmemCap = 200 * 1024;     // dynamic shared memory of a CTA
constexpr size_t kSmemCapPre = 224 * 1024;  // PRE stages are four times as large: all an SM has (227 KB less the static variables)

static bool make_geom_impl(int w, int h, int bs, int R, bool pre, TmaGeom* g) {
  if (!(bs == 8 || bs == 16 || bs == 32)) return false;
  if (R < 1) return false;
  memset(g, 0, sizeof(*g));
  const int n = 2 * R + 1;
  // the 32-bit key holds SAD << KS | rank; larger rank spaces use the 64-bit-key instantiations (16x16 and 32x32 only)
  const bool k64 = n * n > (bs >= 32 ? (1 << 14) : (1 << 16)) - 1;
  if (k64 && bs == 8) return false;
  if ((long long)n * n > 0x7fffffffLL) return false;
  const int seg = pick_seg(bs, R, pre);
  // staged box: starts at the 16-byte aligned column at or below (x2 - R); a lane reads words
  // (off + o) >> 2 ... + bs/4 inclusive, with off <= 15 and o <= 2R
  const int words = ((15 + 2 * R) >> 2) + bs / 4 + 1;
  // the row pitch is a template parameter of the kernel: round up to the next instantiated class
  // (K64 kernels are instantiated for pitches 40 and 64 only)
  static const int kPitchClasses[6] = {16, 24, 32, 40, 48, 64};
  int pww = 0, two_box = 0;
  for (int i = 0; i < 6 && !pww; ++i)
    if (words <= kPitchClasses[i] && (!k64 || kPitchClasses[i] == 40 || kPitchClasses[i] == 64)) pww = kPitchClasses[i];
  if (!pww) {
    // wider than one 256-byte TMA box: two boxes of pitch pww overlapping by 4 words (K64 kernels only)
    for (int i = 0; i < 6 && !pww; ++i)
      if (2 * kPitchClasses[i] - 4 >= words && (kPitchClasses[i] == 40 || kPitchClasses[i] == 64)) pww = kPitchClasses[i];
    if (!pww || bs == 8) return false;
    two_box = 1;
  }
  const bool use_k64 = k64 || two_box;
  if (pre && (use_k64 || (pww != 24 && pww != 40))) return false;  // copy stride = pitch: only these two spread the banks
  const size_t cap = pre ? kSmemCapPre : kSmemCap;
  const int box_w = pww * 4;
  const int segs_total = (n + seg - 1) / seg;
  const int blk_bytes = bs * (bs >= 16 ? bs : 16);
  // per stage: 24 KB, or what the CTA's 200 KB leave per stage next to the spiral-rank table
  const size_t rank_bytes = use_k64 ? 0 : (((size_t)(n + seg) * (4 * pww) * 2 + 127) / 128) * 128;  // rows of 4 * pww entries
  size_t budget = (cap - rank_bytes) / kStages / 128 * 128;
  if (!pre && budget > 24 * 1024) budget = 24 * 1024;
  int spb = segs_total;
  size_t one_box = 0;
  for (;;) {
    const int box_h = spb * seg + bs - 1;
    one_box = (((size_t)box_h * box_w) + 127) / 128 * 128;
    const size_t win_bytes = pre ? (((size_t)box_h * box_w * 4) + 127) / 128 * 128 : one_box * (two_box ? 2 : 1);
    const size_t stage = win_bytes + ((blk_bytes + 127) / 128) * 128;
    if ((stage <= budget && box_h <= 256) || spb == 1) {
      if (stage > (pre ? budget : (size_t)64 * 1024) || box_h > 256) return false;
      g->box_h = box_h;
      g->a.win_bytes = (int)win_bytes;
      g->a.stage_bytes = (int)stage;
      break;
    }
    --spb;
  }
  g->seg = seg;
  g->k64 = use_k64 ? 1 : 0;
  g->box_w = box_w;
  g->a.w = w; g->a.h = h;
  g->a.gw = w / bs; g->a.gh = h / bs;
  g->a.R = R; g->a.n = n;
  g->a.segs_total = segs_total;
  g->a.segs_per_band = spb;
  g->a.nbands = (segs_total + spb - 1) / spb;
  g->a.band_rows = spb * seg;
  g->a.wi_max = (n * spb + 31) / 32;
  g->a.pww = box_w / 4;
  g->a.box_bytes = g->box_h * box_w;
  g->a.blk_bytes = blk_bytes;
  g->a.box1_word = two_box ? pww - 4 : 0;
  g->a.box1_off_words = two_box ? (int)(one_box / 4) : 0;
  // as deep a ring as the shared memory holds (the stage size above was chosen for kStages stages)
  {
    int st = (int)((cap - rank_bytes) / (size_t)g->a.stage_bytes);
    // sixteen consumer warps want ~16 work items ready: a deep ring where a unit holds few items (32x32 / +-16: three), the
    // kStages that the large-window geometries were tuned with elsewhere (config 2 measured the same at 5, 8 and 16)
    const int items_per_unit = (n * spb + 31) / 32;
    if (pre && items_per_unit * kStages < 32 && st < 8 && !getenv("BBME_SEARCH_PRE_FORCE")) return false;  // few items per unit want the deep ring more than the copies
    if (items_per_unit * kStages >= 32 || use_k64 || pww > 32) st = kStages;  // DEEP instantiations exist for pitch classes <= 32
    if (const char* e = getenv("BBME_SEARCH_STAGES")) st = atoi(e);             // tuning runs
    g->deep = (st >= 8 && !use_k64 && pww <= 32) ? 1 : 0;
    g->a.stages = !g->deep ? kStages : (st >= 16 ? 16 : 8);
    g->a.stage_shift = g->a.stages == 16 ? 4 : (g->a.stages == 8 ? 3 : 0);
  }
  g->a.rank_off = g->a.stages * g->a.stage_bytes;
  g->smem = (size_t)g->a.rank_off + rank_bytes;
  if (g->smem > cap) return false;  // ring + rank table must fit the CTA's shared memory: generic kernel instead
  g->pre = pre ? 1 : 0;
  return true;
}

static bool pre_enabled() {
  const char* e = getenv("BBME_SEARCH_PRE");  // tuning / A-B runs: 0 keeps the funnel-shift kernels
  return !(e && atoi(e) == 0);
}

static bool make_geom(int w, int h, int bs, int R, bool want_pre, TmaGeom* g) {
  if (want_pre && pre_enabled() && make_geom_impl(w, h, bs, R, true, g)) return true;
  return make_geom_impl(w, h, bs, R, false, g);
}

void tma_div_magic(unsigned d, uint32_t* m, uint32_t* sh) {
  unsigned lg = 0;
  while ((1ull << lg) < d) ++lg;
  const unsigned p = 31 + lg;
  *m = (uint32_t)(((1ull << p) + d - 1) / d);
  *sh = p - 32;
}

int tma_search_geometry(int w, int h, int bs, int R, int allow_copies, TmaGeomInfo* out) {
  TmaGeom g;
  memset(out, 0, sizeof(*out));
  if (!make_geom(w, h, bs, R, allow_copies != 0, &g)) return 0;
  out->planned = 1;
  out->copies = g.pre;
  out->deep_ring = g.deep;
  out->key64 = g.k64;
  out->rows_per_lane = g.seg;
  out->pitch_words = g.a.pww;
  out->stages = g.a.stages;
  out->stage_bytes = g.a.stage_bytes;
  out->smem_bytes = (int)g.smem;
  out->bands = g.a.nbands;
  out->segments_per_band = g.a.segs_per_band;
  out->box_w = g.box_w;
  out->box_h = g.box_h;
  out->two_boxes = g.a.box1_word != 0;
  out->lanes_per_unit = g.a.n * g.a.segs_per_band > 32 ? g.a.n * g.a.segs_per_band : 32;
  return 1;
}

int tma_search_wants_pre(int w, int h, int bs, int R) {
  TmaGeom g;
  return make_geom(w, h, bs, R, true, &g) && g.pre;
}

// One instantiation: a == nullptr prepares it (shared-memory opt-in, once per plan), else launches it.
template <int BS, int SEG, int PWW, bool K64, bool DEEP, bool PRE>
static int run_inst(const TmaSearchPlan& plan, const TmaSearchArgs* a, int grid, cudaStream_t s) {
  if (!a)
    return cudaFuncSetAttribute(k_search_tma<BS, SEG, PWW, K64, DEEP, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(PRE ? kSmemCapPre : kSmemCap)) == cudaSuccess ? 1 : -1;
  k_search_tma<BS, SEG, PWW, K64, DEEP, PRE><<<grid, kThreads, plan.smem_bytes, s>>>(plan.map_win, plan.map_blk, *a);
  return 1;
}

// Finds the instantiation of (block size, rows per lane, window pitch class, key width).  Returns 1 = done, 0 = there is none
// (the caller falls back to the generic kernel at plan time; never silently at launch time), -1 = CUDA error.
static int dispatch(const TmaSearchPlan& plan, int k64, int pww, int deep, const TmaSearchArgs* a, int grid, cudaStream_t s) {
  const int pre = plan.pre;
#define BBME_CASE(BS_, SEG_, PWW_) \
  if (!k64 && !deep && !pre && plan.bs == BS_ && plan.seg == SEG_ && pww == PWW_) return run_inst<BS_, SEG_, PWW_, false, false, false>(plan, a, grid, s);
#define BBME_CASED(BS_, SEG_, PWW_) /* small windows also exist with the deep ring */ \
  BBME_CASE(BS_, SEG_, PWW_) \
  if (!k64 && deep && !pre && plan.bs == BS_ && plan.seg == SEG_ && pww == PWW_) return run_inst<BS_, SEG_, PWW_, false, true, false>(plan, a, grid, s);
#define BBME_CASEP(BS_, SEG_, PWW_, DEEP_) /* pitch classes 24 and 40 also exist over byte-shifted copies */ \
  if (!k64 && deep == DEEP_ && pre && plan.bs == BS_ && plan.seg == SEG_ && pww == PWW_) return run_inst<BS_, SEG_, PWW_, false, DEEP_ != 0, true>(plan, a, grid, s);
#define BBME_CASE64(BS_, SEG_, PWW_) \
  if (k64 && plan.bs == BS_ && plan.seg == SEG_ && pww == PWW_) return run_inst<BS_, SEG_, PWW_, true, false, false>(plan, a, grid, s);
#define BBME_CASES(BS_, SEG_) \
  BBME_CASED(BS_, SEG_, 16) BBME_CASED(BS_, SEG_, 24) BBME_CASED(BS_, SEG_, 32) BBME_CASE(BS_, SEG_, 40) \
  BBME_CASE(BS_, SEG_, 48) BBME_CASE(BS_, SEG_, 64) \
  BBME_CASEP(BS_, SEG_, 24, 0) BBME_CASEP(BS_, SEG_, 24, 1) BBME_CASEP(BS_, SEG_, 40, 0)
  BBME_CASES(8, 13) BBME_CASES(8, 11) BBME_CASES(8, 26) BBME_CASES(8, 43) BBME_CASES(16, 13) BBME_CASES(16, 11) BBME_CASES(32, 13) BBME_CASES(32, 11)
  BBME_CASE64(16, 13, 40) BBME_CASE64(16, 13, 64) BBME_CASE64(16, 11, 40) BBME_CASE64(16, 11, 64)
  BBME_CASE64(32, 13, 40) BBME_CASE64(32, 13, 64) BBME_CASE64(32, 11, 40) BBME_CASE64(32, 11, 64)
#undef BBME_CASE64
#undef BBME_CASEP
#undef BBME_CASES
#undef BBME_CASED
#undef BBME_CASE
  return 0;
}

int tma_search_plan(TmaSearchPlan* plan, const uint8_t* img1, const uint8_t* img2, const uint8_t* img2_shift4, int w, int h,
                    int pitch, size_t plane, int n_planes, int bs, int R, char* err, size_t errlen) {
  memset(plan, 0, sizeof(*plan));
  TmaGeom g;
  bool want_pre = img2_shift4 != nullptr;
again:
  if (!make_geom(w, h, bs, R, want_pre, &g)) return 0;  // not supported: caller uses the generic kernel
  if (!get_encode()) {
    if (err) snprintf(err, errlen, "cuTensorMapEncodeTiled entry point not available");
    return -1;
  }
  if ((g.pre ? encode_u8_4copies(&plan->map_win, img2_shift4, w, h, pitch, plane, n_planes, g.box_w, g.box_h)
             : encode_u8_3d(&plan->map_win, img2, w, h, pitch, plane, n_planes, g.box_w, g.box_h)) != 0 ||
      encode_u8_3d(&plan->map_blk, img1, w, h, pitch, plane, n_planes, bs >= 16 ? bs : 16, bs) != 0) {
    if (err) snprintf(err, errlen, "cuTensorMapEncodeTiled failed (w=%d h=%d pitch=%d box=%dx%d)", w, h, pitch, g.box_w, g.box_h);
    return -1;
  }
  plan->bs = bs;
  plan->R = R;
  plan->pre = g.pre;
  plan->seg = g.seg;
  plan->band_rows = g.a.band_rows;
  plan->box_w = g.box_w;
  plan->box_h = g.box_h;
  plan->n_box_x = 1;
  plan->threads = kThreads;
  plan->stages = g.a.stages;
  plan->smem_bytes = g.smem;
  // the kernel for this geometry must exist NOW: a geometry without an instantiation runs the generic kernel
  const int have = dispatch(*plan, g.k64, g.a.pww, g.deep, nullptr, 0, nullptr);
  if (have == 0 && g.pre) {  // no PRE instantiation for this (block, rows per lane, pitch): the funnel-shift kernel of the geometry
    want_pre = false;
    goto again;
  }
  if (have < 0) {
    if (err) snprintf(err, errlen, "cudaFuncSetAttribute(max dynamic shared memory) failed for the search kernel");
    return -1;
  }
  plan->supported = have;
  return 0;
}

int launch_search_tma(const TmaSearchPlan& plan, ImgView i1, ImgView i2, MvView mv, int n,
                      unsigned long long* counters, unsigned int* work_ctr, int sm_count, cudaStream_t s) {
  (void)i2;
  TmaGeom g;
  if (!plan.supported || !make_geom(i1.w, i1.h, plan.bs, plan.R, plan.pre != 0, &g)) return -1;
  // the geometry is recomputed here; a tuning variable changed between plan and launch must not pair one geometry's arguments
  // with another's instantiation or shared-memory size
  if (g.pre != plan.pre || g.seg != plan.seg || g.a.stages != plan.stages || g.smem != plan.smem_bytes || g.box_h != plan.box_h) return -1;
  TmaSearchArgs a = g.a;
  a.n_pairs = n;
  a.mv = mv.p;
  a.mv_plane = mv.plane;
  a.counters = counters;
  const int total = a.gw * a.gh * n;
  int grid = sm_count * kMinCtas;
  if (grid > total) grid = total;
  {
    // exact multiply-high division (three integer divisions per 32-lane item otherwise): for d >= 2 and p = 31 + ceil(log2 d),
    // m = ceil(2^p / d) < 2^32 and floor(x * m / 2^p) == x / d for every 0 <= x < 2^31
    const unsigned long long iu = (unsigned long long)(a.n * a.segs_per_band > 32 ? a.n * a.segs_per_band : 32);
    tma_div_magic((unsigned)a.n, &a.n_magic, &a.n_shift);
    tma_div_magic((unsigned)iu, &a.iu_magic, &a.iu_shift);
    // lanes of one CTA are numbered in 31 bits: with the counter one CTA could in principle take every block
    const bool fits = (unsigned long long)total * a.nbands * iu + 64 < (1ull << 31);
    static const bool force_static = getenv("BBME_SEARCH_STATIC") != nullptr;  // A-B runs: blocks strided by CTA index
    a.work_ctr = fits && !force_static ? work_ctr : nullptr;
    if (!fits && (((unsigned long long)total + grid - 1) / grid * a.nbands * iu + 64 >= (1ull << 31))) a.iu_magic = 0;  // divide
    if (a.work_ctr && cudaMemsetAsync(a.work_ctr, 0, sizeof(unsigned int), s) != cudaSuccess) return -1;
  }
  return dispatch(plan, g.k64, a.pww, g.deep, &a, grid, s) == 1 ? 0 : -1;
}

}  // namespace bbme
