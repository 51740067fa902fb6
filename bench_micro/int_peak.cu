// int_peak.cu — issue-rate micro-benchmark for the packed-byte SAD instruction on sm_100a.
//
// MEASURED_PEAKS.json has no integer entry; the search kernel's roofline denominator is
// 4 x (VABSDIFF4.U8.ACC warp-instructions per second) and the kernel design depends on which
// other instructions share its pipe.  Every variant runs register-only, dependence-free
// chains on all SMs and reports thread-ops per clock per SM (clock64) and per second (events).
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o int_peak int_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2);} } while (0)

__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t lop(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t shf(uint32_t lo, uint32_t hi, uint32_t s) {
  uint32_t d;
  asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(s));
  return d;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) {
  uint32_t d;
  asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
  return d;
}
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t iadd3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm volatile("{ .reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %3; }" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t vmin(uint32_t a, uint32_t b) {
  uint32_t d;
  asm volatile("min.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

enum Mix { SAD_ONLY = 0, LOP_ONLY, IMAD_ONLY, SHF_ONLY, PRMT_ONLY, SAD_LOP, SAD_IMAD, SAD_SHF, SAD_PRMT,
           SAD_LDS, SAD4_SHF1, SAD4_LDS1, MIN_ONLY, SAD_MIN, SAD16_MIX, NMIX };
static const char* kMixName[NMIX] = {"sad", "lop3", "imad", "shf", "prmt", "sad+lop3", "sad+imad", "sad+shf",
                                     "sad+prmt", "sad+lds", "4sad+1shf", "4sad+1lds", "min", "sad+min",
                                     "16sad+5lds+4shf"};

constexpr int CH = 16;      // independent chains per thread
constexpr int ITERS = 2048; // loop trips

template <int MIX>
__global__ void __launch_bounds__(1024) k_mix(uint32_t* out, uint32_t seed, long long* cyc) {
  __shared__ uint32_t sm[1024 + 64];
  sm[threadIdx.x] = threadIdx.x * 2654435761u + seed;
  if (threadIdx.x < 64) sm[1024 + threadIdx.x] = seed + threadIdx.x;
  __syncthreads();
  uint32_t a[CH], b[CH], acc[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    a[i] = (threadIdx.x + 1) * 0x01010101u * (i + 1) + seed;
    b[i] = a[i] ^ 0x5a5a5a5au;
    acc[i] = i;
  }
  uint32_t v = seed | 1u;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    if (MIX == SAD_ONLY) {
#pragma unroll
      for (int i = 0; i < CH; ++i) acc[i] = sad4(a[i], b[i], acc[i]);
    } else if (MIX == LOP_ONLY) {
#pragma unroll
      for (int i = 0; i < CH; ++i) acc[i] = lop(a[i], b[i], acc[i]);
    } else if (MIX == IMAD_ONLY) {
#pragma unroll
      for (int i = 0; i < CH; ++i) acc[i] = imad(a[i], b[i], acc[i]);
    } else if (MIX == SHF_ONLY) {
#pragma unroll
      for (int i = 0; i < CH; ++i) acc[i] = shf(acc[i], b[i], a[i]);
    } else if (MIX == PRMT_ONLY) {
#pragma unroll
      for (int i = 0; i < CH; ++i) acc[i] = prmt(acc[i], b[i], a[i]);
    } else if (MIX == MIN_ONLY) {
#pragma unroll
      for (int i = 0; i < CH; ++i) acc[i] = vmin(acc[i] + 0u, b[i]) ^ 0u, b[i] = vmin(b[i], a[i]);
    } else if (MIX == SAD_LOP) {
#pragma unroll
      for (int i = 0; i < CH; i += 2) { acc[i] = sad4(a[i], b[i], acc[i]); acc[i + 1] = lop(a[i + 1], b[i + 1], acc[i + 1]); }
    } else if (MIX == SAD_IMAD) {
#pragma unroll
      for (int i = 0; i < CH; i += 2) { acc[i] = sad4(a[i], b[i], acc[i]); acc[i + 1] = imad(a[i + 1], b[i + 1], acc[i + 1]); }
    } else if (MIX == SAD_SHF) {
#pragma unroll
      for (int i = 0; i < CH; i += 2) { acc[i] = sad4(a[i], b[i], acc[i]); acc[i + 1] = shf(acc[i + 1], b[i + 1], a[i + 1]); }
    } else if (MIX == SAD_PRMT) {
#pragma unroll
      for (int i = 0; i < CH; i += 2) { acc[i] = sad4(a[i], b[i], acc[i]); acc[i + 1] = prmt(acc[i + 1], b[i + 1], a[i + 1]); }
    } else if (MIX == SAD_MIN) {
#pragma unroll
      for (int i = 0; i < CH; i += 2) { acc[i] = sad4(a[i], b[i], acc[i]); acc[i + 1] = vmin(acc[i + 1], b[i + 1]) + 1u; }
    } else if (MIX == SAD_LDS) {
#pragma unroll
      for (int i = 0; i < CH; i += 2) {
        acc[i] = sad4(a[i], b[i], acc[i]);
        acc[i + 1] = sm[(acc[i + 1] + threadIdx.x) & 1023];
      }
    } else if (MIX == SAD4_SHF1) {
#pragma unroll
      for (int i = 0; i < CH; i += 4) {
        b[i] = shf(b[i], a[i], v);
        acc[i] = sad4(a[i], b[i], acc[i]);
        acc[i + 1] = sad4(a[i + 1], b[i], acc[i + 1]);
        acc[i + 2] = sad4(a[i + 2], b[i], acc[i + 2]);
        acc[i + 3] = sad4(a[i + 3], b[i], acc[i + 3]);
      }
    } else if (MIX == SAD4_LDS1) {
#pragma unroll
      for (int i = 0; i < CH; i += 4) {
        uint32_t w = sm[(threadIdx.x + it + i) & 1023];
        acc[i] = sad4(a[i], w, acc[i]);
        acc[i + 1] = sad4(a[i + 1], w, acc[i + 1]);
        acc[i + 2] = sad4(a[i + 2], w, acc[i + 2]);
        acc[i + 3] = sad4(a[i + 3], w, acc[i + 3]);
      }
    } else if (MIX == SAD16_MIX) {
      // the shape of the search kernel's inner row: 5 LDS.32 + 4 funnel shifts feed 16 x 4 SADs
      uint32_t w[5];
      const uint32_t* p = &sm[(threadIdx.x >> 2) + (it & 31)];
#pragma unroll
      for (int k = 0; k < 5; ++k) w[k] = p[k];
      uint32_t s[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) s[k] = shf(w[k], w[k + 1], v);
#pragma unroll
      for (int y = 0; y < 16; ++y) {
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[y] = sad4(a[(y + k) & 15], s[k], acc[y]);
      }
    }
    v += 8;
  }
  long long t1 = clock64();
  uint32_t r = v;
#pragma unroll
  for (int i = 0; i < CH; ++i) r ^= acc[i] ^ b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MIX>
static void run(int sms, int blocks_per_sm, int threads, uint32_t* d_out, long long* d_cyc, double sad_per_iter,
                double all_per_iter) {
  int blocks = sms * blocks_per_sm;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  k_mix<MIX><<<blocks, threads>>>(d_out, 12345u, d_cyc);  // warm-up
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k_mix<MIX><<<blocks, threads>>>(d_out, 777u, d_cyc);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  long long* h = (long long*)malloc(sizeof(long long) * blocks);
  CK(cudaMemcpy(h, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
  double avg = 0;
  long long mx = 0;
  for (int i = 0; i < blocks; ++i) { avg += (double)h[i]; if (h[i] > mx) mx = h[i]; }
  avg /= blocks;
  free(h);
  double thr_iters = (double)threads * blocks_per_sm * ITERS;  // per SM
  double sad_clk = thr_iters * sad_per_iter / (double)mx;      // SAD thread-ops / clk / SM
  double all_clk = thr_iters * all_per_iter / (double)mx;
  double sad_s = (double)blocks * threads * ITERS * sad_per_iter / (ms * 1e-3);
  printf("{\"mix\": \"%s\", \"blocks_per_sm\": %d, \"threads\": %d, \"ms\": %.4f, \"cycles_max\": %lld, "
         "\"cycles_avg\": %.0f, \"sad_lane_ops_per_clk_per_sm\": %.2f, \"all_lane_ops_per_clk_per_sm\": %.2f, "
         "\"sad_warp_instr_per_s\": %.4e, \"absdiff_per_s\": %.4e, \"eff_mhz\": %.0f}\n",
         kMixName[MIX], blocks_per_sm, threads, ms, mx, avg, sad_clk, all_clk, sad_s / 32.0, sad_s * 4.0,
         (double)mx / (ms * 1e3));
  fflush(stdout);
}

int main(int argc, char** argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, dev));
  int sms = p.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"cc\": \"%d.%d\"}\n", p.name, sms, p.clockRate, p.major,
         p.minor);
  uint32_t* d_out;
  long long* d_cyc;
  CK(cudaMalloc(&d_out, sizeof(uint32_t) * sms * 4 * 1024));
  CK(cudaMalloc(&d_cyc, sizeof(long long) * sms * 4));
  for (int pass = 0; pass < 2; ++pass) {
    int bps = pass == 0 ? 1 : 2;
    int th = pass == 0 ? 1024 : 512;
    run<SAD_ONLY>(sms, bps, th, d_out, d_cyc, CH, CH);
    run<LOP_ONLY>(sms, bps, th, d_out, d_cyc, 0, CH);
    run<IMAD_ONLY>(sms, bps, th, d_out, d_cyc, 0, CH);
    run<SHF_ONLY>(sms, bps, th, d_out, d_cyc, 0, CH);
    run<PRMT_ONLY>(sms, bps, th, d_out, d_cyc, 0, CH);
    run<MIN_ONLY>(sms, bps, th, d_out, d_cyc, 0, 2 * CH);
    run<SAD_LOP>(sms, bps, th, d_out, d_cyc, CH / 2, CH);
    run<SAD_IMAD>(sms, bps, th, d_out, d_cyc, CH / 2, CH);
    run<SAD_SHF>(sms, bps, th, d_out, d_cyc, CH / 2, CH);
    run<SAD_PRMT>(sms, bps, th, d_out, d_cyc, CH / 2, CH);
    run<SAD_MIN>(sms, bps, th, d_out, d_cyc, CH / 2, CH + CH / 2);
    run<SAD_LDS>(sms, bps, th, d_out, d_cyc, CH / 2, CH);
    run<SAD4_SHF1>(sms, bps, th, d_out, d_cyc, CH, CH + CH / 4);
    run<SAD4_LDS1>(sms, bps, th, d_out, d_cyc, CH, CH + CH / 4);
    run<SAD16_MIX>(sms, bps, th, d_out, d_cyc, 64, 64 + 9);
  }
  // low-occupancy point: 8 warps / SM (what a 150-register search kernel would run at)
  run<SAD_ONLY>(sms, 1, 256, d_out, d_cyc, CH, CH);
  run<SAD16_MIX>(sms, 1, 256, d_out, d_cyc, 64, 64 + 9);
  run<SAD_ONLY>(sms, 1, 128, d_out, d_cyc, CH, CH);
  return 0;
}
