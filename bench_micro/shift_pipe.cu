// shift_pipe.cu -- which pipe pays for byte alignment next to a VABSDIFF4 stream?  (sm_100a micro-benchmark)
// Per loop iteration and thread: 16 VABSDIFF4 (the search kernel's row: 4 words x 4 candidate rows) plus one of
//   0 nothing            1 4 x SHF.R.W                 2 4 x (IMAD.HI.U32 + IMAD)      3 5 x IMAD.WIDE.U32 + 4 x IMAD
//   4 4 x IMAD.HI.U32    5 4 x IMAD                    6 5 x IMAD.WIDE.U32             7 4 x PRMT
// Prints ms and the cost of the extra instructions in VABSDIFF4-equivalents.   nvcc -arch=sm_100a -O3 -o shift_pipe shift_pipe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

template <int V>
__global__ void __launch_bounds__(512, 1) k(uint32_t* out, const uint32_t* in, int iters, uint32_t s, uint32_t M, uint32_t one) {
  uint32_t x[5], acc[4] = {0, 0, 0, 0}, A[4];
  for (int i = 0; i < 5; ++i) x[i] = in[threadIdx.x * 5 + i];
  for (int i = 0; i < 4; ++i) A[i] = in[4096 + threadIdx.x * 4 + i];
#pragma unroll 4
  for (int it = 0; it < iters; ++it) {
    uint32_t w[4] = {x[0], x[1], x[2], x[3]};
    if (V == 1) {
      for (int i = 0; i < 4; ++i) asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(w[i]) : "r"(x[i]), "r"(x[i + 1]), "r"(s));
    } else if (V == 2) {
      for (int i = 0; i < 4; ++i) {
        uint32_t t;
        asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(t) : "r"(x[i]), "r"(M));
        asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(w[i]) : "r"(x[i + 1]), "r"(M), "r"(t));
      }
    } else if (V == 3) {
      uint64_t X[5];
      for (int i = 0; i < 5; ++i) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(X[i]) : "r"(x[i]), "r"(M));
      for (int i = 0; i < 4; ++i)
        asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(w[i]) : "r"((uint32_t)(X[i] >> 32)), "r"(one), "r"((uint32_t)X[i + 1]));
    } else if (V == 4) {
      for (int i = 0; i < 4; ++i) asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(w[i]) : "r"(x[i]), "r"(M));
    } else if (V == 5) {
      for (int i = 0; i < 4; ++i) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(w[i]) : "r"(x[i + 1]), "r"(M), "r"(x[i]));
    } else if (V == 6) {
      uint64_t X[5];
      for (int i = 0; i < 5; ++i) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(X[i]) : "r"(x[i]), "r"(M));
      for (int i = 0; i < 4; ++i) w[i] = (uint32_t)(X[i] >> 32);
      acc[0] += (uint32_t)X[4] & 0;  // keep the fifth alive at no cost (folded away if the compiler sees the & 0 -- checked in SASS)
    } else if (V == 7) {
      for (int i = 0; i < 4; ++i) asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(w[i]) : "r"(x[i]), "r"(x[i + 1]), "r"(s));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[c] = sad4(A[i], w[i], acc[c]);
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = w[i];  // the aligned words feed the next iteration's alignment: nothing can be hoisted
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
}

template <int V>
float run(uint32_t* out, const uint32_t* in, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<V><<<148, 512>>>(out, in, iters, 8, 1u << 24, 1);
  cudaEventRecord(e0);
  k<V><<<148, 512>>>(out, in, iters, 8, 1u << 24, 1);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  uint32_t *out, *in;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&in, 8192 * 4);
  cudaMemset(in, 0x5a, 8192 * 4);
  const int iters = 1 << 16;
  float t[8] = {run<0>(out, in, iters), run<1>(out, in, iters), run<2>(out, in, iters), run<3>(out, in, iters),
                run<4>(out, in, iters), run<5>(out, in, iters), run<6>(out, in, iters), run<7>(out, in, iters)};
  const char* nm[8] = {"16 VABSDIFF4", "+4 SHF", "+4 IMAD.HI +4 IMAD", "+5 IMAD.WIDE +4 IMAD", "+4 IMAD.HI", "+4 IMAD", "+5 IMAD.WIDE", "+4 PRMT"};
  for (int v = 0; v < 8; ++v)
    printf("{\"variant\": \"%s\", \"ms\": %.3f, \"extra_in_vabsdiff4_slots\": %.2f}\n", nm[v], t[v], (t[v] / t[0] - 1.0) * 16.0);
  if (cudaDeviceSynchronize() != cudaSuccess) return 1;
  return 0;
}
