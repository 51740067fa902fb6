// tma_probe.cu -- bisecting probe for the TMA producer path (debug helper, not part of the library).
// usage: tma_probe <variant>   (each variant in a fresh process: CUDA errors are sticky)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error: %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_probe(const __grid_constant__ CUtensorMap map, uint8_t* out, int box_bytes, int ncopies, int x0, int y0,
                        int skew, int copy_stride, int use_fence) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    if (use_fence) {
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(box_bytes * ncopies) : "memory");
    for (int s = 0; s < ncopies; ++s) {
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
          ::"r"(smem_u32(smem + (size_t)s * copy_stride)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bar)),
            "r"(x0 + s), "r"(y0 - s * skew), "r"(0)
          : "memory");
    }
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  }
  for (int i = threadIdx.x; i < copy_stride * ncopies; i += blockDim.x) out[i] = smem[i];
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int w = 160, h = 96, pitch = 192;
  uint8_t* himg = (uint8_t*)malloc((size_t)pitch * h);
  for (int y = 0; y < h; ++y) for (int x = 0; x < pitch; ++x) himg[y * pitch + x] = (uint8_t)(x < w ? (x * 7 + y * 13) & 0xff : 0xEE);
  uint8_t* dimg; CK(cudaMalloc(&dimg, (size_t)pitch * h + 256)); CK(cudaMemcpy(dimg, himg, (size_t)pitch * h, cudaMemcpyHostToDevice));
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  PFN_encodeTiled enc = (PFN_encodeTiled)fp;
  int box_w = 48, box_h = 40, ncopies = 1, x0 = 16, y0 = 8, skew = 0, use_fence = 1;
  if (variant == 1) { ncopies = 4; }
  if (variant == 2) { ncopies = 4; x0 = -5; y0 = -7; }
  if (variant == 3) { ncopies = 4; x0 = -5; y0 = -7; skew = 2; }
  if (variant == 4) { box_w = 80; box_h = 86; ncopies = 4; x0 = -16; y0 = -16; skew = 2; }
  if (variant == 5) { ncopies = 1; use_fence = 0; }
  if (variant == 6) { box_w = 16; box_h = 16; ncopies = 1; x0 = 32; y0 = 32; }
  if (variant == 7) { box_w = 96; box_h = 80; ncopies = 1; x0 = -16; y0 = -7; }
  if (variant == 8) { box_w = 96; box_h = 80; ncopies = 1; x0 = 112; y0 = 61; }
  CUtensorMap map;
  cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, 1};
  cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * h};
  cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, dimg, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d: encode rc=%d box=%dx%d copies=%d origin=(%d,%d) skew=%d\n", variant, (int)r, box_w, box_h, ncopies, x0, y0, skew);
  const int box_bytes = box_w * box_h;
  const int copy_stride = (box_bytes + 127) / 128 * 128;
  uint8_t* dout; CK(cudaMalloc(&dout, (size_t)copy_stride * ncopies));
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  k_probe<<<1, 128, copy_stride * ncopies>>>(map, dout, box_bytes, ncopies, x0, y0, skew, copy_stride, use_fence);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  uint8_t* hout = (uint8_t*)malloc((size_t)copy_stride * ncopies);
  CK(cudaMemcpy(hout, dout, (size_t)copy_stride * ncopies, cudaMemcpyDeviceToHost));
  long bad = 0;
  for (int s = 0; s < ncopies; ++s)
    for (int y = 0; y < box_h; ++y)
      for (int x = 0; x < box_w; ++x) {
        int gx = x0 + s + x, gy = y0 - s * skew + y;
        uint8_t want = (gx >= 0 && gx < w && gy >= 0 && gy < h) ? himg[gy * pitch + gx] : 0;
        if (hout[(size_t)s * copy_stride + y * box_w + x] != want) ++bad;
      }
  printf("variant %d: %ld mismatching bytes\n", variant, bad);
  return bad ? 1 : 0;
}
