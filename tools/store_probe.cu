// Per-SM global store / load throughput probe: one CTA per SM (forced by shared memory), W warps, each warp-instruction
// stores 512 contiguous bytes (STG.128) -- the access shape of the TB4 gate stores.  Reports GB/s per SM for 8 and 120 CTAs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(float4* buf, long long per_cta_f4, int iters, int mode) {
  extern __shared__ uint8_t sm[];
  float4* base = buf + (long long)blockIdx.x * per_cta_f4;
  const int nth = blockDim.x;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int it = 0; it < iters; ++it) {
    for (long long i = threadIdx.x; i < per_cta_f4; i += nth) {
      if (mode == 0) base[i] = make_float4((float)it, 1.f, 2.f, 3.f);
      else if (mode == 1) { float4 v = base[i]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
      else { float4 v = base[i]; v.x += 1.f; base[i] = v; }
    }
  }
  if (mode == 1 && acc.x == 12345.f) base[0] = acc;
}
int main() {
  const long long per_cta = 8ll << 20;  // 8 MB per CTA per pass (streams through L2)
  float4* buf; cudaMalloc(&buf, per_cta * 148);
  cudaMemset(buf, 0, per_cta * 148);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[3] = {"store", "load", "load+store in place"};
  for (int mode = 0; mode < 3; ++mode)
    for (int threads : {256, 512, 1024})
      for (int ctas : {8, 120, 148}) {
        probe<<<ctas, threads, 200 * 1024>>>(buf, per_cta / 16, 1, mode);
        cudaEventRecord(e0);
        probe<<<ctas, threads, 200 * 1024>>>(buf, per_cta / 16, 4, mode);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double)per_cta * 4 * (mode == 2 ? 2 : 1);
        printf("%-20s threads %4d ctas %3d: %7.1f GB/s per SM, %8.1f GB/s total (%s)\n", names[mode], threads, ctas,
               bytes / ms / 1e6, bytes * ctas / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
