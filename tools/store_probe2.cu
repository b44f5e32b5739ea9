// Per-SM bulk-store (cp.async.bulk shared -> global) throughput, alone and together with LSU stores (STG.128).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// mode 0: bulk only; 1: LSU only; 2: even warps bulk, odd warps LSU (half the bytes each)
__global__ void probe(uint8_t* buf, long long per_cta, int chunk, int mode) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  uint8_t* base = buf + (long long)blockIdx.x * per_cta;
  uint8_t* stage = sm + warp * chunk * 2;
  // fill staging once (content irrelevant)
  for (int i = lane * 16; i < 2 * chunk; i += 512) *reinterpret_cast<float4*>(stage + i) = make_float4(1, 2, 3, 4);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const bool bulk = mode == 0 || (mode == 2 && (warp & 1) == 0);
  long long n = 0;
  for (long long off = (long long)warp * chunk; off < per_cta; off += (long long)nw * chunk, ++n) {
    if (bulk) {
      if (lane == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + off),
                     "r"(smem_u32(stage + (n & 1) * chunk)), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      }
      __syncwarp();
    } else {
      for (int i = lane * 16; i < chunk; i += 512) *reinterpret_cast<float4*>(base + off + i) = make_float4(1, 2, 3, (float)n);
    }
  }
  if (bulk && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
int main() {
  const long long per_cta = 32ll << 20;
  uint8_t* buf; cudaMalloc(&buf, per_cta * 148);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[3] = {"bulk", "lsu", "bulk+lsu"};
  for (int mode = 0; mode < 3; ++mode)
    for (int chunk : {2048, 4096, 8192})
      for (int threads : {256, 512})
        for (int ctas : {8, 120}) {
          if ((threads >> 5) * chunk * 2 > 200 * 1024) continue;
          probe<<<ctas, threads, 200 * 1024>>>(buf, per_cta, chunk, mode);
          cudaEventRecord(e0);
          probe<<<ctas, threads, 200 * 1024>>>(buf, per_cta, chunk, mode);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          printf("%-9s chunk %5d threads %4d ctas %3d: %7.1f GB/s per SM, %8.1f GB/s total (%s)\n", names[mode], chunk, threads,
                 ctas, per_cta / ms / 1e6, (double)per_cta * ctas / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
  return 0;
}
