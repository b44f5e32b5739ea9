// Hardware probe for the tcgen05 path (not part of the library): one CTA computes
// D[128, N] = A[128, K] * B[N, K]^T on the 5th-gen tensor cores with A staged in TMEM (TS mode)
// and B in shared memory, to establish on a real B200
//   (1) that the descriptor / TMEM layouts used by wf_tc_gemm are right,
//   (2) whether kind::tf32 truncates or rounds the low 13 mantissa bits of fp32 inputs,
//   (3) the accuracy of the 3xTF32 split (A_hi*B_hi + A_lo*B_hi + A_hi*B_lo).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tools/tc_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // version = 1 (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ uint32_t make_idesc(int M, int N, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;            // C format F32
  d |= 2u << 7;            // A format TF32
  d |= 2u << 10;           // B format TF32
  d |= (uint32_t)b_mn_major << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
               "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                  "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
                  "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
                  "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr) : "memory");
}

// mode: 0 = one product with raw fp32 operands, 1 = 3xTF32 split.  bmaj: 0 = B K-major, 1 = B N-major.
template <int N, int K>
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                    float* __restrict__ D, int mode, int bmaj) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int KB = K / 32;
  constexpr int TILE = N * 128;            // bytes of one [N x 32] tf32 tile
  float* b_hi = (float*)smem;              // KB tiles
  float* b_lo = (float*)(smem + KB * TILE);
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  // ---- B -> shared memory in the canonical SWIZZLE_128B layout
  for (int idx = tid; idx < N * K; idx += 128) {
    int n = idx / K, k = idx - n * K;
    float v = B[(size_t)n * K + k];
    float hi = mode ? __uint_as_float(__float_as_uint(v) & 0xFFFFE000u) : v;
    float lo = v - hi;
    int kb = k >> 5, kk = k & 31;
    uint32_t off;
    if (!bmaj) {   // K-major: atom = 8 n-rows x 128 B of k
      off = kb * TILE + (n >> 3) * 1024 + (n & 7) * 128 + (((kk >> 2) ^ (n & 7)) << 4) + (kk & 3) * 4;
    } else {       // N-major: atom = 8 k-rows x 128 B of n; panels of 32 n, 32 k-rows (4 atoms) per k-block
      off = kb * TILE + (n >> 5) * 4096 + (kk >> 3) * 1024 + (kk & 7) * 128 + ((((n & 31) >> 2) ^ (kk & 7)) << 4) + (n & 3) * 4;
    }
    *(float*)((uint8_t*)b_hi + off) = hi;
    *(float*)((uint8_t*)b_lo + off) = lo;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" :: "r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  // generic-proxy smem writes must be visible to the async (tensor core) proxy
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t D_COL = 0, AHI_COL = 256, ALO_COL = 256 + K;   // K <= 128

  // ---- A -> TMEM: thread = row, 32 k values per store
  for (int kb = 0; kb < KB; ++kb) {
    uint32_t hi[32], lo[32];
    for (int j = 0; j < 32; ++j) {
      float v = A[(size_t)tid * K + kb * 32 + j];
      float h = mode ? __uint_as_float(__float_as_uint(v) & 0xFFFFE000u) : v;
      hi[j] = __float_as_uint(h);
      lo[j] = __float_as_uint(v - h);
    }
    tmem_st32(tbase + lane_base + AHI_COL + kb * 32, hi);
    tmem_st32(tbase + lane_base + ALO_COL + kb * 32, lo);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, N, bmaj);
    const uint32_t lbo = bmaj ? 4096 : 16, sbo = 1024;
    uint32_t acc = 0;
    const int nprod = mode ? 3 : 1;
    for (int p = 0; p < nprod; ++p) {
      // p = 0: A_hi B_hi ; p = 1: A_lo B_hi ; p = 2: A_hi B_lo
      const uint32_t a_col = (p == 1) ? ALO_COL : AHI_COL;
      const float* bt = (p == 2) ? b_lo : b_hi;
      for (int kb = 0; kb < KB; ++kb)
        for (int k8 = 0; k8 < 4; ++k8) {
          uint32_t baddr = smem_u32((const uint8_t*)bt + kb * TILE) + (bmaj ? k8 * 1024 : k8 * 32);
          mma_ts(tbase + D_COL, tbase + a_col + kb * 32 + k8 * 8, make_desc(baddr, lbo, sbo), idesc, acc);
          acc = 1;
        }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(&mbar)) : "memory");
  }
  // everyone waits for the MMAs
  {
    uint32_t ok = 0;
    for (long long spin = 0; !ok && spin < (1LL << 24); ++spin) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  for (int c = 0; c < N; c += 32) {
    uint32_t v[32];
    tmem_ld32(tbase + lane_base + D_COL + c, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" :: "r"(tbase) : "memory");
}

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float rn_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0xFFFu + ((u >> 13) & 1u); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

template <int N, int K>
static int run(int mode, int bmaj) {
  const int M = 128;
  std::vector<float> A(M * K), B(N * K), D(M * N);
  srand(1 + mode * 7 + bmaj * 13 + N);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xFF, D.size() * 4));
  size_t smem = 2 * (K / 32) * N * 128 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<N, K><<<1, 128, smem>>>(dA, dB, dD, mode, bmaj);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double e_exact = 0, e_trunc = 0, e_rn = 0, mx = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ex = 0, tr = 0, rn = 0;
      for (int k = 0; k < K; ++k) {
        float a = A[m * K + k], b = B[n * K + k];
        ex += (double)a * b;
        tr += (double)trunc_tf32(a) * trunc_tf32(b);
        rn += (double)rn_tf32(a) * rn_tf32(b);
      }
      double d = D[m * N + n];
      e_exact = fmax(e_exact, fabs(d - ex)); e_trunc = fmax(e_trunc, fabs(d - tr)); e_rn = fmax(e_rn, fabs(d - rn));
      mx = fmax(mx, fabs(ex));
    }
  printf("N=%d K=%d mode=%d bmaj=%d : max|D|=%.3f  err vs exact %.3e  vs trunc-model %.3e  vs rn-model %.3e\n", N, K, mode,
         bmaj, mx, e_exact / mx, e_trunc / mx, e_rn / mx);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return 0;
}

int main() {
  run<128, 64>(0, 0);
  run<128, 64>(1, 0);
  run<256, 64>(0, 0);
  run<256, 64>(1, 0);
  run<128, 64>(0, 1);
  run<128, 64>(1, 1);
  run<128, 128>(1, 0);
  run<128, 128>(1, 1);
  return 0;
}
