// Hardware probe for the building blocks of the persistent LSTM kernels (not part of the library):
//   (1) SS-mode tcgen05.mma.kind::f16 with fp16 hi/lo operands written to shared memory by ordinary
//       threads in the canonical K-major SWIZZLE_128B layout, half of A written by the PEER CTA of a
//       2-CTA cluster through distributed shared memory (st.shared::cluster + remote mbarrier arrive);
//   (2) TS-mode kind::f16 with bf16 hi/lo A operands packed two per 32-bit TMEM column
//       (which half holds the even k is what the probe establishes).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/f16_probe tools/f16_probe.cu
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {  // K-major SWIZZLE_128B, 8-row atoms 1024 B apart
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D = F32, A/B format fmt (0 = F16, 1 = BF16), K-major both, M = 128
__device__ __forceinline__ uint32_t make_idesc_f16(int N, uint32_t fmt) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n"
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ bool mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  for (long long spin = 0; !ok && spin < (1LL << 22); ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t caddr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};\n" :: "r"(caddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void arrive_cluster(uint32_t cbar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" :: "r"(cbar) : "memory");
}

// fp16 hi/lo split of 8 consecutive floats -> two 16-byte chunks
__device__ __forceinline__ void split8_f16(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
  for (int i = 0; i < 4; ++i) {
    __half h0 = __float2half_rn(v[2 * i]), h1 = __float2half_rn(v[2 * i + 1]);
    __half l0 = __float2half_rn(v[2 * i] - __half2float(h0)), l1 = __float2half_rn(v[2 * i + 1] - __half2float(h1));
    h[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
    l[i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// ------------------------------------------------------------------------------------------------
// (1) 2-CTA cluster, SS mode, fp16.  D[rank][128 x 256] = A[128 x 128] * B[rank][256 x 128]^T
// A k-block `rank` (64 columns) is written into BOTH CTAs' A buffers by CTA `rank`.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
ss_cluster_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int* err, int iters,
                  long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int N = 256, K = 128;
  uint8_t* a_hi = smem;                    // [2 kb][128 rows][128 B]
  uint8_t* a_lo = smem + 32768;
  uint8_t* b_hi = smem + 65536;            // [2 kb][256 rows][128 B]
  uint8_t* b_lo = smem + 65536 + 65536;
  __shared__ uint64_t a_ready, dfull;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
  const uint32_t peer = rank ^ 1;
  const float* Br = B + (size_t)rank * N * K;

  for (int idx = tid; idx < N * (K / 8); idx += 128) {   // B -> canonical layout, 16-byte chunks
    int n = idx / (K / 8), c = idx - n * (K / 8);        // c: chunk of 8 k
    float v[8];
    for (int j = 0; j < 8; ++j) v[j] = Br[(size_t)n * K + c * 8 + j];
    uint4 hi, lo;
    split8_f16(v, hi, lo);
    int kb = c >> 3, cc = c & 7;
    uint32_t off = kb * (N * 128) + (n >> 3) * 1024 + (n & 7) * 128 + ((cc ^ (n & 7)) << 4);
    *(uint4*)(b_hi + off) = hi;
    *(uint4*)(b_lo + off) = lo;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;\n" :: "r"(smem_u32(&a_ready)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(smem_u32(&dfull)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" :: "r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");   // peer's barriers are initialised
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  bool ok = true;
  long long t0 = clock64();

  for (int it = 0; it < iters; ++it) {
    // ---- my 64 columns of A for row `tid` -> both CTAs (k-block `rank`)
    for (int c = 0; c < 8; ++c) {
      float v[8];
      for (int j = 0; j < 8; ++j) v[j] = A[(size_t)tid * K + rank * 64 + c * 8 + j];
      uint4 hi, lo;
      split8_f16(v, hi, lo);
      uint32_t off = rank * 16384 + (tid >> 3) * 1024 + (tid & 7) * 128 + ((c ^ (tid & 7)) << 4);
      *(uint4*)(a_hi + off) = hi;
      *(uint4*)(a_lo + off) = lo;
      st_cluster_v4(mapa(smem_u32(a_hi + off), peer), hi);
      st_cluster_v4(mapa(smem_u32(a_lo + off), peer), lo);
    }
    asm volatile("fence.proxy.async;\n" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      arrive_cluster(mapa(smem_u32(&a_ready), rank));
      arrive_cluster(mapa(smem_u32(&a_ready), peer));
    }
    if (tid == 0) {
      if (ok && !mbar_wait_cluster(&a_ready, it & 1)) { atomicExch(err, 1); ok = false; }
      asm volatile("fence.proxy.async;\n" ::: "memory");
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint32_t idesc = make_idesc_f16(N, 0);
      uint32_t acc = 0;
      for (int p = 0; p < 3; ++p) {
        const uint8_t* as = p == 1 ? a_lo : a_hi;
        const uint8_t* bs = p == 2 ? b_lo : b_hi;
        for (int kb = 0; kb < 2; ++kb)
          for (int k16 = 0; k16 < 4; ++k16) {
            mma_ss(tbase, make_desc(smem_u32(as + kb * 16384) + k16 * 32), make_desc(smem_u32(bs + kb * (N * 128)) + k16 * 32),
                   idesc, acc);
            acc = 1;
          }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(&dfull)) : "memory");
    }
    if (ok && !mbar_wait_cluster(&dfull, it & 1)) { atomicExch(err, 2); ok = false; }
    ok = __syncthreads_and(ok);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    // both CTAs must be done reading A before anybody overwrites it in the next iteration
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
  }
  long long t1 = clock64();
  if (tid == 0) cycles[rank] = t1 - t0;
  if (ok)
    for (int c = 0; c < N; c += 32) {
      uint32_t v[32];
      tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + c, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      for (int j = 0; j < 32; ++j) D[((size_t)rank * 128 + tid) * N + c + j] = __uint_as_float(v[j]);
    }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" :: "r"(tbase) : "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");   // no CTA exits while the peer may still write to it
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// (2) single CTA, TS mode, bf16 hi/lo.  D[128 x 128] = A[128 x 256] * B[128 x 256]^T, A packed in TMEM.
// order = 0: even k in the low half of each 32-bit column; 1: even k in the high half.
__global__ void __launch_bounds__(128, 1)
ts_bf16_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int order, int* err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int N = 128, K = 256;
  uint8_t* b_hi = smem;                    // [4 kb][128 rows][128 B]
  uint8_t* b_lo = smem + 65536;
  __shared__ uint64_t dfull;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int idx = tid; idx < N * (K / 8); idx += 128) {
    int n = idx / (K / 8), c = idx - n * (K / 8);
    uint32_t h[4], l[4];
    for (int i = 0; i < 4; ++i) {
      float v0 = B[(size_t)n * K + c * 8 + 2 * i], v1 = B[(size_t)n * K + c * 8 + 2 * i + 1];
      __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
      __nv_bfloat16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
      h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      l[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    int kb = c >> 3, cc = c & 7;
    uint32_t off = kb * (N * 128) + (n >> 3) * 1024 + (n & 7) * 128 + ((cc ^ (n & 7)) << 4);
    *(uint4*)(b_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *(uint4*)(b_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(smem_u32(&dfull)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" :: "r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  const uint32_t tl = tbase + ((uint32_t)(warp * 32) << 16);
  const uint32_t D_COL = 0, AHI = 128, ALO = 256;
  for (int c = 0; c < K / 16; ++c) {   // 16 k values -> 8 packed columns
    uint32_t h[8], l[8];
    for (int i = 0; i < 8; ++i) {
      float v0 = A[(size_t)tid * K + c * 16 + 2 * i], v1 = A[(size_t)tid * K + c * 16 + 2 * i + 1];
      __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
      __nv_bfloat16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
      uint32_t e_h = __bfloat16_as_ushort(h0), o_h = __bfloat16_as_ushort(h1), e_l = __bfloat16_as_ushort(l0), o_l = __bfloat16_as_ushort(l1);
      h[i] = order == 0 ? (e_h | (o_h << 16)) : (o_h | (e_h << 16));
      l[i] = order == 0 ? (e_l | (o_l << 16)) : (o_l | (e_l << 16));
    }
    tmem_st8(tl + AHI + c * 8, h);
    tmem_st8(tl + ALO + c * 8, l);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(N, 1);
    uint32_t acc = 0;
    for (int p = 0; p < 3; ++p) {
      const uint32_t ac = p == 1 ? ALO : AHI;
      const uint8_t* bs = p == 2 ? b_lo : b_hi;
      for (int kb = 0; kb < 4; ++kb)
        for (int k16 = 0; k16 < 4; ++k16) {
          mma_ts(tbase + D_COL, tbase + ac + kb * 32 + k16 * 8, make_desc(smem_u32(bs + kb * (N * 128)) + k16 * 32), idesc, acc);
          acc = 1;
        }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(&dfull)) : "memory");
  }
  if (!mbar_wait_cluster(&dfull, 0)) atomicExch(err, 3);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  for (int c = 0; c < N; c += 32) {
    uint32_t v[32];
    tmem_ld32(tl + D_COL + c, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" :: "r"(tbase) : "memory");
}

static void fill(std::vector<float>& v, float scale, unsigned seed) {
  srand(seed);
  for (auto& x : v) x = ((float)rand() / RAND_MAX * 2.f - 1.f) * scale;
}

int main() {
  int* derr; long long* dcyc;
  CK(cudaMalloc(&derr, 4)); CK(cudaMemset(derr, 0, 4));
  CK(cudaMalloc(&dcyc, 16));
  {  // ---- (1)
    const int N = 256, K = 128;
    std::vector<float> A(128 * K), B(2 * N * K), D(2 * 128 * N);
    fill(A, 1.0f, 1); fill(B, 0.09f, 2);
    for (int i = 0; i < 64; ++i) A[i * 7] *= 1e-4f;   // a few tiny values (fp16 subnormal range)
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    size_t smem = 65536 + 131072 + 1024;
    CK(cudaFuncSetAttribute(ss_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int iters : {1, 50}) {
      CK(cudaMemset(dD, 0xFF, D.size() * 4));
      ss_cluster_kernel<<<2, 128, smem>>>(dA, dB, dD, derr, iters, dcyc);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      int herr; long long cyc[2];
      CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(cyc, dcyc, 16, cudaMemcpyDeviceToHost));
      double emax = 0, mx = 0;
      for (int r = 0; r < 2; ++r)
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < N; ++n) {
            double ex = 0;
            for (int k = 0; k < K; ++k) ex += (double)A[m * K + k] * B[((size_t)r * N + n) * K + k];
            emax = fmax(emax, fabs(D[((size_t)r * 128 + m) * N + n] - ex));
            mx = fmax(mx, fabs(ex));
          }
      printf("SS fp16 hi/lo, 2-CTA cluster DSMEM A exchange, iters=%d: err flag %d, max|D|=%.4f, err vs exact %.3e (rel to max), cycles/iter %.0f / %.0f\n",
             iters, herr, mx, emax / mx, (double)cyc[0] / iters, (double)cyc[1] / iters);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  {  // ---- (2)
    const int N = 128, K = 256;
    std::vector<float> A(128 * K), B(N * K), D(128 * N);
    fill(A, 1e-6f, 3); fill(B, 0.09f, 4);   // gradient-sized A
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    size_t smem = 131072 + 1024;
    CK(cudaFuncSetAttribute(ts_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int order = 0; order < 2; ++order) {
      CK(cudaMemset(dD, 0xFF, D.size() * 4));
      ts_bf16_kernel<<<1, 128, smem>>>(dA, dB, dD, order, derr);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      int herr;
      CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
      double emax = 0, mx = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          double ex = 0;
          for (int k = 0; k < K; ++k) ex += (double)A[m * K + k] * B[(size_t)n * K + k];
          emax = fmax(emax, fabs(D[(size_t)m * N + n] - ex));
          mx = fmax(mx, fabs(ex));
        }
      printf("TS bf16 hi/lo, packing order %d: err flag %d, max|D|=%.4e, err vs exact %.3e (rel to max)\n", order, herr, mx, emax / mx);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  return 0;
}
