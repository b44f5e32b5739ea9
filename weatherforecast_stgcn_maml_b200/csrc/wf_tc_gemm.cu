// tcgen05 3xTF32 GEMM:  C[g][m, n] = sum_k A[g][m, k] * W[g][n, k]  (+ bias, + bias2, relu)
//
// Warp-specialised, one 128 x BN output tile per CTA, 2-stage TMA pipeline over K (32 per stage):
//   warp 0      TMA producer: A tile [128 x 32] fp32 and the weight tiles W_hi / W_lo [BN x 32]
//               (SWIZZLE_128B) into shared memory, completion on `full[s]`
//   warps 2..5  converters: thread = row; read the row's 32 fp32 from shared memory (swizzle-aware,
//               conflict free), split into hi / lo (wf_tc.cuh) and tcgen05.st both into TMEM;
//               arrive on `aready[s]`.  Afterwards the same warps run the epilogue.
//   warp 1      MMA issuer: 12 x tcgen05.mma.kind::tf32 per stage (A from TMEM, B from smem:
//               A_hi W_hi + A_lo W_hi + A_hi W_lo), tcgen05.commit -> `empty[s]` (frees the smem
//               stage and the TMEM A stage), final commit -> `dfull`
//   epilogue    tcgen05.ld 32 columns at a time, bias / ReLU, 128-bit stores.
// TMEM: accumulator columns [0, BN), A stages at [BN, BN + 128).
//
// Used for every K-major contraction with K % 32 == 0 and N % 128 == 0: the 256 -> 256 Theta
// transforms of the GCN (model.py:24-26 via GCNConv.lin), the LSTM input projections
// (hybrid_model.py:42-49), and -- with pre-transposed weights -- dX = dG W.
#include "wf_common.cuh"
#include "wf_tc.cuh"

using namespace wftc;

struct TcGemmArgs {
  float* C;
  int ldc;
  long long c_group_rows;  // row stride between groups in C
  int rows_g;              // valid rows per group
  int tiles_g;             // ceil(rows_g / 128)
  int a_group_rows;        // row stride between groups in the A tensor map
  int N, K;
  const float* bias;
  const float* bias2;
  long long bias_gstride;
  int relu;
  int* err;
};

template <int BN>
__global__ void __launch_bounds__(192, 1)
wf_tc_gemm_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                     const __grid_constant__ CUtensorMap tmBlo, const TcGemmArgs a) {
  constexpr int A_BYTES = 128 * 128, B_BYTES = BN * 128, STAGE = A_BYTES + 2 * B_BYTES, NST = 2;
  constexpr uint32_t TMEM_COLS = BN == 256 ? 512 : 256, A_COL = BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[NST], aready[NST], empty[NST], dfull;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.y / a.tiles_g, tile = blockIdx.y - g * a.tiles_g;
  const int n0 = blockIdx.x * BN;
  const int row0 = g * a.a_group_rows + tile * 128;
  const int nkb = a.K / 32;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&aready[s], 4); mbar_init(&empty[s], 1); }
    mbar_init(&dfull, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmBhi); tma_prefetch_desc(&tmBlo);
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb & 1, ph = (kb >> 1) & 1;
        if (!mbar_wait(&empty[s], ph ^ 1)) { atomicExch(a.err, 1); break; }
        uint8_t* st = smem + s * STAGE;
        mbar_expect_tx(&full[s], STAGE);
        tma_load_2d(st, &tmA, &full[s], kb * 32, row0);
        tma_load_3d(st + A_BYTES, &tmBhi, &full[s], kb * 32, n0, g);
        tma_load_3d(st + A_BYTES + B_BYTES, &tmBlo, &full[s], kb * 32, n0, g);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_tf32(BN);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb & 1, ph = (kb >> 1) & 1;
      if (!mbar_wait(&full[s], ph) || !mbar_wait(&aready[s], ph)) { if (lane == 0) atomicExch(a.err, 2); break; }
      tc_fence_after();
      if (lane == 0) {
        const uint32_t bhi = smem_u32(smem + s * STAGE + A_BYTES), blo = bhi + B_BYTES;
        const uint32_t acol = tbase + A_COL + s * 64;
#pragma unroll
        for (int p = 0; p < 3; ++p) {  // A_hi W_hi, A_lo W_hi, A_hi W_lo
          const uint32_t ac = acol + (p == 1 ? 32 : 0);
          const uint32_t bs = p == 2 ? blo : bhi;
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8)
            umma_ts(tbase, ac + k8 * 8, umma_desc_k_sw128(bs + k8 * 32), idesc, (kb | p | k8) ? 1u : 0u);
        }
        umma_commit(&empty[s]);
        if (kb == nkb - 1) umma_commit(&dfull);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;     // tile row == TMEM lane
    const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
    bool ok = true;
    for (int kb = 0; kb < nkb && ok; ++kb) {
      const int s = kb & 1, ph = (kb >> 1) & 1;
      if (!mbar_wait(&full[s], ph) || !mbar_wait(&empty[s], ph ^ 1)) { if (lane == 0) atomicExch(a.err, 3); ok = false; break; }
      const uint8_t* arow = smem + s * STAGE + row * 128;
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(arow + ((c ^ (row & 7)) << 4));
        split_tf32(v.x, hi[4 * c + 0], lo[4 * c + 0]);
        split_tf32(v.y, hi[4 * c + 1], lo[4 * c + 1]);
        split_tf32(v.z, hi[4 * c + 2], lo[4 * c + 2]);
        split_tf32(v.w, hi[4 * c + 3], lo[4 * c + 3]);
      }
      tmem_st32(tlane + A_COL + s * 64, hi);
      tmem_st32(tlane + A_COL + s * 64 + 32, lo);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&aready[s]);
    }
    // ---- epilogue
    if (ok && !mbar_wait(&dfull, 0)) { if (lane == 0) atomicExch(a.err, 4); ok = false; }
    if (ok) {
      tc_fence_after();
      const int grow = tile * 128 + row;
      const bool valid = grow < a.rows_g;
      float* crow = a.C + ((long long)g * a.c_group_rows + grow) * a.ldc + n0;
      const float* b1 = a.bias ? a.bias + g * a.bias_gstride + n0 : nullptr;
      const float* b2 = a.bias2 ? a.bias2 + g * a.bias_gstride + n0 : nullptr;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld32(tlane + c, v);
        tmem_wait_ld();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                   __uint_as_float(v[j + 3]));
            if (b1) { const float4 b = __ldg(reinterpret_cast<const float4*>(b1 + c + j)); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
            if (b2) { const float4 b = __ldg(reinterpret_cast<const float4*>(b2 + c + j)); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
            if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4*>(crow + c + j) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, TMEM_COLS);
}

// ------------------------------------------------------------------ split helper
__global__ void wf_split_lo_kernel(const float4* __restrict__ src, float4* __restrict__ dst, long long quads) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= quads) return;
  float4 v = src[i], o;
  o.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
  o.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  o.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
  o.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  dst[i] = o;
}

// dst = src - trunc_tf32(src): the `lo` half of the 3xTF32 split for a weight buffer.
extern "C" int wf_split_lo(const float* src, float* dst, long long n, void* stream) {
  WF_REQUIRE(n > 0 && n % 4 == 0, "split_lo: n must be a positive multiple of 4");
  wf_split_lo_kernel<<<wf_cdiv(n / 4, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)src, (float4*)dst, n / 4);
  WF_CHECK_LAUNCH("split_lo");
  return WF_OK;
}

// ------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int wf_encode_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, int swizzle_128b) {
  EncodeTiledFn fn = get_encode();
  if (!fn) return wf_fail(WF_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t d[5], s[5];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, e,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_128b ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return wf_fail(WF_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return WF_OK;
}

// ------------------------------------------------------------------ launcher
// A: dense [a_rows_total, K] (row stride lda); group g's rows start at g * a_group_rows, rows_g valid.
// W_hi / W_lo: [G][N, K] with row stride ldb and group stride b_gstride (elements).
int wf_launch_tc_gemm_nt(const float* A, long long a_rows_total, int lda, int a_group_rows, int rows_g, int G, int K,
                         const float* Whi, const float* Wlo, int ldb, long long b_gstride, long long blo_gstride, int N,
                         const float* bias, const float* bias2, long long bias_gstride, int relu, float* C, int ldc,
                         long long c_group_rows, int* err, cudaStream_t st) {
  WF_REQUIRE(K % 32 == 0 && K >= 32, "tc_gemm: K=%d must be a multiple of 32", K);
  WF_REQUIRE(N % 128 == 0, "tc_gemm: N=%d must be a multiple of 128", N);
  WF_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0, "tc_gemm: leading dimensions must be multiples of 4");
  WF_REQUIRE(b_gstride % 4 == 0 && blo_gstride % 4 == 0, "tc_gemm: weight group strides must be multiples of 4");
  WF_REQUIRE(((uintptr_t)A | (uintptr_t)Whi | (uintptr_t)Wlo | (uintptr_t)C) % 16 == 0, "tc_gemm: pointers must be 16-byte aligned");
  const int BN = (N % 256 == 0) ? 256 : 128;
  CUtensorMap tmA, tmBhi, tmBlo;
  int rc;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)a_rows_total};
    uint64_t str[1] = {(uint64_t)lda * 4};
    uint32_t box[2] = {32, 128};
    if ((rc = wf_encode_tensor_map(&tmA, A, 2, dims, str, box, 1))) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)K, (uint64_t)N, (uint64_t)G};
    uint32_t box[3] = {32, (uint32_t)BN, 1};
    uint64_t str[2] = {(uint64_t)ldb * 4, (uint64_t)(G > 1 ? b_gstride : (long long)N * ldb) * 4};
    if ((rc = wf_encode_tensor_map(&tmBhi, Whi, 3, dims, str, box, 1))) return rc;
    uint64_t str2[2] = {(uint64_t)ldb * 4, (uint64_t)(G > 1 ? blo_gstride : (long long)N * ldb) * 4};
    if ((rc = wf_encode_tensor_map(&tmBlo, Wlo, 3, dims, str2, box, 1))) return rc;
  }
  TcGemmArgs a;
  a.C = C; a.ldc = ldc; a.c_group_rows = c_group_rows; a.rows_g = rows_g; a.tiles_g = wf_cdiv(rows_g, 128);
  a.a_group_rows = a_group_rows; a.N = N; a.K = K; a.bias = bias; a.bias2 = bias2; a.bias_gstride = bias_gstride;
  a.relu = relu; a.err = err;
  dim3 grid(N / BN, a.tiles_g * G);
  if (BN == 256) {
    const int smem = 2 * (128 * 128 + 2 * 256 * 128) + 1024;
    static bool set256 = false;
    if (!set256) {
      if (cudaFuncSetAttribute(wf_tc_gemm_nt_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
        return wf_fail(WF_ECUDA, "tc_gemm: cannot raise dynamic shared memory to %d", smem);
      set256 = true;
    }
    wf_tc_gemm_nt_kernel<256><<<grid, 192, smem, st>>>(tmA, tmBhi, tmBlo, a);
  } else {
    const int smem = 2 * (128 * 128 + 2 * 128 * 128) + 1024;
    static bool set128 = false;
    if (!set128) {
      if (cudaFuncSetAttribute(wf_tc_gemm_nt_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
        return wf_fail(WF_ECUDA, "tc_gemm: cannot raise dynamic shared memory to %d", smem);
      set128 = true;
    }
    wf_tc_gemm_nt_kernel<128><<<grid, 192, smem, st>>>(tmA, tmBhi, tmBlo, a);
  }
  WF_CHECK_LAUNCH("tc_gemm_nt");
  return WF_OK;
}

// Test / general entry point: C[g] = A[g] W[g]^T (+ bias + bias2, relu) on the tensor cores.
// W_lo must hold wf_split_lo(W).  err: one device int, set non-zero if a pipeline wait timed out.
extern "C" int wf_tc_gemm_nt(const float* A, int rows_g, int G, int K, const float* W, const float* W_lo,
                             long long w_group_stride, int N, const float* bias, const float* bias2,
                             long long bias_group_stride, int relu, float* C, int* err, void* stream) {
  WF_REQUIRE(rows_g > 0 && G > 0, "tc_gemm_nt: empty problem");
  return wf_launch_tc_gemm_nt(A, (long long)rows_g * G, K, rows_g, rows_g, G, K, W, W_lo, K, w_group_stride,
                              w_group_stride, N, bias, bias2, bias_group_stride, relu, C, N, rows_g, err,
                              (cudaStream_t)stream);
}
