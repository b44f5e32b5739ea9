// FP32 SIMT GEMM kernels (NT / NN / split-K TN) and column sums.  See wf_gemm.cuh.
#include "wf_gemm.cuh"

// ------------------------------------------------------------------ generic epilogue
__device__ __forceinline__ void wf_store_tile(const float (&acc)[8][8], const GemmArgs& a, float* C,
                                              const float* bias, const float* bias2, int m0, int n0, int tid) {
  const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + wf_acc_row(ty, i);
    if (m >= a.M) continue;
    long long o = row_off(a.cm, m);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int n = n0 + h * 64 + tx * 4;
      if (n >= a.N) continue;
      float4 v = make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
      if (bias) {
        float4 b = wf_ldg4(bias + n);
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
      }
      if (bias2) {
        float4 b = wf_ldg4(bias2 + n);
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
      }
      float4* dst = reinterpret_cast<float4*>(C + o + n);
      if (a.accumulate) {
        float4 c = *dst;
        v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
      }
      if (a.relu) {
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      }
      *dst = v;
    }
  }
}

// ------------------------------------------------------------------ NT: C = op(A) * W^T
template <bool CSR>
__global__ void __launch_bounds__(WF_GEMM_THREADS) wf_gemm_nt_kernel(GemmArgs a) {
  __shared__ GemmSmem sm;
  const int tid = threadIdx.x, g = blockIdx.z;
  const int m0 = blockIdx.y * WF_BM, n0 = blockIdx.x * WF_BN;
  LoadRowsK<CSR> la;
  la.P = a.A + g * a.gA; la.map = a.am; la.M = a.M; la.klim = a.K;
  if (la.map.blk_off) la.map.blk_off += (long long)g * a.gAmBlk;
  la.rowptr = CSR ? a.rowptr + g * a.gRowptr : nullptr;
  la.col = CSR ? a.col + g * a.gCsr : nullptr;
  la.val = CSR ? a.val + g * a.gCsr : nullptr;
  la.R = a.R;
  la.init(m0, tid);
  LoadWeightK<false> lb;
  lb.P = a.B + g * a.gB; lb.ldb = a.ldb; lb.N = a.N; lb.klim = a.K;
  lb.init(n0, tid);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  wf_gemm_mainloop(acc, la, lb, 0, a.K, sm, tid);
  wf_store_tile(acc, a, a.C + g * a.gC, a.bias ? a.bias + g * a.gBias : nullptr,
                a.bias2 ? a.bias2 + g * a.gBias : nullptr, m0, n0, tid);
}

// ------------------------------------------------------------------ NN: C = op(A) * W[K,N]
template <bool CSR>
__global__ void __launch_bounds__(WF_GEMM_THREADS) wf_gemm_nn_kernel(GemmArgs a) {
  __shared__ GemmSmem sm;
  const int tid = threadIdx.x, g = blockIdx.z;
  const int m0 = blockIdx.y * WF_BM, n0 = blockIdx.x * WF_BN;
  LoadRowsK<CSR> la;
  la.P = a.A + g * a.gA; la.map = a.am; la.M = a.M; la.klim = a.K;
  if (la.map.blk_off) la.map.blk_off += (long long)g * a.gAmBlk;
  la.rowptr = CSR ? a.rowptr + g * a.gRowptr : nullptr;
  la.col = CSR ? a.col + g * a.gCsr : nullptr;
  la.val = CSR ? a.val + g * a.gCsr : nullptr;
  la.R = a.R;
  la.init(m0, tid);
  LoadColsMajor lb;
  lb.P = a.B + g * a.gB; lb.map = a.bm; lb.ncols = a.N; lb.klim = a.K;
  lb.init(n0, tid);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  wf_gemm_mainloop(acc, la, lb, 0, a.K, sm, tid);
  wf_store_tile(acc, a, a.C + g * a.gC, a.bias ? a.bias + g * a.gBias : nullptr,
                a.bias2 ? a.bias2 + g * a.gBias : nullptr, m0, n0, tid);
}

// ------------------------------------------------------------------ TN: C[n1,n2] = sum_k A[k,n1] B[k,n2]
__global__ void __launch_bounds__(WF_GEMM_THREADS) wf_gemm_tn_kernel(GemmArgs a) {
  __shared__ GemmSmem sm;
  const int tid = threadIdx.x;
  const int g = blockIdx.z / a.splits, s = blockIdx.z - g * a.splits;
  const int m0 = blockIdx.y * WF_BM, n0 = blockIdx.x * WF_BN;
  const int kbeg = s * a.kchunk;
  const int kend = min(a.K, kbeg + a.kchunk);
  LoadColsMajor la;
  la.P = a.A + g * a.gA; la.map = a.am; la.ncols = a.M; la.klim = kend;
  la.init(m0, tid);
  LoadColsMajor lb;
  lb.P = a.B + g * a.gB; lb.map = a.bm; lb.ncols = a.N; lb.klim = kend;
  lb.init(n0, tid);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  wf_gemm_mainloop(acc, la, lb, kbeg, kend, sm, tid);
  if (a.splits == 1) {
    wf_store_tile(acc, a, a.C + g * a.gC, nullptr, nullptr, m0, n0, tid);
  } else {
    float* P = a.partial + (long long)blockIdx.z * a.M * a.N;
    const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int m = m0 + wf_acc_row(ty, i);
      if (m >= a.M) continue;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int n = n0 + h * 64 + tx * 4;
        if (n >= a.N) continue;
        *reinterpret_cast<float4*>(P + (long long)m * a.N + n) =
            make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
      }
    }
  }
}

__global__ void wf_splitk_reduce_kernel(GemmArgs a) {
  const int g = blockIdx.y;
  const int quads = a.M * a.N / 4;
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= quads) return;
  int e = q * 4;
  int m = e / a.N, n = e - m * a.N;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int sp = 0; sp < a.splits; ++sp) {
    float4 v = *reinterpret_cast<const float4*>(a.partial + ((long long)(g * a.splits + sp) * a.M + m) * a.N + n);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  float4* dst = reinterpret_cast<float4*>(a.C + g * a.gC + row_off(a.cm, m) + n);
  if (a.accumulate) {
    float4 c = *dst;
    s.x += c.x; s.y += c.y; s.z += c.z; s.w += c.w;
  }
  *dst = s;
}

// ------------------------------------------------------------------ column sums (bias gradients)
// partial[g][chunk][c] = sum over the chunk's rows; then out[g][c] = sum over chunks.
__global__ void wf_colsum_partial_kernel(const float* A, RowMap am, long long gA, int rows, int cols,
                                         int rows_per_chunk, float* partial) {
  __shared__ float sh[8][33];
  const int g = blockIdx.z, chunk = blockIdx.y;
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int r0 = chunk * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
  const float* P = A + g * gA;
  float s = 0.f;
  if (c < cols)
    for (int r = r0 + threadIdx.y; r < r1; r += 8) s += __ldg(P + row_off(am, r) + c);
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += sh[y][threadIdx.x];
    partial[((long long)g * gridDim.y + chunk) * cols + c] = t;
  }
}

__global__ void wf_colsum_final_kernel(const float* partial, int chunks, int cols, float* out, float* out2,
                                       long long gOut) {
  const int g = blockIdx.y;
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int k = 0; k < chunks; ++k) s += partial[((long long)g * chunks + k) * cols + c];
  out[g * gOut + c] = s;
  if (out2) out2[g * gOut + c] = s;
}

// ------------------------------------------------------------------ host launchers
static int check_common(const GemmArgs& a, const char* name) {
  WF_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "%s: empty problem M=%d N=%d K=%d", name, a.M, a.N, a.K);
  WF_REQUIRE(a.N % 4 == 0, "%s: N=%d must be a multiple of 4", name, a.N);
  return WF_OK;
}

int wf_launch_gemm_nt(const GemmArgs& a, int groups, bool csr, cudaStream_t st) {
  int rc = check_common(a, "gemm_nt");
  if (rc) return rc;
  WF_REQUIRE(a.K % 4 == 0 && a.ldb % 4 == 0 && a.am.ld % 4 == 0, "gemm_nt: K/ld must be multiples of 4");
  WF_REQUIRE(!csr || a.am.rows_per_blk == a.R, "gemm_nt: CSR needs rows_per_blk == R");
  dim3 grid(wf_cdiv(a.N, WF_BN), wf_cdiv(a.M, WF_BM), groups);
  if (csr) wf_gemm_nt_kernel<true><<<grid, WF_GEMM_THREADS, 0, st>>>(a);
  else wf_gemm_nt_kernel<false><<<grid, WF_GEMM_THREADS, 0, st>>>(a);
  WF_CHECK_LAUNCH("gemm_nt");
  return WF_OK;
}

int wf_launch_gemm_nn(const GemmArgs& a, int groups, bool csr, cudaStream_t st) {
  int rc = check_common(a, "gemm_nn");
  if (rc) return rc;
  WF_REQUIRE(a.K % 4 == 0 && a.am.ld % 4 == 0 && a.bm.ld % 4 == 0, "gemm_nn: K/ld must be multiples of 4");
  WF_REQUIRE(!csr || a.am.rows_per_blk == a.R, "gemm_nn: CSR needs rows_per_blk == R");
  dim3 grid(wf_cdiv(a.N, WF_BN), wf_cdiv(a.M, WF_BM), groups);
  if (csr) wf_gemm_nn_kernel<true><<<grid, WF_GEMM_THREADS, 0, st>>>(a);
  else wf_gemm_nn_kernel<false><<<grid, WF_GEMM_THREADS, 0, st>>>(a);
  WF_CHECK_LAUNCH("gemm_nn");
  return WF_OK;
}

int wf_tn_splits(int M, int N, int K, int groups) {
  long long tiles = (long long)wf_cdiv(M, WF_BM) * wf_cdiv(N, WF_BN) * groups;
  long long want = (296 + tiles - 1) / tiles;
  long long maxs = K / 64;  // >= 64 rows of K per split (the head gradient has K = Bw * N = 441 and 15 tiles)
  if (maxs < 1) maxs = 1;
  if (want > maxs) want = maxs;
  if (want > 64) want = 64;
  if (want < 1) want = 1;
  return (int)want;
}

int wf_launch_gemm_tn(GemmArgs a, int groups, size_t partial_floats, cudaStream_t st) {
  int rc = check_common(a, "gemm_tn");
  if (rc) return rc;
  WF_REQUIRE(a.M % 4 == 0 && a.am.ld % 4 == 0 && a.bm.ld % 4 == 0, "gemm_tn: M/ld must be multiples of 4");
  int splits = a.partial ? wf_tn_splits(a.M, a.N, a.K, groups) : 1;
  while (splits > 1 && (size_t)splits * groups * a.M * a.N > partial_floats) --splits;
  a.splits = splits;
  a.kchunk = wf_cdiv(wf_cdiv(a.K, splits), WF_BK) * WF_BK;
  dim3 grid(wf_cdiv(a.N, WF_BN), wf_cdiv(a.M, WF_BM), groups * splits);
  wf_gemm_tn_kernel<<<grid, WF_GEMM_THREADS, 0, st>>>(a);
  WF_CHECK_LAUNCH("gemm_tn");
  if (splits > 1) {
    int quads = a.M * a.N / 4;
    dim3 g2(wf_cdiv(quads, 256), groups);
    wf_splitk_reduce_kernel<<<g2, 256, 0, st>>>(a);
    WF_CHECK_LAUNCH("splitk_reduce");
  }
  return WF_OK;
}

int wf_launch_colsum(const float* A, RowMap am, long long gA, int rows, int cols, float* out, float* out2,
                     long long gOut, int groups, float* ws, size_t ws_floats, cudaStream_t st) {
  WF_REQUIRE(rows > 0 && cols > 0, "colsum: empty");
  int chunks = wf_cdiv(rows, 512);
  if (chunks > 64) chunks = 64;
  while (chunks > 1 && (size_t)chunks * groups * cols > ws_floats) --chunks;
  WF_REQUIRE((size_t)chunks * groups * cols <= ws_floats, "colsum: workspace too small");
  int rpc = wf_cdiv(rows, chunks);
  dim3 grid(wf_cdiv(cols, 32), chunks, groups), block(32, 8);
  wf_colsum_partial_kernel<<<grid, block, 0, st>>>(A, am, gA, rows, cols, rpc, ws);
  WF_CHECK_LAUNCH("colsum_partial");
  dim3 g2(wf_cdiv(cols, 128), groups);
  wf_colsum_final_kernel<<<g2, 128, 0, st>>>(ws, chunks, cols, out, out2, gOut);
  WF_CHECK_LAUNCH("colsum_final");
  return WF_OK;
}
