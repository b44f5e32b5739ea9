// FP32 SIMT GEMM core: 128x128x16 CTA tile, 256 threads, 8x8 register tile per thread,
// register-prefetch double buffering through shared memory.
//
// This is the exact-FP32 path (CUDA-core FFMA).  It backs every contraction that is too
// small or too oddly shaped for the tcgen05 path (K = 24, N = 96, the per-step recurrent
// tiles, the split-K weight-gradient reductions) and serves as the bit-stable reference
// the tensor-core kernels are checked against on the GPU.
//
// Operand addressing is abstracted by small "loader" structs so one main loop serves:
//   NT  C[M,N]  = A[M,K] * W[N,K]^T      (Theta transform, LSTM input projection, head)
//   NN  C[M,N]  = A[M,K] * W[K,N]        (dX = dG * W_ih, dh = dG * W_hh, GCN dX)
//   TN  C[n1,n2] = sum_m A[m,n1]*B[m,n2] (all weight gradients; split-K over rows)
// and the A loader can gather-aggregate rows through a CSR (the fused GCN neighbour
// aggregation: out = (A_hat X) W^T, mathematically equal to A_hat (X W^T)).
#pragma once
#include "wf_common.cuh"

#define WF_BM 128
#define WF_BN 128
#define WF_BK 16
#define WF_SPAD 4
#define WF_GEMM_THREADS 256

typedef float (*wf_tile_t)[WF_BM + WF_SPAD];  // [WF_BK][128 + pad]

__device__ __forceinline__ float4 wf_ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ---- K-major operand rows addressed through a RowMap; optional CSR gather ----------------
template <bool CSR>
struct LoadRowsK {
  const float* P;
  RowMap map;
  int M;       // logical rows available
  int klim;    // exclusive K limit
  const int* rowptr;
  const int* col;
  const float* val;
  int R;       // rows per window (CSR only; == map.rows_per_blk)
  int tid;
  long long off[2];
  int p0[2], p1[2];
  bool ok[2];
  float4 r[2];

  __device__ __forceinline__ void init(int m0, int tid_) {
    tid = tid_;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int m = m0 + (tid >> 2) + 64 * i;
      ok[i] = m < M;
      off[i] = 0; p0[i] = 0; p1[i] = 0;
      if (ok[i]) {
        if (!CSR) {
          off[i] = row_off(map, m);
        } else {
          int w = m / R;
          int rr = m - w * R;
          p0[i] = rowptr[rr];
          p1[i] = rowptr[rr + 1];
          off[i] = row_off(map, w * R);  // element offset of row 0 of this window
        }
      }
    }
  }
  __device__ __forceinline__ void fetch(int k0) {
    int k = k0 + (tid & 3) * 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok[i] && k < klim) {
        if (!CSR) {
          a = wf_ldg4(P + off[i] + k);
        } else {
          for (int p = p0[i]; p < p1[i]; ++p) {
            float v = __ldg(val + p);
            float4 x = wf_ldg4(P + off[i] + (long long)__ldg(col + p) * map.ld + k);
            a.x = fmaf(v, x.x, a.x); a.y = fmaf(v, x.y, a.y);
            a.z = fmaf(v, x.z, a.z); a.w = fmaf(v, x.w, a.w);
          }
        }
      }
      r[i] = a;
    }
  }
  __device__ __forceinline__ void store(wf_tile_t S) const {
    int kq = (tid & 3) * 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int ml = (tid >> 2) + 64 * i;
      S[kq + 0][ml] = r[i].x; S[kq + 1][ml] = r[i].y;
      S[kq + 2][ml] = r[i].z; S[kq + 3][ml] = r[i].w;
    }
  }
};

// ---- K-major weight rows W[n, k] with a plain leading dimension ---------------------------
// LSTM_PERM: tile column nl <-> W row gate*L + unit with gate = nl & 3,
// unit = u0 + 2*((nl & 63) >> 2) + (nl >> 6), so that one thread's 8 accumulator columns
// are the four gates (i, f, g, o) of two adjacent hidden units.
template <bool LSTM_PERM>
struct LoadWeightK {
  const float* P;
  int ldb;
  int N;     // rows of W (or hidden size L when LSTM_PERM)
  int klim;
  int tid;
  const float* rowp[2];
  bool ok[2];
  float4 r[2];

  __device__ __forceinline__ void init(int n0, int tid_) {
    tid = tid_;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int nl = (tid >> 2) + 64 * i;
      int n;
      if (LSTM_PERM) {
        int unit = n0 + 2 * ((nl & 63) >> 2) + (nl >> 6);
        ok[i] = unit < N;
        n = (nl & 3) * N + unit;
      } else {
        n = n0 + nl;
        ok[i] = n < N;
      }
      rowp[i] = P + (long long)n * ldb;
    }
  }
  __device__ __forceinline__ void fetch(int k0) {
    int k = k0 + (tid & 3) * 4;
#pragma unroll
    for (int i = 0; i < 2; ++i)
      r[i] = (ok[i] && k < klim) ? wf_ldg4(rowp[i] + k) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __device__ __forceinline__ void store(wf_tile_t S) const {
    int kq = (tid & 3) * 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int nl = (tid >> 2) + 64 * i;
      S[kq + 0][nl] = r[i].x; S[kq + 1][nl] = r[i].y;
      S[kq + 2][nl] = r[i].z; S[kq + 3][nl] = r[i].w;
    }
  }
};

// ---- operand whose tile columns are contiguous in memory: element (c, k) at row k, col c ---
// (TN: both activations, rows = reduction index through a RowMap; NN: weights W[K, N]).
struct LoadColsMajor {
  const float* P;
  RowMap map;
  int ncols;  // valid columns
  int c0;     // first column of this tile
  int klim;
  int tid;
  float4 r[2];

  __device__ __forceinline__ void init(int c0_, int tid_) { c0 = c0_; tid = tid_; }
  __device__ __forceinline__ void fetch(int k0) {
    int c = c0 + (tid & 31) * 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int k = k0 + (tid >> 5) + 8 * i;
      r[i] = (k < klim && c < ncols) ? wf_ldg4(P + row_off(map, k) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __device__ __forceinline__ void store(wf_tile_t S) const {
#pragma unroll
    for (int i = 0; i < 2; ++i)
      *reinterpret_cast<float4*>(&S[(tid >> 5) + 8 * i][(tid & 31) * 4]) = r[i];
  }
};

struct GemmSmem {
  float As[2][WF_BK][WF_BM + WF_SPAD];
  float Bs[2][WF_BK][WF_BN + WF_SPAD];
};

// acc[i][j]: row = (i < 4 ? ty*4 + i : 64 + ty*4 + i - 4), col = (j < 4 ? tx*4 + j : 64 + tx*4 + j - 4)
template <class LA, class LB>
__device__ __forceinline__ void wf_gemm_mainloop(float (&acc)[8][8], LA& la, LB& lb, int kbeg, int kend,
                                                 GemmSmem& sm, int tid) {
  const int tx = tid & 15, ty = tid >> 4;
  const int nk = (kend - kbeg + WF_BK - 1) / WF_BK;
  if (nk <= 0) return;
  la.fetch(kbeg);
  lb.fetch(kbeg);
  la.store(sm.As[0]);
  lb.store(sm.Bs[0]);
  __syncthreads();
  for (int it = 0; it < nk; ++it) {
    const int cur = it & 1;
    if (it + 1 < nk) {
      la.fetch(kbeg + (it + 1) * WF_BK);
      lb.fetch(kbeg + (it + 1) * WF_BK);
    }
#pragma unroll
    for (int kk = 0; kk < WF_BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&sm.As[cur][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&sm.As[cur][kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&sm.Bs[cur][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&sm.Bs[cur][kk][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (it + 1 < nk) {
      la.store(sm.As[cur ^ 1]);
      lb.store(sm.Bs[cur ^ 1]);
    }
    __syncthreads();
  }
}

__device__ __forceinline__ int wf_acc_row(int ty, int i) { return i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4); }

struct GemmArgs {
  const float* A; RowMap am; long long gA; int gAmBlk;  // gAmBlk: per-group stride into am.blk_off
  const float* B; RowMap bm; int ldb; long long gB;
  float* C; RowMap cm; long long gC;
  const float* bias; const float* bias2; long long gBias;
  int M, N, K;
  int relu, accumulate;
  const int* rowptr; const int* col; const float* val; long long gRowptr, gCsr; int R;
  int splits; int kchunk; float* partial;
};

int wf_launch_gemm_nt(const GemmArgs& a, int groups, bool csr, cudaStream_t st);
int wf_launch_gemm_nn(const GemmArgs& a, int groups, bool csr, cudaStream_t st);
// TN: C[M = A cols, N = B cols], K = logical rows.  Chooses split-K itself when a.partial != null.
int wf_launch_gemm_tn(GemmArgs a, int groups, size_t partial_floats, cudaStream_t st);
int wf_tn_splits(int M, int N, int K, int groups);
int wf_launch_colsum(const float* A, RowMap am, long long gA, int rows, int cols, float* out, float* out2,
                     long long gOut, int groups, float* ws, size_t ws_floats, cudaStream_t st);
