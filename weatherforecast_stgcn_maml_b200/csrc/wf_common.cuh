// Shared helpers for the wf_stgcn C-ABI library (sm_100a).
//
// Conventions used by every kernel in this directory:
//  * activations are row-major f32 [rows, channels]; one window has R = T*N rows,
//    time-major (row = t*N + node), exactly the layout dataset.py:36-37 produces;
//  * work is batched over G groups (MAML tasks: own graph, own fast weights) of Bw
//    windows each; window w = g*Bw + b;
//  * launchers never allocate and never synchronise; they return 0 or a negative
//    WF_E* code and record a message retrievable with wf_last_error().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define WF_OK 0
#define WF_EINVAL -1
#define WF_ECUDA -2
#define WF_EWORKSPACE -3

extern thread_local char wf_err_msg[512];
int wf_fail(int code, const char* fmt, ...);

#define WF_CHECK_LAUNCH(name)                                                       \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) return wf_fail(WF_ECUDA, "%s: %s", name, cudaGetErrorString(e__)); \
  } while (0)

#define WF_REQUIRE(cond, ...)                         \
  do {                                                \
    if (!(cond)) return wf_fail(WF_EINVAL, __VA_ARGS__); \
  } while (0)

static inline int wf_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Maps a logical row index m to the element offset of that row's first element.
//   blk = m / rows_per_blk, i = m % rows_per_blk
//   off = base + (blk_off ? blk_off[blk] : blk * blk_stride) + i * ld
// "blk" is a window; blk_off lets layer 1 read windows straight out of a resident
// features[time, N, C] tensor (a window is a contiguous slice, dataset.py:33-37).
struct RowMap {
  long long base;
  long long blk_stride;
  const long long* blk_off;
  int rows_per_blk;
  int ld;
};

__device__ __forceinline__ long long row_off(const RowMap& r, int m) {
  int blk = m / r.rows_per_blk;
  int i = m - blk * r.rows_per_blk;
  long long bo = r.blk_off ? r.blk_off[blk] : (long long)blk * r.blk_stride;
  return r.base + bo + (long long)i * r.ld;
}

static inline RowMap make_rowmap(long long base, int rows_per_blk, long long blk_stride, int ld,
                                 const long long* blk_off = nullptr) {
  RowMap r;
  r.base = base;
  r.blk_stride = blk_stride;
  r.blk_off = blk_off;
  r.rows_per_blk = rows_per_blk > 0 ? rows_per_blk : 1;
  r.ld = ld;
  return r;
}

__device__ __forceinline__ float wf_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in every thread.  blockDim.x must be a multiple of 32, <= 1024.
__device__ __forceinline__ float block_sum(float v, float* sh /* >= 33 floats */) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}
