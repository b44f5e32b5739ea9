// Feature assembly on the device: the arithmetic of prepare_model_input
// (featurePreprocessor.py:84-177) as two streaming passes over weather[time * N, 12] f32:
//
//   wf_feature_stats     per variable: NaN count and nanmean (the fill value, :104-109), then mean and population
//                        standard deviation of the FILLED array over (time, nodes) (:133-136); f64 accumulation,
//                        fixed-order two-level reduction (deterministic)
//   wf_assemble_features out[row, 0:12]  = (x or fill - mean) / std      (:146; fp32 or fp64 arithmetic, see below)
//                        out[row, 12:16] = time features of the row's time step (:114, :164-165)
//                        out[row, 16:24] = the region's Koppen embedding row (:168-172)
//                        NaN anywhere -> 0 (:178-180)
//
// Both are HBM-bound streams (48 B in, or 48 B in + 96 B out, per row); one thread per row, 16-byte accesses.
// The reference computes (x - mean) / std in fp32 when it derives the statistics itself (numpy f32 arrays) and in fp64
// when statistics are passed in (python lists -> f64 arrays, cast to f32 afterwards): `f64_arith` selects which.
#include <math.h>

#include "wf_common.cuh"

namespace {

constexpr int FEAT_W = 12, FEAT_T = 4, FEAT_K = 8, FEAT_C = FEAT_W + FEAT_T + FEAT_K;
constexpr int STAT_THREADS = 256, STAT_BLOCKS_MAX = 592;  // 4 x 148

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block partials of 12 per-variable f64 sums -> part[blockIdx.x][which][12] (fixed order: lane tree, then warp order)
__device__ __forceinline__ void block_reduce12(double* acc, double* sh /* [8][12] */, double* out) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int v = 0; v < FEAT_W; ++v) {
    const double s = warp_sum_d(acc[v]);
    if (lane == 0) sh[w * FEAT_W + v] = s;
  }
  __syncthreads();
  if (threadIdx.x < FEAT_W) {
    double s = 0.0;
    for (int i = 0; i < STAT_THREADS / 32; ++i) s += sh[i * FEAT_W + threadIdx.x];
    out[threadIdx.x] = s;
  }
  __syncthreads();
}

__device__ __forceinline__ void load_row12(const float* w, long long row, float* x) {
  const float4* p = reinterpret_cast<const float4*>(w + row * FEAT_W);
  const float4 a = p[0], b = p[1], c = p[2];
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
  x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  x[8] = c.x; x[9] = c.y; x[10] = c.z; x[11] = c.w;
}

// pass 1: part[b][0][v] = sum of the non-NaN values, part[b][1][v] = their count
__global__ void __launch_bounds__(STAT_THREADS) feat_sum_kernel(const float* __restrict__ w, long long rows, double* part) {
  __shared__ double sh[8 * FEAT_W];
  double s[FEAT_W], n[FEAT_W];
#pragma unroll
  for (int v = 0; v < FEAT_W; ++v) { s[v] = 0.0; n[v] = 0.0; }
  for (long long r = (long long)blockIdx.x * STAT_THREADS + threadIdx.x; r < rows; r += (long long)gridDim.x * STAT_THREADS) {
    float x[FEAT_W];
    load_row12(w, r, x);
#pragma unroll
    for (int v = 0; v < FEAT_W; ++v)
      if (!isnan(x[v])) { s[v] += (double)x[v]; n[v] += 1.0; }
  }
  block_reduce12(s, sh, part + (long long)blockIdx.x * 2 * FEAT_W);
  block_reduce12(n, sh, part + (long long)blockIdx.x * 2 * FEAT_W + FEAT_W);
}

// between the passes (one block): totals -> fill (f32 nanmean, 0 if a variable is all NaN), mean of the filled array
__global__ void feat_mean_kernel(const double* part, int blocks, long long rows, float* fill, double* mean, long long* nan_count) {
  const int v = threadIdx.x;
  if (v >= FEAT_W) return;
  double s = 0.0, n = 0.0;
  for (int b = 0; b < blocks; ++b) { s += part[(long long)b * 2 * FEAT_W + v]; n += part[(long long)b * 2 * FEAT_W + FEAT_W + v]; }
  const float f = n > 0.0 ? (float)(s / n) : 0.0f;  // np.nanmean of f32 data is f32; all-NaN -> 0 (:106-108)
  fill[v] = f;
  const double nn = (double)rows - n;
  mean[v] = rows > 0 ? (s + nn * (double)f) / (double)rows : 0.0;
  nan_count[v] = (long long)nn;
}

// pass 2: part[b][0][v] = sum of squared deviations of the filled values from mean[v]
__global__ void __launch_bounds__(STAT_THREADS) feat_ssq_kernel(const float* __restrict__ w, long long rows, const float* fill,
                                                                const double* mean, double* part) {
  __shared__ double sh[8 * FEAT_W];
  double q[FEAT_W], m[FEAT_W];
  float f[FEAT_W];
#pragma unroll
  for (int v = 0; v < FEAT_W; ++v) { q[v] = 0.0; m[v] = mean[v]; f[v] = fill[v]; }
  for (long long r = (long long)blockIdx.x * STAT_THREADS + threadIdx.x; r < rows; r += (long long)gridDim.x * STAT_THREADS) {
    float x[FEAT_W];
    load_row12(w, r, x);
#pragma unroll
    for (int v = 0; v < FEAT_W; ++v) {
      const double d = (double)(isnan(x[v]) ? f[v] : x[v]) - m[v];
      q[v] += d * d;
    }
  }
  block_reduce12(q, sh, part + (long long)blockIdx.x * 2 * FEAT_W);
}

__global__ void feat_std_kernel(const double* part, int blocks, long long rows, double* stdev) {
  const int v = threadIdx.x;
  if (v >= FEAT_W) return;
  double q = 0.0;
  for (int b = 0; b < blocks; ++b) q += part[(long long)b * 2 * FEAT_W + v];
  stdev[v] = rows > 0 ? sqrt(q / (double)rows) : 0.0;  // population std (numpy ddof = 0); the caller adds the 1e-8 (:136)
}

struct AssembleArgs {
  const float* w; long long rows; int N;
  float fill[FEAT_W]; double mean[FEAT_W]; double stdev[FEAT_W];
  int normalize, f64_arith;
  const float* timefeat;  // [time, 4]
  float koppen[FEAT_K];
  float* out;
};

__global__ void __launch_bounds__(256) feat_assemble_kernel(const AssembleArgs a) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.rows) return;
  float x[FEAT_W], y[FEAT_C];
  load_row12(a.w, r, x);
#pragma unroll
  for (int v = 0; v < FEAT_W; ++v) {
    const float xf = isnan(x[v]) ? a.fill[v] : x[v];
    float o = xf;
    if (a.normalize) {
      if (a.f64_arith) o = (float)(((double)xf - a.mean[v]) / a.stdev[v]);
      else o = __fdiv_rn(__fsub_rn(xf, (float)a.mean[v]), (float)a.stdev[v]);  // numpy f32: one rounding per operation
    }
    y[v] = o;
  }
  const float4 tf = *reinterpret_cast<const float4*>(a.timefeat + (r / a.N) * FEAT_T);
  y[12] = tf.x; y[13] = tf.y; y[14] = tf.z; y[15] = tf.w;
#pragma unroll
  for (int k = 0; k < FEAT_K; ++k) y[16 + k] = a.koppen[k];
  float4* o4 = reinterpret_cast<float4*>(a.out + r * FEAT_C);
#pragma unroll
  for (int i = 0; i < FEAT_C / 4; ++i) {
    float4 v4 = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
    if (isnan(v4.x)) v4.x = 0.f;  // torch.nan_to_num(combined, nan=0.0) (:178-180)
    if (isnan(v4.y)) v4.y = 0.f;
    if (isnan(v4.z)) v4.z = 0.f;
    if (isnan(v4.w)) v4.w = 0.f;
    o4[i] = v4;
  }
}

int stat_blocks(long long rows) {
  const long long b = (rows + STAT_THREADS - 1) / STAT_THREADS;
  return (int)(b < 1 ? 1 : (b > STAT_BLOCKS_MAX ? STAT_BLOCKS_MAX : b));
}

}  // namespace

extern "C" size_t wf_feature_stats_workspace_bytes(long long rows) {
  return sizeof(double) * 2 * FEAT_W * (size_t)stat_blocks(rows);
}

extern "C" int wf_feature_stats(const float* weather, long long rows, float* fill, double* mean, double* stdev,
                                long long* nan_count, void* workspace, size_t workspace_bytes, void* stream) {
  WF_REQUIRE(weather != nullptr && rows >= 0, "feature_stats: bad input");
  WF_REQUIRE(((uintptr_t)weather & 15) == 0, "feature_stats: weather must be 16-byte aligned");
  const int blocks = stat_blocks(rows);
  if (workspace_bytes < wf_feature_stats_workspace_bytes(rows))
    return wf_fail(WF_EWORKSPACE, "feature_stats: workspace %zu < %zu", workspace_bytes, wf_feature_stats_workspace_bytes(rows));
  cudaStream_t st = (cudaStream_t)stream;
  double* part = (double*)workspace;
  feat_sum_kernel<<<blocks, STAT_THREADS, 0, st>>>(weather, rows, part);
  WF_CHECK_LAUNCH("feat_sum");
  feat_mean_kernel<<<1, 32, 0, st>>>(part, blocks, rows, fill, mean, nan_count);
  WF_CHECK_LAUNCH("feat_mean");
  feat_ssq_kernel<<<blocks, STAT_THREADS, 0, st>>>(weather, rows, fill, mean, part);
  WF_CHECK_LAUNCH("feat_ssq");
  feat_std_kernel<<<1, 32, 0, st>>>(part, blocks, rows, stdev);
  WF_CHECK_LAUNCH("feat_std");
  return WF_OK;
}

extern "C" int wf_assemble_features(const float* weather, long long time_steps, int N, const float* fill_host,
                                    const double* mean_host, const double* std_host, int normalize, int f64_arith,
                                    const float* timefeat, const float* koppen_host, float* out, void* stream) {
  WF_REQUIRE(weather != nullptr && out != nullptr && timefeat != nullptr && time_steps >= 0 && N > 0, "assemble_features: bad input");
  WF_REQUIRE((((uintptr_t)weather | (uintptr_t)out | (uintptr_t)timefeat) & 15) == 0, "assemble_features: buffers must be 16-byte aligned");
  WF_REQUIRE(!normalize || (mean_host != nullptr && std_host != nullptr), "assemble_features: normalize needs mean and std");
  AssembleArgs a;
  a.w = weather; a.rows = time_steps * N; a.N = N; a.normalize = normalize; a.f64_arith = f64_arith;
  a.timefeat = timefeat; a.out = out;
  for (int v = 0; v < FEAT_W; ++v) {
    a.fill[v] = fill_host ? fill_host[v] : 0.0f;
    a.mean[v] = mean_host ? mean_host[v] : 0.0;
    a.stdev[v] = std_host ? std_host[v] : 1.0;
  }
  for (int k = 0; k < FEAT_K; ++k) a.koppen[k] = koppen_host ? koppen_host[k] : 0.0f;
  if (a.rows == 0) return WF_OK;
  feat_assemble_kernel<<<wf_cdiv(a.rows, 256), 256, 0, (cudaStream_t)stream>>>(a);
  WF_CHECK_LAUNCH("feat_assemble");
  return WF_OK;
}
