// Stand-alone dropout pass and the RNG pass counter (C-ABI).  The fused kernels (GCN epilogue, LSTM recurrence,
// dX epilogue) apply the same masks in place; this pass serves the sites that are too small to fuse (head input:
// hybrid_model.py:108), the exact-FP32 paths, and tests that read a mask back (apply it to ones).
#include "wf_common.cuh"
#include "wf_rng.cuh"

namespace {

__global__ void __launch_bounds__(256) wf_dropout_apply_kernel(const float* __restrict__ in, long long in_blk_stride,
                                                               int rows_per_blk, int in_ld, long long rows, int cols,
                                                               DropCfg d, float* __restrict__ out) {
  const int c4n = cols >> 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * c4n) return;
  const long long r = i / c4n;
  const int c4 = (int)(i - r * c4n);
  const long long blk = r / rows_per_blk;
  const int ri = (int)(r - blk * rows_per_blk);
  const float4 v = *reinterpret_cast<const float4*>(in + blk * in_blk_stride + (long long)ri * in_ld + 4 * c4);
  float m[4] = {1.f, 1.f, 1.f, 1.f};
  if (d.rng != nullptr) {
    const DropState s = wf_drop_state(d);
    wf_drop4(s, (unsigned long long)i, m);  // e / 4 with e = r * cols + 4 * c4
  }
  *reinterpret_cast<float4*>(out + r * cols + 4 * c4) = make_float4(v.x * m[0], v.y * m[1], v.z * m[2], v.w * m[3]);
}

__global__ void wf_rng_advance_kernel(unsigned long long* rng) { rng[1] += 1ULL; }

}  // namespace

// out[r, c] = in[row r][c] * keep(site, r*cols + c) / (1 - p).  Input row r lives at
// in + (r / rows_per_blk) * in_blk_stride + (r % rows_per_blk) * in_ld (a strided gather, e.g. the last time slice
// of every window); out is dense [rows, cols] and may alias a dense input.  p <= 0 or rng == NULL: plain copy.
extern "C" int wf_dropout_apply(const float* in, long long in_blk_stride, int rows_per_blk, int in_ld, long long rows,
                                int cols, float p, const unsigned long long* rng, int site, float* out, void* stream) {
  WF_REQUIRE(rows > 0 && cols > 0 && cols % 4 == 0 && in_ld % 4 == 0 && in_blk_stride % 4 == 0,
             "dropout_apply: cols, in_ld and in_blk_stride must be multiples of 4");
  WF_REQUIRE(p >= 0.f && p < 1.f, "dropout_apply: p=%f outside [0, 1)", (double)p);
  WF_REQUIRE(rows_per_blk > 0, "dropout_apply: rows_per_blk must be positive");
  const long long n4 = rows * (cols >> 2);
  wf_dropout_apply_kernel<<<wf_cdiv(n4, 256), 256, 0, (cudaStream_t)stream>>>(in, in_blk_stride, rows_per_blk, in_ld, rows,
                                                                              cols, wf_drop_cfg(p, rng, site), out);
  WF_CHECK_LAUNCH("dropout_apply");
  return WF_OK;
}

// rng[1] += 1 on the stream: the next forward pass draws fresh masks (also inside a replayed CUDA graph).
extern "C" int wf_rng_advance(unsigned long long* rng, void* stream) {
  WF_REQUIRE(rng != nullptr, "rng_advance: null state");
  wf_rng_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(rng);
  WF_CHECK_LAUNCH("rng_advance");
  return WF_OK;
}
