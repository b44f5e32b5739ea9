// Output head + MSE (forward and backward seed) and the fused clip + SGD / Adam(W) steps (C-ABI).
//
// Head: hybrid_model.py:108-115 (Linear(L -> H*12), view(N, H, 12).reshape(-1, 12)) and
// nn.MSELoss as used at train_hybrid_maml_v5.py:133,167 / adapt_hybrid_v5.py:197.
// LAYOUT TRAP reproduced on purpose: prediction rows are node-major
// (row = node*H + h) while target rows are horizon-major (row = h*N + node, dataset.py:45-48);
// the reference compares the two flat buffers element by element, and so does this kernel.
//
// Optimisers: clip_grad_norm_(max_norm) + SGD (train_hybrid_maml_v5.py:116-118,135-139),
// + AdamW (train_hybrid_maml_v5.py:245-249) / Adam with coupled L2 (adaptive_scheduler.py:89-93).
#include "wf_gemm.cuh"

// ------------------------------------------------------------------ MSE
// One cluster of CTAs per window.  Targets either come from an explicit y buffer [G*Bw, N*O] or are read in
// place from the resident features tensor: y_flat[(h*N + n)*nw + c] = feat[tgt_off[w] + h*N*feat_ld
// + n*feat_ld + c], tgt_off[w] = element offset of features[idx + W + 1, 0, 0] (dataset.py:40-44).
// One cluster of MSE_CTAS blocks per window (a single block per window left 133 SMs idle for 32 us): every block sums a
// slice, the partial sums meet in rank 0's shared memory through DSMEM and are added in rank order (deterministic).
constexpr int MSE_CTAS = 8, MSE_THREADS = 256;
__global__ void __cluster_dims__(MSE_CTAS, 1, 1) __launch_bounds__(MSE_THREADS)
wf_mse_kernel(const float* __restrict__ pred, const float* __restrict__ y, const float* __restrict__ feat,
              const long long* __restrict__ tgt_off, int feat_ld, int N, int O, int nw, float grad_scale,
              float* __restrict__ dpred, float* __restrict__ loss) {
  __shared__ float sh[33];
  __shared__ float part[MSE_CTAS];
  const int w = blockIdx.x / MSE_CTAS, rank = blockIdx.x % MSE_CTAS;
  const int per = N * O;
  const float* p = pred + (long long)w * per;
  const float inv = 1.0f / (float)per;
  const long long toff = y ? 0 : tgt_off[w];
  float s = 0.f;
  for (int i = rank * MSE_THREADS + threadIdx.x; i < per; i += MSE_CTAS * MSE_THREADS) {
    float t;
    if (y) {
      t = y[(long long)w * per + i];
    } else {
      const int ry = i / nw, c = i - ry * nw;
      t = feat[toff + (long long)ry * feat_ld + c];  // ry = h*N + n and time rows are N*feat_ld apart
    }
    const float d = p[i] - t;
    s = fmaf(d, d, s);
    if (dpred) dpred[(long long)w * per + i] = 2.0f * d * inv * grad_scale;
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) {  // partial -> rank 0's shared memory
    uint32_t local = (uint32_t)__cvta_generic_to_shared(&part[rank]), remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(remote) : "r"(local), "r"(0));
    asm volatile("st.shared::cluster.f32 [%0], %1;\n" ::"r"(remote), "f"(s) : "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
  if (rank == 0 && threadIdx.x == 0) {
    float tot = 0.f;
    for (int r = 0; r < MSE_CTAS; ++r) tot += part[r];
    loss[w] = tot * inv;
  }
}

static void head_offsets(int layers, int F, int L, int O, long long* hw, long long* hb) {
  long long off = 0;
  for (int l = 0; l < layers; ++l) off += 4LL * L * (l == 0 ? F : L) + 4LL * L * L + 8LL * L;
  *hw = off;
  *hb = off + (long long)O * L;
}

extern "C" size_t wf_head_workspace_bytes(int L, int O, int N, int G, int Bw) {
  (void)N; (void)Bw;
  size_t part = (size_t)64 * G * O * (L > 32 ? L : 32);  // split-K / colsum partials
  return sizeof(float) * part + 256;
}

static int head_check(int layers, int L, int O, int T, int N, int G, int Bw) {
  WF_REQUIRE(layers >= 1 && layers <= 8 && L % 4 == 0 && O % 4 == 0, "head: bad dims L=%d O=%d", L, O);
  WF_REQUIRE(T > 0 && N > 0 && G > 0 && Bw > 0, "head: empty batch");
  return WF_OK;
}

// pred [G*Bw*N, O] = h_top[last step] W_o^T + b_o.   h_top: top layer hidden states [G*Bw*R, L].
extern "C" int wf_head_fwd(const float* h_top, const float* params, long long params_group_stride, int layers,
                           int F, int L, int O, int T, int N, int G, int Bw, float* pred, void* stream) {
  int rc = head_check(layers, L, O, T, N, G, Bw);
  if (rc) return rc;
  long long hw, hb;
  head_offsets(layers, F, L, O, &hw, &hb);
  const long long R = (long long)T * N, rows = (long long)Bw * R;
  const int M = Bw * N;
  GemmArgs a = {};
  a.A = h_top; a.am = make_rowmap((long long)(T - 1) * N * L, N, R * L, L); a.gA = rows * L;
  a.B = params + hw; a.ldb = L; a.gB = params_group_stride;
  a.C = pred; a.cm = make_rowmap(0, M, 0, O); a.gC = (long long)M * O;
  a.bias = params + hb; a.gBias = params_group_stride;
  a.M = M; a.N = O; a.K = L;
  return wf_launch_gemm_nt(a, G, false, (cudaStream_t)stream);
}

// loss[w] = mean((pred - y)^2) per window; dpred (optional) = 2 (pred - y) / (N*O) * grad_scale.
extern "C" int wf_mse_fwd_bwd(const float* pred, const float* y, const float* feat, const long long* tgt_off,
                              int feat_ld, int num_weather, int N, int O, int windows, float grad_scale,
                              float* loss, float* dpred, void* stream) {
  WF_REQUIRE(y != nullptr || (feat != nullptr && tgt_off != nullptr), "mse: no targets given");
  WF_REQUIRE(num_weather > 0 && O % num_weather == 0 && windows > 0, "mse: bad dims");
  WF_REQUIRE((long long)N * O < (1LL << 31), "mse: N * O must fit 32 bits");
  wf_mse_kernel<<<windows * MSE_CTAS, MSE_THREADS, 0, (cudaStream_t)stream>>>(pred, y, feat, tgt_off, feat_ld, N, O,
                                                                              num_weather, grad_scale, dpred, loss);
  WF_CHECK_LAUNCH("mse");
  return WF_OK;
}

// dlast [G*Bw*N, L] = dpred W_o ; grads (optional): dW_o = dpred^T h_last, db_o = colsum(dpred).
extern "C" int wf_head_bwd(const float* dpred, const float* h_top, const float* params,
                           long long params_group_stride, int layers, int F, int L, int O, int T, int N, int G,
                           int Bw, float* dlast, float* grads, long long grads_group_stride, void* workspace,
                           size_t workspace_bytes, void* stream) {
  int rc = head_check(layers, L, O, T, N, G, Bw);
  if (rc) return rc;
  if (workspace_bytes < wf_head_workspace_bytes(L, O, N, G, Bw))
    return wf_fail(WF_EWORKSPACE, "head_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  long long hw, hb;
  head_offsets(layers, F, L, O, &hw, &hb);
  const long long R = (long long)T * N, rows = (long long)Bw * R;
  const int M = Bw * N;
  float* part = (float*)workspace;
  size_t partf = (workspace_bytes - 256) / sizeof(float);
  RowMap dpm = make_rowmap(0, M, 0, O);
  if (dlast) {   // (NULL: only the parameter gradients -- the two halves can run on different streams)
    GemmArgs a = {};
    a.A = dpred; a.am = dpm; a.gA = (long long)M * O;
    a.B = params + hw; a.bm = make_rowmap(0, O, 0, L); a.gB = params_group_stride;
    a.C = dlast; a.cm = make_rowmap(0, M, 0, L); a.gC = (long long)M * L;
    a.M = M; a.N = L; a.K = O;
    rc = wf_launch_gemm_nn(a, G, false, st);
    if (rc) return rc;
  }
  if (grads) {
    GemmArgs a = {};
    a.A = dpred; a.am = dpm; a.gA = (long long)M * O;
    a.B = h_top; a.bm = make_rowmap((long long)(T - 1) * N * L, N, R * L, L); a.gB = rows * L;
    a.C = grads + hw; a.cm = make_rowmap(0, O, 0, L); a.gC = grads_group_stride;
    a.M = O; a.N = L; a.K = M;
    a.partial = part;
    rc = wf_launch_gemm_tn(a, G, partf, st);
    if (rc) return rc;
    rc = wf_launch_colsum(dpred, dpm, (long long)M * O, M, O, grads + hb, nullptr, grads_group_stride, G, part,
                          partf, st);
    if (rc) return rc;
  }
  return WF_OK;
}

// ------------------------------------------------------------------ optimisers
#define WF_NORM_BLOCKS 64

__global__ void wf_sumsq_partial_kernel(const float* __restrict__ grad, long long gstride, long long P,
                                        float* __restrict__ partial) {
  __shared__ float sh[33];
  const int g = blockIdx.y;
  const float4* p = reinterpret_cast<const float4*>(grad + g * gstride);
  const long long quads = P / 4;
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < quads; i += (long long)gridDim.x * blockDim.x) {
    float4 v = p[i];
    s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) partial[g * WF_NORM_BLOCKS + blockIdx.x] = s;
}

__device__ __forceinline__ float wf_clip_coef(const float* partial, int g, float max_norm, float* sh, float* norm_out) {
  float v = threadIdx.x < WF_NORM_BLOCKS ? partial[g * WF_NORM_BLOCKS + threadIdx.x] : 0.f;
  float tot = block_sum(v, sh);
  float nrm = sqrtf(tot);
  if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) norm_out[g] = nrm;
  if (max_norm <= 0.f) return 1.0f;
  return fminf(1.0f, max_norm / (nrm + 1e-6f));  // torch clip_grad_norm_: clamp(max_norm/(norm+1e-6), max=1)
}

__global__ void wf_clip_sgd_apply_kernel(float* __restrict__ theta, long long tstride, const float* __restrict__ grad,
                                         long long gstride, long long P, const float* __restrict__ partial, float lr,
                                         float max_norm, float* norm_out) {
  __shared__ float sh[33];
  const int g = blockIdx.y;
  const float coef = wf_clip_coef(partial, g, max_norm, sh, norm_out);
  float4* t = reinterpret_cast<float4*>(theta + g * tstride);
  const float4* q = reinterpret_cast<const float4*>(grad + g * gstride);
  const long long quads = P / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < quads; i += (long long)gridDim.x * blockDim.x) {
    float4 w = t[i], d = q[i];
    w.x -= lr * (d.x * coef); w.y -= lr * (d.y * coef); w.z -= lr * (d.z * coef); w.w -= lr * (d.w * coef);
    t[i] = w;
  }
}

// hyper (device, 8 floats): lr, beta1, beta2, eps, weight_decay, bias_correction1, bias_correction2, grad_scale
__global__ void wf_clip_adam_apply_kernel(float* __restrict__ theta, const float* __restrict__ grad, float* __restrict__ m,
                                          float* __restrict__ v, long long P, const float* __restrict__ partial,
                                          const float* __restrict__ hyper, float max_norm, int decoupled,
                                          float* norm_out) {
  __shared__ float sh[33];
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
  const float bc1 = hyper[5], bc2 = hyper[6], gs = hyper[7];
  // the norm partials were taken on the unscaled buffer: ||gs * g|| = |gs| * ||g||
  float vsum = threadIdx.x < WF_NORM_BLOCKS ? partial[threadIdx.x] : 0.f;
  float nrm = sqrtf(block_sum(vsum, sh)) * fabsf(gs);
  if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) norm_out[0] = nrm;
  const float coef = (max_norm > 0.f ? fminf(1.0f, max_norm / (nrm + 1e-6f)) : 1.0f) * gs;
  const float step_size = lr / bc1;
  const float bc2s = sqrtf(bc2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
    float w = theta[i], gr = grad[i] * coef;
    if (decoupled) w *= 1.0f - lr * wd;   // AdamW: p.mul_(1 - lr*wd)
    else gr = fmaf(wd, w, gr);            // Adam: grad.add(p, alpha=wd)
    float mi = m[i], vi = v[i];
    mi = mi + (gr - mi) * (1.0f - b1);    // exp_avg.lerp_(grad, 1 - beta1)
    vi = vi * b2 + (1.0f - b2) * gr * gr;  // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
    float denom = sqrtf(vi) / bc2s + eps;
    w -= step_size * (mi / denom);
    theta[i] = w; m[i] = mi; v[i] = vi;
  }
}

__global__ void wf_sum_groups_kernel(const float* __restrict__ src, long long gstride, int G, long long P,
                                     float* __restrict__ dst, int accumulate) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  float s = accumulate ? dst[i] : 0.f;
  for (int g = 0; g < G; ++g) s += src[g * gstride + i];
  dst[i] = s;
}

extern "C" size_t wf_optim_workspace_bytes(int G) { return sizeof(float) * (size_t)G * WF_NORM_BLOCKS + 64; }

// theta[g] -= lr * clip(grad[g]); one independent clip norm per group (task).  norms [G] optional.
extern "C" int wf_clip_sgd_step(float* theta, long long theta_group_stride, const float* grad,
                                long long grad_group_stride, long long P, int G, float lr, float max_norm,
                                float* norms, void* workspace, size_t workspace_bytes, void* stream) {
  WF_REQUIRE(P > 0 && P % 4 == 0 && G > 0, "clip_sgd: P=%lld must be a positive multiple of 4", P);
  WF_REQUIRE(theta_group_stride % 4 == 0 && grad_group_stride % 4 == 0, "clip_sgd: strides must be multiples of 4");
  if (workspace_bytes < wf_optim_workspace_bytes(G)) return wf_fail(WF_EWORKSPACE, "clip_sgd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = (float*)workspace;
  wf_sumsq_partial_kernel<<<dim3(WF_NORM_BLOCKS, G), 256, 0, st>>>(grad, grad_group_stride, P, partial);
  int nb = wf_cdiv(P / 4, 256 * 4);
  if (nb > 148 * 4) nb = 148 * 4;
  wf_clip_sgd_apply_kernel<<<dim3(nb, G), 256, 0, st>>>(theta, theta_group_stride, grad, grad_group_stride, P, partial,
                                                       lr, max_norm, norms);
  WF_CHECK_LAUNCH("clip_sgd");
  return WF_OK;
}

// Single parameter set.  decoupled = 1: AdamW; 0: Adam with L2 folded into the gradient.
extern "C" int wf_clip_adam_step(float* theta, const float* grad, float* exp_avg, float* exp_avg_sq, long long P,
                                 const float* hyper_dev, float max_norm, int decoupled, float* norm_out,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  WF_REQUIRE(P > 0 && P % 4 == 0, "clip_adam: P must be a positive multiple of 4");
  if (workspace_bytes < wf_optim_workspace_bytes(1)) return wf_fail(WF_EWORKSPACE, "clip_adam: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = (float*)workspace;
  wf_sumsq_partial_kernel<<<dim3(WF_NORM_BLOCKS, 1), 256, 0, st>>>(grad, 0, P, partial);
  int nb = wf_cdiv(P, 256 * 4);
  if (nb > 148 * 4) nb = 148 * 4;
  wf_clip_adam_apply_kernel<<<nb, 256, 0, st>>>(theta, grad, exp_avg, exp_avg_sq, P, partial, hyper_dev, max_norm,
                                               decoupled, norm_out);
  WF_CHECK_LAUNCH("clip_adam");
  return WF_OK;
}

extern "C" int wf_sum_groups(const float* src, long long src_group_stride, int G, long long P, float* dst,
                             int accumulate, void* stream) {
  WF_REQUIRE(P > 0 && G > 0, "sum_groups: empty");
  wf_sum_groups_kernel<<<wf_cdiv(P, 256), 256, 0, (cudaStream_t)stream>>>(src, src_group_stride, G, P, dst, accumulate);
  WF_CHECK_LAUNCH("sum_groups");
  return WF_OK;
}

// ------------------------------------------------------------------ error plumbing
#include <stdarg.h>
thread_local char wf_err_msg[512] = "";
int wf_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(wf_err_msg, sizeof(wf_err_msg), fmt, ap);
  va_end(ap);
  return code;
}
extern "C" const char* wf_last_error(void) { return wf_err_msg; }
extern "C" int wf_abi_version(void) { return 2; }
