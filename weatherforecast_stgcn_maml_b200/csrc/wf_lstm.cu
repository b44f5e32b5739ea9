// Multi-layer LSTM forward and BPTT, batched over nodes x windows x MAML tasks (C-ABI).
//
// Replaces the reference's per-node Python loop -- hybrid_model.py:93-105 calls nn.LSTM once
// per node with batch 1 (441 calls per window), and autograd replays 441 cuDNN RNN backward
// graphs (train_hybrid_maml_v5.py:134).  Here every (task, window, node) is one sequence of
// a single batched recurrence; tasks carry their own fast weights (group stride on the
// parameter pointer) so one launch serves the whole MAML inner step.
//
// Layout: all per-layer buffers keep the reference's time-major rows (row = t*N + node inside
// a window, hybrid_model.py:89-90 only *views* this), so the step-t operand of sequence
// (w, node) is row  w*R + t*N + node  -- no transpose is ever materialised.
//
//   gates [layers][G*Bw*R, 4L]  forward: x-projection then activated (i, f, g, o);
//                               backward: overwritten in place by dL/d(pre-activation)
//   h, c  [layers][G*Bw*R, L]
//
// Parameters: the trainable flat buffer in state_dict order:
//   per layer  weight_ih [4L, Kin], weight_hh [4L, L], bias_ih [4L], bias_hh [4L];
//   then output_layer.weight [O, L], output_layer.bias [O].
#include "wf_gemm.cuh"
#include "wf_rng.cuh"

#include "wf_layout.cuh"

extern "C" long long wf_param_count(int layers, int F, int L, int O) {
  if (layers < 1 || layers > 8) return -1;
  return lstm_layout(layers, F, L, O).total;
}

struct LstmStepArgs {
  float* H;          // [G*Bw*R, L]
  float* C;          // [G*Bw*R, L]
  float* XG;         // [G*Bw*R, 4L]
  const float* Whh;  // [4L, L]
  long long gW;      // parameter group stride
  const float* ext;  // backward: dL/dh from above
  float* DC;         // backward: running dL/dc  [G*Bw*N, L]
  int ext_last_only;
  int t, T, N, L, Bw;
};

// ------------------------------------------------------------------ forward step
// gates_pre = x-projection[t] + h[t-1] W_hh^T ; cell update fused in the GEMM epilogue.
// grid (L/32, ceil(Bw*N/128), G); each CTA: 128 sequences x (32 hidden units x 4 gates).
__global__ void __launch_bounds__(WF_GEMM_THREADS) wf_lstm_step_fwd_kernel(LstmStepArgs s) {
  __shared__ GemmSmem sm;
  const int tid = threadIdx.x, g = blockIdx.z;
  const int m0 = blockIdx.y * WF_BM, u0 = blockIdx.x * 32;
  const int R = s.T * s.N, M = s.Bw * s.N, L = s.L;
  const long long grow = (long long)g * s.Bw * R;  // first row of this group
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  if (s.t > 0) {
    LoadRowsK<false> la;
    la.P = s.H + grow * L;
    la.map = RowMap{(long long)(s.t - 1) * s.N * L, (long long)R * L, nullptr, s.N, L};
    la.M = M; la.klim = L; la.rowptr = nullptr; la.col = nullptr; la.val = nullptr; la.R = 0;
    la.init(m0, tid);
    LoadWeightK<true> lb;
    lb.P = s.Whh + g * s.gW; lb.ldb = L; lb.N = L; lb.klim = L;
    lb.init(u0, tid);
    wf_gemm_mainloop(acc, la, lb, 0, L, sm, tid);
  }
  const int tx = tid & 15, ty = tid >> 4;
  const int unit = u0 + 2 * tx;  // this thread owns units (unit, unit+1), all four gates
  if (unit >= L) return;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + wf_acc_row(ty, i);
    if (m >= M) continue;
    int blk = m / s.N, n = m - blk * s.N;
    long long ridx = grow + (long long)blk * R + (long long)s.t * s.N + n;
    float* xg = s.XG + ridx * 4 * L + unit;
    float2 pi = *reinterpret_cast<float2*>(xg);
    float2 pf = *reinterpret_cast<float2*>(xg + L);
    float2 pg = *reinterpret_cast<float2*>(xg + 2 * L);
    float2 po = *reinterpret_cast<float2*>(xg + 3 * L);
    float2 cp = make_float2(0.f, 0.f);
    if (s.t > 0) cp = *reinterpret_cast<const float2*>(s.C + (ridx - s.N) * L + unit);
    // acc[i][0..3] = gates i,f,g,o of `unit`; acc[i][4..7] = same for unit + 1
    float i0 = wf_sigmoid(pi.x + acc[i][0]), i1 = wf_sigmoid(pi.y + acc[i][4]);
    float f0 = wf_sigmoid(pf.x + acc[i][1]), f1 = wf_sigmoid(pf.y + acc[i][5]);
    float g0 = tanhf(pg.x + acc[i][2]), g1 = tanhf(pg.y + acc[i][6]);
    float o0 = wf_sigmoid(po.x + acc[i][3]), o1 = wf_sigmoid(po.y + acc[i][7]);
    float c0 = fmaf(f0, cp.x, i0 * g0), c1 = fmaf(f1, cp.y, i1 * g1);
    float h0 = o0 * tanhf(c0), h1 = o1 * tanhf(c1);
    *reinterpret_cast<float2*>(xg) = make_float2(i0, i1);
    *reinterpret_cast<float2*>(xg + L) = make_float2(f0, f1);
    *reinterpret_cast<float2*>(xg + 2 * L) = make_float2(g0, g1);
    *reinterpret_cast<float2*>(xg + 3 * L) = make_float2(o0, o1);
    *reinterpret_cast<float2*>(s.C + ridx * L + unit) = make_float2(c0, c1);
    *reinterpret_cast<float2*>(s.H + ridx * L + unit) = make_float2(h0, h1);
  }
}

// ------------------------------------------------------------------ backward step
// dh[t] = ext[t] + dG[t+1] W_hh ; gate gradients fused in the GEMM epilogue; dG[t] overwrites
// the activated gates in place.  grid (ceil(L/128), ceil(Bw*N/128), G).
__global__ void __launch_bounds__(WF_GEMM_THREADS) wf_lstm_step_bwd_kernel(LstmStepArgs s) {
  __shared__ GemmSmem sm;
  const int tid = threadIdx.x, g = blockIdx.z;
  const int m0 = blockIdx.y * WF_BM, n0 = blockIdx.x * WF_BN;
  const int R = s.T * s.N, M = s.Bw * s.N, L = s.L;
  const long long grow = (long long)g * s.Bw * R;
  const bool last = s.t == s.T - 1;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  if (!last) {
    LoadRowsK<false> la;
    la.P = s.XG + grow * 4 * L;
    la.map = RowMap{(long long)(s.t + 1) * s.N * 4 * L, (long long)R * 4 * L, nullptr, s.N, 4 * L};
    la.M = M; la.klim = 4 * L; la.rowptr = nullptr; la.col = nullptr; la.val = nullptr; la.R = 0;
    la.init(m0, tid);
    LoadColsMajor lb;
    lb.P = s.Whh + g * s.gW;
    lb.map = RowMap{0, 0, nullptr, 4 * L, L};
    lb.ncols = L; lb.klim = 4 * L;
    lb.init(n0, tid);
    wf_gemm_mainloop(acc, la, lb, 0, 4 * L, sm, tid);
  }
  const int tx = tid & 15, ty = tid >> 4;
  const bool use_ext = s.ext != nullptr && (!s.ext_last_only || last);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + wf_acc_row(ty, i);
    if (m >= M) continue;
    int blk = m / s.N, n = m - blk * s.N;
    long long ridx = grow + (long long)blk * R + (long long)s.t * s.N + n;
    long long sidx = (long long)g * M + m;  // compact per-sequence index
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int unit = n0 + h * 64 + tx * 4;
      if (unit >= L) continue;
      float dh[4] = {acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]};
      if (use_ext) {
        float4 e = *reinterpret_cast<const float4*>(s.ext + (s.ext_last_only ? sidx : ridx) * L + unit);
        dh[0] += e.x; dh[1] += e.y; dh[2] += e.z; dh[3] += e.w;
      }
      float* xg = s.XG + ridx * 4 * L + unit;
      float4 gi = *reinterpret_cast<float4*>(xg);
      float4 gf = *reinterpret_cast<float4*>(xg + L);
      float4 gg = *reinterpret_cast<float4*>(xg + 2 * L);
      float4 go = *reinterpret_cast<float4*>(xg + 3 * L);
      float4 cc = *reinterpret_cast<const float4*>(s.C + ridx * L + unit);
      float4 cp = make_float4(0.f, 0.f, 0.f, 0.f);
      if (s.t > 0) cp = *reinterpret_cast<const float4*>(s.C + (ridx - s.N) * L + unit);
      float4 dcn = make_float4(0.f, 0.f, 0.f, 0.f);
      float* dcp = s.DC + sidx * L + unit;
      if (!last) dcn = *reinterpret_cast<float4*>(dcp);
      const float vi[4] = {gi.x, gi.y, gi.z, gi.w}, vf[4] = {gf.x, gf.y, gf.z, gf.w};
      const float vg[4] = {gg.x, gg.y, gg.z, gg.w}, vo[4] = {go.x, go.y, go.z, go.w};
      const float vc[4] = {cc.x, cc.y, cc.z, cc.w}, vp[4] = {cp.x, cp.y, cp.z, cp.w};
      const float vn[4] = {dcn.x, dcn.y, dcn.z, dcn.w};
      float di[4], df[4], dg[4], dO[4], dcprev[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float tc = tanhf(vc[j]);
        float dc = vn[j] + dh[j] * vo[j] * (1.f - tc * tc);
        dO[j] = dh[j] * tc * vo[j] * (1.f - vo[j]);
        di[j] = dc * vg[j] * vi[j] * (1.f - vi[j]);
        df[j] = dc * vp[j] * vf[j] * (1.f - vf[j]);
        dg[j] = dc * vi[j] * (1.f - vg[j] * vg[j]);
        dcprev[j] = dc * vf[j];
      }
      *reinterpret_cast<float4*>(xg) = make_float4(di[0], di[1], di[2], di[3]);
      *reinterpret_cast<float4*>(xg + L) = make_float4(df[0], df[1], df[2], df[3]);
      *reinterpret_cast<float4*>(xg + 2 * L) = make_float4(dg[0], dg[1], dg[2], dg[3]);
      *reinterpret_cast<float4*>(xg + 3 * L) = make_float4(dO[0], dO[1], dO[2], dO[3]);
      *reinterpret_cast<float4*>(dcp) = make_float4(dcprev[0], dcprev[1], dcprev[2], dcprev[3]);
    }
  }
}

// ------------------------------------------------------------------ C-ABI
static int check_dims(const char* fn, int layers, int F, int L, int T, int N, int G, int Bw) {
  WF_REQUIRE(layers >= 1 && layers <= 8, "%s: layers=%d out of range", fn, layers);
  WF_REQUIRE(F % 4 == 0 && L % 32 == 0, "%s: F=%d must be %%4, L=%d must be %%32", fn, F, L);
  WF_REQUIRE(T > 0 && N > 0 && G > 0 && Bw > 0, "%s: empty batch", fn);
  return WF_OK;
}

extern "C" int wf_dropout_apply(const float* in, long long in_blk_stride, int rows_per_blk, int in_ld, long long rows,
                                int cols, float p, const unsigned long long* rng, int site, float* out, void* stream);

extern "C" int wf_lstm_fwd(const float* x, const float* params, long long params_group_stride, int layers, int F,
                           int L, int O, int T, int N, int G, int Bw, float* gates, float* h, float* c,
                           float p_drop, const unsigned long long* rng, float* h_masked, void* stream) {
  int rc = check_dims("lstm_fwd", layers, F, L, T, N, G, Bw);
  if (rc) return rc;
  WF_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "lstm_fwd: p_drop=%f outside [0, 1)", (double)p_drop);
  WF_REQUIRE(p_drop == 0.f || layers == 1 || (rng != nullptr && h_masked != nullptr), "lstm_fwd: dropout needs rng and h_masked");
  const bool drop = p_drop > 0.f && layers > 1;
  cudaStream_t st = (cudaStream_t)stream;
  const LstmLayout P = lstm_layout(layers, F, L, O);
  const long long R = (long long)T * N, rows = (long long)Bw * R, allrows = rows * G;
  for (int l = 0; l < layers; ++l) {
    const int kin = l == 0 ? F : L;
    float* XG = gates + (long long)l * allrows * 4 * L;
    float* H = h + (long long)l * allrows * L;
    float* C = c + (long long)l * allrows * L;
    GemmArgs a = {};
    // inter-layer dropout (hybrid_model.py:47): layer l >= 1 reads mask * h[l-1] / (1 - p), kept in h_masked for BPTT
    a.A = l == 0 ? x : (drop ? h_masked : h) + (long long)(l - 1) * allrows * L;
    a.am = make_rowmap(0, (int)rows, 0, kin); a.gA = rows * kin;
    a.B = params + P.w_ih[l]; a.ldb = kin; a.gB = params_group_stride;
    a.C = XG; a.cm = make_rowmap(0, (int)rows, 0, 4 * L); a.gC = rows * 4 * L;
    a.bias = params + P.b_ih[l]; a.bias2 = params + P.b_hh[l]; a.gBias = params_group_stride;
    a.M = (int)rows; a.N = 4 * L; a.K = kin;
    rc = wf_launch_gemm_nt(a, G, false, st);
    if (rc) return rc;
    LstmStepArgs s = {};
    s.H = H; s.C = C; s.XG = XG; s.Whh = params + P.w_hh[l]; s.gW = params_group_stride;
    s.T = T; s.N = N; s.L = L; s.Bw = Bw;
    dim3 grid(L / 32, wf_cdiv((long long)Bw * N, WF_BM), G);
    for (int t = 0; t < T; ++t) {
      s.t = t;
      wf_lstm_step_fwd_kernel<<<grid, WF_GEMM_THREADS, 0, st>>>(s);
    }
    WF_CHECK_LAUNCH("lstm_step_fwd");
    if (drop && l + 1 < layers) {
      rc = wf_dropout_apply(H, 0, (int)(allrows > 0x7fffffff ? 0x7fffffff : allrows), L, allrows, L, p_drop, rng, WF_SITE_LSTM + l,
                            h_masked + (long long)l * allrows * L, stream);
      if (rc) return rc;
    }
  }
  return WF_OK;
}

static size_t lstm_bwd_partial_floats(int F, int L, int T, int N, int G, int Bw) {
  int kin = F > L ? F : L;
  long long K = (long long)Bw * T * N;
  size_t a = (size_t)wf_tn_splits(4 * L, kin, (int)K, G) * G * 4 * L * kin;
  size_t b = (size_t)64 * G * 4 * L;  // colsum partials
  return a > b ? a : b;
}

extern "C" size_t wf_lstm_bwd_workspace_bytes(int layers, int F, int L, int T, int N, int G, int Bw) {
  size_t dc = (size_t)G * Bw * N * L;
  size_t dx = (size_t)G * Bw * T * N * L;
  return sizeof(float) * (dc + dx + lstm_bwd_partial_floats(F, L, T, N, G, Bw)) + 256;
}

// gates (in: activations from wf_lstm_fwd; out: pre-activation gradients), dlast = dL/dh of the
// top layer at the last step [G*Bw*N, L]; grads [G, grads_group_stride] receives the LSTM part
// (same offsets as `params`).  x-gradient is not produced: the GCN features are detached
// (hybrid_model.py:63).
extern "C" int wf_lstm_bwd(const float* x, const float* params, long long params_group_stride, int layers, int F,
                           int L, int O, int T, int N, int G, int Bw, float* gates, const float* h, const float* c,
                           const float* dlast, float* grads, long long grads_group_stride, float p_drop,
                           const unsigned long long* rng, const float* h_masked, void* workspace,
                           size_t workspace_bytes, void* stream) {
  int rc = check_dims("lstm_bwd", layers, F, L, T, N, G, Bw);
  if (rc) return rc;
  WF_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "lstm_bwd: p_drop=%f outside [0, 1)", (double)p_drop);
  WF_REQUIRE(p_drop == 0.f || layers == 1 || (rng != nullptr && h_masked != nullptr), "lstm_bwd: dropout needs rng and h_masked");
  const bool drop = p_drop > 0.f && layers > 1;
  if (workspace_bytes < wf_lstm_bwd_workspace_bytes(layers, F, L, T, N, G, Bw))
    return wf_fail(WF_EWORKSPACE, "lstm_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const LstmLayout P = lstm_layout(layers, F, L, O);
  const long long R = (long long)T * N, rows = (long long)Bw * R, allrows = rows * G;
  float* DC = (float*)workspace;
  float* DX = DC + (size_t)G * Bw * N * L;
  float* part = DX + (size_t)allrows * L;
  const size_t partf = lstm_bwd_partial_floats(F, L, T, N, G, Bw);
  for (int l = layers - 1; l >= 0; --l) {
    const int kin = l == 0 ? F : L;
    float* XG = gates + (long long)l * allrows * 4 * L;
    const float* H = h + (long long)l * allrows * L;
    const float* C = c + (long long)l * allrows * L;
    const float* Xl = l == 0 ? x : (drop ? h_masked : h) + (long long)(l - 1) * allrows * L;
    LstmStepArgs s = {};
    s.H = const_cast<float*>(H); s.C = const_cast<float*>(C); s.XG = XG;
    s.Whh = params + P.w_hh[l]; s.gW = params_group_stride;
    s.ext = l == layers - 1 ? dlast : DX;
    s.ext_last_only = l == layers - 1 ? 1 : 0;
    s.DC = DC;
    s.T = T; s.N = N; s.L = L; s.Bw = Bw;
    dim3 grid(wf_cdiv(L, WF_BN), wf_cdiv((long long)Bw * N, WF_BM), G);
    for (int t = T - 1; t >= 0; --t) {
      s.t = t;
      wf_lstm_step_bwd_kernel<<<grid, WF_GEMM_THREADS, 0, st>>>(s);
    }
    WF_CHECK_LAUNCH("lstm_step_bwd");
    // bias gradients: bias_ih and bias_hh receive the same column sums of dG
    RowMap gm = make_rowmap(0, (int)rows, 0, 4 * L);
    rc = wf_launch_colsum(XG, gm, rows * 4 * L, (int)rows, 4 * L, grads + P.b_ih[l], grads + P.b_hh[l],
                          grads_group_stride, G, part, partf, st);
    if (rc) return rc;
    {  // dW_ih = dG^T X_l
      GemmArgs a = {};
      a.A = XG; a.am = gm; a.gA = rows * 4 * L;
      a.B = Xl; a.bm = make_rowmap(0, (int)rows, 0, kin); a.gB = rows * kin;
      a.C = grads + P.w_ih[l]; a.cm = make_rowmap(0, 4 * L, 0, kin); a.gC = grads_group_stride;
      a.M = 4 * L; a.N = kin; a.K = (int)rows;
      a.partial = part;
      rc = wf_launch_gemm_tn(a, G, partf, st);
      if (rc) return rc;
    }
    if (T > 1) {  // dW_hh = sum_{t>=1} dG[t]^T h[t-1]
      GemmArgs a = {};
      const int rpb = (T - 1) * N;
      a.A = XG; a.am = make_rowmap((long long)N * 4 * L, rpb, R * 4 * L, 4 * L); a.gA = rows * 4 * L;
      a.B = H; a.bm = make_rowmap(0, rpb, R * L, L); a.gB = rows * L;
      a.C = grads + P.w_hh[l]; a.cm = make_rowmap(0, 4 * L, 0, L); a.gC = grads_group_stride;
      a.M = 4 * L; a.N = L; a.K = Bw * rpb;
      a.partial = part;
      rc = wf_launch_gemm_tn(a, G, partf, st);
      if (rc) return rc;
    } else {
      for (int g = 0; g < G; ++g)
        cudaMemsetAsync(grads + g * grads_group_stride + P.w_hh[l], 0, sizeof(float) * 4 * L * L, st);
    }
    if (l > 0) {  // dL/d(input of layer l) = dG W_ih  -> ext of layer l-1
      GemmArgs a = {};
      a.A = XG; a.am = gm; a.gA = rows * 4 * L;
      a.B = params + P.w_ih[l]; a.bm = make_rowmap(0, 4 * L, 0, kin); a.gB = params_group_stride;
      a.C = DX; a.cm = make_rowmap(0, (int)rows, 0, kin); a.gC = rows * kin;
      a.M = (int)rows; a.N = kin; a.K = 4 * L;
      rc = wf_launch_gemm_nn(a, G, false, st);
      if (rc) return rc;
      if (drop) {  // ... through the mask of layer l-1's output
        rc = wf_dropout_apply(DX, 0, (int)(allrows > 0x7fffffff ? 0x7fffffff : allrows), L, allrows, L, p_drop, rng,
                              WF_SITE_LSTM + l - 1, DX, stream);
        if (rc) return rc;
      }
    }
  }
  return WF_OK;
}
