// C-ABI entry points of the tensor-core (tcgen05, 3xTF32) path: weight preparation, GCN layer,
// LSTM forward and BPTT.  Same mathematics and buffer contracts as the FP32 SIMT entry points in
// wf_gcn.cu / wf_lstm.cu, plus the transposed activation copies the weight-gradient products read.
//
// Transposed copies: [(G*Bw)][channels][RT] with column (t, node) = t*Np + node, Np = N rounded up
// to a multiple of 4 and RT = T*Np (wf_transposed_pitch).  The padding columns must be zero; the
// kernels never write them, so a buffer zeroed once at allocation stays valid.
#include "wf_gemm.cuh"
#include "wf_layout.cuh"

int wf_launch_tc_rows(const float* A, long long a_rows_total, int lda, int a_group_rows, int rows_g, int G, int K,
                      const float* Whi, const float* Wlo, int ldb, long long b_gstride, long long blo_gstride,
                      int b_shared, int N, const float* bias, const float* bias2, long long bias_gstride, int relu,
                      float* C, int ldc, long long c_gstride, const int* rowptr, const int* col, const float* val,
                      long long g_rowptr, long long g_csr, int R, int Bw, float* ct, float* ct_lo, int Nn, int* err,
                      cudaStream_t st);
int wf_launch_tc_wgrad(const float* AT, int M, const float* BT, const float* BT_lo, int N, int R, int Bw, int G,
                       int a_k0, int b_k0, int klen, float* dW, long long dw_gstride, int* err, cudaStream_t st,
                       float* partials, size_t partial_floats);
int wf_launch_tc_lstm_fwd(float* H, float* Cst, float* XG, float* HT, float* HT_lo, const float* Whh,
                          const float* Whh_lo, long long w_gstride, long long wlo_gstride, int L, int T, int Nn, int Bw,
                          int G, int t, int* err, cudaStream_t st);
int wf_launch_tc_lstm_bwd(float* XG, const float* Cst, float* DGT, float* DC, const float* ext, int ext_last_only,
                          const float* WhhT, const float* WhhT_lo, long long wt_gstride, int L, int T, int Nn, int Bw,
                          int G, int t, int* err, cudaStream_t st);
int wf_launch_transpose_split(const float* in, long long in_gstride, int rows, int cols, float* out, float* out_lo,
                              long long out_gstride, int G, cudaStream_t st);
extern "C" int wf_split_lo(const float* src, float* dst, long long n, void* stream);

extern "C" long long wf_param_count_transposed(int layers, int F, int L, int O) {
  if (layers < 1 || layers > 8) return -1;
  return lstm_layout(layers, F, L, O).totalT;
}

// Row pitch RT of the transposed activation copies for a window of T steps over N nodes.
extern "C" long long wf_transposed_pitch(int T, int N) { return (long long)T * ((N + 3) & ~3); }

// params [G, P] -> params_lo [G, P] (lo halves) and paramsT / paramsT_lo [G, PT]: W_hh^T for every layer,
// W_ih^T for layers >= 1 (the operands of dh = dG W_hh and dX = dG W_ih).  Run after every update of
// the fast weights (it replaces nothing in the reference: this is operand staging for 3xTF32).
extern "C" int wf_prep_weights_tc(const float* params, long long params_group_stride, int layers, int F, int L, int O,
                                  int G, float* params_lo, float* paramsT, float* paramsT_lo, void* stream) {
  WF_REQUIRE(layers >= 1 && layers <= 8 && G > 0, "prep_weights: bad dims");
  const LstmLayout P = lstm_layout(layers, F, L, O);
  WF_REQUIRE(G == 1 || params_group_stride == P.total, "prep_weights: parameter sets must be contiguous");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = wf_split_lo(params, params_lo, P.total * G, stream);
  if (rc) return rc;
  for (int l = 0; l < layers; ++l) {
    rc = wf_launch_transpose_split(params + P.w_hh[l], params_group_stride, 4 * L, L, paramsT + P.whhT[l],
                                   paramsT_lo + P.whhT[l], P.totalT, G, st);
    if (rc) return rc;
    if (l > 0) {
      rc = wf_launch_transpose_split(params + P.w_ih[l], params_group_stride, 4 * L, L, paramsT + P.wihT[l],
                                     paramsT_lo + P.wihT[l], P.totalT, G, st);
      if (rc) return rc;
    }
  }
  return WF_OK;
}

// GCNConv + ReLU on the tensor cores: Y = relu((A_hat X) W^T + b), X dense [G*Bw*R, Cin], W shared.
// YT / YT_lo (optional): transposed copies [(G*Bw)][Cout][RT] for the LSTM layer-0 weight gradient
// (N = nodes per time slice, R = T*N).
extern "C" int wf_gcn_layer_fwd_tc(const float* X, const float* W, const float* W_lo, const float* bias,
                                   const int* rowptr, const int* col, const float* val, long long rowptr_group_stride,
                                   long long csr_group_stride, int R, int N, int Cin, int Cout, int G, int Bw, int relu,
                                   float* Y, float* YT, float* YT_lo, int* err, void* stream) {
  WF_REQUIRE(G > 0 && Bw > 0 && R > 0 && N > 0, "gcn_layer_fwd_tc: bad batch");
  const long long rows_g = (long long)Bw * R;
  return wf_launch_tc_rows(X, rows_g * G, Cin, (int)rows_g, (int)rows_g, G, Cin, W, W_lo, Cin, 0, 0, 1, Cout, bias,
                           nullptr, 0, relu, Y, Cout, rows_g * Cout, rowptr, col, val, rowptr_group_stride,
                           csr_group_stride, R, Bw, YT, YT_lo, N, err, (cudaStream_t)stream);
}

// LSTM forward on the tensor cores.  hT / hT_lo [layers][(G*Bw)][L][RT] are written when non-null
// (training); everything else as wf_lstm_fwd.
extern "C" int wf_lstm_fwd_tc(const float* x, const float* params, const float* params_lo, long long params_group_stride,
                              int layers, int F, int L, int O, int T, int N, int G, int Bw, float* gates, float* h,
                              float* c, float* hT, float* hT_lo, int* err, void* stream) {
  WF_REQUIRE(layers >= 1 && layers <= 8 && F % 32 == 0 && L % 64 == 0, "lstm_fwd_tc: F=%d must be %%32, L=%d %%64", F, L);
  WF_REQUIRE(T > 0 && N > 0 && G > 0 && Bw > 0, "lstm_fwd_tc: empty batch");
  cudaStream_t st = (cudaStream_t)stream;
  const LstmLayout P = lstm_layout(layers, F, L, O);
  const long long R = (long long)T * N, rows = (long long)Bw * R, allrows = rows * G;
  const long long tsz = (long long)G * Bw * L * wf_transposed_pitch(T, N);  // one layer of h^T
  int rc;
  for (int l = 0; l < layers; ++l) {
    const int kin = l == 0 ? F : L;
    float* XG = gates + (long long)l * allrows * 4 * L;
    float* H = h + (long long)l * allrows * L;
    float* C = c + (long long)l * allrows * L;
    float* HT = hT ? hT + l * tsz : nullptr;
    float* HTlo = hT ? hT_lo + l * tsz : nullptr;
    const float* Xl = l == 0 ? x : h + (long long)(l - 1) * allrows * L;
    rc = wf_launch_tc_rows(Xl, allrows, kin, (int)rows, (int)rows, G, kin, params + P.w_ih[l], params_lo + P.w_ih[l], kin,
                           params_group_stride, params_group_stride, 0, 4 * L, params + P.b_ih[l], params + P.b_hh[l],
                           params_group_stride, 0, XG, 4 * L, rows * 4 * L, nullptr, nullptr, nullptr, 0, 0, (int)R, Bw,
                           nullptr, nullptr, N, err, st);
    if (rc) return rc;
    for (int t = 0; t < T; ++t) {
      rc = wf_launch_tc_lstm_fwd(H, C, XG, HT, HTlo, params + P.w_hh[l], params_lo + P.w_hh[l], params_group_stride,
                                 params_group_stride, L, T, N, Bw, G, t, err, st);
      if (rc) return rc;
    }
  }
  return WF_OK;
}

extern "C" size_t wf_lstm_bwd_tc_workspace_bytes(int layers, int F, int L, int T, int N, int G, int Bw) {
  (void)layers; (void)F;
  size_t rows = (size_t)G * Bw * T * N;
  size_t dc = (size_t)G * Bw * N * L;
  size_t dx = rows * L;
  size_t part = (size_t)64 * G * 4 * L;  // colsum partials
  return sizeof(float) * (dc + dx + part) + 256;
}

// BPTT on the tensor cores.  xT / xT_lo: transposed layer-0 input [(G*Bw)][F][RT]; hT / hT_lo from
// wf_lstm_fwd_tc; paramsT / paramsT_lo from wf_prep_weights_tc; dgT: scratch [(G*Bw)][4L][RT] whose
// padding columns are zero.  Other arguments as wf_lstm_bwd.
extern "C" int wf_lstm_bwd_tc(const float* xT, const float* xT_lo, const float* paramsT, const float* paramsT_lo,
                              int layers, int F, int L, int O, int T, int N, int G, int Bw, float* gates, const float* c,
                              const float* hT, const float* hT_lo, float* dgT, const float* dlast, float* grads,
                              long long grads_group_stride, void* workspace, size_t workspace_bytes, int* err,
                              void* stream) {
  WF_REQUIRE(layers >= 1 && layers <= 8 && L == 128 && F % 128 == 0, "lstm_bwd_tc: needs L == 128 and F %% 128 == 0");
  if (workspace_bytes < wf_lstm_bwd_tc_workspace_bytes(layers, F, L, T, N, G, Bw))
    return wf_fail(WF_EWORKSPACE, "lstm_bwd_tc: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const LstmLayout P = lstm_layout(layers, F, L, O);
  const long long R = (long long)T * N, rows = (long long)Bw * R, allrows = rows * G;
  const int Np = (N + 3) & ~3;
  const int RT = T * Np;
  const long long tsz = (long long)G * Bw * L * RT;
  float* DC = (float*)workspace;
  float* DX = DC + (size_t)G * Bw * N * L;
  float* part = DX + (size_t)allrows * L;
  const size_t partf = (size_t)64 * G * 4 * L;
  int rc;
  for (int l = layers - 1; l >= 0; --l) {
    const int kin = l == 0 ? F : L;
    float* XG = gates + (long long)l * allrows * 4 * L;
    const float* C = c + (long long)l * allrows * L;
    const float* HT = hT + l * tsz;
    const float* HTlo = hT_lo + l * tsz;
    const float* XT = l == 0 ? xT : hT + (l - 1) * tsz;
    const float* XTlo = l == 0 ? xT_lo : hT_lo + (l - 1) * tsz;
    const float* ext = l == layers - 1 ? dlast : DX;
    for (int t = T - 1; t >= 0; --t) {
      rc = wf_launch_tc_lstm_bwd(XG, C, dgT, DC, ext, l == layers - 1 ? 1 : 0, paramsT + P.whhT[l],
                                 paramsT_lo + P.whhT[l], P.totalT, L, T, N, Bw, G, t, err, st);
      if (rc) return rc;
    }
    RowMap gm = make_rowmap(0, (int)rows, 0, 4 * L);
    rc = wf_launch_colsum(XG, gm, rows * 4 * L, (int)rows, 4 * L, grads + P.b_ih[l], grads + P.b_hh[l],
                          grads_group_stride, G, part, partf, st);
    if (rc) return rc;
    // dW_ih = dG^T X_l  as  (dG^T)(X_l^T)^T over all columns of every window (padding columns are zero)
    rc = wf_launch_tc_wgrad(dgT, 4 * L, XT, XTlo, kin, RT, Bw, G, 0, 0, RT, grads + P.w_ih[l], grads_group_stride, err, st,
                            nullptr, 0);
    if (rc) return rc;
    if (T > 1) {  // dW_hh = sum_{t>=1} dG[t]^T h[t-1]: dG^T columns [Np, RT) against h^T columns [0, RT-Np)
      rc = wf_launch_tc_wgrad(dgT, 4 * L, HT, HTlo, L, RT, Bw, G, Np, 0, RT - Np, grads + P.w_hh[l], grads_group_stride,
                              err, st, nullptr, 0);
      if (rc) return rc;
    } else {
      for (int g = 0; g < G; ++g)
        cudaMemsetAsync(grads + g * grads_group_stride + P.w_hh[l], 0, sizeof(float) * 4 * L * L, st);
    }
    if (l > 0) {  // dL/d(input of layer l) = dG W_ih  ->  ext of layer l-1
      rc = wf_launch_tc_rows(XG, allrows, 4 * L, (int)rows, (int)rows, G, 4 * L, paramsT + P.wihT[l],
                             paramsT_lo + P.wihT[l], 4 * L, P.totalT, P.totalT, 0, L, nullptr, nullptr, 0, 0, DX, L,
                             rows * L, nullptr, nullptr, nullptr, 0, 0, (int)R, Bw, nullptr, nullptr, N, err, st);
      if (rc) return rc;
    }
  }
  return WF_OK;
}

// Test entry point for the weight-gradient contraction (see wf_launch_tc_wgrad); R = row pitch.
extern "C" int wf_tc_wgrad(const float* AT, int M, const float* BT, const float* BT_lo, int N, int R, int Bw, int G,
                           int a_k0, int b_k0, int klen, float* dW, long long dw_group_stride, int* err, void* stream) {
  return wf_launch_tc_wgrad(AT, M, BT, BT_lo, N, R, Bw, G, a_k0, b_k0, klen, dW, dw_group_stride, err,
                            (cudaStream_t)stream, nullptr, 0);
}
