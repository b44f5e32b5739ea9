/* wf_stgcn.h -- C ABI of libwf_stgcn.so, the sm_100a implementation of the
 * Hybrid MAML-STGCN-LSTM v5 hot path (Yalt8826/WeatherForecast_STGCN_MAML).
 *
 * The reference has no FFI layer of its own (it is 13 Python files that call
 * torch / torch_geometric / scipy); each entry point below names the reference
 * lines whose work it replaces, and weatherforecast_stgcn_maml_b200/_lib.py is
 * the ctypes binding a maintainer would add.
 *
 * Rules common to every launcher:
 *   - all pointers are DEVICE pointers unless stated otherwise, f32 row-major;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised;
 *   - nothing is allocated: scratch comes from the caller, sized by *_workspace_bytes;
 *   - returns 0, or <0 (WF_EINVAL -1, WF_ECUDA -2, WF_EWORKSPACE -3) with a message
 *     available from wf_last_error() (thread-local);
 *   - batching: G groups (MAML tasks: own graph + own fast weights) x Bw windows,
 *     window w = g*Bw + b; a window has R = T*N rows, time-major (row = t*N + node),
 *     the layout dataset.py:36-37 produces.
 */
#ifndef WF_STGCN_H
#define WF_STGCN_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

int wf_abi_version(void);
const char* wf_last_error(void);

/* Number of trainable f32 values in the flat parameter buffer used by the LSTM/head entry
 * points: per layer weight_ih[4L,Kin] weight_hh[4L,L] bias_ih[4L] bias_hh[4L] (Kin = F for
 * layer 0, else L), then output_layer.weight[O,L], output_layer.bias[O] -- the order of the
 * reference state_dict (hybrid_model.py:42-55).  606,304 for the v5 model. */
long long wf_param_count(int layers, int F, int L, int O);

/* graphBuilder.py:9-47 build_spatial_graph.  lats f64[nlat], lons f64[nlon] (device);
 * node id = ilat*nlon + ilon.  Writes edge_index i64[2, N*k]: row 0 = node (each k times),
 * row 1 = its k nearest neighbours ordered by (squared distance, index).  monotonic != 0
 * (both axes strictly monotonic) enables the exhaustive (2k+1)^2 stencil search. */
int wf_knn_grid_build(const double* lats, int nlat, const double* lons, int nlon, int k,
                      int monotonic, long long* edge_index, void* stream);

/* PyG gcn_norm (invoked by every GCNConv.forward: model.py:31-40, hybrid_model.py:65-74),
 * computed once per region over R = window*N rows: self loops dropped, one added per row,
 * deg = in-degree by target, w = deg[src]^-1/2 * deg[dst]^-1/2.  Emits A_hat as CSR by
 * target (rowptr i32[R+1], col/val capacity E+R) and A_hat^T as CSR by source. */
size_t wf_gcn_norm_workspace_bytes(long long E, int R);
int wf_gcn_norm_csr(const long long* edge_index, long long E, int R, int* rowptr, int* col,
                    float* val, int* rowptr_t, int* col_t, float* val_t, void* workspace,
                    size_t workspace_bytes, void* stream);

/* GCNConv + ReLU (model.py:31-42, hybrid_model.py:65-75): Y = relu((A_hat X) W^T + b).
 * Window w of X starts at element x_win_off[w] (or w*x_win_stride), rows x_ld apart.
 * rowptr == NULL means identity aggregation. */
int wf_gcn_layer_fwd(const float* X, int x_ld, long long x_win_stride, const long long* x_win_off,
                     const float* W, const float* bias, long long w_group_stride,
                     long long b_group_stride, const int* rowptr, const int* col, const float* val,
                     long long rowptr_group_stride, long long csr_group_stride, int R, int Cin,
                     int Cout, int G, int Bw, int relu, float* Y, void* stream);

/* Backward of the above (autograd through STGCN.forward, model.py:30-52).  dY is overwritten
 * with dY*(Y>0).  dX/dW/db may be NULL. */
size_t wf_gcn_layer_bwd_workspace_bytes(int R, int Cin, int Cout, int G, int Bw);
int wf_gcn_layer_bwd(const float* X, int x_ld, long long x_win_stride, const long long* x_win_off,
                     const float* Y, float* dY, const float* W, long long w_group_stride,
                     const int* rowptr, const int* col, const float* val, const int* rowptr_t,
                     const int* col_t, const float* val_t, long long rowptr_group_stride,
                     long long csr_group_stride, int R, int Cin, int Cout, int G, int Bw, int relu,
                     float* dX, float* dW, float* db, long long dw_group_stride,
                     long long db_group_stride, void* workspace, size_t workspace_bytes, void* stream);

/* nn.LSTM(F -> L, `layers`, batch_first) over every (task, window, node) sequence
 * (hybrid_model.py:42-49, 93-105).  x [G*Bw*R, F]; gates [layers][G*Bw*R,4L],
 * h, c [layers][G*Bw*R, L] are outputs kept for BPTT.  p_drop > 0: inter-layer dropout
 * (hybrid_model.py:47) with the masks of wf_dropout_apply (site 16 + layer); h_masked
 * [layers-1][G*Bw*R, L] receives the masked outputs the next layer reads (kept for BPTT). */
int wf_lstm_fwd(const float* x, const float* params, long long params_group_stride, int layers,
                int F, int L, int O, int T, int N, int G, int Bw, float* gates, float* h, float* c,
                float p_drop, const unsigned long long* rng, float* h_masked, void* stream);

/* BPTT (loss.backward(): train_hybrid_maml_v5.py:134,169; adapt_hybrid_v5.py:198). */
size_t wf_lstm_bwd_workspace_bytes(int layers, int F, int L, int T, int N, int G, int Bw);
int wf_lstm_bwd(const float* x, const float* params, long long params_group_stride, int layers,
                int F, int L, int O, int T, int N, int G, int Bw, float* gates, const float* h,
                const float* c, const float* dlast, float* grads, long long grads_group_stride,
                float p_drop, const unsigned long long* rng, const float* h_masked,
                void* workspace, size_t workspace_bytes, void* stream);

/* Linear head (hybrid_model.py:108-115): pred[G*Bw*N, O] = h_top[last step] W_o^T + b_o.
 * Row n of a window's pred, viewed as [H, 12], is the reference's rows n*H .. n*H+H-1. */
int wf_head_fwd(const float* h_top, const float* params, long long params_group_stride, int layers,
                int F, int L, int O, int T, int N, int G, int Bw, float* pred, void* stream);

/* nn.MSELoss per window (train_hybrid_maml_v5.py:133,167; adapt_hybrid_v5.py:197) and its
 * gradient seed scaled by grad_scale.  Targets: y [windows, N*O] or, when y == NULL, read in
 * place from feat with tgt_off[w] = element offset of features[idx+T+1] (dataset.py:40-48).
 * The node-major prediction buffer is compared flat against the horizon-major target buffer,
 * exactly as the reference does. */
int wf_mse_fwd_bwd(const float* pred, const float* y, const float* feat, const long long* tgt_off,
                   int feat_ld, int num_weather, int N, int O, int windows, float grad_scale,
                   float* loss, float* dpred, void* stream);

/* Head backward: dlast[G*Bw*N, L] = dpred W_o; grads (optional) gets dW_o, db_o. */
size_t wf_head_workspace_bytes(int L, int O, int N, int G, int Bw);
int wf_head_bwd(const float* dpred, const float* h_top, const float* params,
                long long params_group_stride, int layers, int F, int L, int O, int T, int N, int G,
                int Bw, float* dlast, float* grads, long long grads_group_stride, void* workspace,
                size_t workspace_bytes, void* stream);

/* clip_grad_norm_(max_norm) + SGD, one clip norm per task (train_hybrid_maml_v5.py:116-118,
 * 135-139).  max_norm <= 0 disables clipping. */
size_t wf_optim_workspace_bytes(int G);
int wf_clip_sgd_step(float* theta, long long theta_group_stride, const float* grad,
                     long long grad_group_stride, long long P, int G, float lr, float max_norm,
                     float* norms, void* workspace, size_t workspace_bytes, void* stream);

/* clip_grad_norm_ + AdamW (decoupled=1, train_hybrid_maml_v5.py:174-178,245-249) or Adam with
 * L2 (decoupled=0, adaptive_scheduler.py:89-93; adapt_hybrid_v5.py:200-201).  hyper_dev: 8
 * device floats {lr, beta1, beta2, eps, weight_decay, 1-beta1^t, 1-beta2^t, grad_scale}. */
int wf_clip_adam_step(float* theta, const float* grad, float* exp_avg, float* exp_avg_sq,
                      long long P, const float* hyper_dev, float max_norm, int decoupled,
                      float* norm_out, void* workspace, size_t workspace_bytes, void* stream);

/* dst[i] (+)= sum_g src[g*stride + i]: the per-task query gradients summed into the
 * meta-gradient buffer (train_hybrid_maml_v5.py:169 accumulates .grad across tasks). */
int wf_sum_groups(const float* src, long long src_group_stride, int G, long long P, float* dst,
                  int accumulate, void* stream);

/* dst = src - trunc_tf32(src): the `lo` half of the 3xTF32 operand split (csrc/wf_tc.cuh). */
int wf_split_lo(const float* src, float* dst, long long n, void* stream);

/* tcgen05 3xTF32 GEMM, C[g] = A[g] W[g]^T (+ bias + bias2, relu): A [G*rows_g, K], W/W_lo
 * [G][N, K] (group stride w_group_stride), C [G*rows_g, N]; K % 32 == 0, N % 128 == 0.
 * The dense contraction behind GCNConv.lin (model.py:23-26) and the LSTM input projections
 * (hybrid_model.py:42-49).  err: one device int, non-zero if a pipeline wait timed out. */
int wf_tc_gemm_nt(const float* A, int rows_g, int G, int K, const float* W, const float* W_lo,
                  long long w_group_stride, int N, const float* bias, const float* bias2,
                  long long bias_group_stride, int relu, float* C, int* err, void* stream);

/* ---- tensor-core (tcgen05, 3xTF32) variants of the hot entry points -------------------------
 * Same mathematics and buffers as the FP32 entry points above, FP32-class accuracy (~2e-6), plus
 * transposed activation copies [(G*Bw)][channels][R] that turn the weight-gradient products into
 * K-major contractions.  err: one device int, set non-zero if a pipeline wait timed out. */

/* Size of the pre-transposed weight buffer (W_hh^T per layer, W_ih^T for layers >= 1). */
long long wf_param_count_transposed(int layers, int F, int L, int O);

/* Operand staging after every update of the (fast) weights: lo halves + transposed copies. */
int wf_prep_weights_tc(const float* params, long long params_group_stride, int layers, int F, int L,
                       int O, int G, float* params_lo, float* paramsT, float* paramsT_lo, void* stream);

/* Row pitch RT of the transposed activation copies [(G*Bw)][channels][RT]: column (t, node) =
 * t*Np + node with Np = N rounded up to 4 (TMA box starts must be 16-byte aligned -- an unaligned
 * inner coordinate faults on B200), RT = T*Np.  Padding columns must be zero and are never written. */
long long wf_transposed_pitch(int T, int N);

/* GCNConv + ReLU (model.py:31-42, hybrid_model.py:65-75) with the neighbour aggregation fused
 * into the A-operand path of the tcgen05 GEMM.  X dense [G*Bw*R, Cin], Cin % 32 == 0,
 * Cout % 128 == 0, W shared by all groups, N nodes per time slice (R = T*N); YT / YT_lo optional. */
int wf_gcn_layer_fwd_tc(const float* X, const float* W, const float* W_lo, const float* bias,
                        const int* rowptr, const int* col, const float* val,
                        long long rowptr_group_stride, long long csr_group_stride, int R, int N, int Cin,
                        int Cout, int G, int Bw, int relu, float* Y, float* YT, float* YT_lo, int* err,
                        void* stream);

/* nn.LSTM forward (hybrid_model.py:42-49, 93-105); hT / hT_lo [layers][(G*Bw)][L][RT] optional. */
int wf_lstm_fwd_tc(const float* x, const float* params, const float* params_lo,
                   long long params_group_stride, int layers, int F, int L, int O, int T, int N, int G,
                   int Bw, float* gates, float* h, float* c, float* hT, float* hT_lo, int* err,
                   void* stream);

/* BPTT (train_hybrid_maml_v5.py:134,169).  xT / xT_lo: transposed layer-0 input [(G*Bw)][F][RT];
 * dgT: scratch [(G*Bw)][4L][RT] with zero padding columns. */
size_t wf_lstm_bwd_tc_workspace_bytes(int layers, int F, int L, int T, int N, int G, int Bw);
int wf_lstm_bwd_tc(const float* xT, const float* xT_lo, const float* paramsT, const float* paramsT_lo,
                   int layers, int F, int L, int O, int T, int N, int G, int Bw, float* gates,
                   const float* c, const float* hT, const float* hT_lo, float* dgT, const float* dlast,
                   float* grads, long long grads_group_stride, void* workspace, size_t workspace_bytes,
                   int* err, void* stream);

/* dW[g][M, N] = sum_w sum_k AT[g*Bw+w][m, a_k0+k] * BT[g*Bw+w][n, b_k0+k], k < klen: the weight
 * gradient dG^T X over transposed activation copies [(G*Bw)][rows][R] (test entry point). */
int wf_tc_wgrad(const float* AT, int M, const float* BT, const float* BT_lo, int N, int R, int Bw, int G,
                int a_k0, int b_k0, int klen, float* dW, long long dw_group_stride, int* err, void* stream);

/* ---- persistent LSTM recurrence (csrc/wf_lstm_seq.cu) ------------------------------------------
 * One launch per layer runs all T steps: a 2-CTA cluster owns a tile of 128 (task, window, node)
 * sequences, keeps the task's W_hh in shared memory as 16-bit hi/lo operands and exchanges h / partial
 * dh through distributed shared memory.  Activations private to this path (gates, cell state, dh
 * between layers) use the TB4 layout: per (window, step, 128-node tile) a block
 * [channels/4][128 rows][4 floats]. */

/* f32 elements of a TB4 buffer with `channels` values per (window, step, node). */
long long wf_tb4_elems(int channels, int T, int N, long long windows);

/* 16-bit elements of EACH of the four recurrent-operand buffers wf_prep_weights_seq writes. */
long long wf_seq_weight_elems(int layers, int L, int G);

/* fp32 -> 16-bit hi / lo operand halves (fmt 0: fp16, 1: bf16): hi = rn(x), lo = rn(x - hi); n % 4 == 0. */
int wf_split16(const float* src, void* hi, void* lo, long long n, int fmt, void* stream);

/* Row pitch RT16 of the 16-bit transposed activation copies [(G*Bw)][channels][RT16] and of the fp32 dG^T
 * scratch that pairs with them: column (t, node) = t*Np + node, Np = N rounded up to 8.  Padding columns zero. */
long long wf_transposed_pitch16(int T, int N);

/* Persistent tcgen05 GEMM with 16-bit hi/lo operand splits (csrc/wf_gemm16.cu), test / general entry point:
 * C[g] = A[g] W[g]^T (+ bias + bias2, relu); A [G*rows_g, K] fp32, W16 hi/lo [G][N, K] from wf_split16(W, fmt);
 * K % 64 == 0, N % 128 == 0. */
int wf_g16_gemm_nt(const float* A, int rows_g, int G, int K, const void* W16_hi, const void* W16_lo,
                   long long w_group_stride, int N, const float* bias, const float* bias2,
                   long long bias_group_stride, int relu, int fmt, float* C, int* err, void* stream);

/* GCNConv + ReLU (model.py:31-42, hybrid_model.py:65-75) on the fp16 hi/lo GEMM, neighbour aggregation fused into
 * the A-operand path.  X dense [G*Bw*R, Cin], Cin % 64 == 0, Cout % 128 == 0; W16 hi/lo = wf_split16(W, 0), shared
 * by all groups; YT hi/lo (optional): bf16 transposed copies [(G*Bw)][Cout][RT16] for the LSTM layer-0 dW.
 * gather_rows (optional, i32 [G][gather_max], -1 padded): rows of a window whose aggregation is not the unit self
 * loop; with agg (scratch f32 [G*Bw*R, Cin]) they are aggregated by a pre-pass instead of inside the GEMM.
 * x_win_off (optional): element offset of every window in X = a resident features tensor with x_rows_total rows of
 * Cin floats (dataset.py:36-37: a window is a contiguous slice); Cin % 8 == 0 suffices then (needs gather_rows + agg).
 * p_drop > 0: nn.Dropout after the ReLU (hybrid_model.py:67-73, model.py:33-42) fused into the epilogue, masks of
 * wf_dropout_apply for site `site` (the layer index) over Y [G*Bw*R, Cout].  err is also set (code 41) when an output
 * reaches the fp16 operand range limit of the next layer (|y| >= 32768 or non-finite). */
int wf_gcn_layer_fwd_g16(const float* X, const long long* x_win_off, long long x_rows_total,
                         const void* W16_hi, const void* W16_lo, const float* bias, const int* rowptr, const int* col, const float* val, long long rowptr_group_stride,
                         long long csr_group_stride, const int* gather_rows, int gather_max,
                         long long gather_group_stride, float* agg, int R, int N, int Cin, int Cout, int G, int Bw,
                         int relu, float* Y, void* YT_hi, void* YT_lo, float p_drop,
                         const unsigned long long* rng, int site, int* err, void* stream);

/* Group stride (16-bit elements) of the p16 (which = 0) / pT16 (which = 1) buffers below: the parameter counts
 * rounded up to 8 so that every group starts 16-byte aligned. */
long long wf_param_stride16(int layers, int F, int L, int O, int which);

/* Operand staging after every weight update: p16 = fp16 hi/lo of the flat parameters [G][stride16(0)]; pT16 = bf16
 * hi/lo of W_ih^T (layers >= 1) at the offsets of wf_param_count_transposed [G][stride16(1)]; f16 = W_hh fp16 hi/lo
 * [G][layers][4L][L]; b16 = W_hh^T regrouped per CTA rank, bf16 hi/lo [G][layers][2][L][2L]. */
int wf_prep_weights_seq(const float* params, long long params_group_stride, int layers, int F, int L,
                        int O, int G, void* p16_hi, void* p16_lo, void* pT16_hi, void* pT16_lo, void* f16_hi,
                        void* f16_lo, void* b16_hi, void* b16_lo, void* stream);

/* 16-bit elements of ONE plane of a TB8 buffer with `channels` values per (window, step, node).  TB8 is the layout of the
 * 16-bit hi / lo activations of the LSTM path: per (window, step, 128-node tile) a block [channels/8][128 rows][8 values];
 * an activation is two such planes (hi, lo) back to back.  The same bytes serve as a K-major operand (rows = M: input
 * projections, dX) and as an MN-major operand (rows = K: weight gradients) of the tensor cores -- nothing is transposed. */
long long wf_tb8_elems(int channels, int T, int N, long long windows);

/* nn.LSTM forward (hybrid_model.py:42-49, 93-105).  x16: layer-0 input as fp16 hi / lo planes, row-major
 * [2][G*Bw*T*N][F] (wf_gcn_layer_fwd_ss output, or wf_split16 of an fp32 tensor); gates (4L channels) and c (L channels)
 * are TB4 fp32, [layers] of them.  Hidden states leave as 16-bit hi / lo plane pairs in the TB8 layout:
 *   h16  [layers-1][2][wf_tb8_elems(L, ..)]  fp16: what the NEXT layer's input projection reads (masked when p_drop > 0);
 *   hb16 [layers][2][...]   (optional: training)  bf16: the plain h, operand of the weight gradients (tcgen05 kind::f16
 *        takes one format for both operands and dG needs bf16's exponent range);
 *   hb16m [layers-1][2][...] (p_drop > 0, training)  bf16: the masked h (dW_ih of the next layer).
 * ZERO-INITIALISE all three once: the padding rows of a node tile are never written and are contracted by the weight
 * gradients.  hlast [G*Bw*N, L] fp32: the top layer's last step (what the head reads).
 * p_drop > 0: inter-layer dropout (hybrid_model.py:47) applied by the recurrence kernel to what the next layer reads
 * (site 16 + layer, element ((z*T + t)*N + node)*L + unit). */
int wf_lstm_fwd_seq(const void* x16, const float* params, const void* p16_hi, const void* p16_lo,
                    long long params_group_stride, const void* f16_hi, const void* f16_lo, int layers,
                    int F, int L, int O, int T, int N, int G, int Bw, float* gates, void* h16, float* c,
                    float* hlast, void* hb16, float p_drop, const unsigned long long* rng, void* hb16m,
                    int* err, void* stream);

/* BPTT (train_hybrid_maml_v5.py:134,169; adapt_hybrid_v5.py:198) over the buffers of wf_lstm_fwd_seq.  xb16: the layer-0
 * input as BF16 hi / lo planes, row-major [2][G*Bw*T*N][F] (wf_gcn_layer_fwd_ss's Yb16, or wf_split16(.., fmt 1)).  dg16: scratch
 * [2][wf_tb8_elems(4L, ..)] holding one layer's dL/d(pre-activation) as bf16 hi / lo planes (TB8; zero-initialise once).
 * Per layer: one persistent recurrence launch, ONE pass over dG for [dW_ih | dW_hh] and both bias gradients
 * (contraction over rows, operands read MN-major from the activations' own layouts), dX for the layer below. */
size_t wf_lstm_bwd_seq_workspace_bytes(int layers, int F, int L, int T, int N, int G, int Bw);
int wf_lstm_bwd_seq(const void* xb16, const void* pT16_hi, const void* pT16_lo, const void* b16_hi,
                    const void* b16_lo, int layers, int F, int L, int O, int T, int N, int G, int Bw,
                    const float* gates, const float* c, const void* hb16, void* dg16, const float* dlast,
                    float* grads, long long grads_group_stride, float p_drop, const unsigned long long* rng,
                    const void* hb16m, void* workspace, size_t workspace_bytes, int* err, void* stream);

/* Single-layer recurrence launches (the persistent kernels wf_lstm_fwd_seq / wf_lstm_bwd_seq issue per layer), for
 * harnesses that time the dominant kernels alone or drive the layers themselves.  *_l pointers address ONE layer. */
int wf_lstm_seq_recur_fwd(float* gates_l, float* c_l, void* h16_l, void* hb16_l, float* hlast, const void* f16_hi,
                          const void* f16_lo, int layer, int layers, int L, int T, int N, int G, int Bw, int* err,
                          void* stream);
int wf_lstm_seq_recur_bwd(const float* gates_l, const float* c_l, void* dg16, const float* ext, int ext_is_dlast,
                          const void* b16_hi, const void* b16_lo, int layer, int layers, int L, int T, int N, int G,
                          int Bw, int* err, void* stream);

/* GCNConv + ReLU (+ train-mode dropout) on PRE-SPLIT fp16 hi / lo activations (model.py:31-42, hybrid_model.py:65-75;
 * csrc/wf_gemm_ss.cu): Y16 = split(dropout(relu((A_hat X) W^T + b))) with X16 / Y16 as planes [2][G*Bw][R][C].  The first
 * layer passes the fp32 windows instead (X32 with x_win_off[w] = element offset of window w in a resident features tensor,
 * dataset.py:36-37, or dense when x_win_off == NULL) plus a scratch xsplit16 [2][G*Bw][R][Cin].  Every row whose
 * aggregation is not the unit self loop must lie in the leading agg_rows rows of its window (a multiple of 128; the t = 0
 * slice for the reference's graphs): they are aggregated into side16 [2][G*Bw][agg_rows][Cin] first.
 * The weight slice stays resident in shared memory, activations stream through TMA, the output leaves through TMA stores.
 * Cin % 8 == 0, Cout % 128 == 0; W16 hi/lo = wf_split16(W, 0), shared by all groups; dropout / err as wf_gcn_layer_fwd_g16.
 * Yb16 (optional): the same output once more as bf16 hi / lo planes (xb16 of wf_lstm_bwd_seq). */
int wf_gcn_layer_fwd_ss(const float* X32, const long long* x_win_off, const void* X16, void* xsplit16,
                        const void* W16_hi, const void* W16_lo, const float* bias, const int* rowptr,
                        const int* col, const float* val, long long rowptr_group_stride,
                        long long csr_group_stride, int agg_rows, void* side16, int R, int Cin, int Cout, int G,
                        int Bw, int relu, void* Y16, void* Yb16, float p_drop, const unsigned long long* rng,
                        int site, int* err, void* stream);

/* loss.backward() through GCNConv + ReLU (+ Dropout) on the tensor cores (model.py:31-42 under autograd: STGCN.forward,
 * SURVEY.md D4; replaces wf_gcn_layer_bwd where the widths allow).  X [Bw*R][Cin], Y / dY [Bw*R][Cout], W [Cout][Cin]:
 * fp32 row-major; CSR by target of A_hat over the R rows of a window and its transpose (shared by the Bw windows).
 * dZ = dY * mask * (Y > 0) is written once as bf16 hi/lo planes and read by both products: dW = dZ^T (A_hat X), db = dZ^T 1
 * (weight-gradient kernel) and dX = A_hat^T (dZ W) (SS GEMM + transposed aggregation).  dX may be NULL.  Cout % 128 == 0,
 * Cin % 8 == 0 and <= 256 (dX: Cin 128 or 256).  The mask is regenerated from (p_drop, rng = {seed, pass} of the forward
 * call, site).  workspace: wf_gcn_layer_bwd_ss_workspace_bytes, 256-byte aligned. */
size_t wf_gcn_layer_bwd_ss_workspace_bytes(int R, int Cin, int Cout, int Bw);
int wf_gcn_layer_bwd_ss(const float* X, const float* Y, const float* dY, const float* W, const int* rowptr,
                        const int* col, const float* val, const int* rowptr_t, const int* col_t, const float* val_t,
                        int R, int Cin, int Cout, int Bw, int relu, float p_drop, const unsigned long long* rng,
                        int site, float* dX, float* dW, float* db, void* workspace, size_t workspace_bytes, int* err,
                        void* stream);

/* The two SS-mode kernels on operands given as plain 16-bit planes (test / general entry points).
 * wf_ss_nodes_gemm: C (TB4 fp32: per (window, step, node tile) a block [Ntot/4][128 rows][4]) = A W^T (+ bias + bias2);
 *   avar 0: A16 row-major planes [2][G*Bw*T][Nn][K], 1: TB8 planes; W16 hi / lo [G][Ntot][K]; bn = 64 / 128 / 256 columns
 *   per CTA with bn * K / k_parts <= 32768 (the weight slice is resident in shared memory); afmt / bfmt 0 fp16, 1 bf16;
 *   k_parts > 1 deals the K range over that many CTAs whose partial products are added into the (cleared) output.
 * wf_ss_wgrad: dst0 [G][512][w0] (+ dst1 [G][512][w1], db [G][512]) = dG^T [B0 | B1] summed over all blocks of a group;
 *   dg16: TB8 bf16 planes with 512 channels; half h of 128 columns: bvar 0 = TB8 bf16 planes with bC channels (bcol0:
 *   first channel), 1 = row-major bf16 planes [2][G*Bw*T][Nn][bC] (bcol0: first column); bshift 1 = pair dG of step t
 *   with B of step t - 1 (dW_hh).  part: scratch, part_floats >= 513 * 512 * G floats per split-K slice. */
int wf_ss_nodes_gemm(int bn, int avar, const void* A16, long long a_plane, int K, int afmt, const void* W16_hi,
                     const void* W16_lo, long long w_group_stride, int Ntot, int bfmt, const float* bias,
                     const float* bias2, long long bias_group_stride, float* C, int T, int Nn, int Bw, int G,
                     int k_parts, int* err, void* stream);
int wf_ss_wgrad(const void* dg16, long long dg_plane, int nh, const void* b0, long long b0_plane, int b0var,
                int b0shift, int b0col0, int b0C, const void* b1, long long b1_plane, int b1var, int b1shift,
                int b1col0, int b1C, int T, int Nn, int Bw, int G, float* part, long long part_floats, float* dst0,
                int ld0, int w0, float* dst1, int ld1, int w1, float* db, long long gstride, int* err,
                void* stream);

/* out[i] = float(hi[i]) + float(lo[i]) (fmt 0: fp16, 1: bf16): the fp32 value of a pair of operand planes. */
int wf_join16(const void* hi, const void* lo, long long n, int fmt, float* out, void* stream);

/* Nodes per node tile of the TB4 activation layout (<= 128 rows per tile in memory; the N nodes of a window are dealt
 * evenly over ceil(N / 128) tiles, rounded up to 8: 441 -> 112).  Node n of a window lives in tile n / wf_tile_rows(N),
 * row n % wf_tile_rows(N). */
int wf_tile_rows(int N);

/* ---- dropout (hybrid_model.py:47,58,67-73,108; model.py:27,33-42) ---------------------------------------------------
 * Counter-based masks (Philox4x32-10, csrc/wf_rng.cuh): element e of site s in forward pass c is kept with probability
 * 1 - p and scaled by 1 / (1 - p); nothing is stored, backward kernels regenerate the mask.  rng: two device words
 * {seed, pass counter}.  Sites: GCN layer i = i, LSTM layer l output = 16 + l, head input = 32.
 * wf_dropout_apply: out[r, c] = in[row r][c] * mask(site, r*cols + c); input row r lives at
 * in + (r / rows_per_blk)*in_blk_stride + (r % rows_per_blk)*in_ld; out dense [rows, cols] (may alias a dense input).
 * Applied to ones it returns the mask itself (tests). */
int wf_dropout_apply(const float* in, long long in_blk_stride, int rows_per_blk, int in_ld, long long rows,
                     int cols, float p, const unsigned long long* rng, int site, float* out, void* stream);
/* rng[1] += 1 on the stream (after a forward + backward pair): fresh masks for the next pass, also under graph replay. */
int wf_rng_advance(unsigned long long* rng, void* stream);

/* ---- feature assembly: prepare_model_input, featurePreprocessor.py:84-177 ----
 * weather: device f32 [time * N, 12] (time-major rows, the reference's reshape at :121-122), may hold NaN.
 * wf_feature_stats: per variable the NaN fill value (f32 nanmean, 0 if all NaN; :104-109), the mean and the population
 * standard deviation of the FILLED array over (time, nodes) in f64 (:133-136; the caller adds the 1e-8) and the NaN
 * count.  Outputs are DEVICE buffers fill[12], mean[12], stdev[12], nan_count[12]. */
size_t wf_feature_stats_workspace_bytes(long long rows);
int wf_feature_stats(const float* weather, long long rows, float* fill, double* mean, double* stdev,
                     long long* nan_count, void* workspace, size_t workspace_bytes, void* stream);
/* out[time * N, 24] = [ (x or fill - mean) / std | time features of the step | Koppen embedding row ], NaN -> 0
 * (:146, :164-180).  fill / mean / std / koppen are HOST arrays (12, 12, 12, 8 values); timefeat is device f32 [time, 4]
 * (embed_utils.py:9-27).  f64_arith = 1 reproduces the reference with statistics passed in (f64 arrays), 0 the
 * statistics it derives itself (f32 arrays). */
int wf_assemble_features(const float* weather, long long time_steps, int N, const float* fill_host,
                         const double* mean_host, const double* std_host, int normalize, int f64_arith,
                         const float* timefeat, const float* koppen_host, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
