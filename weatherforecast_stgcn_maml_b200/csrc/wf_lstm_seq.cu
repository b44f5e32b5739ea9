// Persistent LSTM recurrence kernels (sm_100a): one launch runs all T steps of one layer.
//
// Replaces the reference's per-node Python loop over nn.LSTM (hybrid_model.py:93-105: 441 calls
// per window) and autograd's BPTT (train_hybrid_maml_v5.py:134,169).  Sequences are independent,
// so a tile of 128 (task, window, node) sequences is owned by one 2-CTA cluster for the whole
// window; the recurrent weights of the tile's task stay in shared memory for all T steps and the
// per-step product runs on the tcgen05 tensor cores with 16-bit hi/lo operand splits
//     a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo   (fp32 accumulate in TMEM)
// fp16 in the forward pass (|h| < 1, |W| ~ 0.1: ~2^-20 relative), bf16 in the backward pass
// (gradients span the fp32 exponent range; ~2^-16 relative against a 1e-3 gradient tolerance).
//
// Forward (N split): CTA r of the pair holds W_hh rows of units [64r, 64r+64) x 4 gates
// (256 x 128, hi + lo = 128 KB), computes D[128 x 256] = h[t-1] W^T from shared memory (SS mode),
// applies the cell non-linearities and writes its 64 units of h[t] as fp16 hi/lo straight into
// the A-operand buffer of BOTH CTAs -- the peer's copy as st.async transactions that complete on the
// peer's mbarrier (no cluster-scope release on the sending warp) -- so the only per-step exchange is
// the operand itself.  The default kernel (wf_lstm_seq_fwd16_kernel, 16 cell warps) keeps every global
// store out of the dependent part of a step: results are staged in TMEM and written under the next
// step's MMA, which is itself issued K step by K step behind the cell mathematics (see the comment on
// that kernel and DESIGN.md section 4).
// Backward (K split): CTA r holds W_hh^T restricted to its own gate rows (128 x 256, 128 KB), keeps
// its own dG[t+1] (128 x 256) in TMEM as the A operand (TS mode) and produces a partial
// dh[128 x 128]; the half belonging to the peer's units goes through distributed shared memory
// (fp32, 32 KB per step as st.async transactions, double buffered).  The global copies of dG[t] are written
// after the hand-over, read back from the TMEM operand.
//
// Activation layout ("TB4", wf_layout.cuh): per (window, step, 128-node tile) a block of
// [channels / 4][128 rows][4 floats], so a warp whose lanes are consecutive rows moves 512
// contiguous bytes per float4 instruction in both the GEMM epilogues and the cell epilogues.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "wf_common.cuh"
#include "wf_layout.cuh"
#include "wf_rng.cuh"
#include "wf_tc.cuh"

using namespace wftc;

int wf_ss_launch_nodes(int bn, int avar, const void* A16, long long a_plane, int K, int afmt, const void* Bhi, const void* Blo,
                       int ldb, long long b_gstride, int Ntot, int bfmt, const float* bias, const float* bias2,
                       long long bias_gstride, float* C, int T, int Nn, int Bw, int G, const DropCfg* drop, int* err,
                       cudaStream_t st, int k_parts = 1);
int wf_ss_launch_wgrad(const void* dg16, long long dg_plane, int nh, const void* const* bsrc, const long long* bplane,
                       const int* bvar, const int* bshift, const int* bcol0, const int* bC, int T, int Nn, int Bw, int G,
                       float* part, size_t part_floats, float* dst0, int ld0, int w0, float* dst1, int ld1, int w1, float* db1,
                       float* db2, long long gstride, int* err, cudaStream_t st, int M = 512);
int wf_launch_split16(const float* src, long long src_gstride, void* hi, void* lo, long long dst_gstride, long long n, int G,
                      int fmt, cudaStream_t st);
int wf_launch_transpose_split16(const float* in, long long in_gstride, int rows, int cols, void* out_hi, void* out_lo,
                                long long out_gstride, int G, cudaStream_t st);
int wf_np(int N);
extern "C" int wf_tile_rows(int N);

namespace {

constexpr int SEQ_THREADS = 256;  // 8 warps = 4 TMEM lane quarters x 2 unit halves; warp 0 also owns TMEM, TMA and MMA issue
                                  // (9 warps would put 3 on one SM sub-partition and cap registers at 168)
constexpr int SEQ_SMEM = 196608 + 1024;

struct SeqArgs {
  float* XG;          // TB4, 4L channels. fwd: input projection in, activated gates out; bwd: gates in, dG out
  float* Cst;         // TB4, L channels: cell state
  // fwd out, all hi / lo plane pairs in the TB8 layout [2][block][L/8][128 rows][8] (optional each):
  uint16_t* H16;      //   fp16: what the NEXT layer's input projection reads (K-major operand) -- masked when dropout is on
  uint16_t* HB16;     //   bf16: the plain h, for dW_hh (and dW_ih of the next layer when dropout is off): MN-major operand
  uint16_t* HB16m;    //   bf16 (dropout on): the masked h, for dW_ih of the next layer
                      //   (tcgen05 kind::f16 wants both operands in ONE format, and dG needs bf16's exponent range)
  long long h16_plane;  // elements between the hi and the lo plane
  float* Hlast;       // fwd out (top layer): h of the last step, row-major [Z*Nn, L] (what the head reads)
  DropCfg drop;       // fwd: inter-layer dropout on this layer's output (hybrid_model.py:47); rng == nullptr: off
  uint16_t* DG16;     // bwd out: dG as bf16 hi / lo planes, TB8 [2][block][4L/8][128 rows][8]
  long long dg16_plane;
  const float* ext;   // bwd: dL/dh from above -- TB4 (L channels), or row-major dlast [Z*Nn, L] if ext_last_only
  int ext_last_only;
  int T, Nn, Bw, tpw, rpt;  // rpt: nodes per node tile (wf_tile_rows)
  int slab0, slab_g;  // weight map z coordinate = slab0 + g * slab_g (+ rank in the backward kernel)
  int* err;
};

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// asynchronous 16-byte store into the peer CTA's shared memory that signals `cbar` (an mbarrier of the SAME peer CTA)
// with the byte count when it lands: the data hand-over needs no release fence on the sending warp
__device__ __forceinline__ void st_async_v4(uint32_t caddr, uint4 v, uint32_t cbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];\n"
               ::"r"(caddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(cbar) : "memory");
}
// no data is published with this arrival (it only says "my tensor core is done reading"): no memory barrier
__device__ __forceinline__ void arrive_cluster_relaxed(uint32_t cbar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cbar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_cta() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
// non-blocking phase test (CTA-scope acquire)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// kind::f16 instruction descriptor: D = F32, A/B = fmt (0 F16, 1 BF16), both K-major, M = 128
__host__ __device__ constexpr uint32_t idesc_16(int n, uint32_t fmt) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_ss_16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts_16(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};\n"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}

__device__ __forceinline__ float fast_tanh(float x) { return 2.0f * __fdividef(1.0f, 1.0f + __expf(-2.0f * x)) - 1.0f; }

// packed conversions (F2FP, ALU pipe); the residuals ra, rb feed the lo half
__device__ __forceinline__ uint32_t pack_f16(float a, float b, float& ra, float& rb) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  ra = a - f.x;
  rb = b - f.y;
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b, float& ra, float& rb) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const uint32_t u = *reinterpret_cast<const uint32_t*>(&h);
  ra = a - __uint_as_float(u << 16);
  rb = b - __uint_as_float(u & 0xFFFF0000u);
  return u;
}

// ================================================================================= forward, 16 warps
// Arranged around the measured per-SM limits:
// an SM writes at most ~62 GB/s to L2 whatever the path (LSU or bulk copy), so the 224 KB a CTA stores per step cost
// 3.6 us and must not sit between the cell mathematics and the hand-over of h[t].  Per step:
//   phase A  D[t] (TMEM) + input projection -> gates, c, h;  h[t] -> both CTAs' A operand;  h / h^T stores;  the
//            activated gates go back into the accumulator columns they came from (tcgen05.st, 256 B/clk);
//   hand-over, MMA[t+1] into the OTHER accumulator buffer;
//   phase B  (under MMA[t+1] and the hand-over latency) gates TMEM -> global in place, cell state -> global.
// 16 warps: 4 TMEM lane quarters x 4 unit groups of 16, chunks of 4 units (one TB4 float4 per gate), <= 128 registers.
#ifdef WF_SEQ_TRACE
__device__ long long wf_seq_trace_buf[32 * 16 * 24];  // [step][warp][point], CTA 0 (tools/trace_fwd.py)
#define WF_TR(pt) do { if (blockIdx.x == 0 && lane == 0 && t < 32) wf_seq_trace_buf[(t * 16 + warp) * 24 + (pt)] = clock64(); } while (0)
#define WF_TRS(pt) do { if (blockIdx.x == 0 && lane == 0 && s < 32) wf_seq_trace_buf[(s * 16 + warp) * 24 + (pt)] = clock64(); } while (0)
#else
#define WF_TR(pt) do { } while (0)
#define WF_TRS(pt) do { } while (0)
#endif
__device__ __forceinline__ void st_async_v2(uint32_t caddr, uint32_t x, uint32_t y, uint32_t cbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];\n"
               ::"r"(caddr), "r"(x), "r"(y), "r"(cbar) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
// 2^(-x * scale), exponent clamped to 29: a product of four denominators 1 + 2^29 stays finite, and sigmoid(-20) = 2e-9
// is below the arithmetic's resolution anyway.  .ftz forms: one MUFU each, no denormal pre-scaling.
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float exp2_neg(float x, float scale) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(fminf(-x * scale, 29.0f)));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}

// h = o tanh(c) for four units from d = 1 + e^(-2c): tanh = 2 / d - 1 with one reciprocal for the four denominators
__device__ __forceinline__ void hidden4(const float* go, const float* dc, float* hh) {
  const float p1 = dc[0] * dc[1], p2 = dc[2] * dc[3], rr = rcp_approx(p1 * p2);
  const float r1 = rr * p2, r2 = rr * p1;
  hh[0] = go[0] * fmaf(2.0f, r1 * dc[1], -1.0f);
  hh[1] = go[1] * fmaf(2.0f, r1 * dc[0], -1.0f);
  hh[2] = go[2] * fmaf(2.0f, r2 * dc[3], -1.0f);
  hh[3] = go[3] * fmaf(2.0f, r2 * dc[2], -1.0f);
}

template <bool DROP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512, 1)
wf_lstm_seq_fwd16_kernel(const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo, const SeqArgs a) {
  constexpr int L = 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* b_hi = smem;            // [2 k-blocks][256 gate rows][128 B]
  uint8_t* b_lo = smem + 65536;
  uint8_t* a_hi = smem + 131072;   // [2 k-blocks][128 rows][128 B]
  uint8_t* a_lo = smem + 163840;
  __shared__ uint64_t wfull, a_ready[4], issued[3], dfull, peer_done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank(), peer = rank ^ 1u;
  const int tile = blockIdx.x >> 1;
  const int z = tile / a.tpw, nt = tile - z * a.tpw, node0 = nt * a.rpt, g = z / a.Bw;
  const int T = a.T;

  if (tid == 0) {
    // a_ready[c], one per chunk of 16 units (= one K step of 16 per k-block): 16 local warps + the expect_tx arrival;
    // the peer's 16 units of h[t] arrive as 8 KB of st.async transactions
    mbar_init(&wfull, 1); mbar_init(&dfull, 4); mbar_init(&peer_done, 1);  // dfull: one commit per K-step issuer
    for (int c = 0; c < 4; ++c) mbar_init(&a_ready[c], 17);
    for (int c = 0; c < 3; ++c) mbar_init(&issued[c], 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmWhi); tma_prefetch_desc(&tmWlo);
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);  // two accumulator buffers of 256 columns
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers exist before anybody arrives on them remotely
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  if (warp == 0 && lane == 0) {  // recurrent weights of this tile's task: resident for all T steps
    if (T > 1)
      for (int c = 0; c < 4; ++c) mbar_expect_tx(&a_ready[c], 8192);  // phase 0: the peer's half of h[0]
    const int slab = a.slab0 + g * a.slab_g;
    mbar_expect_tx(&wfull, 131072);
    for (int kb = 0; kb < 2; ++kb) {
      tma_load_4d(b_hi + kb * 32768, &tmWhi, &wfull, kb * 64, 64 * (int)rank, 0, slab);
      tma_load_4d(b_lo + kb * 32768, &tmWlo, &wfull, kb * 64, 64 * (int)rank, 0, slab);
    }
  }
  const int q = warp & 3, ug = warp >> 2;
  const int r = q * 32 + lane;            // tile row == TMEM lane
  const int ub = ug * 4;                  // chunk c: units 16 c + ub .. + 4 of the CTA's 64, so that ALL warps finish
  const int u0 = 64 * (int)rank + ub;     // K step c of h[t] together and MMA[t+1] can start behind chunk 0
  const int node = node0 + r;
  const bool valid = r < a.rpt && node < a.Nn;
  const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16) + (uint32_t)ub;
  const long long R = (long long)T * a.Nn;
  float4* const xg4 = reinterpret_cast<float4*>(a.XG);
  float4* const c4 = reinterpret_cast<float4*>(a.Cst);
  const uint32_t pd_remote = mapa_u32(smem_u32(&peer_done), peer);
  const uint32_t ar_remote = mapa_u32(smem_u32(&a_ready[0]), peer);
  bool ok = true;

  float cst[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) cst[j] = 0.f;
  DropState dstate;
  if (DROP) dstate = wf_drop_state(a.drop);
  float4 xq[2][4];  // two chunks of 4 units in flight: one float4 per gate
  auto xg_index = [&](int t, int c, int gate) -> long long {
    const long long blk = ((long long)z * T + t) * a.tpw + nt;
    return (blk * 128 + gate * 32 + (u0 >> 2) + 4 * c) * 128 + r;
  };
  auto load_chunk = [&](int t, int c, float4* dst) {
#pragma unroll
    for (int gate = 0; gate < 4; ++gate)  // padding rows of a tile: no traffic
      dst[gate] = valid ? xg4[xg_index(t, c, gate)] : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  load_chunk(0, 0, xq[0]);
  load_chunk(0, 1, xq[1]);

  // MMA[t+1] = h[t] W_hh^T, one K step of 16 units (x 2 k-blocks x 3 hi/lo products = 6 instructions) at a time, as soon
  // as that slice of h[t] is complete in BOTH CTAs: only the last K step is left when phase A ends.  tcgen05.mma blocks
  // the issuing thread while the pipe is busy (~128 cycles per instruction here), so the four K steps are issued by four
  // different warps (K step j by warp j), chained through issued[j] so that they enter the pipe in order.
  bool mdone = true;  // warps 0-3: my K step of the pending MMA has been issued (or none is pending)
  auto mma_step = [&](int t, bool block) {
    if (mdone) return;
    const int ks = warp;
    if (ks == 0 && t == 0 && ok && !mbar_wait(&wfull, 0)) { ok = false; if (lane == 0) atomicExch(a.err, 11); }
    if (block) {
      // the peer's slice arrived as st.async transactions counted by this barrier: a CTA-scope acquire is enough
      if (ok && !mbar_wait(&a_ready[ks], t & 1)) { ok = false; if (lane == 0) atomicExch(a.err, 12); }
      if (ks > 0 && ok && !mbar_wait(&issued[ks - 1], t & 1)) { ok = false; if (lane == 0) atomicExch(a.err, 15); }
    } else {
      bool rdy = mbar_test(&a_ready[ks], t & 1);
      if (ks > 0) rdy = rdy && mbar_test(&issued[ks - 1], t & 1);
      if (!__all_sync(0xffffffffu, rdy)) return;
    }
    if (lane == 0 && t + 2 < T) mbar_expect_tx(&a_ready[ks], 8192);  // next phase: the peer's slice of h[t+1]
    fence_proxy_async_cta();
    tc_fence_after();
    if (lane == 0 && ok) {
      const uint32_t idesc = idesc_16(256, 0);
      const uint32_t dnext = tbase + (uint32_t)((t + 1) & 1) * 256u;
#pragma unroll
      for (int p = 0; p < 3; ++p) {  // h_hi W_hi, h_lo W_hi, h_hi W_lo
        const uint32_t as = smem_u32(p == 1 ? a_lo : a_hi), bs = smem_u32(p == 2 ? b_lo : b_hi);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
          umma_ss_16(dnext, umma_desc_k_sw128(as + kb * 16384 + ks * 32), umma_desc_k_sw128(bs + kb * 32768 + ks * 32),
                     idesc, (ks | p | kb) != 0 ? 1u : 0u);
      }
      umma_commit(&dfull);  // a commit tracks the issuing thread's own instructions: D[t+1] is complete after all four
      if (ks < 3) { tc_fence_before(); mbar_arrive(&issued[ks]); }
    }
    __syncwarp();
    mdone = true;
  };

  for (int t = 0; t < T; ++t) {
    const uint32_t dcol = tlane + (uint32_t)(t & 1) * 256u;  // this step's accumulator buffer (and gate staging)
    if (t > 0) {
      if (ok && !mbar_wait(&dfull, (t - 1) & 1)) { ok = false; if (lane == 0) atomicExch(a.err, 13); }
      tc_fence_after();
      if (warp == 0 && lane == 0) arrive_cluster_relaxed(pd_remote);  // my MMA no longer reads my A buffer
    }
    WF_TR(0);
    mdone = !(warp < 4 && t + 1 < T);
    const long long blk = ((long long)z * T + t) * a.tpw + nt;
    const long long hrow = ((long long)z * R + (long long)t * a.Nn + node) * L + u0;
    // ------------------------------------------------------------ phase A: cell mathematics, h[t] out
    uint32_t acc[4][4];
    auto issue_acc = [&](int c) {
      __syncwarp();
#pragma unroll
      for (int gate = 0; gate < 4; ++gate) tmem_ld4(dcol + gate * 64 + 16 * c, acc[gate]);
    };
    if (t > 0) issue_acc(0);
    else {
#pragma unroll
      for (int gate = 0; gate < 4; ++gate)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[gate][j] = 0u;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (t > 0) tmem_wait_ld();
      WF_TR(8 + 4 * c);
      const float4* x = xq[c & 1];
      const float xi[4] = {x[0].x, x[0].y, x[0].z, x[0].w}, xf[4] = {x[1].x, x[1].y, x[1].z, x[1].w};
      const float xgv[4] = {x[2].x, x[2].y, x[2].z, x[2].w}, xo[4] = {x[3].x, x[3].y, x[3].z, x[3].w};
      uint32_t gt[4][4];
      float hh[4], go[4], dc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // i, f, o = 1 / (1 + e^-p), g = 2 / (1 + e^-2p) - 1 with ONE reciprocal for the four denominators
        const float di = 1.0f + exp2_neg(xi[j] + __uint_as_float(acc[0][j]), kLog2e);
        const float df = 1.0f + exp2_neg(xf[j] + __uint_as_float(acc[1][j]), kLog2e);
        const float dg = 1.0f + exp2_neg(xgv[j] + __uint_as_float(acc[2][j]), 2.0f * kLog2e);
        const float dq = 1.0f + exp2_neg(xo[j] + __uint_as_float(acc[3][j]), kLog2e);
        const float p1 = di * df, p2 = dg * dq, rr = rcp_approx(p1 * p2);
        const float r1 = rr * p2, r2 = rr * p1;  // 1 / (di df), 1 / (dg dq)
        const float gi = r1 * df, gf = r1 * di, gg = fmaf(2.0f, r2 * dq, -1.0f);
        go[j] = r2 * dg;
        const float cc = fmaf(gf, cst[4 * c + j], gi * gg);
        cst[4 * c + j] = cc;
        dc[j] = 1.0f + exp2_neg(cc, 2.0f * kLog2e);
        gt[0][j] = __float_as_uint(gi); gt[1][j] = __float_as_uint(gf);
        gt[2][j] = __float_as_uint(gg); gt[3][j] = __float_as_uint(go[j]);
      }
      hidden4(go, dc, hh);
      __syncwarp();
#pragma unroll
      for (int gate = 0; gate < 4; ++gate) tmem_st4(dcol + gate * 64 + 16 * c, gt[gate]);
      if (t > 0 && c < 3) issue_acc(c + 1);  // in flight under this chunk's stores
      WF_TR(9 + 4 * c);
      if (c < 2) load_chunk(t, c + 2, xq[c & 1]);  // chunks 2, 3 of this step
      if (t + 1 < T) {
        // h[t] as fp16 hi/lo -> the A operand of step t+1 in both CTAs (k-block `rank`, 16-byte chunk (16c + ub) / 8)
        float ra, rb, rc, rd, d0, d1;
        const uint32_t h0 = pack_f16(hh[0], hh[1], ra, rb), h1 = pack_f16(hh[2], hh[3], rc, rd);
        const uint32_t l0 = pack_f16(ra, rb, d0, d1), l1 = pack_f16(rc, rd, d0, d1);
        const uint32_t off = rank * 16384u + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u +
                             ((uint32_t)((2 * c + (ug >> 1)) ^ (r & 7)) << 4) + 8u * (uint32_t)(ug & 1);
        if (c == 0 && t > 0) {  // the peer's MMA of this step must be done with the peer's A buffer (no data is
          // acquired here, so a CTA-scope wait: the cluster-scope form costs a CCTL.IVALL per warp and step)
          if (ok && !mbar_wait(&peer_done, (t - 1) & 1)) { ok = false; if (lane == 0) atomicExch(a.err, 14); }
        }
        sts64(smem_u32(a_hi + off), h0, h1);
        sts64(smem_u32(a_lo + off), l0, l1);
        st_async_v2(mapa_u32(smem_u32(a_hi + off), peer), h0, h1, ar_remote + 8u * c);
        st_async_v2(mapa_u32(smem_u32(a_lo + off), peer), l0, l1, ar_remote + 8u * c);
        // K step c of h[t] is complete in this warp: my generic-proxy operand writes -> visible to the tensor core
        // (async proxy), my TMEM reads of the other accumulator buffer (phase B of step t-1) precede MMA[t+1]
        fence_proxy_async_cta();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[c]);
        if (warp < 4 && c > warp) mma_step(t, false);  // my K step, if its operands have arrived everywhere
      }
      WF_TR(10 + 4 * c);
    }
    tmem_wait_st();  // the staged gates are in TMEM before phase B reads them back
    WF_TR(1);
    if (t + 1 < T) {
      load_chunk(t + 1, 0, xq[0]);  // ahead of phase B's stores: the memory pipeline is in order
      load_chunk(t + 1, 1, xq[1]);
      WF_TR(2);
      if (warp < 4) mma_step(t, true);
      WF_TR(4);
    }
    // ------------------------------------------------------------ phase B: gates (TMEM) and cell state -> global
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t gt[4][4];
      __syncwarp();
#pragma unroll
      for (int gate = 0; gate < 4; ++gate) tmem_ld4(dcol + gate * 64 + 16 * c, gt[gate]);
      tmem_wait_ld();
      // h[t] = o tanh(c[t]) again (bit-identical to phase A: same function, same inputs) instead of 16 more registers
      float hh[4];
      {
        float go[4], dc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          go[j] = __uint_as_float(gt[3][j]);
          dc[j] = 1.0f + exp2_neg(cst[4 * c + j], 2.0f * kLog2e);
        }
        hidden4(go, dc, hh);
      }
      if (valid) {
#pragma unroll
        for (int gate = 0; gate < 4; ++gate)
          xg4[xg_index(t, c, gate)] = make_float4(__uint_as_float(gt[gate][0]), __uint_as_float(gt[gate][1]),
                                                  __uint_as_float(gt[gate][2]), __uint_as_float(gt[gate][3]));
        c4[(blk * 32 + (u0 >> 2) + 4 * c) * 128 + r] = make_float4(cst[4 * c], cst[4 * c + 1], cst[4 * c + 2], cst[4 * c + 3]);
        // h[t] leaves as 16-bit hi/lo planes in the TB8 layout (8 bytes per plane here; the warp of the neighbouring unit
        // group fills the other half of each 16-byte chunk).  Inter-layer dropout (hybrid_model.py:47): the NEXT layer
        // reads mask * h / (1 - p); the recurrence of this layer (the operand handed over in phase A) and dW_hh keep h.
        const int unit = u0 + 16 * c;
        const long long h16o = ((blk * (L / 8) + (unit >> 3)) * 128 + r) * 8 + (unit & 7);
        float hm[4] = {hh[0], hh[1], hh[2], hh[3]};
        if (DROP) {
          float m[4];
          wf_drop4(dstate, (unsigned long long)(hrow + 16 * c) >> 2, m);
#pragma unroll
          for (int j = 0; j < 4; ++j) hm[j] *= m[j];
        }
        if (a.H16 != nullptr) {
          float ra, rb, rc, rd, d0, d1;
          const uint32_t h0 = pack_f16(hm[0], hm[1], ra, rb), h1 = pack_f16(hm[2], hm[3], rc, rd);
          const uint32_t l0 = pack_f16(ra, rb, d0, d1), l1 = pack_f16(rc, rd, d0, d1);
          *reinterpret_cast<uint2*>(a.H16 + h16o) = make_uint2(h0, h1);
          *reinterpret_cast<uint2*>(a.H16 + a.h16_plane + h16o) = make_uint2(l0, l1);
        }
        if (a.HB16 != nullptr) {
          float ra, rb, rc, rd, d0, d1;
          const uint32_t h0 = pack_bf16(hh[0], hh[1], ra, rb), h1 = pack_bf16(hh[2], hh[3], rc, rd);
          const uint32_t l0 = pack_bf16(ra, rb, d0, d1), l1 = pack_bf16(rc, rd, d0, d1);
          *reinterpret_cast<uint2*>(a.HB16 + h16o) = make_uint2(h0, h1);
          *reinterpret_cast<uint2*>(a.HB16 + a.h16_plane + h16o) = make_uint2(l0, l1);
        }
        if (DROP && a.HB16m != nullptr) {
          float ra, rb, rc, rd, d0, d1;
          const uint32_t h0 = pack_bf16(hm[0], hm[1], ra, rb), h1 = pack_bf16(hm[2], hm[3], rc, rd);
          const uint32_t l0 = pack_bf16(ra, rb, d0, d1), l1 = pack_bf16(rc, rd, d0, d1);
          *reinterpret_cast<uint2*>(a.HB16m + h16o) = make_uint2(h0, h1);
          *reinterpret_cast<uint2*>(a.HB16m + a.h16_plane + h16o) = make_uint2(l0, l1);
        }
        if (a.Hlast != nullptr && t == T - 1)
          *reinterpret_cast<float4*>(a.Hlast + ((long long)z * a.Nn + node) * L + unit) = make_float4(hh[0], hh[1], hh[2], hh[3]);
      }
    }
    WF_TR(5);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
  cluster_sync_all();  // nobody exits while the peer may still address this CTA's shared memory
}

// ================================================================================= backward
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SEQ_THREADS, 1)
wf_lstm_seq_bwd_kernel(const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo, const SeqArgs a) {
  constexpr int L = 128;
  constexpr uint32_t A_HI = 128, A_LO = 256;  // TMEM columns: D [0,128), dG hi [128,256), dG lo [256,384)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* b_hi = smem;            // [4 k-blocks][128 unit rows][128 B]
  uint8_t* b_lo = smem + 65536;
  uint8_t* xbuf = smem + 131072;   // 2 x [2 halves][8 quads][128 rows][16 B]: the peer's partial dh for my units
  __shared__ uint64_t wfull, a_ready, issued, dfull, x_ready[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank(), peer = rank ^ 1u;
  const int tile = blockIdx.x >> 1;
  const int z = tile / a.tpw, nt = tile - z * a.tpw, node0 = nt * a.rpt, g = z / a.Bw;
  const int T = a.T;

  if (tid == 0) {
    mbar_init(&wfull, 1); mbar_init(&a_ready, 8); mbar_init(&dfull, 2); mbar_init(&issued, 1); mbar_init(&x_ready[0], 1); mbar_init(&x_ready[1], 1);  // 1 = the expect_tx arrival; data = 32 KB of st.async
    mbar_fence_init();
    tma_prefetch_desc(&tmWhi); tma_prefetch_desc(&tmWlo);
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  if (warp == 0 && lane == 0) {
    const int slab = a.slab0 + g * a.slab_g + (int)rank;
    mbar_expect_tx(&wfull, 131072);
    for (int kb = 0; kb < 4; ++kb) {
      tma_load_3d(b_hi + kb * 16384, &tmWhi, &wfull, kb * 64, 0, slab);
      tma_load_3d(b_lo + kb * 16384, &tmWlo, &wfull, kb * 64, 0, slab);
    }
  }
  {
    const int q = warp & 3, half = warp >> 2;
    const int r = q * 32 + lane;
    const int ub = half * 32;
    const int u0 = 64 * (int)rank + ub;
    const int node = node0 + r;
    const bool valid = r < a.rpt && node < a.Nn;
    const float vm = valid ? 1.0f : 0.0f;
    const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
    const float4* const xg4 = reinterpret_cast<const float4*>(a.XG);
    const float4* const c4 = reinterpret_cast<const float4*>(a.Cst);
    const float4* const e4 = reinterpret_cast<const float4*>(a.ext);
    const uint32_t xr_remote0 = mapa_u32(smem_u32(&x_ready[0]), peer), xr_remote1 = mapa_u32(smem_u32(&x_ready[1]), peer);
    const uint32_t xbuf_remote = mapa_u32(smem_u32(xbuf), peer);
    bool ok = true;

    float dcs[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) dcs[j] = 0.f;
    // per chunk: gates 8 float4, c[t] 2, c[t-1] 2, ext 2
    float4 ld[2][14];
    auto blk_of = [&](int t) -> long long { return ((long long)z * T + t) * a.tpw + nt; };
    auto load_chunk = [&](int t, int c, float4* dst) {
      const long long blk = blk_of(t);
      const int uq = (u0 + 8 * c) >> 2;
      if (!valid) {  // padding rows of a tile: no traffic, zero gates (their dG is masked by vm anyway)
#pragma unroll
        for (int i2 = 0; i2 < 14; ++i2) dst[i2] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
      }
#pragma unroll
      for (int gate = 0; gate < 4; ++gate) {
        dst[gate * 2] = xg4[(blk * 128 + gate * 32 + uq) * 128 + r];
        dst[gate * 2 + 1] = xg4[(blk * 128 + gate * 32 + uq + 1) * 128 + r];
      }
      dst[8] = c4[(blk * 32 + uq) * 128 + r];
      dst[9] = c4[(blk * 32 + uq + 1) * 128 + r];
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t > 0) {
        const long long blkp = blk_of(t - 1);
        dst[10] = c4[(blkp * 32 + uq) * 128 + r];
        dst[11] = c4[(blkp * 32 + uq + 1) * 128 + r];
      } else {
        dst[10] = zero; dst[11] = zero;
      }
      if (a.ext_last_only) {
        if (t == T - 1 && valid) {
          const float4* p = reinterpret_cast<const float4*>(a.ext + ((long long)z * a.Nn + node) * L + u0 + 8 * c);
          dst[12] = p[0]; dst[13] = p[1];
        } else {
          dst[12] = zero; dst[13] = zero;
        }
      } else {
        dst[12] = e4[(blk * 32 + uq) * 128 + r];
        dst[13] = e4[(blk * 32 + uq + 1) * 128 + r];
      }
    };
    load_chunk(T - 1, 0, ld[0]);

    // MMA of step s: partial dh[128 x 128] = dG[t+1][:, my gate rows] W_hh[my gate rows, :], 48 instructions.  tcgen05.mma
    // blocks the issuing thread for about as long as the queued instructions run, and the issuer is a cell warp with its
    // own deferred stores still to do, so the work is split: warp 0 issues the k-blocks of gates i, f right after the
    // hand-over, warp 4 (the other warp on TMEM lanes 0-31) those of g, o after the first half of ITS stores, chained
    // through `issued` so that they enter the pipe in order; both commit on dfull (count 2).  Warp 0 also arms the barrier
    // on which the peer's partial dh for my units will arrive as st.async transactions (phase (s - 1) / 2 of x_ready[s & 1]).
    auto issue_mma = [&](int s, int part) {
      if (part == 0 && lane == 0) mbar_expect_tx(&x_ready[s & 1], 32768);
      if (ok && s == 1 && !mbar_wait(&wfull, 0)) { ok = false; if (lane == 0) atomicExch(a.err, 21); }
      if (ok && !mbar_wait(&a_ready, (s - 1) & 1)) { ok = false; if (lane == 0) atomicExch(a.err, 22); }
      if (part == 1 && ok && !mbar_wait(&issued, (s - 1) & 1)) { ok = false; if (lane == 0) atomicExch(a.err, 25); }
      tc_fence_after();
      if (lane == 0 && ok) {
        const uint32_t idesc = idesc_16(128, 1);
#pragma unroll
        for (int p = 0; p < 3; ++p) {  // dG_hi W_hi, dG_lo W_hi, dG_hi W_lo
          const uint32_t ac = tbase + (p == 1 ? A_LO : A_HI), bs = smem_u32(p == 2 ? b_lo : b_hi);
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const int kb = 2 * part + kk;
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16)
              umma_ts_16(tbase, ac + kb * 32 + k16 * 8, umma_desc_k_sw128(bs + kb * 16384 + k16 * 32), idesc,
                         (part | p | kk | k16) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&dfull);
        if (part == 0) { tc_fence_before(); mbar_arrive(&issued); }
      }
      __syncwarp();
    };

    for (int s = 0; s < T; ++s) {
      const int t = T - 1 - s;
      const int xb = s & 1;
      if (s > 0) {
        // the peer's partial dh for my units arrives as st.async transactions on x_ready[xb] (phase (s - 1) / 2)
        if (ok && !mbar_wait(&dfull, (s - 1) & 1)) { ok = false; if (lane == 0) atomicExch(a.err, 23); }
        tc_fence_after();
        WF_TRS(0);
        // ---- the partial dh of the peer's units -> peer (same (row, half, quad) slot the peer's twin thread reads)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[8];
          __syncwarp();
          tmem_ld8(tlane + 64 * peer + ub + 8 * c, v);
          tmem_wait_ld();
          const uint32_t dst = xbuf_remote + (uint32_t)xb * 32768u + (uint32_t)(((half * 8 + 2 * c) * 128 + r) * 16);
          st_async_v4(dst, make_uint4(v[0], v[1], v[2], v[3]), xb ? xr_remote1 : xr_remote0);
          st_async_v4(dst + 2048u, make_uint4(v[4], v[5], v[6], v[7]), xb ? xr_remote1 : xr_remote0);
        }
        WF_TRS(1);
        // the data arrive as st.async transactions counted by this barrier: a CTA-scope acquire is enough
        if (ok && !mbar_wait(&x_ready[xb], ((s - 1) >> 1) & 1)) { ok = false; if (lane == 0) atomicExch(a.err, 24); }
        WF_TRS(2);
      }
      const long long blk = blk_of(t);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        // prefetch the next chunk (possibly of the next step) while this one is processed
        {
          const int cn = (c + 1) & 3, sn = s + ((c + 1) >> 2);
          if (sn < T) load_chunk(T - 1 - sn, cn, ld[(c + 1) & 1]);
        }
        float dh[8];
        if (s > 0) {
          uint32_t v[8];
          __syncwarp();
          tmem_ld8(tlane + 64 * rank + ub + 8 * c, v);
          tmem_wait_ld();
          const uint32_t src = smem_u32(xbuf + xb * 32768 + ((half * 8 + 2 * c) * 128 + r) * 16);
          const float4 p0 = lds128(src), p1 = lds128(src + 2048);
          dh[0] = __uint_as_float(v[0]) + p0.x; dh[1] = __uint_as_float(v[1]) + p0.y;
          dh[2] = __uint_as_float(v[2]) + p0.z; dh[3] = __uint_as_float(v[3]) + p0.w;
          dh[4] = __uint_as_float(v[4]) + p1.x; dh[5] = __uint_as_float(v[5]) + p1.y;
          dh[6] = __uint_as_float(v[6]) + p1.z; dh[7] = __uint_as_float(v[7]) + p1.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) dh[j] = 0.f;
        }
        const float4* x = ld[c & 1];
        float di[8], df[8], dg[8], dO[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int hsel = j >> 2, cmp = j & 3;
          auto pick = [&](const float4& v4) { return cmp == 0 ? v4.x : cmp == 1 ? v4.y : cmp == 2 ? v4.z : v4.w; };
          const float vi = pick(x[0 + hsel]), vf = pick(x[2 + hsel]), vg = pick(x[4 + hsel]), vo = pick(x[6 + hsel]);
          const float vc = pick(x[8 + hsel]), vp = pick(x[10 + hsel]), ve = pick(x[12 + hsel]);
          const float dhj = (dh[j] + ve) * vm;
          const float tc = fast_tanh(vc);
          const float dc = dcs[8 * c + j] + dhj * vo * (1.f - tc * tc);
          dO[j] = dhj * tc * vo * (1.f - vo);
          di[j] = dc * vg * vi * (1.f - vi);
          df[j] = dc * vp * vf * (1.f - vf);
          dg[j] = dc * vi * (1.f - vg * vg);
          dcs[8 * c + j] = dc * vf;
        }
        {
          // dG[t] as bf16 hi/lo -> TMEM A operand of the next step: k = gate*64 + (ub + 8c + j), two k per column.
          // The global copies of dG[t] are read back from here AFTER the hand-over (hi + lo is exactly what the GEMMs
          // downstream split it into again), so nothing is stored to global memory in this loop: an SM writes ~62 GB/s
          // at most, and the 256 KB of a step would otherwise sit on the recurrence's critical path for 4 us.
          const float* gsrc[4] = {di, df, dg, dO};
#pragma unroll
          for (int gate = 0; gate < 4; ++gate) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float ra, rb, d0, d1;
              hi[i] = pack_bf16(gsrc[gate][2 * i], gsrc[gate][2 * i + 1], ra, rb);
              lo[i] = pack_bf16(ra, rb, d0, d1);
            }
            const uint32_t col = (uint32_t)((gate * 64 + ub + 8 * c) >> 1);
            tmem_st4(tlane + A_HI + col, hi);
            tmem_st4(tlane + A_LO + col, lo);
          }
        }
        WF_TRS(3 + c);
      }
      tmem_wait_st();
      WF_TRS(7);
      if (s + 1 < T) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready);
        if (warp == 0) issue_mma(s + 1, 0);  // the first half of the MMA before warp 0's own stores
        WF_TRS(8);  // warp 0 feeds the tensor core first, then stores like everybody else
      }
      // ---- deferred stores of step t (under MMA[s+1]): dG leaves exactly as it sits in the TMEM operand -- bf16 hi / lo
      // planes, TB8 layout: per (gate, 8 units) one 16-byte chunk per plane and row, 512 contiguous bytes per warp store,
      // no arithmetic.  dX and the weight gradients read these planes directly (K-major and MN-major views of the same bytes).
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        __syncwarp();
#pragma unroll
        for (int gate = 0; gate < 4; ++gate) {
          uint32_t hi[4], lo[4];
          const uint32_t col = (uint32_t)((gate * 64 + ub + 8 * c) >> 1);
          tmem_ld4(tlane + A_HI + col, hi);
          tmem_ld4(tlane + A_LO + col, lo);
          tmem_wait_ld();
          if (valid) {
            const long long o = ((blk * (4 * L / 8) + ((gate * L + u0 + 8 * c) >> 3)) * 128 + r) * 8;
            *reinterpret_cast<uint4*>(a.DG16 + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(a.DG16 + a.dg16_plane + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        if (c == 1 && warp == 4 && s + 1 < T) issue_mma(s + 1, 1);  // the second half, between warp 4's store halves
      }
      WF_TRS(9);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
  cluster_sync_all();
}

// ================================================================================= helpers
// W_hh [4L, L] fp32 of every (task, layer) -> the operand copies the two kernels keep in shared memory:
//   f16_hi / f16_lo [G][layers][4L][L]            (forward: rows = gate rows, K = hidden units)
//   b16_hi / b16_lo [G][layers][2][L][2L]         (backward: per CTA rank, rows = hidden units n,
//                                                  K = gate*64 + (unit - 64*rank) over the rank's own gate rows)
// One block = a 32 (gate rows j) x 32 (hidden units n) tile: the forward copy keeps the orientation, the backward copy
// is its transpose, staged through shared memory so that both are written with consecutive lanes on consecutive addresses.
__global__ void __launch_bounds__(256) wf_prep_seq_kernel(const float* __restrict__ params, long long gstride, LstmLayout P,
                                                          int layers, int L, __half* __restrict__ f_hi, __half* __restrict__ f_lo,
                                                          __nv_bfloat16* __restrict__ b_hi, __nv_bfloat16* __restrict__ b_lo) {
  __shared__ float tile[32][33];
  const int g = blockIdx.z, l = blockIdx.y;
  const int tiles_n = L / 32, j0 = (blockIdx.x / tiles_n) * 32, n0 = (blockIdx.x % tiles_n) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const long long slab = (long long)g * layers + l;
  const float* W = params + g * gstride + P.w_hh[l];
  for (int i = ty; i < 32; i += 8) {
    const int idx = (j0 + i) * L + n0 + tx;
    const float v = W[idx];
    tile[i][tx] = v;
    const __half h = __float2half_rn(v);
    f_hi[slab * 4 * L * L + idx] = h;
    f_lo[slab * 4 * L * L + idx] = __float2half_rn(v - __half2float(h));
  }
  __syncthreads();
  // 32 consecutive gate rows j0 .. j0+31 share gate and rank (j0 is a multiple of 32, gates / ranks change every 128 / 64)
  const int gate = j0 / L, u0 = j0 - gate * L, rk = u0 >> 6, ul0 = u0 & 63;
  for (int i = ty; i < 32; i += 8) {
    const float v = tile[tx][i];                              // element (j0 + tx, n0 + i)
    const long long o = ((slab * 2 + rk) * L + n0 + i) * (2 * L) + gate * 64 + ul0 + tx;
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    b_hi[o] = b;
    b_lo[o] = __float2bfloat16_rn(v - __bfloat162float(b));
  }
}

int seq_maps_fwd(CUtensorMap* hi, CUtensorMap* lo, const void* f_hi, const void* f_lo, int L, int slabs) {
  uint64_t dims[4] = {(uint64_t)L, (uint64_t)L, 4, (uint64_t)slabs};
  uint64_t str[3] = {(uint64_t)L * 2, (uint64_t)L * L * 2, (uint64_t)4 * L * L * 2};
  uint32_t box[4] = {64, 64, 4, 1};
  int rc = wf_encode_tensor_map(hi, f_hi, 4, dims, str, box, 1, 1);
  if (rc) return rc;
  return wf_encode_tensor_map(lo, f_lo, 4, dims, str, box, 1, 1);
}
int seq_maps_bwd(CUtensorMap* hi, CUtensorMap* lo, const void* b_hi, const void* b_lo, int L, int slabs) {
  uint64_t dims[3] = {(uint64_t)2 * L, (uint64_t)L, (uint64_t)slabs * 2};
  uint64_t str[2] = {(uint64_t)2 * L * 2, (uint64_t)L * 2 * L * 2};
  uint32_t box[3] = {64, 128, 1};
  int rc = wf_encode_tensor_map(hi, b_hi, 3, dims, str, box, 1, 2);
  if (rc) return rc;
  return wf_encode_tensor_map(lo, b_lo, 3, dims, str, box, 1, 2);
}

template <typename K>
int seq_configure(K kernel) {
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SEQ_SMEM) != cudaSuccess)
    return wf_fail(WF_ECUDA, "lstm_seq: cannot raise dynamic shared memory to %d", SEQ_SMEM);
  return WF_OK;
}

}  // namespace

// Elements of a TB4 buffer with `channels` per (window, step, node): windows*T*ceil(N/128) blocks of channels*128.
extern "C" long long wf_tb4_elems(int channels, int T, int N, long long windows) {
  return windows * T * wf_cdiv(N, 128) * (long long)channels * 128;
}

// Group stride (16-bit elements) of the p16 / pT16 operand buffers: the parameter counts rounded up to 8 so that every
// group starts 16-byte aligned (TMA).  which = 0: flat parameters, 1: transposed W_ih buffer.
static long long stride16(long long n) { return (n + 7) & ~7LL; }
extern "C" long long wf_param_stride16(int layers, int F, int L, int O, int which) {
  if (layers < 1 || layers > 8) return -1;
  const LstmLayout P = lstm_layout(layers, F, L, O);
  return stride16(which ? P.totalT : P.total);
}

// Elements (16-bit each) of ONE of the four recurrent-operand buffers written by wf_prep_weights_seq.
extern "C" long long wf_seq_weight_elems(int layers, int L, int G) { return (long long)G * layers * 4 * L * L; }

// Operand staging after every update of the (fast) weights -- it replaces nothing in the reference:
//   p16_hi / p16_lo   fp16 hi/lo of the whole flat parameter buffer, [G][wf_param_stride16(.., 0)] (W_ih: projections)
//   pT16_hi / pT16_lo bf16 hi/lo of W_ih^T, layers >= 1, at the offsets of wf_param_count_transposed,
//                     [G][wf_param_stride16(.., 1)] (dX)
//   f16_hi / f16_lo   W_hh as fp16 hi/lo [G][layers][4L][L] (forward recurrence)
//   b16_hi / b16_lo   W_hh^T regrouped per CTA rank as bf16 hi/lo [G][layers][2][L][2L] (backward recurrence)
extern "C" int wf_prep_weights_seq(const float* params, long long params_group_stride, int layers, int F, int L, int O,
                                   int G, void* p16_hi, void* p16_lo, void* pT16_hi, void* pT16_lo, void* f16_hi, void* f16_lo,
                                   void* b16_hi, void* b16_lo, void* stream) {
  WF_REQUIRE(layers >= 1 && layers <= 8 && L == 128 && G > 0, "prep_weights_seq: needs L == 128");
  const LstmLayout P = lstm_layout(layers, F, L, O);
  WF_REQUIRE(G == 1 || params_group_stride == P.total, "prep_weights_seq: parameter sets must be contiguous");
  WF_REQUIRE(P.total % 4 == 0, "prep_weights_seq: parameter count must be a multiple of 4");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = wf_launch_split16(params, params_group_stride, p16_hi, p16_lo, stride16(P.total), P.total, G, 0, st);
  if (rc) return rc;
  for (int l = 1; l < layers; ++l) {
    rc = wf_launch_transpose_split16(params + P.w_ih[l], params_group_stride, 4 * L, L, (uint16_t*)pT16_hi + P.wihT[l],
                                     (uint16_t*)pT16_lo + P.wihT[l], stride16(P.totalT), G, st);
    if (rc) return rc;
  }
  dim3 grid((4 * L / 32) * (L / 32), layers, G);
  wf_prep_seq_kernel<<<grid, 256, 0, st>>>(params, params_group_stride, P, layers, L, (__half*)f16_hi, (__half*)f16_lo,
                                           (__nv_bfloat16*)b16_hi, (__nv_bfloat16*)b16_lo);
  WF_CHECK_LAUNCH("prep_weights_seq");
  return WF_OK;
}

// 16-bit elements of ONE plane of a TB8 buffer with `channels` values per (window, step, node): the hi and the lo plane
// of an activation are two such planes back to back.
extern "C" long long wf_tb8_elems(int channels, int T, int N, long long windows) {
  return windows * T * wf_cdiv(N, 128) * (long long)channels * 128;
}

// nn.LSTM forward (hybrid_model.py:42-49, 93-105): per layer one input-projection GEMM + one persistent launch.
//   x16: layer-0 input, fp16 hi / lo planes row-major [2][G*Bw*T*N][F] (wf_gcn_layer_fwd_ss / wf_split16);
//   params: flat fp32 weights (biases); p16 / f16 operands from wf_prep_weights_seq;
//   gates TB4 [layers][4L ch], c TB4 [layers][L ch];
//   h16 [layers-1][2][wf_tb8_elems(L, ..)]: fp16 hi / lo planes of what the NEXT layer reads (masked when p_drop > 0);
//   hb16 [layers][2][..] (optional: training): bf16 hi / lo planes of the plain h, the weight gradients' operand;
//   hb16m [layers-1][2][..] (p_drop > 0 and training): bf16 planes of the masked h (dW_ih of the next layer).
//   All TB8 and ZERO-INITIALISED by the caller: the padding rows of a node tile are never written and must read as zero
//   where rows are contracted.  hlast [G*Bw*N, L] fp32: the top layer's last step (what the head reads).
extern "C" int wf_lstm_fwd_seq(const void* x16, const float* params, const void* p16_hi, const void* p16_lo,
                               long long params_group_stride, const void* f16_hi, const void* f16_lo, int layers, int F, int L,
                               int O, int T, int N, int G, int Bw, float* gates, void* h16, float* c, float* hlast, void* hb16,
                               float p_drop, const unsigned long long* rng, void* hb16m, int* err, void* stream) {
  WF_REQUIRE(layers >= 1 && layers <= 8 && L == 128 && F % 32 == 0, "lstm_fwd_seq: needs L == 128, F %% 32 == 0");
  WF_REQUIRE(T > 0 && N > 0 && G > 0 && Bw > 0, "lstm_fwd_seq: empty batch");
  WF_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "lstm_fwd_seq: p_drop=%f outside [0, 1)", (double)p_drop);
  WF_REQUIRE(p_drop == 0.f || layers == 1 || rng != nullptr, "lstm_fwd_seq: dropout needs the rng state");
  WF_REQUIRE(p_drop == 0.f || layers == 1 || hb16 == nullptr || hb16m != nullptr, "lstm_fwd_seq: training with dropout needs hb16m");
  cudaStream_t st = (cudaStream_t)stream;
  static bool configured = false;
  if (!configured) {
    int rc = seq_configure(wf_lstm_seq_fwd16_kernel<false>);
    if (rc == WF_OK) rc = seq_configure(wf_lstm_seq_fwd16_kernel<true>);
    if (rc) return rc;
    configured = true;
  }
  const LstmLayout P = lstm_layout(layers, F, L, O);
  const long long Z = (long long)G * Bw;
  const int tpw = wf_cdiv(N, 128);
  const long long g_elems = wf_tb4_elems(4 * L, T, N, Z), c_elems = wf_tb4_elems(L, T, N, Z);
  const long long hp = wf_tb8_elems(L, T, N, Z);   // one plane of h
  CUtensorMap tmhi, tmlo;
  int rc = seq_maps_fwd(&tmhi, &tmlo, f16_hi, f16_lo, L, G * layers);
  if (rc) return rc;
  for (int l = 0; l < layers; ++l) {
    const int kin = l == 0 ? F : L;
    float* XG = gates + l * g_elems;
    // input projection x W_ih^T + b_ih + b_hh -> the gate pre-activations' TB4 block (4L channels)
    if (l == 0)
      rc = wf_ss_launch_nodes(kin <= 128 ? 256 : 128, 0, x16, Z * T * N * (long long)F, kin, 0, (const uint16_t*)p16_hi + P.w_ih[l],
                              (const uint16_t*)p16_lo + P.w_ih[l], kin, stride16(P.total), 4 * L, 0, params + P.b_ih[l],
                              params + P.b_hh[l], params_group_stride, XG, T, N, Bw, G, nullptr, err, st);
    else
      rc = wf_ss_launch_nodes(256, 1, (const uint16_t*)h16 + (long long)(l - 1) * 2 * hp, hp, kin, 0,
                              (const uint16_t*)p16_hi + P.w_ih[l], (const uint16_t*)p16_lo + P.w_ih[l], kin, stride16(P.total),
                              4 * L, 0, params + P.b_ih[l], params + P.b_hh[l], params_group_stride, XG, T, N, Bw, G, nullptr,
                              err, st);
    if (rc) return rc;
    SeqArgs a;
    memset(&a, 0, sizeof(a));
    a.XG = XG; a.Cst = c + l * c_elems; a.h16_plane = hp;
    a.H16 = l + 1 < layers ? (uint16_t*)h16 + (long long)l * 2 * hp : nullptr;
    a.HB16 = hb16 != nullptr ? (uint16_t*)hb16 + (long long)l * 2 * hp : nullptr;
    a.Hlast = l + 1 == layers ? hlast : nullptr;
    a.T = T; a.Nn = N; a.Bw = Bw; a.tpw = tpw; a.rpt = wf_tile_rows(N);
    a.slab0 = l; a.slab_g = layers; a.err = err;
    const bool drop = p_drop > 0.f && l + 1 < layers;  // nn.LSTM: dropout on the outputs of every layer but the last
    if (drop) {
      a.drop = wf_drop_cfg(p_drop, rng, WF_SITE_LSTM + l);
      a.HB16m = hb16m != nullptr ? (uint16_t*)hb16m + (long long)l * 2 * hp : nullptr;
      wf_lstm_seq_fwd16_kernel<true><<<dim3((unsigned)(2 * Z * tpw)), 512, SEQ_SMEM, st>>>(tmhi, tmlo, a);
    } else {
      wf_lstm_seq_fwd16_kernel<false><<<dim3((unsigned)(2 * Z * tpw)), 512, SEQ_SMEM, st>>>(tmhi, tmlo, a);
    }
    WF_CHECK_LAUNCH("lstm_seq_fwd");
  }
  return WF_OK;
}

// workspace = dh between layers (TB4, L channels) + split-K partials of the weight gradients
static size_t seq_partial_floats(int F, int L, int G) { return (size_t)(8 * G > 40 ? 8 * G : 40) * 4 * L * 257; }
extern "C" size_t wf_lstm_bwd_seq_workspace_bytes(int layers, int F, int L, int T, int N, int G, int Bw) {
  (void)layers;
  return sizeof(float) * ((size_t)wf_tb4_elems(L, T, N, (long long)G * Bw) + seq_partial_floats(F, L, G)) + 256;
}

// BPTT (train_hybrid_maml_v5.py:134,169) with one persistent launch per layer.  gates / c / hb16 (/ hb16m) from
// wf_lstm_fwd_seq; xb16: the layer-0 input as BF16 hi / lo planes row-major [2][G*Bw*T*N][F]; dg16: scratch
// [2][wf_tb8_elems(4L, ..)] for one layer's dG (bf16 hi / lo, ZERO-INITIALISED once by the caller: padding rows stay zero);
// pT16 / b16 operands from wf_prep_weights_seq; dlast [G*Bw*N, L].
extern "C" int wf_lstm_bwd_seq(const void* xb16, const void* pT16_hi, const void* pT16_lo, const void* b16_hi, const void* b16_lo,
                               int layers, int F, int L, int O, int T, int N, int G, int Bw, const float* gates, const float* c,
                               const void* hb16, void* dg16, const float* dlast, float* grads, long long grads_group_stride,
                               float p_drop, const unsigned long long* rng, const void* hb16m, void* workspace,
                               size_t workspace_bytes, int* err, void* stream) {
  WF_REQUIRE(layers >= 1 && layers <= 8 && L == 128 && F % 128 == 0 && F <= 256, "lstm_bwd_seq: needs L == 128 and F in {128, 256}");
  WF_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "lstm_bwd_seq: p_drop=%f outside [0, 1)", (double)p_drop);
  WF_REQUIRE(p_drop == 0.f || layers == 1 || (rng != nullptr && hb16m != nullptr), "lstm_bwd_seq: dropout needs rng and hb16m");
  WF_REQUIRE(hb16 != nullptr && xb16 != nullptr && dg16 != nullptr, "lstm_bwd_seq: missing operand planes");
  if (workspace_bytes < wf_lstm_bwd_seq_workspace_bytes(layers, F, L, T, N, G, Bw))
    return wf_fail(WF_EWORKSPACE, "lstm_bwd_seq: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  static bool configured = false;
  if (!configured) { int rc = seq_configure(wf_lstm_seq_bwd_kernel); if (rc) return rc; configured = true; }
  const bool drop = p_drop > 0.f && layers > 1;
  const LstmLayout P = lstm_layout(layers, F, L, O);
  const long long Z = (long long)G * Bw;
  const int tpw = wf_cdiv(N, 128);
  const long long g_elems = wf_tb4_elems(4 * L, T, N, Z), c_elems = wf_tb4_elems(L, T, N, Z);
  const long long hp = wf_tb8_elems(L, T, N, Z), dgp = wf_tb8_elems(4 * L, T, N, Z);
  float* DX = (float*)workspace;
  float* partials = DX + wf_tb4_elems(L, T, N, Z);
  const size_t partial_floats = seq_partial_floats(F, L, G);
  const uint16_t* H16 = (const uint16_t*)hb16;
  const uint16_t* H16m = (const uint16_t*)hb16m;
  CUtensorMap tmhi, tmlo;
  int rc = seq_maps_bwd(&tmhi, &tmlo, b16_hi, b16_lo, L, G * layers);
  if (rc) return rc;
  for (int l = layers - 1; l >= 0; --l) {
    SeqArgs a;
    memset(&a, 0, sizeof(a));
    a.XG = const_cast<float*>(gates) + l * g_elems; a.Cst = const_cast<float*>(c) + l * c_elems;
    a.DG16 = (uint16_t*)dg16; a.dg16_plane = dgp;
    a.ext = l == layers - 1 ? dlast : DX; a.ext_last_only = l == layers - 1 ? 1 : 0;
    a.T = T; a.Nn = N; a.Bw = Bw; a.tpw = tpw; a.rpt = wf_tile_rows(N);
    a.slab0 = 2 * l; a.slab_g = 2 * layers; a.err = err;
    wf_lstm_seq_bwd_kernel<<<dim3((unsigned)(2 * Z * tpw)), SEQ_THREADS, SEQ_SMEM, st>>>(tmhi, tmlo, a);
    WF_CHECK_LAUNCH("lstm_seq_bwd");
    // weight gradients, one pass over dG: [dW_ih | dW_hh] = dG^T [x | h(t-1)], the bias gradients = row sums of dG^T
    const void* Hl = H16 + (long long)l * 2 * hp;
    if (l > 0) {
      const void* Xl = (drop ? H16m : H16) + (long long)(l - 1) * 2 * hp;   // the (masked) output of the layer below
      const void* src[2] = {Xl, Hl};
      const long long plane[2] = {hp, hp};
      const int var[2] = {0, 0}, shift[2] = {0, 1}, col0[2] = {0, 0}, ch[2] = {L, L};
      rc = wf_ss_launch_wgrad(dg16, dgp, T > 1 ? 2 : 1, src, plane, var, shift, col0, ch, T, N, Bw, G, partials, partial_floats,
                              grads + P.w_ih[l], L, L, T > 1 ? grads + P.w_hh[l] : nullptr, L, L, grads + P.b_ih[l],
                              grads + P.b_hh[l], grads_group_stride, err, st);
      if (rc) return rc;
    } else {
      for (int f0 = 0; f0 < F; f0 += 256) {   // layer 0: the features are row-major planes, 256 columns per pass
        const int w = F - f0 < 256 ? F - f0 : 256;
        const void* src[2] = {xb16, xb16};
        const long long plane[2] = {Z * T * N * (long long)F, Z * T * N * (long long)F};
        const int var[2] = {1, 1}, shift[2] = {0, 0}, col0[2] = {f0, f0 + 128}, ch[2] = {F, F};
        rc = wf_ss_launch_wgrad(dg16, dgp, w > 128 ? 2 : 1, src, plane, var, shift, col0, ch, T, N, Bw, G, partials, partial_floats,
                                grads + P.w_ih[0] + f0, F, w, nullptr, 0, 0, f0 == 0 ? grads + P.b_ih[0] : nullptr,
                                f0 == 0 ? grads + P.b_hh[0] : nullptr, grads_group_stride, err, st);
        if (rc) return rc;
      }
      if (T > 1) {
        const void* src[2] = {Hl, Hl};
        const long long plane[2] = {hp, hp};
        const int var[2] = {0, 0}, shift[2] = {1, 1}, col0[2] = {0, 0}, ch[2] = {L, L};
        rc = wf_ss_launch_wgrad(dg16, dgp, 1, src, plane, var, shift, col0, ch, T, N, Bw, G, partials, partial_floats,
                                grads + P.w_hh[0], L, L, nullptr, 0, 0, nullptr, nullptr, grads_group_stride, err, st);
        if (rc) return rc;
      }
    }
    if (T == 1)
      for (int g = 0; g < G; ++g) cudaMemsetAsync(grads + g * grads_group_stride + P.w_hh[l], 0, sizeof(float) * 4 * L * L, st);
    if (l > 0) {  // dL/d(input of layer l) = dG W_ih -> ext of layer l-1 (TB4 out), through layer l-1's mask
      const DropCfg dc = wf_drop_cfg(p_drop, rng, WF_SITE_LSTM + l - 1);
      // WF_DX_KPARTS=2: K = 4L split over two CTAs (each keeps its half of W_ih^T resident, issues full-width N = L
      // instructions and reads dG once instead of once per column part; the halves meet through red.global.add in a cleared
      // output).  Measured 115 us against 116 us for two 64-column parts -- the A ring's bytes in flight bind both -- so
      // the plain form stays the default.
      static const int dx_kparts = getenv("WF_DX_KPARTS") ? atoi(getenv("WF_DX_KPARTS")) : 1;
      rc = wf_ss_launch_nodes(dx_kparts > 1 ? L : 64, 1, dg16, dgp, 4 * L, 1, (const uint16_t*)pT16_hi + P.wihT[l],
                              (const uint16_t*)pT16_lo + P.wihT[l], 4 * L, stride16(P.totalT), L, 1, nullptr, nullptr, 0, DX, T, N, Bw,
                              G, drop ? &dc : nullptr, err, st, dx_kparts > 1 ? dx_kparts : 1);
      if (rc) return rc;
    }
  }
  return WF_OK;
}

// ---- single-layer recurrence entry points: exactly the persistent launches the two functions above issue per layer,
// exposed so that a harness can time the dominant kernels alone (bench.py roofline) or drive layers itself.
extern "C" int wf_lstm_seq_recur_fwd(float* gates_l, float* c_l, void* h16_l, void* hb16_l, float* hlast, const void* f16_hi,
                                     const void* f16_lo, int layer, int layers, int L, int T, int N, int G, int Bw, int* err,
                                     void* stream) {
  WF_REQUIRE(layers >= 1 && layers <= 8 && layer >= 0 && layer < layers && L == 128, "lstm_seq_recur_fwd: bad layer / L");
  static bool configured = false;
  if (!configured) { int rc = seq_configure(wf_lstm_seq_fwd16_kernel<false>); if (rc) return rc; configured = true; }
  const long long Z = (long long)G * Bw;
  CUtensorMap tmhi, tmlo;
  int rc = seq_maps_fwd(&tmhi, &tmlo, f16_hi, f16_lo, L, G * layers);
  if (rc) return rc;
  SeqArgs a;
  memset(&a, 0, sizeof(a));
  a.XG = gates_l; a.Cst = c_l; a.H16 = (uint16_t*)h16_l; a.HB16 = (uint16_t*)hb16_l; a.h16_plane = wf_tb8_elems(L, T, N, Z);
  a.Hlast = hlast;
  a.T = T; a.Nn = N; a.Bw = Bw; a.tpw = wf_cdiv(N, 128); a.rpt = wf_tile_rows(N);
  a.slab0 = layer; a.slab_g = layers; a.err = err;
  wf_lstm_seq_fwd16_kernel<false><<<dim3((unsigned)(2 * Z * a.tpw)), 512, SEQ_SMEM, (cudaStream_t)stream>>>(tmhi, tmlo, a);
  WF_CHECK_LAUNCH("lstm_seq_recur_fwd");
  return WF_OK;
}

extern "C" int wf_lstm_seq_recur_bwd(const float* gates_l, const float* c_l, void* dg16, const float* ext, int ext_is_dlast,
                                     const void* b16_hi, const void* b16_lo, int layer, int layers, int L, int T, int N, int G,
                                     int Bw, int* err, void* stream) {
  WF_REQUIRE(layers >= 1 && layers <= 8 && layer >= 0 && layer < layers && L == 128, "lstm_seq_recur_bwd: bad layer / L");
  static bool configured = false;
  if (!configured) { int rc = seq_configure(wf_lstm_seq_bwd_kernel); if (rc) return rc; configured = true; }
  const long long Z = (long long)G * Bw;
  CUtensorMap tmhi, tmlo;
  int rc = seq_maps_bwd(&tmhi, &tmlo, b16_hi, b16_lo, L, G * layers);
  if (rc) return rc;
  SeqArgs a;
  memset(&a, 0, sizeof(a));
  a.XG = const_cast<float*>(gates_l); a.Cst = const_cast<float*>(c_l); a.DG16 = (uint16_t*)dg16;
  a.dg16_plane = wf_tb8_elems(4 * L, T, N, Z); a.ext = ext; a.ext_last_only = ext_is_dlast ? 1 : 0;
  a.T = T; a.Nn = N; a.Bw = Bw; a.tpw = wf_cdiv(N, 128); a.rpt = wf_tile_rows(N);
  a.slab0 = 2 * layer; a.slab_g = 2 * layers; a.err = err;
  wf_lstm_seq_bwd_kernel<<<dim3((unsigned)(2 * Z * a.tpw)), SEQ_THREADS, SEQ_SMEM, (cudaStream_t)stream>>>(tmhi, tmlo, a);
  WF_CHECK_LAUNCH("lstm_seq_recur_bwd");
  return WF_OK;
}

#ifdef WF_SEQ_TRACE
extern "C" int wf_seq_trace_read(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, wf_seq_trace_buf, sizeof(long long) * 32 * 16 * 24);
}
#endif
