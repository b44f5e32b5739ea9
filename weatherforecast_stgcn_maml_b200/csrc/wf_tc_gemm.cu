// tcgen05 3xTF32 contraction kernel shared by every dense product on the hot path.
//
//   D[128 x BN] (TMEM, fp32) = sum over K of  A[128 x 32] * B[BN x 32]^T      (3xTF32, wf_tc.cuh)
//
// Warp-specialised, one output tile per CTA, 2-stage TMA pipeline over K (32 per stage):
//   warp 0      TMA producer: A tile [128 x 32] fp32 plus the B tiles B_hi / B_lo [BN x 32]
//               (SWIZZLE_128B) into shared memory, completion on `full[s]`
//   warps 2..5  converters: thread = tile row; read the row's 32 fp32 from shared memory
//               (swizzle-aware, conflict free) -- or, for GCN rows with neighbours, gather-
//               aggregate them through the CSR straight from global memory -- split into hi / lo
//               and tcgen05.st both into TMEM; arrive on `aready[s]`.  Then they run the epilogue.
//   warp 1      MMA issuer: 12 x tcgen05.mma.kind::tf32 per stage (A from TMEM, B from smem:
//               A_hi B_hi + A_lo B_hi + A_hi B_lo), tcgen05.commit -> `empty[s]` (frees the smem
//               stage and the TMEM A stage); final commit -> `dfull`
// TMEM: accumulator columns [0, BN), A stages at [BN, BN + 128).
//
// Modes (template EPI):
//   EPI_STORE     C = D (+ bias, + bias2, relu); optional transposed copies C^T and lo(C^T) for the
//                 weight-gradient products.  GCN Theta transform with fused neighbour
//                 aggregation (model.py:23-26), LSTM input projections (hybrid_model.py:42-49),
//                 dX = dG W (pre-transposed W), and the weight gradients dW = dG^T X computed as
//                 (dG^T)(X^T)^T over the transposed activation copies (K = rows, in window segments).
//   EPI_LSTM_FWD  one LSTM time step: D = h[t-1] W_hh^T (BN = 4 gates x 64 units), epilogue adds the
//                 input projection, applies the cell non-linearities and writes gates / c / h
//                 (+ h^T and lo(h^T)) -- hybrid_model.py:98.
//   EPI_LSTM_BWD  one BPTT step: D = dG[t+1] W_hh (BN = 128 units), epilogue forms the gate
//                 gradients in place (+ dG^T).
#include "wf_common.cuh"
#include "wf_tc.cuh"

using namespace wftc;

enum { EPI_STORE = 0, EPI_LSTM_FWD = 1, EPI_LSTM_BWD = 2 };
enum { TILE_ROWS = 0, TILE_WGRAD = 1, TILE_STEP = 2 };

struct TcArgs {
  // ---- tiling
  int tile_mode;        // TILE_ROWS / TILE_WGRAD / TILE_STEP
  int tiles_g;          // M tiles per group (blockIdx.y = g * tiles_g + tile)
  int rows_g;           // valid rows per group (TILE_ROWS) / gate rows (TILE_WGRAD)
  int a_group_rows;     // TILE_ROWS: row stride between groups in the A map
  int Bw, R, Nn, T, t;  // windows per group, rows per window, nodes, steps, current step (TILE_STEP / WGRAD)
  int Np, RT;           // transposed copies: column of (t, node) = t*Np + node, row pitch RT = T*Np, Np = Nn rounded
                        // up to 4 -- TMA needs 16-byte aligned box starts (an unaligned inner coordinate faults)
  int nkb;              // k-blocks per segment
  int nseg;             // K segments (windows) accumulated into one tile (TILE_WGRAD: Bw, else 1)
  int nkb_split;        // TILE_WGRAD split-K: k-blocks per blockIdx.z slice (partials c_sstride apart)
  long long c_sstride;
  int a_k0, b_k0;       // first K coordinate of a segment in the A / B maps
  int b_rank;           // 3: (k, n, z)   4: (k, unit, gate, group)
  int b_gmul;           // 0: B shared by all groups, 1: per-group B
  // ---- CSR gather on A (GCN aggregation), TILE_ROWS only
  const float* a_raw; int lda;
  const int* rowptr; const int* col; const float* val; long long g_rowptr, g_csr;
  // ---- store epilogue
  float* C; int ldc; long long c_gstride;
  const float* bias; const float* bias2; long long bias_gstride; int relu;
  float* ct; float* ct_lo; int ct_cols;     // transposed copies [(g*Bw + w)][ct_cols][R]
  // ---- LSTM epilogues
  float* XG; float* Cst; float* H; float* HT; float* HT_lo; float* DGT; float* DC; const float* ext;
  int ext_last_only; int L;
  int* err;
};

__device__ __forceinline__ float tc_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int BN, int EPI>
__global__ void __launch_bounds__(192, 1)
wf_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
             const __grid_constant__ CUtensorMap tmBlo, const TcArgs a) {
  constexpr int A_BYTES = 128 * 128, B_BYTES = BN * 128, STAGE = A_BYTES + 2 * B_BYTES, NST = BN == 128 ? 4 : 2;
  constexpr uint32_t TMEM_COLS = 512, A_COL = BN;  // D [0, BN), A stages (hi 32 + lo 32 columns each) behind it
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[NST], aready[NST], empty[NST], dfull;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.y / a.tiles_g, tile = blockIdx.y - g * a.tiles_g;
  const int n0 = blockIdx.x * BN;

  // ---- tile -> operand coordinates
  int a_row, a_z0 = 0, a_zstep = 0, b_z0 = g * a.b_gmul, b_zstep = 0;
  int win = 0, node0 = 0;  // TILE_STEP
  if (a.tile_mode == TILE_ROWS) {
    a_row = g * a.a_group_rows + tile * 128;
  } else if (a.tile_mode == TILE_WGRAD) {
    a_row = tile * 128;           // gate rows of dG^T
    a_z0 = g * a.Bw; a_zstep = 1;
    b_z0 = g * a.Bw; b_zstep = 1;
  } else {
    const int tpw = (a.Nn + 127) / 128;
    win = tile / tpw;
    node0 = (tile - win * tpw) * 128;
    const int tt = EPI == EPI_LSTM_FWD ? a.t - 1 : a.t + 1;
    a_row = (g * a.Bw + win) * a.R + tt * a.Nn + node0;
  }
  const int kb0 = a.tile_mode == TILE_WGRAD ? (int)blockIdx.z * a.nkb_split : 0;
  const int nkb_loc = a.tile_mode == TILE_WGRAD ? min(a.nkb - kb0, a.nkb_split) : a.nkb;
  const int nkb_total = nkb_loc * a.nseg;
  __shared__ int abort_s;  // a pipeline already timed out somewhere: do not pile up waits (CTA-uniform decision)
  if (threadIdx.x == 0) abort_s = *reinterpret_cast<volatile int*>(a.err);
  __syncthreads();
  if (abort_s != 0) return;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&aready[s], 4); mbar_init(&empty[s], 1); }
    mbar_init(&dfull, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmBhi); tma_prefetch_desc(&tmBlo);
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int seg = 0; seg < a.nseg; ++seg)
        for (int kb = kb0; kb < kb0 + nkb_loc; ++kb, ++it) {
          const int s = it % NST, ph = (it / NST) & 1;
          if (!mbar_wait(&empty[s], ph ^ 1)) { atomicExch(a.err, 1); seg = a.nseg; break; }
          uint8_t* st = smem + s * STAGE;
          mbar_expect_tx(&full[s], STAGE);
          tma_load_3d(st, &tmA, &full[s], a.a_k0 + kb * 32, a_row, a_z0 + seg * a_zstep);
          if (a.b_rank == 4) {
            tma_load_4d(st + A_BYTES, &tmBhi, &full[s], a.b_k0 + kb * 32, blockIdx.x * (BN / 4), 0, g);
            tma_load_4d(st + A_BYTES + B_BYTES, &tmBlo, &full[s], a.b_k0 + kb * 32, blockIdx.x * (BN / 4), 0, g);
          } else {
            tma_load_3d(st + A_BYTES, &tmBhi, &full[s], a.b_k0 + kb * 32, n0, b_z0 + seg * b_zstep);
            tma_load_3d(st + A_BYTES + B_BYTES, &tmBlo, &full[s], a.b_k0 + kb * 32, n0, b_z0 + seg * b_zstep);
          }
        }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_tf32(BN);
    for (int it = 0; it < nkb_total; ++it) {
      const int s = it % NST, ph = (it / NST) & 1;
      if (!mbar_wait(&full[s], ph) || !mbar_wait(&aready[s], ph)) { if (lane == 0) atomicExch(a.err, 2); break; }
      tc_fence_after();
      if (lane == 0) {
        const uint32_t bhi = smem_u32(smem + s * STAGE + A_BYTES), blo = bhi + B_BYTES;
        const uint32_t acol = tbase + A_COL + s * 64;
#pragma unroll
        for (int p = 0; p < 3; ++p) {  // A_hi B_hi, A_lo B_hi, A_hi B_lo
          const uint32_t ac = acol + (p == 1 ? 32 : 0);
          const uint32_t bs = p == 2 ? blo : bhi;
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8)
            umma_ts(tbase, ac + k8 * 8, umma_desc_k_sw128(bs + k8 * 32), idesc, (it | p | k8) ? 1u : 0u);
        }
        umma_commit(&empty[s]);
        if (it == nkb_total - 1) umma_commit(&dfull);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;     // tile row == TMEM lane
    const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
    bool ok = true;

    // CSR gather set-up (GCN): rows whose aggregation is not the unit self loop
    int p0 = 0, p1 = 0;
    long long wbase = 0;
    bool gather = false;
    if (EPI == EPI_STORE && a.rowptr != nullptr) {
      const int grow = tile * 128 + row;
      if (grow < a.rows_g) {
        const int w = grow / a.R, rr = grow - w * a.R;
        const int* rp = a.rowptr + g * a.g_rowptr;
        p0 = rp[rr]; p1 = rp[rr + 1];
        wbase = ((long long)g * a.a_group_rows + (long long)w * a.R) * a.lda;
        const int* cl = a.col + g * a.g_csr;
        const float* vl = a.val + g * a.g_csr;
        gather = !(p1 - p0 == 1 && cl[p0] == rr && vl[p0] == 1.0f);
      }
    }

    for (int it = 0; it < nkb_total && ok; ++it) {
      const int s = it % NST, ph = (it / NST) & 1;
      if (!mbar_wait(&full[s], ph) || !mbar_wait(&empty[s], ph ^ 1)) { if (lane == 0) atomicExch(a.err, 3); ok = false; break; }
      uint32_t hi[32], lo[32];
      if (!gather) {
        const uint8_t* arow = smem + s * STAGE + row * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(arow + ((c ^ (row & 7)) << 4));
          split_tf32(v.x, hi[4 * c + 0], lo[4 * c + 0]);
          split_tf32(v.y, hi[4 * c + 1], lo[4 * c + 1]);
          split_tf32(v.z, hi[4 * c + 2], lo[4 * c + 2]);
          split_tf32(v.w, hi[4 * c + 3], lo[4 * c + 3]);
        }
      } else {
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0.f;
        const int* cl = a.col + g * a.g_csr;
        const float* vl = a.val + g * a.g_csr;
        const int k0 = a.a_k0 + (it % a.nkb) * 32;
        for (int p = p0; p < p1; ++p) {
          const float v = __ldg(vl + p);
          const float4* src = reinterpret_cast<const float4*>(a.a_raw + wbase + (long long)__ldg(cl + p) * a.lda + k0);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 x = __ldg(src + c);
            acc[4 * c + 0] = fmaf(v, x.x, acc[4 * c + 0]); acc[4 * c + 1] = fmaf(v, x.y, acc[4 * c + 1]);
            acc[4 * c + 2] = fmaf(v, x.z, acc[4 * c + 2]); acc[4 * c + 3] = fmaf(v, x.w, acc[4 * c + 3]);
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) split_tf32(acc[j], hi[j], lo[j]);
      }
      __syncwarp();  // gather / non-gather lanes diverged above
      tmem_st32(tlane + A_COL + s * 64, hi);
      tmem_st32(tlane + A_COL + s * 64 + 32, lo);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&aready[s]);
    }

    // ------------------------------------------------------------------ epilogue
    const bool have_acc = nkb_total > 0;
    if (ok && have_acc && !mbar_wait(&dfull, 0)) { if (lane == 0) atomicExch(a.err, 4); ok = false; }
    if (ok) {
      tc_fence_after();
      if (EPI == EPI_STORE) {
        const int grow = tile * 128 + row;
        const bool valid = grow < a.rows_g;
        float* crow = a.C + g * a.c_gstride + blockIdx.z * a.c_sstride + (long long)grow * a.ldc + n0;
        const float* b1 = a.bias ? a.bias + g * a.bias_gstride + n0 : nullptr;
        const float* b2 = a.bias2 ? a.bias2 + g * a.bias_gstride + n0 : nullptr;
        long long ctbase = 0;
        if (a.ct != nullptr && valid) {
          const int w = grow / a.R, rr = grow - w * a.R;
          const int tt = rr / a.Nn, nn = rr - tt * a.Nn;
          ctbase = ((long long)(g * a.Bw + w) * a.ct_cols + n0) * a.RT + (long long)tt * a.Np + nn;
        }
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld32(tlane + c, v);
          tmem_wait_ld();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                     __uint_as_float(v[j + 3]));
              if (b1) { const float4 b = __ldg(reinterpret_cast<const float4*>(b1 + c + j)); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
              if (b2) { const float4 b = __ldg(reinterpret_cast<const float4*>(b2 + c + j)); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
              if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              *reinterpret_cast<float4*>(crow + c + j) = o;
              if (a.ct != nullptr) {  // lanes of a warp hold consecutive rows -> coalesced transposed stores
                const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const long long ti = ctbase + (long long)(c + j + e) * a.RT;
                  a.ct[ti] = ov[e];
                  if (a.ct_lo) a.ct_lo[ti] = ov[e] - __uint_as_float(__float_as_uint(ov[e]) & 0xFFFFE000u);
                }
              }
            }
          }
        }
      } else if (EPI == EPI_LSTM_FWD) {
        // BN = 256 columns = gates (i, f, g, o) x 64 units [u0, u0 + 64)
        const int L = a.L, u0 = blockIdx.x * 64;
        const bool valid = node0 + row < a.Nn;
        const int node = valid ? node0 + row : a.Nn - 1;  // clamp: tcgen05.ld below is warp-collective
        {
          const long long z = (long long)g * a.Bw + win;
          const long long ridx = z * a.R + (long long)a.t * a.Nn + node;
          float* xg = a.XG + ridx * 4 * L + u0;
          const float* cprev = a.t > 0 ? a.Cst + (ridx - a.Nn) * L + u0 : nullptr;
          float* cout = a.Cst + ridx * L + u0;
          float* hout = a.H + ridx * L + u0;
          const long long tcol = (long long)a.t * a.Np + node;
#pragma unroll 1
          for (int uc = 0; uc < 64; uc += 16) {
            uint32_t vi[16], vf[16], vg[16], vo[16];
            if (have_acc) {
              __syncwarp();
              tmem_ld16(tlane + uc, vi); tmem_ld16(tlane + 64 + uc, vf);
              tmem_ld16(tlane + 128 + uc, vg); tmem_ld16(tlane + 192 + uc, vo);
              tmem_wait_ld();
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) { vi[j] = 0; vf[j] = 0; vg[j] = 0; vo[j] = 0; }
            }
            if (!valid) continue;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 xi = *reinterpret_cast<const float4*>(xg + uc + j);
              const float4 xf = *reinterpret_cast<const float4*>(xg + L + uc + j);
              const float4 xgg = *reinterpret_cast<const float4*>(xg + 2 * L + uc + j);
              const float4 xo = *reinterpret_cast<const float4*>(xg + 3 * L + uc + j);
              float4 cp = make_float4(0.f, 0.f, 0.f, 0.f);
              if (cprev) cp = *reinterpret_cast<const float4*>(cprev + uc + j);
              const float pi[4] = {xi.x, xi.y, xi.z, xi.w}, pf[4] = {xf.x, xf.y, xf.z, xf.w};
              const float pg[4] = {xgg.x, xgg.y, xgg.z, xgg.w}, po[4] = {xo.x, xo.y, xo.z, xo.w};
              const float pc[4] = {cp.x, cp.y, cp.z, cp.w};
              float gi[4], gf[4], gg[4], go[4], cc[4], hh[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                gi[e] = tc_sigmoid(pi[e] + __uint_as_float(vi[j + e]));
                gf[e] = tc_sigmoid(pf[e] + __uint_as_float(vf[j + e]));
                gg[e] = tanhf(pg[e] + __uint_as_float(vg[j + e]));
                go[e] = tc_sigmoid(po[e] + __uint_as_float(vo[j + e]));
                cc[e] = fmaf(gf[e], pc[e], gi[e] * gg[e]);
                hh[e] = go[e] * tanhf(cc[e]);
              }
              *reinterpret_cast<float4*>(xg + uc + j) = make_float4(gi[0], gi[1], gi[2], gi[3]);
              *reinterpret_cast<float4*>(xg + L + uc + j) = make_float4(gf[0], gf[1], gf[2], gf[3]);
              *reinterpret_cast<float4*>(xg + 2 * L + uc + j) = make_float4(gg[0], gg[1], gg[2], gg[3]);
              *reinterpret_cast<float4*>(xg + 3 * L + uc + j) = make_float4(go[0], go[1], go[2], go[3]);
              *reinterpret_cast<float4*>(cout + uc + j) = make_float4(cc[0], cc[1], cc[2], cc[3]);
              *reinterpret_cast<float4*>(hout + uc + j) = make_float4(hh[0], hh[1], hh[2], hh[3]);
              if (a.HT != nullptr) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const long long ti = (z * L + u0 + uc + j + e) * a.RT + tcol;
                  a.HT[ti] = hh[e];
                  a.HT_lo[ti] = hh[e] - __uint_as_float(__float_as_uint(hh[e]) & 0xFFFFE000u);
                }
              }
            }
          }
        }
      } else {  // EPI_LSTM_BWD: BN = 128 = hidden units
        const int L = a.L;
        const bool valid = node0 + row < a.Nn;
        const int node = valid ? node0 + row : a.Nn - 1;  // clamp: tcgen05.ld below is warp-collective
        const bool last = a.t == a.T - 1;
        {
          const long long z = (long long)g * a.Bw + win;
          const long long ridx = z * a.R + (long long)a.t * a.Nn + node;
          const long long sidx = z * a.Nn + node;  // compact per-sequence index
          float* xg = a.XG + ridx * 4 * L + n0;
          const float* cc_p = a.Cst + ridx * L + n0;
          const float* cp_p = a.t > 0 ? a.Cst + (ridx - a.Nn) * L + n0 : nullptr;
          float* dcp = a.DC + sidx * L + n0;
          const bool use_ext = a.ext != nullptr && (!a.ext_last_only || last);
          const float* ext = use_ext ? a.ext + (a.ext_last_only ? sidx : ridx) * L + n0 : nullptr;
          const long long tcol = (long long)a.t * a.Np + node;
#pragma unroll 1
          for (int uc = 0; uc < 128; uc += 16) {
            uint32_t vd[16];
            __syncwarp();  // lanes of ragged tiles `continue` below: reconverge before the warp-collective load
            if (have_acc) { tmem_ld16(tlane + uc, vd); tmem_wait_ld(); }
            else {
#pragma unroll
              for (int j = 0; j < 16; ++j) vd[j] = 0;
            }
            if (!valid) continue;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 gi4 = *reinterpret_cast<const float4*>(xg + uc + j);
              const float4 gf4 = *reinterpret_cast<const float4*>(xg + L + uc + j);
              const float4 gg4 = *reinterpret_cast<const float4*>(xg + 2 * L + uc + j);
              const float4 go4 = *reinterpret_cast<const float4*>(xg + 3 * L + uc + j);
              const float4 cc4 = *reinterpret_cast<const float4*>(cc_p + uc + j);
              float4 cp4 = make_float4(0.f, 0.f, 0.f, 0.f), dn4 = cp4, ex4 = cp4;
              if (cp_p) cp4 = *reinterpret_cast<const float4*>(cp_p + uc + j);
              if (!last) dn4 = *reinterpret_cast<const float4*>(dcp + uc + j);
              if (ext) ex4 = *reinterpret_cast<const float4*>(ext + uc + j);
              const float vi[4] = {gi4.x, gi4.y, gi4.z, gi4.w}, vf[4] = {gf4.x, gf4.y, gf4.z, gf4.w};
              const float vg[4] = {gg4.x, gg4.y, gg4.z, gg4.w}, vo[4] = {go4.x, go4.y, go4.z, go4.w};
              const float vc[4] = {cc4.x, cc4.y, cc4.z, cc4.w}, vp[4] = {cp4.x, cp4.y, cp4.z, cp4.w};
              const float vn[4] = {dn4.x, dn4.y, dn4.z, dn4.w}, ve[4] = {ex4.x, ex4.y, ex4.z, ex4.w};
              float di[4], df[4], dg[4], dO[4], dcv[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float dh = __uint_as_float(vd[j + e]) + ve[e];
                const float tc = tanhf(vc[e]);
                const float dc = vn[e] + dh * vo[e] * (1.f - tc * tc);
                dO[e] = dh * tc * vo[e] * (1.f - vo[e]);
                di[e] = dc * vg[e] * vi[e] * (1.f - vi[e]);
                df[e] = dc * vp[e] * vf[e] * (1.f - vf[e]);
                dg[e] = dc * vi[e] * (1.f - vg[e] * vg[e]);
                dcv[e] = dc * vf[e];
              }
              *reinterpret_cast<float4*>(xg + uc + j) = make_float4(di[0], di[1], di[2], di[3]);
              *reinterpret_cast<float4*>(xg + L + uc + j) = make_float4(df[0], df[1], df[2], df[3]);
              *reinterpret_cast<float4*>(xg + 2 * L + uc + j) = make_float4(dg[0], dg[1], dg[2], dg[3]);
              *reinterpret_cast<float4*>(xg + 3 * L + uc + j) = make_float4(dO[0], dO[1], dO[2], dO[3]);
              *reinterpret_cast<float4*>(dcp + uc + j) = make_float4(dcv[0], dcv[1], dcv[2], dcv[3]);
              if (a.DGT != nullptr) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const long long u = n0 + uc + j + e;
                  float* base = a.DGT + (z * 4 * L + u) * a.RT + tcol;
                  base[0] = di[e];
                  base[(long long)L * a.RT] = df[e];
                  base[2LL * L * a.RT] = dg[e];
                  base[3LL * L * a.RT] = dO[e];
                }
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, TMEM_COLS);
}

// ------------------------------------------------------------------ weight preparation
__global__ void wf_split_lo_kernel(const float4* __restrict__ src, float4* __restrict__ dst, long long quads) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= quads) return;
  float4 v = src[i], o;
  o.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
  o.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  o.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
  o.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  dst[i] = o;
}

// dst = src - trunc_tf32(src): the `lo` half of the 3xTF32 split for a weight buffer.
extern "C" int wf_split_lo(const float* src, float* dst, long long n, void* stream) {
  WF_REQUIRE(n > 0 && n % 4 == 0, "split_lo: n must be a positive multiple of 4");
  wf_split_lo_kernel<<<wf_cdiv(n / 4, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)src, (float4*)dst, n / 4);
  WF_CHECK_LAUNCH("split_lo");
  return WF_OK;
}

// out[g][c][r] = in[g][r][c] (and its lo half): W[rows, cols] -> W^T[cols, rows] for every group.
__global__ void wf_transpose_split_kernel(const float* __restrict__ in, long long in_gstride, int rows, int cols,
                                          float* __restrict__ out, float* __restrict__ out_lo, long long out_gstride) {
  __shared__ float t[32][33];
  const int g = blockIdx.z;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    int r = r0 + i, c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (r < rows && c < cols) ? in[g * in_gstride + (long long)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) {
      float v = t[threadIdx.x][i];
      long long o = g * out_gstride + (long long)c * rows + r;
      out[o] = v;
      out_lo[o] = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    }
  }
}

int wf_launch_transpose_split(const float* in, long long in_gstride, int rows, int cols, float* out, float* out_lo,
                              long long out_gstride, int G, cudaStream_t st) {
  dim3 grid(wf_cdiv(cols, 32), wf_cdiv(rows, 32), G), block(32, 8);
  wf_transpose_split_kernel<<<grid, block, 0, st>>>(in, in_gstride, rows, cols, out, out_lo, out_gstride);
  WF_CHECK_LAUNCH("transpose_split");
  return WF_OK;
}

// ------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int wf_encode_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, int swizzle_128b, int dtype) {
  EncodeTiledFn fn = get_encode();
  if (!fn) return wf_fail(WF_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t d[5], s[5];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  const CUtensorMapDataType dt = dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                               : dtype == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, e,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_128b == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_128b == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                  : swizzle_128b == 3 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return wf_fail(WF_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return WF_OK;
}

// 3-D map (k, rows, z) over a row-major matrix (z = 1) or a stack of them.
static int map3(CUtensorMap* m, const float* base, uint64_t k, uint64_t rows, uint64_t z, uint64_t ld, uint64_t zstride,
                uint32_t box_rows) {
  uint64_t dims[3] = {k, rows, z};
  uint64_t str[2] = {ld * 4, zstride * 4};
  uint32_t box[3] = {32, box_rows, 1};
  return wf_encode_tensor_map(m, base, 3, dims, str, box, 1);
}

template <int BN, int EPI>
static int launch_variant(const CUtensorMap& tmA, const CUtensorMap& tmBhi, const CUtensorMap& tmBlo, const TcArgs& a,
                          dim3 grid, cudaStream_t st) {
  constexpr int smem = (BN == 128 ? 4 : 2) * (128 * 128 + 2 * BN * 128) + 1024;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(wf_tc_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return wf_fail(WF_ECUDA, "tc kernel: cannot raise dynamic shared memory to %d", smem);
    configured = true;
  }
  wf_tc_kernel<BN, EPI><<<grid, 192, smem, st>>>(tmA, tmBhi, tmBlo, a);
  WF_CHECK_LAUNCH("tc_kernel");
  static const bool debug_sync = getenv("WF_DEBUG_SYNC") != nullptr;  // localise device faults (never in production)
  if (debug_sync) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess)
      return wf_fail(WF_ECUDA, "tc_kernel<BN=%d,EPI=%d> tile_mode=%d t=%d nkb=%d nseg=%d grid=(%d,%d): %s", BN, EPI,
                     a.tile_mode, a.t, a.nkb, a.nseg, grid.x, grid.y, cudaGetErrorString(e));
  }
  return WF_OK;
}

static void tc_defaults(TcArgs& a) {
  memset(&a, 0, sizeof(a));
  a.nseg = 1; a.b_rank = 3; a.b_gmul = 1; a.Bw = 1; a.R = 1; a.nkb_split = 1 << 30;
}

// ---- C[g] = (A_hat[g]) A[g] W[g]^T (+bias, +bias2, relu), rows tiled, optional CSR gather on A and
// optional transposed copies of C.
int wf_launch_tc_rows(const float* A, long long a_rows_total, int lda, int a_group_rows, int rows_g, int G, int K,
                      const float* Whi, const float* Wlo, int ldb, long long b_gstride, long long blo_gstride,
                      int b_shared, int N, const float* bias, const float* bias2, long long bias_gstride, int relu,
                      float* C, int ldc, long long c_gstride, const int* rowptr, const int* col, const float* val,
                      long long g_rowptr, long long g_csr, int R, int Bw, float* ct, float* ct_lo, int Nn, int* err,
                      cudaStream_t st) {
  WF_REQUIRE(K % 32 == 0 && K >= 32, "tc_rows: K=%d must be a multiple of 32", K);
  WF_REQUIRE(N % 128 == 0, "tc_rows: N=%d must be a multiple of 128", N);
  WF_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0, "tc_rows: leading dimensions must be multiples of 4");
  WF_REQUIRE(((uintptr_t)A | (uintptr_t)Whi | (uintptr_t)Wlo | (uintptr_t)C) % 16 == 0, "tc_rows: pointers must be 16-byte aligned");
  const int BN = (N % 256 == 0) ? 256 : 128;
  const int Gb = b_shared ? 1 : G;
  CUtensorMap tmA, tmBhi, tmBlo;
  int rc;
  if ((rc = map3(&tmA, A, K, a_rows_total, 1, lda, (uint64_t)a_rows_total * lda, 128))) return rc;
  if ((rc = map3(&tmBhi, Whi, K, N, Gb, ldb, Gb > 1 ? b_gstride : (long long)N * ldb, BN))) return rc;
  if ((rc = map3(&tmBlo, Wlo, K, N, Gb, ldb, Gb > 1 ? blo_gstride : (long long)N * ldb, BN))) return rc;
  TcArgs a;
  tc_defaults(a);
  a.tile_mode = TILE_ROWS; a.tiles_g = wf_cdiv(rows_g, 128); a.rows_g = rows_g; a.a_group_rows = a_group_rows;
  a.nkb = K / 32; a.b_gmul = b_shared ? 0 : 1;
  a.a_raw = A; a.lda = lda; a.rowptr = rowptr; a.col = col; a.val = val; a.g_rowptr = g_rowptr; a.g_csr = g_csr;
  a.R = R > 0 ? R : rows_g; a.Bw = Bw > 0 ? Bw : 1;
  a.Nn = Nn > 0 ? Nn : a.R; a.Np = (a.Nn + 3) & ~3; a.RT = (a.R / a.Nn) * a.Np;
  WF_REQUIRE(ct == nullptr || a.R % a.Nn == 0, "tc_rows: transposed copies need R to be a multiple of the node count");
  a.C = C; a.ldc = ldc; a.c_gstride = c_gstride; a.bias = bias; a.bias2 = bias2; a.bias_gstride = bias_gstride;
  a.relu = relu; a.ct = ct; a.ct_lo = ct_lo; a.ct_cols = N; a.err = err;
  dim3 grid(N / BN, a.tiles_g * G);
  return BN == 256 ? launch_variant<256, EPI_STORE>(tmA, tmBhi, tmBlo, a, grid, st)
                   : launch_variant<128, EPI_STORE>(tmA, tmBhi, tmBlo, a, grid, st);
}

// Sum of split-K partials: out[g][i] = sum_s part[s][g][i], i < count (deterministic order).
__global__ void wf_sum_splits_kernel(const float4* __restrict__ part, int splits, long long count4, long long part_sstride4,
                                     float4* __restrict__ out, long long out_gstride4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int g = blockIdx.y;
  if (i >= count4) return;
  float4 acc = part[g * count4 + i];
  for (int sidx = 1; sidx < splits; ++sidx) {
    const float4 v = part[sidx * part_sstride4 + g * count4 + i];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  out[g * out_gstride4 + i] = acc;
}

// Split factor for the weight-gradient contraction: fill the SMs when there are few output tiles.
int wf_wgrad_splits(int M, int N, int G, int klen) {
  const int BN = (N % 256 == 0) ? 256 : 128;
  const int tiles = (M / 128) * (N / BN) * G, nkb = wf_cdiv(klen, 32);
  int splits = 148 / (tiles > 0 ? tiles : 1);
  if (splits < 1) splits = 1;
  if (splits > 8) splits = 8;
  if (splits > nkb) splits = nkb;
  const int per = wf_cdiv(nkb, splits);
  return wf_cdiv(nkb, per);
}

int wf_launch_tc_wgrad(const float* AT, int M, const float* BT, const float* BT_lo, int N, int R, int Bw, int G,
                       int a_k0, int b_k0, int klen, float* dW, long long dw_gstride, int* err, cudaStream_t st,
                       float* partials, size_t partial_floats) {
  WF_REQUIRE(M % 128 == 0 && N % 128 == 0 && R % 4 == 0, "tc_wgrad: M=%d N=%d must be multiples of 128, R=%d of 4", M, N, R);
  // measured on B200: a TMA box whose inner start coordinate is not 16-byte aligned faults (illegal instruction)
  WF_REQUIRE(a_k0 % 4 == 0 && b_k0 % 4 == 0, "tc_wgrad: K offsets (%d, %d) must be multiples of 4", a_k0, b_k0);
  const int BN = (N % 256 == 0) ? 256 : 128;
  CUtensorMap tmA, tmBhi, tmBlo;
  int rc;
  const uint64_t Z = (uint64_t)G * Bw;
  // the K extent seen through the maps ends at a_k0 + klen / b_k0 + klen: everything beyond is zero-filled
  if ((rc = map3(&tmA, AT, (uint64_t)a_k0 + klen, M, Z, R, (uint64_t)M * R, 128))) return rc;
  if ((rc = map3(&tmBhi, BT, (uint64_t)b_k0 + klen, N, Z, R, (uint64_t)N * R, BN))) return rc;
  if ((rc = map3(&tmBlo, BT_lo, (uint64_t)b_k0 + klen, N, Z, R, (uint64_t)N * R, BN))) return rc;
  int splits = wf_wgrad_splits(M, N, G, klen);
  const long long per_split = (long long)G * M * N;
  if (partials == nullptr || (long long)partial_floats < per_split * splits || (dw_gstride % 4) != 0) splits = 1;
  TcArgs a;
  tc_defaults(a);
  a.tile_mode = TILE_WGRAD; a.tiles_g = M / 128; a.rows_g = M; a.Bw = Bw; a.R = R;
  a.nkb = wf_cdiv(klen, 32); a.nseg = Bw; a.a_k0 = a_k0; a.b_k0 = b_k0;
  a.nkb_split = wf_cdiv(a.nkb, splits);
  a.ldc = N; a.err = err;
  if (splits > 1) { a.C = partials; a.c_gstride = (long long)M * N; a.c_sstride = per_split; }
  else { a.C = dW; a.c_gstride = dw_gstride; a.c_sstride = 0; }
  dim3 grid(N / BN, a.tiles_g * G, splits);
  rc = BN == 256 ? launch_variant<256, EPI_STORE>(tmA, tmBhi, tmBlo, a, grid, st)
                 : launch_variant<128, EPI_STORE>(tmA, tmBhi, tmBlo, a, grid, st);
  if (rc || splits == 1) return rc;
  const long long count4 = (long long)M * N / 4;
  wf_sum_splits_kernel<<<dim3(wf_cdiv(count4, 256), G), 256, 0, st>>>((const float4*)partials, splits, count4, per_split / 4,
                                                                     (float4*)dW, dw_gstride / 4);
  WF_CHECK_LAUNCH("sum_splits");
  return WF_OK;
}

// ---- one LSTM time step, forward.  H [G*Bw*R, L] row-major; W_hh hi (raw) / lo: [G][4L, L].
int wf_launch_tc_lstm_fwd(float* H, float* Cst, float* XG, float* HT, float* HT_lo, const float* Whh,
                          const float* Whh_lo, long long w_gstride, long long wlo_gstride, int L, int T, int Nn, int Bw,
                          int G, int t, int* err, cudaStream_t st) {
  WF_REQUIRE(L % 64 == 0 && L % 32 == 0, "tc_lstm_fwd: L=%d must be a multiple of 64", L);
  const int R = T * Nn;
  CUtensorMap tmA, tmBhi, tmBlo;
  int rc;
  if ((rc = map3(&tmA, H, L, (uint64_t)G * Bw * R, 1, L, (uint64_t)G * Bw * R * L, 128))) return rc;
  {
    uint64_t dims[4] = {(uint64_t)L, (uint64_t)L, 4, (uint64_t)G};
    uint32_t box[4] = {32, 64, 4, 1};
    uint64_t s1[3] = {(uint64_t)L * 4, (uint64_t)L * L * 4, (uint64_t)(G > 1 ? w_gstride : 4LL * L * L) * 4};
    uint64_t s2[3] = {(uint64_t)L * 4, (uint64_t)L * L * 4, (uint64_t)(G > 1 ? wlo_gstride : 4LL * L * L) * 4};
    if ((rc = wf_encode_tensor_map(&tmBhi, Whh, 4, dims, s1, box, 1))) return rc;
    if ((rc = wf_encode_tensor_map(&tmBlo, Whh_lo, 4, dims, s2, box, 1))) return rc;
  }
  TcArgs a;
  tc_defaults(a);
  a.tile_mode = TILE_STEP; a.tiles_g = Bw * wf_cdiv(Nn, 128); a.Bw = Bw; a.R = R; a.Nn = Nn; a.T = T; a.t = t;
  a.nkb = t > 0 ? L / 32 : 0; a.b_rank = 4;
  a.Np = (Nn + 3) & ~3; a.RT = T * a.Np;
  a.XG = XG; a.Cst = Cst; a.H = H; a.HT = HT; a.HT_lo = HT_lo; a.L = L; a.err = err;
  dim3 grid(L / 64, a.tiles_g * G);
  return launch_variant<256, EPI_LSTM_FWD>(tmA, tmBhi, tmBlo, a, grid, st);
}

// ---- one BPTT step.  A = dG rows of step t+1 (XG buffer); WhhT hi / lo: [G][L, 4L] (pre-transposed).
int wf_launch_tc_lstm_bwd(float* XG, const float* Cst, float* DGT, float* DC, const float* ext, int ext_last_only,
                          const float* WhhT, const float* WhhT_lo, long long wt_gstride, int L, int T, int Nn, int Bw,
                          int G, int t, int* err, cudaStream_t st) {
  WF_REQUIRE(L == 128, "tc_lstm_bwd: L=%d (only 128 is supported by the tensor-core step)", L);
  const int R = T * Nn;
  CUtensorMap tmA, tmBhi, tmBlo;
  int rc;
  if ((rc = map3(&tmA, XG, 4 * L, (uint64_t)G * Bw * R, 1, 4 * L, (uint64_t)G * Bw * R * 4 * L, 128))) return rc;
  if ((rc = map3(&tmBhi, WhhT, 4 * L, L, G, 4 * L, G > 1 ? wt_gstride : 4LL * L * L, 128))) return rc;
  if ((rc = map3(&tmBlo, WhhT_lo, 4 * L, L, G, 4 * L, G > 1 ? wt_gstride : 4LL * L * L, 128))) return rc;
  TcArgs a;
  tc_defaults(a);
  a.tile_mode = TILE_STEP; a.tiles_g = Bw * wf_cdiv(Nn, 128); a.Bw = Bw; a.R = R; a.Nn = Nn; a.T = T; a.t = t;
  a.nkb = t < T - 1 ? 4 * L / 32 : 0;
  a.Np = (Nn + 3) & ~3; a.RT = T * a.Np;
  a.XG = XG; a.Cst = const_cast<float*>(Cst); a.DGT = DGT; a.DC = DC; a.ext = ext; a.ext_last_only = ext_last_only;
  a.L = L; a.err = err;
  dim3 grid(1, a.tiles_g * G);
  return launch_variant<128, EPI_LSTM_BWD>(tmA, tmBhi, tmBlo, a, grid, st);
}

// Test / general entry point: C[g] = A[g] W[g]^T (+ bias + bias2, relu) on the tensor cores.
// W_lo must hold wf_split_lo(W).  err: one device int, set non-zero if a pipeline wait timed out.
extern "C" int wf_tc_gemm_nt(const float* A, int rows_g, int G, int K, const float* W, const float* W_lo,
                             long long w_group_stride, int N, const float* bias, const float* bias2,
                             long long bias_group_stride, int relu, float* C, int* err, void* stream) {
  WF_REQUIRE(rows_g > 0 && G > 0, "tc_gemm_nt: empty problem");
  return wf_launch_tc_rows(A, (long long)rows_g * G, K, rows_g, rows_g, G, K, W, W_lo, K, w_group_stride, w_group_stride,
                           0, N, bias, bias2, bias_group_stride, relu, C, N, (long long)rows_g * N, nullptr, nullptr,
                           nullptr, 0, 0, 0, 1, nullptr, nullptr, 0, err, (cudaStream_t)stream);
}
