// Persistent tcgen05 GEMM with 16-bit hi/lo operand splits (sm_100a) -- the dense products around the LSTM:
// GCN Theta transform with fused neighbour aggregation (model.py:23-26, hybrid_model.py:65-74), LSTM input
// projections (hybrid_model.py:42-49), dX = dG W_ih and the weight gradients dW = dG^T X of loss.backward()
// (train_hybrid_maml_v5.py:134,169).
//
//   D[128 x 128] (TMEM, fp32) += A[128 x 64] * B[128 x 64]^T   per k-block, as three kind::f16 products
//   a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo  (fp16 hi/lo in the forward pass: ~2^-20; bf16 hi/lo for gradient
//   operands, which need the fp32 exponent range: ~2^-16).  Twice the MMA rate and half the B bytes of 3xTF32.
//
// One CTA per SM, static round-robin over output tiles (n fastest, so the CTAs working on one A row block
// at the same time share it in L2).  Warp roles:
//   warp 0      TMA producer: A tile [128 x 64] fp32 (two SWIZZLE_128B boxes, or two TB4 boxes) + B_hi / B_lo
//               [128 x 64] 16-bit into a 3-stage shared-memory ring
//   warp 1      MMA issuer (TS mode: A from TMEM, B from smem); owns TMEM: 2 accumulator stages x 128 columns
//               + 3 A stages x 64 columns
//   warps 2..9  converters: thread = tile row x half a k-block; fp32 from smem (GCN rows with neighbours: the
//               pre-aggregated row, or a CSR gather-aggregate from global) -> hi/lo 16-bit pairs (packed F2FP
//               conversions) -> tcgen05.st into the TMEM A stage
//   warps 10..13 epilogue of the PREVIOUS tile while the next one is multiplied: tcgen05.ld -> bias / ReLU ->
//               row-major, TB4, or split-K partial stores (+ transposed bf16 hi/lo copies for the weight gradients)
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "wf_common.cuh"
#include "wf_rng.cuh"
#include "wf_tc.cuh"

using namespace wftc;

namespace {

enum { G16_ROWS = 0, G16_NODES = 1, G16_WGRAD = 2 };
constexpr int G16_THREADS = 448;  // producer, MMA, 8 converter warps, 4 epilogue warps
constexpr int G16_NST = 3;
constexpr int G16_A_BYTES = 32768, G16_B_BYTES = 16384, G16_STAGE = G16_A_BYTES + 2 * G16_B_BYTES;
constexpr int G16_SMEM = G16_NST * G16_STAGE + 1024;
// fp16 hi/lo operand splits hold |x| < 65520 (beyond, hi rounds to inf and lo = x - inf is NaN).  Layer outputs are
// flagged from 32768 up: the next layer's aggregation may still add neighbours of the same size.
constexpr float WF_F16_RANGE_LIMIT = 32768.0f;

struct G16Args {
  int mode;
  int m_tiles, n_tiles, G, splits;   // tile id -> (split, g, m tile, n tile), n fastest
  int rows_g, a_group_rows;          // ROWS: valid rows per group / row stride between groups in the A map
  const long long* a_win_off;        // ROWS, optional: element offset of every window (g*Bw + w) in A -- windows read in place
  int win_tiles;                     //   from a resident features tensor (dataset.py:36-37); tiles per window
  int Bw, R, Nn, T, tpw, rpt, Np, RT;  // rpt: nodes per TB4 node tile (wf_tile_rows)
  int nkb, nseg, nkb_split;          // k-blocks (64 wide) per segment; segments (wgrad: windows); split-K slice
  int a_k0, b_k0;
  int b_gmul;                        // 0: B shared by all groups
  int a_tb4;
  const float* a_raw; int lda;       // CSR gather on A (GCN aggregation), ROWS only
  int kreal;                         // real K of the operands (<= 64 * nkb; columns beyond are zero)
  const int* rowptr; const int* col; const float* val; long long g_rowptr, g_csr;
  const float* agg;                  // optional: pre-aggregated rows (same indexing as a_raw) for rows with neighbours
  float* C; int ldc; long long c_gstride, c_sstride; int c_cols;
  const float* bias; const float* bias2; long long bias_gstride; int relu;
  __nv_bfloat16* ct_hi; __nv_bfloat16* ct_lo;   // transposed copies [(g*Bw + w)][c_cols][RT]
  float* rowsum_part;                // WGRAD: per (split, k-half) partial row sums of A [parts][G][M] (bias gradients), or null
  DropCfg drop;                      // ROWS: dropout on the output (GCN sites); NODES: on the TB4 output (dX = mask of the LSTM site)
  float range_limit;                 // ROWS: > 0 flags |output| >= limit in err (the next layer splits it into fp16 hi/lo)
  int* err;
};

__host__ __device__ constexpr uint32_t g16_idesc(uint32_t fmt) {  // D = F32, A/B = fmt (0 F16, 1 BF16), K-major, M = N = 128
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void g16_mma(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

// (a, b) -> packed 16-bit hi pair and lo pair (a in the low half).  Packed conversions (F2FP) run on the ALU pipe;
// the scalar cvt.rn.f16.f32 goes through the quarter-rate conversion unit.
template <int FMT>
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  if (FMT == 0) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  } else {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xFFFF0000u));
    lo = *reinterpret_cast<const uint32_t*>(&l);
  }
}

struct TileCoord {
  int g, mt, ntile, split;
  int a_row, a_z0, b_z0, zstep;   // TMA coordinates
  int zt, blk, node0;             // NODES
  int row0, rlim;                 // ROWS / WGRAD: first output row of the tile in its group, exclusive row limit
  int kb0, nkb_loc;
};

__device__ __forceinline__ TileCoord decode_tile(const G16Args& a, int tile) {
  TileCoord c;
  c.ntile = tile % a.n_tiles; tile /= a.n_tiles;
  c.mt = tile % a.m_tiles; tile /= a.m_tiles;
  c.g = tile % a.G;
  c.split = tile / a.G;
  c.a_z0 = 0; c.zstep = 0; c.b_z0 = c.g * a.b_gmul; c.zt = 0; c.blk = 0; c.node0 = 0;
  c.kb0 = 0; c.nkb_loc = a.nkb;
  c.row0 = c.mt * 128; c.rlim = a.rows_g;
  if (a.mode == G16_ROWS && a.a_win_off != nullptr) {
    const int w = c.mt / a.win_tiles, mtw = c.mt - w * a.win_tiles;
    c.a_row = (int)(a.a_win_off[c.g * a.Bw + w] / a.lda) + mtw * 128;
    c.row0 = w * a.R + mtw * 128;
    c.rlim = w * a.R + a.R;
  } else if (a.mode == G16_ROWS) {
    c.a_row = c.g * a.a_group_rows + c.mt * 128;
  } else if (a.mode == G16_NODES) {
    const int ztl = c.mt / a.tpw, nt = c.mt - ztl * a.tpw;
    c.node0 = nt * a.rpt;
    c.zt = c.g * a.Bw * a.T + ztl;
    c.blk = c.zt * a.tpw + nt;
    c.a_row = c.node0;
  } else {
    c.a_row = c.mt * 128;
    c.a_z0 = c.g * a.Bw; c.b_z0 = c.g * a.Bw; c.zstep = 1;
    c.kb0 = c.split * a.nkb_split;
    c.nkb_loc = min(a.nkb - c.kb0, a.nkb_split);
  }
  return c;
}

template <int FMT, bool DROP>
__global__ void __launch_bounds__(G16_THREADS, 1)
wf_g16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
              const __grid_constant__ CUtensorMap tmBlo, const G16Args a) {
  constexpr int NST = G16_NST, STAGE = G16_STAGE, A_BYTES = G16_A_BYTES, B_BYTES = G16_B_BYTES;
  constexpr uint32_t A_COL = 256;  // TMEM: D stage ds at [128 ds, +128); A stage s at [256 + 64 s, +64) (hi 32 | lo 32)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[NST], aready[NST], empty[NST], dfull[2], dempty[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = a.n_tiles * a.m_tiles * a.G * a.splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&aready[s], 8); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&dfull[s], 1); mbar_init(&dempty[s], 4); }
    mbar_fence_init();
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmBhi); tma_prefetch_desc(&tmBlo);
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      // cursor over this CTA's (tile, segment, k-block) sequence.  (An L2 prefetch cursor running ahead of it was
      // measured and made things slightly slower: the ring is not waiting on HBM latency.)
      struct Cursor { int tile, seg, kb; TileCoord c; };
      auto cur_init = [&](Cursor& u, int tile) { u.tile = tile; u.seg = 0; if (tile < total) { u.c = decode_tile(a, tile); u.kb = u.c.kb0; } };
      auto cur_next = [&](Cursor& u) {
        if (++u.kb == u.c.kb0 + u.c.nkb_loc) { u.kb = u.c.kb0; if (++u.seg == a.nseg) cur_init(u, u.tile + gridDim.x); }
      };
      Cursor cu;
      cur_init(cu, blockIdx.x);
      int it = 0;
      while (cu.tile < total) {
        const int s = it % NST, ph = (it / NST) & 1;
        if (!mbar_wait(&empty[s], ph ^ 1)) { atomicExch(a.err, 31); break; }
        const TileCoord& c = cu.c;
        uint8_t* st = smem + s * STAGE;
        mbar_expect_tx(&full[s], STAGE);
        const int ka = a.a_k0 + cu.kb * 64, kbb = a.b_k0 + cu.kb * 64;
        if (a.mode == G16_NODES && a.a_tb4) {
          tma_load_4d(st, &tmA, &full[s], 0, 0, ka >> 2, c.blk);
          tma_load_4d(st + 16384, &tmA, &full[s], 0, 0, (ka >> 2) + 8, c.blk);
        } else {
          const int zc = a.mode == G16_NODES ? c.zt : c.a_z0 + cu.seg * c.zstep;
          tma_load_3d(st, &tmA, &full[s], ka, c.a_row, zc);
          tma_load_3d(st + 16384, &tmA, &full[s], ka + 32, c.a_row, zc);
        }
        const int zb = c.b_z0 + cu.seg * c.zstep;
        tma_load_3d(st + A_BYTES, &tmBhi, &full[s], kbb, c.ntile * 128, zb);
        tma_load_3d(st + A_BYTES + B_BYTES, &tmBlo, &full[s], kbb, c.ntile * 128, zb);
        cur_next(cu);
        ++it;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = g16_idesc(FMT);
    int it = 0, lt = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < total && ok; tile += gridDim.x, ++lt) {
      const TileCoord c = decode_tile(a, tile);
      const int ds = lt & 1, nk = c.nkb_loc * a.nseg;
      if (!mbar_wait(&dempty[ds], ((lt >> 1) & 1) ^ 1)) { if (lane == 0) atomicExch(a.err, 32); ok = false; break; }
      tc_fence_after();
      for (int k = 0; k < nk; ++k, ++it) {
        const int s = it % NST, ph = (it / NST) & 1;
        if (!mbar_wait(&full[s], ph) || !mbar_wait(&aready[s], ph)) { if (lane == 0) atomicExch(a.err, 33); ok = false; break; }
        tc_fence_after();
        if (lane == 0) {
          const uint32_t bhi = smem_u32(smem + s * STAGE + A_BYTES), blo = bhi + B_BYTES;
          const uint32_t acol = tbase + A_COL + s * 64;
#pragma unroll
          for (int p = 0; p < 3; ++p) {  // A_hi B_hi, A_lo B_hi, A_hi B_lo
            const uint32_t ac = acol + (p == 1 ? 32 : 0);
            const uint32_t bs = p == 2 ? blo : bhi;
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16)
              g16_mma(tbase + ds * 128, ac + k16 * 8, umma_desc_k_sw128(bs + k16 * 32), idesc, (k | p | k16) ? 1u : 0u);
          }
          umma_commit(&empty[s]);
          if (k == nk - 1) umma_commit(&dfull[ds]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------ converters (A: fp32 -> hi/lo 16-bit -> TMEM)
    // warps 2..5 take the first 32 k of every k-block, warps 6..9 the second 32 (TMEM lane quarter = warp & 3)
    const int q = warp & 3, row = q * 32 + lane, h = (warp - 2) >> 2;
    const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
    int it = 0;
    bool ok = true;
    // CSR row descriptor of this thread's row, fetched one tile ahead (two dependent global loads)
    int n_p0 = 0, n_p1 = 0, n_col = -1;
    float n_val = 0.f;
    auto fetch_row = [&](int tile) {
      n_p0 = 0; n_p1 = 0; n_col = -1; n_val = 0.f;
      if (a.mode == G16_ROWS && a.rowptr != nullptr && tile < total) {
        const TileCoord c = decode_tile(a, tile);
        const int grow = c.row0 + row;
        if (grow < c.rlim) {
          const int rr = grow % a.R;
          const int* rp = a.rowptr + c.g * a.g_rowptr;
          n_p0 = __ldg(rp + rr); n_p1 = __ldg(rp + rr + 1);
          n_col = __ldg(a.col + c.g * a.g_csr + n_p0);
          n_val = __ldg(a.val + c.g * a.g_csr + n_p0);
        }
      }
    };
    fetch_row(blockIdx.x);
    for (int tile = blockIdx.x; tile < total && ok; tile += gridDim.x) {
      const TileCoord c = decode_tile(a, tile);
      const int p0 = n_p0, p1 = n_p1;
      long long wbase = 0;
      int grow_in_w = 0;
      bool gather = false;
      if (a.mode == G16_ROWS && a.rowptr != nullptr) {  // rows whose aggregation is not the unit self loop
        const int grow = c.row0 + row;
        if (grow < c.rlim) {
          const int w = grow / a.R, rr = grow - w * a.R;
          grow_in_w = rr;
          wbase = ((long long)c.g * a.a_group_rows + (long long)w * a.R) * a.lda;
          gather = !(p1 - p0 == 1 && n_col == rr && n_val == 1.0f);
        }
      }
      fetch_row(tile + gridDim.x);
      const int nk = c.nkb_loc * a.nseg;
      const bool want_sum = a.rowsum_part != nullptr && c.ntile == 0;  // sum_k A[m, k]: the LSTM bias gradient, for free
      float rsum = 0.f;
      for (int k = 0; k < nk; ++k, ++it) {
        const int s = it % NST, ph = (it / NST) & 1;
        if (!mbar_wait(&full[s], ph) || !mbar_wait(&empty[s], ph ^ 1)) { if (lane == 0) atomicExch(a.err, 34); ok = false; break; }
        uint32_t hi[16], lo[16];
        if (!gather) {
          const uint32_t st = smem_u32(smem + s * STAGE + h * 16384);
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            // TB4 stage: [channel group][row][4 floats]; otherwise a SWIZZLE_128B box of 32 floats per row
            const float4 v = lds128(st + (a.a_tb4 ? ch * 2048 + row * 16 : row * 128 + ((ch ^ (row & 7)) << 4)));
            if (want_sum) rsum += (v.x + v.y) + (v.z + v.w);
            split_pair<FMT>(v.x, v.y, hi[2 * ch], lo[2 * ch]);
            split_pair<FMT>(v.z, v.w, hi[2 * ch + 1], lo[2 * ch + 1]);
          }
        } else {
          const int* cl = a.col + c.g * a.g_csr;
          const float* vl = a.val + c.g * a.g_csr;
          const int k0 = a.a_k0 + (c.kb0 + k % c.nkb_loc) * 64 + 32 * h;
          float acc[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = 0.f;
          if (a.agg != nullptr) {  // aggregated by wf_agg_rows_kernel beforehand: one row read
            const float4* src = reinterpret_cast<const float4*>(a.agg + wbase + (long long)grow_in_w * a.lda + k0);
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              const float4 x = k0 + 4 * ch < a.kreal ? __ldg(src + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
              acc[4 * ch + 0] = x.x; acc[4 * ch + 1] = x.y; acc[4 * ch + 2] = x.z; acc[4 * ch + 3] = x.w;
            }
          } else
          for (int p = p0; p < p1; ++p) {
            const float v = __ldg(vl + p);
            const float4* src = reinterpret_cast<const float4*>(a.a_raw + wbase + (long long)__ldg(cl + p) * a.lda + k0);
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              const float4 x = __ldg(src + ch);
              acc[4 * ch + 0] = fmaf(v, x.x, acc[4 * ch + 0]); acc[4 * ch + 1] = fmaf(v, x.y, acc[4 * ch + 1]);
              acc[4 * ch + 2] = fmaf(v, x.z, acc[4 * ch + 2]); acc[4 * ch + 3] = fmaf(v, x.w, acc[4 * ch + 3]);
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) split_pair<FMT>(acc[2 * j], acc[2 * j + 1], hi[j], lo[j]);
        }
        __syncwarp();  // gather / non-gather lanes diverged above
        tmem_st16(tlane + A_COL + s * 64 + 16 * h, hi);
        tmem_st16(tlane + A_COL + s * 64 + 32 + 16 * h, lo);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&aready[s]);
      }
      if (want_sum && ok)
        a.rowsum_part[((long long)(c.split * 2 + h) * a.G + c.g) * a.rows_g + c.mt * 128 + row] = rsum;
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
    int lt = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < total && ok; tile += gridDim.x, ++lt) {
      const TileCoord c = decode_tile(a, tile);
      const int ds = lt & 1, n0 = c.ntile * 128;
      if (!mbar_wait(&dfull[ds], (lt >> 1) & 1)) { if (lane == 0) atomicExch(a.err, 35); ok = false; break; }
      tc_fence_after();
      const float* b1 = a.bias ? a.bias + c.g * a.bias_gstride + n0 : nullptr;
      const float* b2 = a.bias2 ? a.bias2 + c.g * a.bias_gstride + n0 : nullptr;
      if (a.mode == G16_NODES) {
        // TB4 block = [c_cols / 4 channel groups][128 rows][4 floats]: 512 contiguous bytes per warp store
        float4* cblk = reinterpret_cast<float4*>(a.C) + ((long long)c.blk * (a.c_cols >> 2) + (n0 >> 2)) * 128 + row;
        DropState dst;
        unsigned long long e4row = 0;  // (element index of this row's column n0) / 4 in the canonical [G*Bw, T, N, c_cols] tensor
        if (DROP) {
          dst = wf_drop_state(a.drop);
          e4row = (((unsigned long long)c.zt * a.Nn + (unsigned)(c.node0 + row)) * (unsigned)a.c_cols + (unsigned)n0) >> 2;
        }
#pragma unroll 1
        for (int cc = 0; cc < 128; cc += 32) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld32(tlane + ds * 128 + cc, v);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            if (b1) { const float4 b = __ldg(reinterpret_cast<const float4*>(b1 + cc + j)); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
            if (b2) { const float4 b = __ldg(reinterpret_cast<const float4*>(b2 + cc + j)); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
            if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (DROP) {
              float m[4];
              wf_drop4(dst, e4row + (unsigned)((cc + j) >> 2), m);
              o.x *= m[0]; o.y *= m[1]; o.z *= m[2]; o.w *= m[3];
            }
            if (row < a.rpt) cblk[(long long)((cc + j) >> 2) * 128] = o;  // rows >= rpt of a node tile are padding
          }
        }
      } else {
        const int grow = c.row0 + row;
        const bool valid = grow < c.rlim;
        float* crow = a.C + c.g * a.c_gstride + c.split * a.c_sstride + (long long)grow * a.ldc + n0;
        long long ctbase = 0;
        if (a.ct_hi != nullptr && valid) {
          const int w = grow / a.R, rr = grow - w * a.R;
          const int tt = rr / a.Nn, nn = rr - tt * a.Nn;
          ctbase = ((long long)(c.g * a.Bw + w) * a.c_cols + n0) * a.RT + (long long)tt * a.Np + nn;
        }
        DropState dst;
        unsigned long long e4row = 0;  // (element index of this row's column n0) / 4 in the canonical [G*rows_g, c_cols] tensor
        if (DROP) {
          dst = wf_drop_state(a.drop);
          e4row = (((unsigned long long)c.g * (unsigned)a.rows_g + (unsigned)grow) * (unsigned)a.c_cols + (unsigned)n0) >> 2;
        }
        float amax = 0.f;
#pragma unroll 1
        for (int cc = 0; cc < 128; cc += 32) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld32(tlane + ds * 128 + cc, v);
          tmem_wait_ld();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
              if (b1) { const float4 b = __ldg(reinterpret_cast<const float4*>(b1 + cc + j)); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
              if (b2) { const float4 b = __ldg(reinterpret_cast<const float4*>(b2 + cc + j)); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
              if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              if (DROP) {
                float m[4];
                wf_drop4(dst, e4row + (unsigned)((cc + j) >> 2), m);
                o.x *= m[0]; o.y *= m[1]; o.z *= m[2]; o.w *= m[3];
              }
              amax = fmaxf(amax, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
              *reinterpret_cast<float4*>(crow + cc + j) = o;
              if (a.ct_hi != nullptr) {  // lanes of a warp hold consecutive rows -> contiguous transposed stores
                const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const long long ti = ctbase + (long long)(cc + j + e) * a.RT;
                  uint32_t ph, pl;
                  split_pair<1>(ov[e], 0.f, ph, pl);
                  reinterpret_cast<uint16_t*>(a.ct_hi)[ti] = (uint16_t)ph;
                  reinterpret_cast<uint16_t*>(a.ct_lo)[ti] = (uint16_t)pl;
                }
              }
            }
          }
        }
        // an activation the next layer cannot split into fp16 hi/lo (|x| >= 65520 rounds to inf), or a non-finite one:
        // flag it instead of propagating NaN silently (engine.check() reports it; precision="fp32" has no such limit)
        if (a.range_limit > 0.f && valid && !(amax < a.range_limit)) atomicExch(a.err, 41);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dempty[ds]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, 512);
}

// fp32 -> 16-bit hi / lo (fmt 0: fp16, 1: bf16), elementwise; blockIdx.y = group (own source / destination stride)
__global__ void wf_split16_kernel(const float4* __restrict__ src, long long src_gstride4, uint2* __restrict__ hi,
                                  uint2* __restrict__ lo, long long dst_gstride4, long long quads, int fmt) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= quads) return;
  const float4 v = src[blockIdx.y * src_gstride4 + i];
  i += blockIdx.y * dst_gstride4;
  uint2 h, l;
  if (fmt == 0) { split_pair<0>(v.x, v.y, h.x, l.x); split_pair<0>(v.z, v.w, h.y, l.y); }
  else { split_pair<1>(v.x, v.y, h.x, l.x); split_pair<1>(v.z, v.w, h.y, l.y); }
  hi[i] = h;
  lo[i] = l;
}

// out1[g][m] = out2[g][m] = sum over parts of part[p][g][m] (fixed order: deterministic)
__global__ void wf_sum_rowsum_parts_kernel(const float* __restrict__ part, int parts, int G, int M, float* out1, float* out2,
                                           long long out_gstride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G * M) return;
  float acc = 0.f;
  for (int p = 0; p < parts; ++p) acc += part[(long long)p * G * M + i];
  const int g = i / M, m = i - g * M;
  out1[g * out_gstride + m] = acc;
  if (out2) out2[g * out_gstride + m] = acc;
}

int map16(CUtensorMap* m, const void* base, uint64_t k, uint64_t rows, uint64_t z, uint64_t ld, uint64_t zstride, int fmt) {
  uint64_t dims[3] = {k, rows, z};
  uint64_t str[2] = {ld * 2, zstride * 2};
  uint32_t box[3] = {64, 128, 1};
  return wf_encode_tensor_map(m, base, 3, dims, str, box, 1, fmt == 0 ? 1 : 2);
}
int map32(CUtensorMap* m, const float* base, uint64_t k, uint64_t rows, uint64_t z, uint64_t ld, uint64_t zstride) {
  uint64_t dims[3] = {k, rows, z};
  uint64_t str[2] = {ld * 4, zstride * 4};
  uint32_t box[3] = {32, 128, 1};
  return wf_encode_tensor_map(m, base, 3, dims, str, box, 1, 0);
}

int g16_launch(int fmt, const CUtensorMap& tmA, const CUtensorMap& tmBhi, const CUtensorMap& tmBlo, const G16Args& a, cudaStream_t st) {
  static bool configured = false;
  static int sms = 148;
  if (!configured) {
    if (cudaFuncSetAttribute(wf_g16_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, G16_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(wf_g16_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, G16_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(wf_g16_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, G16_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(wf_g16_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, G16_SMEM) != cudaSuccess)
      return wf_fail(WF_ECUDA, "g16 kernel: cannot raise dynamic shared memory to %d", G16_SMEM);
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    configured = true;
  }
  const long long total = (long long)a.n_tiles * a.m_tiles * a.G * a.splits;
  const int grid = (int)(total < sms ? total : sms);
  const bool drop = a.drop.rng != nullptr && a.mode != G16_WGRAD;
  if (fmt == 0 && !drop) wf_g16_kernel<0, false><<<grid, G16_THREADS, G16_SMEM, st>>>(tmA, tmBhi, tmBlo, a);
  else if (fmt == 0) wf_g16_kernel<0, true><<<grid, G16_THREADS, G16_SMEM, st>>>(tmA, tmBhi, tmBlo, a);
  else if (!drop) wf_g16_kernel<1, false><<<grid, G16_THREADS, G16_SMEM, st>>>(tmA, tmBhi, tmBlo, a);
  else wf_g16_kernel<1, true><<<grid, G16_THREADS, G16_SMEM, st>>>(tmA, tmBhi, tmBlo, a);
  WF_CHECK_LAUNCH("g16_kernel");
  return WF_OK;
}

void g16_defaults(G16Args& a) {
  memset(&a, 0, sizeof(a));
  a.G = 1; a.splits = 1; a.nseg = 1; a.b_gmul = 1; a.Bw = 1; a.R = 1; a.Nn = 1; a.T = 1; a.tpw = 1; a.nkb_split = 1 << 30;
}

}  // namespace

// Column pitch of one time slice in the transposed copies: N rounded up to 8, so 16-bit rows stay 16-byte aligned
// for TMA whatever T is.
int wf_np(int N) { return (N + 7) & ~7; }

// Nodes per TB4 node tile.  A window's N nodes occupy ceil(N / 128) tiles of 128 ROWS each in memory; the nodes are dealt
// evenly (rounded up to 8) instead of 128 per tile with a short last one: 441 nodes = 4 x 112 (- 7), so every CTA of the
// recurrence kernels moves 12.5 % fewer bytes per step and none waits for a full-tile neighbour.  Rows >= wf_tile_rows of
// a tile are padding: never read as data, written only by GEMM epilogues that store whole tiles.
extern "C" int wf_tile_rows(int N) {
  const int tiles = (N + 127) / 128;
  if (tiles <= 0) return 128;
  const int r = (((N + tiles - 1) / tiles) + 7) & ~7;
  return r < 128 ? r : 128;
}

// G groups of n values each; group g reads src + g*src_gstride and writes hi/lo + g*dst_gstride (elements).
int wf_launch_split16(const float* src, long long src_gstride, void* hi, void* lo, long long dst_gstride, long long n, int G,
                      int fmt, cudaStream_t st) {
  WF_REQUIRE(n > 0 && n % 4 == 0 && src_gstride % 4 == 0 && dst_gstride % 4 == 0, "split16: sizes must be multiples of 4");
  wf_split16_kernel<<<dim3(wf_cdiv(n / 4, 256), G), 256, 0, st>>>((const float4*)src, src_gstride / 4, (uint2*)hi, (uint2*)lo,
                                                                dst_gstride / 4, n / 4, fmt);
  WF_CHECK_LAUNCH("split16");
  return WF_OK;
}

// AGG[z][rr] = sum_p val[p] * X[z][col[p]] for the listed rows rr of every window z (rows whose aggregation is not
// the unit self loop; -1 entries are padding).  grid (ceil(nlist / 8), G*Bw), one warp per (row, window).
static __global__ void __launch_bounds__(256) wf_agg_rows_kernel(const float* __restrict__ X, const long long* __restrict__ x_win_off,
                                                                 float* __restrict__ AGG, int C, int R, int Bw,
                                                                 const int* __restrict__ rowptr, const int* __restrict__ col,
                                                                 const float* __restrict__ val, long long g_rowptr, long long g_csr,
                                                                 const int* __restrict__ list, int nlist, long long g_list) {
  const int z = blockIdx.y, g = z / Bw, i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= nlist) return;
  const int rr = list[g * g_list + i];
  if (rr < 0) return;
  const int* rp = rowptr + g * g_rowptr;
  const int p0 = rp[rr], p1 = rp[rr + 1];
  const float* Xz = X + (x_win_off ? x_win_off[z] : (long long)z * R * C);
  for (int c4 = lane; c4 < (C >> 2); c4 += 32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = p0; p < p1; ++p) {
      const float v = __ldg(val + g * g_csr + p);
      const float4 x = __ldg(reinterpret_cast<const float4*>(Xz + (long long)__ldg(col + g * g_csr + p) * C) + c4);
      acc.x = fmaf(v, x.x, acc.x); acc.y = fmaf(v, x.y, acc.y); acc.z = fmaf(v, x.z, acc.z); acc.w = fmaf(v, x.w, acc.w);
    }
    reinterpret_cast<float4*>(AGG + ((long long)z * R + rr) * C)[c4] = acc;
  }
}

// ---- C[g] = (A_hat[g]) A[g] W[g]^T (+bias, +bias2, relu): rows tiled, optional CSR gather on A and transposed
// bf16 hi/lo copies of C.  W16 hi/lo: [Gb][N, K] 16-bit (fmt).
int wf_launch_g16_rows(int fmt, const float* A, long long a_rows_total, int lda, int a_group_rows, int rows_g, int G, int K,
                       const void* Whi, const void* Wlo, int ldb, long long b_gstride, int b_shared, int N, const float* bias,
                       const float* bias2, long long bias_gstride, int relu, float* C, int ldc, long long c_gstride,
                       const int* rowptr, const int* col, const float* val, long long g_rowptr, long long g_csr, int R, int Bw,
                       void* ct_hi, void* ct_lo, int Nn, const float* agg, int* err, cudaStream_t st,
                       const long long* a_win_off = nullptr, int kpad = 0, const DropCfg* drop = nullptr,
                       float range_limit = 0.f) {
  // kpad: the K extent seen by the k-loop when the operands' real K (their row length) is shorter and not a multiple of
  // 64: TMA zero-fills the columns beyond K (the 24-channel first GCN layer)
  if (kpad > 0) { WF_REQUIRE(K % 8 == 0 && kpad % 64 == 0 && kpad >= K, "g16_rows: bad K padding %d -> %d", K, kpad); }
  else WF_REQUIRE(K % 64 == 0 && K >= 64, "g16_rows: K=%d must be a multiple of 64", K);
  WF_REQUIRE(a_win_off == nullptr || (rowptr == nullptr || agg != nullptr), "g16_rows: windowed A needs pre-aggregated rows");
  WF_REQUIRE(N % 128 == 0, "g16_rows: N=%d must be a multiple of 128", N);
  WF_REQUIRE(lda % 4 == 0 && ldb % 8 == 0 && ldc % 4 == 0, "g16_rows: leading dimensions must keep 16-byte alignment");
  WF_REQUIRE(((uintptr_t)A | (uintptr_t)Whi | (uintptr_t)Wlo | (uintptr_t)C) % 16 == 0, "g16_rows: pointers must be 16-byte aligned");
  const int Gb = b_shared ? 1 : G;
  CUtensorMap tmA, tmBhi, tmBlo;
  int rc;
  if ((rc = map32(&tmA, A, K, a_rows_total, 1, lda, (uint64_t)a_rows_total * lda))) return rc;
  if ((rc = map16(&tmBhi, Whi, K, N, Gb, ldb, Gb > 1 ? b_gstride : (long long)N * ldb, fmt))) return rc;
  if ((rc = map16(&tmBlo, Wlo, K, N, Gb, ldb, Gb > 1 ? b_gstride : (long long)N * ldb, fmt))) return rc;
  G16Args a;
  g16_defaults(a);
  a.mode = G16_ROWS; a.m_tiles = wf_cdiv(rows_g, 128); a.n_tiles = N / 128; a.G = G;
  a.rows_g = rows_g; a.a_group_rows = a_group_rows; a.nkb = (kpad > 0 ? kpad : K) / 64; a.b_gmul = b_shared ? 0 : 1;
  a.a_raw = A; a.lda = lda; a.kreal = K; a.rowptr = rowptr; a.col = col; a.val = val; a.g_rowptr = g_rowptr; a.g_csr = g_csr; a.agg = agg;
  a.R = R > 0 ? R : rows_g; a.Bw = Bw > 0 ? Bw : 1;
  a.Nn = Nn > 0 ? Nn : a.R; a.Np = wf_np(a.Nn); a.RT = (a.R / a.Nn) * a.Np;
  WF_REQUIRE(ct_hi == nullptr || a.R % a.Nn == 0, "g16_rows: transposed copies need R to be a multiple of the node count");
  if (a_win_off != nullptr) { a.a_win_off = a_win_off; a.win_tiles = wf_cdiv(a.R, 128); a.m_tiles = a.Bw * a.win_tiles; }
  a.C = C; a.ldc = ldc; a.c_gstride = c_gstride; a.c_cols = N; a.bias = bias; a.bias2 = bias2; a.bias_gstride = bias_gstride;
  a.relu = relu; a.ct_hi = (__nv_bfloat16*)ct_hi; a.ct_lo = (__nv_bfloat16*)ct_lo; a.err = err;
  if (drop != nullptr) a.drop = *drop;
  a.range_limit = range_limit;
  return g16_launch(fmt, tmA, tmBhi, tmBlo, a, st);
}

// ---- C = A W^T (+bias + bias2) tiled per (window, step, 128 nodes), output in the TB4 layout.
// A: row-major [G*Bw*T*Nn, K] (a_tb4 == 0) or a TB4 buffer with K channels (a_tb4 == 1).
int wf_launch_g16_nodes(int fmt, const float* A, int a_tb4, int K, const void* Whi, const void* Wlo, int ldb, long long b_gstride,
                        int N, const float* bias, const float* bias2, long long bias_gstride, float* C, int T, int Nn, int Bw,
                        int G, int* err, cudaStream_t st, const DropCfg* drop) {
  WF_REQUIRE(K % 64 == 0 && K >= 64, "g16_nodes: K=%d must be a multiple of 64", K);
  WF_REQUIRE(N % 128 == 0 && ldb % 8 == 0, "g16_nodes: N=%d must be a multiple of 128, ldb of 8", N);
  WF_REQUIRE(((uintptr_t)A | (uintptr_t)Whi | (uintptr_t)Wlo | (uintptr_t)C) % 16 == 0, "g16_nodes: pointers must be 16-byte aligned");
  const int tpw = wf_cdiv(Nn, 128);
  const long long ZT = (long long)G * Bw * T;
  CUtensorMap tmA, tmBhi, tmBlo;
  int rc;
  if (a_tb4) {
    uint64_t dims[4] = {256, 2, (uint64_t)(K / 4), (uint64_t)(ZT * tpw)};
    uint64_t str[3] = {256 * 4, 512 * 4, (uint64_t)K * 128 * 4};
    uint32_t box[4] = {256, 2, 8, 1};
    if ((rc = wf_encode_tensor_map(&tmA, A, 4, dims, str, box, 0, 0))) return rc;
  } else {
    if ((rc = map32(&tmA, A, K, Nn, ZT, K, (uint64_t)Nn * K))) return rc;
  }
  if ((rc = map16(&tmBhi, Whi, K, N, G, ldb, G > 1 ? b_gstride : (long long)N * ldb, fmt))) return rc;
  if ((rc = map16(&tmBlo, Wlo, K, N, G, ldb, G > 1 ? b_gstride : (long long)N * ldb, fmt))) return rc;
  G16Args a;
  g16_defaults(a);
  a.mode = G16_NODES; a.tpw = tpw; a.rpt = wf_tile_rows(Nn); a.a_tb4 = a_tb4; a.m_tiles = Bw * T * tpw; a.n_tiles = N / 128; a.G = G;
  a.Bw = Bw; a.T = T; a.Nn = Nn; a.nkb = K / 64;
  a.C = C; a.c_cols = N; a.bias = bias; a.bias2 = bias2; a.bias_gstride = bias_gstride; a.err = err;
  if (drop != nullptr) a.drop = *drop;
  return g16_launch(fmt, tmA, tmBhi, tmBlo, a, st);
}

// Sum of split-K partials: out[g][i] = sum_s part[s][g][i] (fixed order: deterministic).
static __global__ void wf_sum_splits16_kernel(const float4* __restrict__ part, int splits, long long count4, long long part_sstride4,
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

// ---- weight gradient: dW[g][M, N] = sum over windows w and columns k of AT[g*Bw+w][m, a_k0+k] * BT[g*Bw+w][n, b_k0+k]
// AT: fp32 [G*Bw][M][R]; BT hi/lo: bf16 [G*Bw][N][R].  Split-K over `partials` when there are few output tiles.
// rowsum1 / rowsum2 (optional): sum over windows and k of AT[m, k] -> [g][M] (the two LSTM bias gradients), computed in
// the converter path; needs 16*G*M extra floats at the end of `partials`.
int wf_launch_g16_wgrad(const float* AT, int M, const void* BT_hi, const void* BT_lo, int N, int R, int Bw, int G, int a_k0,
                        int b_k0, int klen, float* dW, long long dw_gstride, int* err, cudaStream_t st, float* partials,
                        size_t partial_floats, float* rowsum1, float* rowsum2, long long rowsum_gstride) {
  WF_REQUIRE(M % 128 == 0 && N % 128 == 0 && R % 8 == 0, "g16_wgrad: M=%d N=%d must be multiples of 128, R=%d of 8", M, N, R);
  WF_REQUIRE(a_k0 % 4 == 0 && b_k0 % 8 == 0, "g16_wgrad: K offsets (%d, %d) must keep 16-byte alignment", a_k0, b_k0);
  CUtensorMap tmA, tmBhi, tmBlo;
  int rc;
  const uint64_t Z = (uint64_t)G * Bw;
  // the K extent seen through the maps ends at a_k0 + klen / b_k0 + klen: everything beyond is zero-filled
  if ((rc = map32(&tmA, AT, (uint64_t)a_k0 + klen, M, Z, R, (uint64_t)M * R))) return rc;
  if ((rc = map16(&tmBhi, BT_hi, (uint64_t)b_k0 + klen, N, Z, R, (uint64_t)N * R, 1))) return rc;
  if ((rc = map16(&tmBlo, BT_lo, (uint64_t)b_k0 + klen, N, Z, R, (uint64_t)N * R, 1))) return rc;
  const int tiles = (M / 128) * (N / 128) * G, nkb = wf_cdiv(klen, 64);
  int splits = 148 / (tiles > 0 ? tiles : 1);
  if (splits < 1) splits = 1;
  if (splits > 8) splits = 8;
  if (splits > nkb) splits = nkb;
  int per = wf_cdiv(nkb, splits);
  splits = wf_cdiv(nkb, per);
  const long long per_split = (long long)G * M * N, rs_floats = rowsum1 ? 16LL * G * M : 0;
  WF_REQUIRE(rowsum1 == nullptr || (partials != nullptr && (long long)partial_floats >= rs_floats),
             "g16_wgrad: row sums need %lld floats of scratch", rs_floats);
  if (partials == nullptr || (long long)partial_floats < per_split * splits + rs_floats || (dw_gstride % 4) != 0) { splits = 1; per = nkb; }
  float* rs_part = rowsum1 ? partials + (partial_floats - rs_floats) : nullptr;
  G16Args a;
  g16_defaults(a);
  a.mode = G16_WGRAD; a.m_tiles = M / 128; a.n_tiles = N / 128; a.G = G; a.splits = splits;
  a.rows_g = M; a.Bw = Bw; a.R = R; a.nkb = nkb; a.nseg = Bw; a.nkb_split = per; a.a_k0 = a_k0; a.b_k0 = b_k0;
  a.ldc = N; a.c_cols = N; a.err = err; a.rowsum_part = rs_part;
  if (splits > 1) { a.C = partials; a.c_gstride = (long long)M * N; a.c_sstride = per_split; }
  else { a.C = dW; a.c_gstride = dw_gstride; a.c_sstride = 0; }
  rc = g16_launch(1, tmA, tmBhi, tmBlo, a, st);
  if (rc) return rc;
  if (rs_part) {
    wf_sum_rowsum_parts_kernel<<<wf_cdiv(G * M, 256), 256, 0, st>>>(rs_part, 2 * splits, G, M, rowsum1, rowsum2, rowsum_gstride);
    WF_CHECK_LAUNCH("sum_rowsum_parts");
  }
  if (splits == 1) return WF_OK;
  const long long count4 = (long long)M * N / 4;
  wf_sum_splits16_kernel<<<dim3(wf_cdiv(count4, 256), G), 256, 0, st>>>((const float4*)partials, splits, count4, per_split / 4,
                                                                     (float4*)dW, dw_gstride / 4);
  WF_CHECK_LAUNCH("sum_splits");
  return WF_OK;
}

// out[g][c][r] = split16(in[g][r][c]): W[rows, cols] -> W^T[cols, rows] as bf16 hi / lo for every group.
static __global__ void wf_transpose_split16_kernel(const float* __restrict__ in, long long in_gstride, int rows, int cols,
                                                   __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
                                                   long long out_gstride) {
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
      const float v = t[threadIdx.x][i];
      const long long o = g * out_gstride + (long long)c * rows + r;
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      out_hi[o] = h;
      out_lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

int wf_launch_transpose_split16(const float* in, long long in_gstride, int rows, int cols, void* out_hi, void* out_lo,
                                long long out_gstride, int G, cudaStream_t st) {
  dim3 grid(wf_cdiv(cols, 32), wf_cdiv(rows, 32), G), block(32, 8);
  wf_transpose_split16_kernel<<<grid, block, 0, st>>>(in, in_gstride, rows, cols, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo,
                                                      out_gstride);
  WF_CHECK_LAUNCH("transpose_split16");
  return WF_OK;
}

// fp32 -> 16-bit hi / lo operand halves (fmt 0: fp16, 1: bf16); n a multiple of 4.
extern "C" int wf_split16(const float* src, void* hi, void* lo, long long n, int fmt, void* stream) {
  WF_REQUIRE(fmt == 0 || fmt == 1, "split16: fmt must be 0 (fp16) or 1 (bf16)");
  return wf_launch_split16(src, 0, hi, lo, 0, n, 1, fmt, (cudaStream_t)stream);
}

// Row pitch of the 16-bit transposed activation copies [(G*Bw)][channels][RT16] (and of the fp32 dG^T scratch that
// pairs with them): column (t, node) = t*Np + node with Np = N rounded up to 8.  Padding columns must be zero.
extern "C" long long wf_transposed_pitch16(int T, int N) { return (long long)T * wf_np(N); }

// GCNConv + ReLU (model.py:31-42, hybrid_model.py:65-75): Y = relu((A_hat X) W^T + b) on the persistent fp16 hi/lo GEMM
// with the neighbour aggregation fused into the A-operand path.  X dense [G*Bw*R, Cin], Cin % 64 == 0, Cout % 128 == 0,
// W16 hi/lo = wf_split16(W, fmt 0), shared by all groups; YT hi/lo (optional): bf16 transposed copies [(G*Bw)][Cout][RT16].
// gather_rows (optional): per group the rows of a window whose aggregation is not the unit self loop, padded with -1 to
// gather_max, with `agg` a scratch [G*Bw*R, Cin]: those rows are aggregated by a small pre-pass and the GEMM reads one
// row instead of walking the CSR inside its operand pipeline.
// x_win_off (optional): element offset of every window in X, a resident features tensor of x_rows_total rows of Cin
// floats (the first layer reads windows in place, dataset.py:36-37); Cin only has to be a multiple of 8 then (TMA
// zero-fills the k-block).
extern "C" int wf_gcn_layer_fwd_g16(const float* X, const long long* x_win_off, long long x_rows_total, const void* W16_hi,
                                    const void* W16_lo, const float* bias, const int* rowptr, const int* col, const float* val,
                                    long long rowptr_group_stride, long long csr_group_stride, const int* gather_rows,
                                    int gather_max, long long gather_group_stride, float* agg, int R, int N, int Cin, int Cout,
                                    int G, int Bw, int relu, float* Y, void* YT_hi, void* YT_lo, float p_drop,
                                    const unsigned long long* rng, int site, int* err, void* stream) {
  WF_REQUIRE(G > 0 && Bw > 0 && R > 0 && N > 0, "gcn_layer_fwd_g16: bad batch");
  WF_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "gcn_layer_fwd_g16: p_drop=%f outside [0, 1)", (double)p_drop);
  WF_REQUIRE(p_drop == 0.f || rng != nullptr, "gcn_layer_fwd_g16: dropout needs the rng state");
  const DropCfg drop = wf_drop_cfg(p_drop, rng, WF_SITE_GCN + site);
  const long long rows_g = (long long)Bw * R;
  cudaStream_t st = (cudaStream_t)stream;
  const bool pre = rowptr != nullptr && gather_rows != nullptr && agg != nullptr && gather_max > 0;
  WF_REQUIRE(x_win_off == nullptr || rowptr == nullptr || pre, "gcn_layer_fwd_g16: windows read in place need gather_rows + agg");
  if (pre) {
    WF_REQUIRE(Cin % 4 == 0, "gcn_layer_fwd_g16: Cin must be a multiple of 4");
    wf_agg_rows_kernel<<<dim3(wf_cdiv(gather_max, 8), G * Bw), 256, 0, st>>>(X, x_win_off, agg, Cin, R, Bw, rowptr, col, val,
                                                                            rowptr_group_stride, csr_group_stride, gather_rows,
                                                                            gather_max, gather_group_stride);
    WF_CHECK_LAUNCH("agg_rows");
  }
  const int kpad = Cin % 64 == 0 ? 0 : (Cin + 63) / 64 * 64;
  return wf_launch_g16_rows(0, X, x_win_off ? x_rows_total : rows_g * G, Cin, (int)rows_g, (int)rows_g, G, Cin, W16_hi, W16_lo, Cin,
                            0, 1, Cout, bias, nullptr, 0, relu, Y, Cout, rows_g * Cout, rowptr, col, val, rowptr_group_stride,
                            csr_group_stride, R, Bw, YT_hi, YT_lo, N, pre ? agg : nullptr, err, st, x_win_off, kpad, &drop,
                            WF_F16_RANGE_LIMIT);
}

// Test / general entry point: C[g] = A[g] W[g]^T (+ bias + bias2, relu), W16 hi/lo [G][N, K] from wf_split16(W, fmt).
extern "C" int wf_g16_gemm_nt(const float* A, int rows_g, int G, int K, const void* W16_hi, const void* W16_lo,
                              long long w_group_stride, int N, const float* bias, const float* bias2, long long bias_group_stride,
                              int relu, int fmt, float* C, int* err, void* stream) {
  WF_REQUIRE(rows_g > 0 && G > 0, "g16_gemm_nt: empty problem");
  WF_REQUIRE(fmt == 0 || fmt == 1, "g16_gemm_nt: fmt must be 0 (fp16) or 1 (bf16)");
  return wf_launch_g16_rows(fmt, A, (long long)rows_g * G, K, rows_g, rows_g, G, K, W16_hi, W16_lo, K, w_group_stride, 0, N, bias,
                            bias2, bias_group_stride, relu, C, N, (long long)rows_g * N, nullptr, nullptr, nullptr, 0, 0, 0, 1,
                            nullptr, nullptr, 0, nullptr, err, (cudaStream_t)stream);
}
