// SS-mode tcgen05 GEMMs on PRE-SPLIT 16-bit hi/lo operands (sm_100a) -- second generation of the dense products around
// the LSTM and of the GCN Theta transform (model.py:23-26, hybrid_model.py:42-49,65-74; loss.backward()).
//
//   D[128 x BN] (TMEM, fp32) += A[128 x 32] * B[BN x 32]^T  per k-block, three kind::f16 products
//   a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo   (fp16 pairs forward, bf16 pairs where an operand is a gradient)
//
// What changed against csrc/wf_gemm16.cu, and why:
//  * every producer kernel writes its output already split (two 16-bit planes, the same bytes as one fp32), so both
//    operands go TMA -> shared memory -> tensor core: no converter warps, no TMEM operand hop, a pipeline of two barriers
//    per stage instead of four;
//  * the weight operand B is RESIDENT in shared memory (BN x K x hi/lo = 128 KB) for all tiles of a CTA: the L2 -> SM path
//    (~42 B/clk/SM) carries the activations once per n-part and nothing else -- re-streaming B per tile was 50 % of the
//    bytes through L2 in the first generation;
//  * CTAs own CONTIGUOUS ranges of row tiles of ONE n-part, so a CTA changes task (reloads B) at most once or twice;
//  * GCN outputs leave through shared memory and cp.async.bulk.tensor stores (UTMASTG), full 128-byte lines.
//
// Activation layouts the operands are read from:
//   row-major hl16   [2 planes][rows][K]                     GCN activations (fp16): K-major, SWIZZLE_64B boxes
//   TB8              [2 planes][block][K/8][128 rows][8]     LSTM h (fp16), dG (bf16); block = (window, step, node tile):
//                    no-swizzle canonical layouts -- K-major for the projections / dX (rows = M), and the SAME bytes
//                    MN-major for the weight gradients (rows = K), so nothing is ever transposed in memory.
#include <type_traits>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "wf_common.cuh"
#include "wf_layout.cuh"
#include "wf_rng.cuh"
#include "wf_tc.cuh"

using namespace wftc;

int wf_np(int N);
extern "C" int wf_tile_rows(int N);

namespace {

constexpr int SS_THREADS = 192;  // warp 0: TMA producer; warp 1: MMA issue + TMEM owner; warps 2-5: epilogue (lane quarters 2,3,0,1)
constexpr int SS_THREADS_HL = 320;  // the hl16 epilogue runs on 8 warps (2-9): a warp's ~12 dependent instructions per value leave
                                    // its scheduler idle most of the time, so two warps per scheduler each take half of the columns
// One TMA tensor load costs ~0.2 us of serialised issue whatever its size (measured: tools/ss_bench.py ablations, and the
// first generation's 4 loads per k-block = 1.3 us), so a stage is ONE load: the hi and the lo plane of a [128 x 64] operand
// tile together (the plane is the outermost box dimension) = 32 KB.
constexpr int SS_BK = 64;
constexpr int SS_A_PLANE = 128 * SS_BK * 2;   // 16 KB: one 16-bit plane of an A stage
constexpr int SS_A_STAGE = 2 * SS_A_PLANE;    // hi + lo
constexpr int SS_B_BYTES = 131072;            // resident B: BN x K x 2 B x 2 planes
constexpr int SS_STG_BYTES = 32768;           // GCN epilogue staging: 8 warps x (hi + lo) x 32 rows x 64 B
constexpr int SS_MAX_STAGES = 6;

enum { SS_A_KS = 0, SS_A_KT = 1 };    // A: K-major SWIZZLE_128B from row-major hl16 / K-major no-swizzle from TB8
enum { SS_E_TB4 = 0, SS_E_HL = 1 };   // epilogue: fp32 TB4 block (+bias) / row-major hl16 through TMA stores (+bias, ReLU)
enum { SS_ROWS = 0, SS_NODES = 1 };   // row tiles: 128 rows of a window / the node tile of one (window, step)

struct SsArgs {
  int mode, avar, epi;
  int n_parts;          // N_total / BN
  int m_tiles_g, G;     // row tiles per group, groups
  int nkb;              // K / 64 (rounded up: TMA zero-fills beyond K)
  int nst;              // A ring depth
  int k_parts;          // >= 1: the K range is dealt over this many CTAs per row-tile range (nkb = k-blocks of ONE part); their
                        // partial products meet in a zeroed TB4 output through red.global.add (two addends: order-free)
  int b_per_group;      // weights differ per group: reload the resident B when a CTA's range crosses into the next group
  int afmt, bfmt;       // 0 fp16, 1 bf16
  // ROWS: windows of R rows; the first agg_rows rows of every window come from the side buffer (aggregated rows)
  int R, Bw, win_tiles, agg_tiles;
  // NODES
  int T, Nn, tpw, rpt;
  // epilogue
  float* C; int c_cols;
  const float* bias; const float* bias2; long long bias_gstride; int relu;
  DropCfg drop; float range_limit;
  int out2;             // E_HL: also write bf16 planes through tmOut2
  int pf;               // L2 prefetch distance in stages (0: off)
  int debug;            // WF_SS_DEBUG (tools/ss_bench.py ablations): 1 no MMA, 2 no epilogue stores
  int* err;
};

__host__ __device__ constexpr uint32_t ss_idesc(int n, uint32_t afmt, uint32_t bfmt, uint32_t a_mn, uint32_t b_mn, int m = 128) {
  return (1u << 4) | (afmt << 7) | (bfmt << 10) | (a_mn << 15) | (b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// shared-memory matrix descriptor: start address, leading / stride byte offsets, layout (0 none, 2 SWIZZLE_128B, 4 SWIZZLE_64B)
__device__ __forceinline__ uint64_t ss_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ void ss_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n"
               ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];\n"
               ::"l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_prefetch_5d(const CUtensorMap* m, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];\n"
               ::"l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n"
               ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// Bounded wait by POLLING (mbarrier.test_wait never suspends the thread).  mbarrier.try_wait parks the thread and its
// wake-up was measured at about a microsecond per hand-over (tools/ss_bench.py): with two hand-overs per pipeline stage that,
// not memory latency, set the stage round trip.  The waiting warps here have nothing else to do.
__device__ __forceinline__ bool ss_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
#define mbar_wait ss_wait

// One elected lane of a converged warp (elect.sync): unlike `lane == 0` the compiler knows the region runs in exactly one
// thread and feeds tcgen05.mma's uniform-register operands without a vote / broadcast loop per instruction.
__device__ __forceinline__ bool ss_elect() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void ss_split_bf16(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xFFFF0000u));
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void ss_split_f16(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

#ifdef WF_SS_TRACE
__device__ long long wf_ss_trace_buf[6 * 256];   // [role event][stage index], CTA 0 (tools/ss_trace.py)
#define SS_TR(ev, i) do { if (blockIdx.x == 0 && (i) < 256) wf_ss_trace_buf[(ev) * 256 + (i)] = clock64(); } while (0)
#else
#define SS_TR(ev, i) do { } while (0)
#endif

struct SsTile { int g, mtg, z, mtw, zt, nt, blk, node0; bool side; };

// row tile `mt` of the flattened (group, tile-in-group) space
__device__ __forceinline__ SsTile ss_decode(const SsArgs& a, int mt) {
  SsTile t;
  t.g = mt / a.m_tiles_g;
  t.mtg = mt - t.g * a.m_tiles_g;
  t.z = 0; t.mtw = 0; t.zt = 0; t.nt = 0; t.blk = 0; t.node0 = 0; t.side = false;
  if (a.mode == SS_ROWS) {
    const int w = t.mtg / a.win_tiles;
    t.mtw = t.mtg - w * a.win_tiles;
    t.z = t.g * a.Bw + w;
    t.side = t.mtw < a.agg_tiles;
  } else {
    const int ztl = t.mtg / a.tpw;
    t.nt = t.mtg - ztl * a.tpw;
    t.node0 = t.nt * a.rpt;
    t.zt = t.g * a.Bw * a.T + ztl;
    t.blk = t.zt * a.tpw + t.nt;
  }
  return t;
}

// D[row tile][n-part] = A B^T with B resident.  tmA: main A source; tmA2: ROWS side buffer (aggregated leading rows);
// tmBhi / tmBlo: weights [G or 1][N_total][K] as two planes; tmOut: E_HL output [2 planes][Z][R][N_total] (fp16);
// tmOut2 (a.out2): the same values once more as bf16 planes (the weight gradients' operand format).
template <int BN, bool DROP>
__global__ void __launch_bounds__(SS_THREADS_HL, 1)
wf_ss_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
             const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
             const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOut2, const SsArgs a) {
  // A tcgen05.mma that accumulates into the columns the previous one wrote waits ~150 clk for it, whatever its size
  // (measured: 150-180 clk per instruction for N = 64 and N = 128 alike), so the products of a tile are dealt round-robin
  // over NC independent accumulators that the epilogue adds up: N = 64: 3 chains, N = 128: 2, N = 256 (128 clk of work
  // per instruction anyway): 1.  TMEM: 2 stages x NC x BN columns.
  constexpr int NC = 1;   // (independent accumulation chains were measured neutral: the ring, not the MMA pipe, binds)
  constexpr int DCOLS = NC * BN;                // columns of one accumulator stage
  constexpr int TCOLS = 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;                           // [plane][k-block][BN rows][64 B]
  uint8_t* sA = smem + SS_B_BYTES;              // ring of A stages: [plane][...]
  uint8_t* sStg = sA + a.nst * SS_A_STAGE;
  __shared__ uint64_t full[SS_MAX_STAGES], empty[SS_MAX_STAGES], dfull[2], dempty[2], bfull, bempty;
  __shared__ uint32_t tmem_base_s;
  // bias + bias2 of this CTA's columns for the current group, two copies (toggled per group change): every epilogue warp
  // fills the copy itself (identical values, no CTA-wide barrier that a timed-out warp could leave hanging), and a warp
  // is never more than two tiles ahead of the slowest one (TMEM double buffer), so the older copy is never rewritten
  // while still being read
  // (256-column parts only: they run two ring stages, the narrower ones have 160 bytes of shared memory left and read
  // the bias through __ldg)
  constexpr bool SBIAS = BN == 256;
  __shared__ __align__(16) float sbias[2][SBIAS ? BN : 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NST = a.nst;

  // this CTA: one n-part, a contiguous range of row tiles
  const int parts = a.n_parts * a.k_parts, pidx = blockIdx.x % parts;
  const int npart = pidx % a.n_parts, kpart = pidx / a.n_parts, slot = blockIdx.x / parts, slots = gridDim.x / parts;
  const int kb0 = kpart * a.nkb;
  const int total_mt = a.m_tiles_g * a.G;
  const int per = (total_mt + slots - 1) / slots;
  const int mt0 = slot * per, mt1 = min(total_mt, mt0 + per);
  const int n0 = npart * BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&dfull[s], 1); mbar_init(&dempty[s], a.epi == SS_E_HL ? 8 : 4); }
    mbar_init(&bfull, 1); mbar_init(&bempty, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmBhi); tma_prefetch_desc(&tmBlo);
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  const uint32_t b_plane = (uint32_t)a.nkb * BN * 128u;   // bytes of one resident B plane: [k-block][BN rows][128 B]

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int it = 0, bload = 0, gprev = -1;
      for (int mt = mt0; mt < mt1; ++mt) {
        const SsTile t = ss_decode(a, mt);
        const int gb = a.b_per_group ? t.g : 0;
        if (gb != gprev) {  // (re)load the resident weight slice of this group
          if (bload > 0 && !mbar_wait(&bempty, (bload - 1) & 1)) { atomicExch(a.err, 51); break; }
          mbar_expect_tx(&bfull, 2 * b_plane);
          for (int kb = 0; kb < a.nkb; ++kb) {
            tma_load_3d(sB + kb * BN * 128, &tmBhi, &bfull, (kb0 + kb) * SS_BK, n0, gb);
            tma_load_3d(sB + b_plane + kb * BN * 128, &tmBlo, &bfull, (kb0 + kb) * SS_BK, n0, gb);
          }
          gprev = gb; ++bload;
        }
        for (int kb = 0; kb < a.nkb; ++kb, ++it) {
          const int s = it % NST, ph = (it / NST) & 1;
          if (!mbar_wait(&empty[s], ph ^ 1)) { atomicExch(a.err, 52); mt = mt1; break; }
          SS_TR(0, it);   // producer: slot free
          uint8_t* st = sA + s * SS_A_STAGE;
          mbar_expect_tx(&full[s], SS_A_STAGE);
          // one load per stage: both planes of the [128 rows x 64 k] tile
          if (a.avar == SS_A_KT)            // TB8 block: [plane][block][K/8][128 rows][8]
            tma_load_5d(st, &tmA, &full[s], 0, 0, (kb0 + kb) * 8, t.blk, 0);
          else if (a.mode == SS_NODES)      // row-major [plane][(window, step)][node][K]
            tma_load_4d(st, &tmA, &full[s], (kb0 + kb) * SS_BK, t.node0, t.zt, 0);
          else                              // row-major [plane][window][row][K]; leading rows from the side buffer
            tma_load_4d(st, t.side ? &tmA2 : &tmA, &full[s], (kb0 + kb) * SS_BK, t.mtw * 128, t.z, 0);
          if (a.pf > 0) {   // the ring is three loads deep and a load from DRAM takes ~2,700 clk: warm L2 `pf` stages ahead
            const int fut = (mt - mt0) * a.nkb + kb + a.pf, fmt_ = mt0 + fut / a.nkb, fkb = kb0 + fut % a.nkb;
            if (fmt_ < mt1) {
              const SsTile f = ss_decode(a, fmt_);
              if (a.avar == SS_A_KT) tma_prefetch_5d(&tmA, 0, 0, fkb * 8, f.blk, 0);
              else if (a.mode == SS_NODES) tma_prefetch_4d(&tmA, fkb * SS_BK, f.node0, f.zt, 0);
              else tma_prefetch_4d(f.side ? &tmA2 : &tmA, fkb * SS_BK, f.mtw * 128, f.z, 0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = ss_idesc(BN, (uint32_t)a.afmt, (uint32_t)a.bfmt, 0, 0);
    int it = 0, lt = 0, bload = 0, gprev = -1;
    bool ok = true;
    for (int mt = mt0; mt < mt1 && ok; ++mt, ++lt) {
      const SsTile t = ss_decode(a, mt);
      const int gb = a.b_per_group ? t.g : 0;
      const int ds = lt & 1;
      if (!mbar_wait(&dempty[ds], ((lt >> 1) & 1) ^ 1)) { if (lane == 0) atomicExch(a.err, 53); ok = false; break; }
      if (gb != gprev) {
        if (!mbar_wait(&bfull, bload & 1)) { if (lane == 0) atomicExch(a.err, 54); ok = false; break; }
        gprev = gb; ++bload;
      }
      tc_fence_after();
      uint32_t issued = 0;   // instructions of this tile so far: chain = issued % NC, the first NC overwrite
      for (int kb = 0; kb < a.nkb; ++kb, ++it) {
        const int s = it % NST, ph = (it / NST) & 1;
        if (!mbar_wait(&full[s], ph)) { if (lane == 0) atomicExch(a.err, 55); ok = false; break; }
        tc_fence_after();
        const bool leader = ss_elect();
        if (leader) SS_TR(1, it);   // MMA warp: data landed
        if (leader && (a.debug & 1)) {
          umma_commit(&empty[s]);
          if (kb == a.nkb - 1) {
            umma_commit(&dfull[ds]);
            const int gnext = mt + 1 < mt1 ? (a.b_per_group ? (mt + 1) / a.m_tiles_g : 0) : -2;
            if (gnext != gb && gnext != -2) umma_commit(&bempty);
          }
        } else if (leader) {
          const uint32_t ahi = smem_u32(sA + s * SS_A_STAGE), bhi = smem_u32(sB + kb * BN * 128);
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16) {
#pragma unroll
            for (int p = 0; p < 3; ++p) {  // A_hi B_hi, A_lo B_hi, A_hi B_lo
              const uint32_t as = ahi + (p == 1 ? SS_A_PLANE : 0), bs = bhi + (p == 2 ? b_plane : 0);
              const uint64_t ad = a.avar == SS_A_KT ? ss_desc(as + k16 * 4096, 2048, 128, 0) : ss_desc(as + k16 * 32, 16, 1024, 2);
              ss_mma(tbase + ds * DCOLS + (issued % NC) * BN, ad, ss_desc(bs + k16 * 32, 16, 1024, 2), idesc, issued >= NC ? 1u : 0u);
              ++issued;
            }
          }
          umma_commit(&empty[s]);
          SS_TR(2, it);   // MMA warp: stage issued
          if (kb == a.nkb - 1) {
            umma_commit(&dfull[ds]);
            // last tile of this group in my range: the resident B may be replaced once these MMAs have completed
            const int gnext = mt + 1 < mt1 ? (a.b_per_group ? (mt + 1) / a.m_tiles_g : 0) : -2;
            if (gnext != gb && gnext != -2) umma_commit(&bempty);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp < 6 || a.epi == SS_E_HL) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
    const int chalf = (warp - 2) >> 2;         // hl16 epilogue: warps 2-5 take the first half of the columns, 6-9 the second
    uint8_t* stg = sStg + (warp - 2) * 4096;   // [plane][32 rows][64 B], 16-byte chunks XOR-swizzled by (row >> 1) & 3 (SWIZZLE_64B)
    int lt = 0, gbias = -1, bsel = 1;
    bool ok = true;
    const bool has_bias = (a.bias != nullptr || a.bias2 != nullptr) && kpart == 0;
    for (int mt = mt0; mt < mt1 && ok; ++mt, ++lt) {
      const SsTile t = ss_decode(a, mt);
      const int ds = lt & 1;
      const float* b1 = has_bias && a.bias ? a.bias + t.g * a.bias_gstride + n0 : nullptr;
      const float* b2 = has_bias && a.bias2 ? a.bias2 + t.g * a.bias_gstride + n0 : nullptr;
      if (SBIAS && has_bias && t.g != gbias) {
        // the summed bias row of this group goes to shared memory once (16 global loads per 8 stores in the column loop
        // cost the store-bound projections 15-20 %)
        bsel ^= 1;
        for (int c = lane; c < BN; c += 32)
          sbias[bsel][c] = (a.bias ? __ldg(a.bias + t.g * a.bias_gstride + n0 + c) : 0.f) +
                           (a.bias2 ? __ldg(a.bias2 + t.g * a.bias_gstride + n0 + c) : 0.f);
        __syncwarp();
        gbias = t.g;
      }
      if (!mbar_wait(&dfull[ds], (lt >> 1) & 1)) { if (lane == 0) atomicExch(a.err, 56); ok = false; break; }
      tc_fence_after();
      if (warp == 2 && lane == 0) SS_TR(4, lt);   // epilogue: accumulator complete
      DropState dst;
      if (DROP) dst = wf_drop_state(a.drop);
      if (a.epi == SS_E_TB4) {
        // TB4 block = [c_cols / 4 channel groups][128 rows][4 floats]: 512 contiguous bytes per warp store
        float4* cblk = reinterpret_cast<float4*>(a.C) + ((long long)t.blk * (a.c_cols >> 2) + (n0 >> 2)) * 128 + row;
        unsigned long long e4row = 0;
        if (DROP) e4row = (((unsigned long long)t.zt * a.Nn + (unsigned)(t.node0 + row)) * (unsigned)a.c_cols + (unsigned)n0) >> 2;
        // (the store / reduction choice is made OUTSIDE the column loop: a branch per store splits the loop body into basic
        // blocks and pins every bias load behind the previous store -- measured 2.5x on the projections)
        auto emit = [&](const uint32_t (&v)[32], int cc, auto atomic_tag) {
          constexpr bool ATOMIC = decltype(atomic_tag)::value;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            if (SBIAS) {
              if (has_bias) { const float4 b = *reinterpret_cast<const float4*>(&sbias[bsel][cc + j]); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
            } else {
              if (b1) { const float4 b = __ldg(reinterpret_cast<const float4*>(b1 + cc + j)); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
              if (b2) { const float4 b = __ldg(reinterpret_cast<const float4*>(b2 + cc + j)); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
            }
            if (DROP) {
              float m[4];
              wf_drop4(dst, e4row + (unsigned)((cc + j) >> 2), m);
              o.x *= m[0]; o.y *= m[1]; o.z *= m[2]; o.w *= m[3];
            }
            if (row < a.rpt && !(a.debug & 2)) {  // rows >= rpt of a node tile are padding
              float4* dstp = cblk + (long long)((cc + j) >> 2) * 128;
              if constexpr (ATOMIC)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dstp), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w));   // no "memory" clobber: it would pin the bias loads of the next columns behind it
              else
                *dstp = o;
            }
          }
        };
        // (keeping the next chunk's TMEM load in flight under these stores was measured neutral without bias pointers and
        // 2x slower with them: the plain load - wait - store sequence stays)
        auto drain = [&](auto atomic_tag) {
#pragma unroll 1
          for (int cc = 0; cc < BN; cc += 32) {
            uint32_t v[32];
            __syncwarp();
            tmem_ld32(tlane + ds * DCOLS + cc, v);
            tmem_wait_ld();
            emit(v, cc, atomic_tag);
          }
        };
        if (a.k_parts > 1) drain(std::true_type{});
        else drain(std::false_type{});
      } else {
        // row-major hl16 through shared memory + TMA stores: per 64 columns and plane a [32 rows][128 B] box per warp;
        // the stores clip at the window's R rows (3-D map), so partial tiles need no masking
        const int grow = t.mtw * 128 + row;
        const bool valid = grow < a.R;
        unsigned long long e4row = 0;
        if (DROP) e4row = ((((unsigned long long)t.z * (unsigned)a.R) + (unsigned)grow) * (unsigned)a.c_cols + (unsigned)n0) >> 2;
        float amax = 0.f;
        const int npass = a.out2 ? 2 : 1;   // pass 1: the same values again as bf16 planes
        const int cbase = chalf * (BN / 2);  // this warp's columns: [cbase, cbase + BN / 2), 32 at a time
#pragma unroll 1
        for (int c32p = 0; c32p < (BN / 64) * npass; ++c32p) {
          const int cc = cbase + (c32p / npass) * 32, pass = c32p % npass;
          if (lane == 0) tma_store_wait_read();   // the previous boxes have been read out of the staging buffer
          __syncwarp();
          uint32_t v[32];
          tmem_ld32(tlane + ds * DCOLS + cc, v);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 8; e += 4) {
              float4 o = make_float4(__uint_as_float(v[j + e]), __uint_as_float(v[j + e + 1]), __uint_as_float(v[j + e + 2]),
                                     __uint_as_float(v[j + e + 3]));
              if (SBIAS) {
                if (has_bias) { const float4 b = *reinterpret_cast<const float4*>(&sbias[bsel][cc + j + e]); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
              } else if (b1) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(b1 + cc + j + e)); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
              }
              if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              if (DROP) {
                float m[4];
                wf_drop4(dst, e4row + (unsigned)((cc + j + e) >> 2), m);
                o.x *= m[0]; o.y *= m[1]; o.z *= m[2]; o.w *= m[3];
              }
              if (valid) amax = fmaxf(amax, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
              if (pass == 0) {
                ss_split_f16(o.x, o.y, hi[e >> 1], lo[e >> 1]);
                ss_split_f16(o.z, o.w, hi[(e >> 1) + 1], lo[(e >> 1) + 1]);
              } else {
                ss_split_bf16(o.x, o.y, hi[e >> 1], lo[e >> 1]);
                ss_split_bf16(o.z, o.w, hi[(e >> 1) + 1], lo[(e >> 1) + 1]);
              }
            }
            const int ch = j >> 3;   // 16-byte chunk of this row's 64-byte line
            const uint32_t off = (uint32_t)(lane * 64 + ((ch ^ ((lane >> 1) & 3)) << 4));
            sts128(smem_u32(stg + off), make_uint4(hi[0], hi[1], hi[2], hi[3]));
            sts128(smem_u32(stg + 2048 + off), make_uint4(lo[0], lo[1], lo[2], lo[3]));
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            const CUtensorMap* om = pass == 0 ? &tmOut : &tmOut2;
            tma_store_4d(om, stg, n0 + cc, t.mtw * 128 + q * 32, t.z, 0);
            tma_store_4d(om, stg + 2048, n0 + cc, t.mtw * 128 + q * 32, t.z, 1);
            tma_store_commit();
          }
        }
        // an activation the next layer cannot hold as fp16 hi/lo (|x| >= 65520 rounds to inf), or a non-finite one
        if (a.range_limit > 0.f && valid && !(amax < a.range_limit)) atomicExch(a.err, 41);
      }
      tc_fence_before();
      __syncwarp();
      if (warp == 2 && lane == 0) SS_TR(5, lt);   // epilogue: tile drained
      if (lane == 0) mbar_arrive(&dempty[ds]);
    }
    if (a.epi == SS_E_HL && lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, TCOLS);
}


// ================================================================================= weight gradients
// dW[g][512 gate rows][nh * 128] = sum over the task's (window, step, node tile) blocks of dG^T [X | H_prev]
// (loss.backward() through nn.LSTM: dW_ih = dG^T x, dW_hh = sum_{t >= 1} dG[t]^T h[t-1]; train_hybrid_maml_v5.py:134,169).
// Both operands are read MN-major straight from the activations' own layouts -- the contraction runs over ROWS:
//   A   = dG, TB8 bf16 hi/lo: block [64 channel groups][128 rows][8]  -> [plane][16 groups of this m tile][64 rows][16 B]
//         per stage, ONE load
//   B_h = half h of the N range, either TB8 fp16 (h of the layer below at the same step, or this layer's h one step
//         EARLIER: `shift`; step 0 has no such term and is skipped) or row-major fp16 features (SWIZZLE_128B boxes).
// Rows >= rpt of a block are padding: never written by any kernel, zero since allocation, so they add nothing.
// The bias gradients (row sums of dG^T) come out of the same pass as a 16-column product with a tile of ones.
// Split-K over blocks; partial tiles go to `part`, summed in fixed order by wf_wg_reduce_kernel (deterministic).
constexpr int WG_THREADS = 192;
struct WgArgs {
  int G, Bw, T, tpw, rpt, Nn;
  int splits, bps;              // blocks per split
  int nh;                       // B halves of 128 columns
  int bvar[2];                  // 0: TB8 (MN-major, no swizzle), 1: row-major hl16 (MN-major, SWIZZLE_128B)
  int bshift[2];                // 1: the block of the previous step (t - 1)
  int bcol0[2];                 // TB8: first channel group of the half; row-major: first column
  int afmt, bfmt;
  int nst;
  float* part; float* bias_part;   // [splits][G][M][nh * 128] (tiles in TB4 order), [splits][G][M]
  int mt;                          // M / 128 row tiles of dG's channels (4: the LSTM's 4L = 512 gate rows)
  int sumw;                        // pair kernel: bias gradient by the summing warps (1) or by 16-column instructions (0)
  int* err;
};

__global__ void __launch_bounds__(WG_THREADS, 1)
wf_wg_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
             const __grid_constant__ CUtensorMap tmB1, const WgArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int STAGE = SS_A_STAGE * (1 + a.nh);   // A (hi, lo) then each B half (hi, lo)
  uint8_t* ones = smem + a.nst * STAGE;        // [2 column groups][16 rows][16 B] of 1.0 in B's format
  __shared__ uint64_t full[SS_MAX_STAGES], empty[SS_MAX_STAGES], dfull, dempty;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NST = a.nst;
  const int blocks_g = a.Bw * a.T * a.tpw;
  const int total = a.splits * a.G * a.mt;
  const int kb_n = (a.rpt + 63) / 64;          // 64-row k-blocks that hold data
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&dfull, 1); mbar_init(&dempty, 4);
    mbar_fence_init();
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB0); tma_prefetch_desc(&tmB1);
  }
  if (threadIdx.x < 128) reinterpret_cast<uint32_t*>(ones)[threadIdx.x] = a.bfmt == 0 ? 0x3C003C00u : 0x3F803F80u;
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int mtile = tile % a.mt, g = (tile / a.mt) % a.G, split = (tile / a.mt) / a.G;
        const int b0 = split * a.bps, b1 = min(blocks_g, b0 + a.bps);
        for (int b = b0; b < b1; ++b) {
          const int t = (b / a.tpw) % a.T, nt = b % a.tpw;
          const int blk = g * blocks_g + b;
          for (int kb = 0; kb < kb_n; ++kb, ++it) {
            const int s = it % NST, ph = (it / NST) & 1;
            if (!mbar_wait(&empty[s], ph ^ 1)) { atomicExch(a.err, 61); tile = total; b = b1; break; }
            SS_TR(0, it);
            uint8_t* st = smem + s * STAGE;
            int bytes = SS_A_STAGE;
            for (int h = 0; h < a.nh; ++h) if (!(a.bshift[h] && t == 0)) bytes += SS_A_STAGE;
            mbar_expect_tx(&full[s], bytes);
            tma_load_5d(st, &tmA, &full[s], 0, 2 * kb, mtile * 16, blk, 0);   // rows [64 kb, +64), both planes
            // B region behind A.  TB8 halves: [hi half 0][hi half 1][lo half 0][lo half 1] (16 KB each), so that the hi
            // (lo) planes of both halves form ONE 256-column operand; row-major source: per 64 columns a box
            // [plane][64 rows][128 B], 16 KB apart
            uint8_t* sbase = st + SS_A_STAGE;
            for (int h = 0; h < a.nh; ++h) {
              if (a.bshift[h] && t == 0) continue;
              const CUtensorMap* m = h == 0 ? &tmB0 : &tmB1;
              if (a.bvar[h] == 0) {
                const int bb = blk - a.bshift[h] * a.tpw;
                tma_load_5d(sbase + h * SS_A_PLANE, m, &full[s], 0, 2 * kb, a.bcol0[h], bb, 0);
                tma_load_5d(sbase + (a.nh + h) * SS_A_PLANE, m, &full[s], 0, 2 * kb, a.bcol0[h], bb, 1);
              } else {
                const int zt = blk / a.tpw, node = nt * a.rpt + kb * 64;
                for (int j = 0; j < 2; ++j)
                  tma_load_4d(sbase + (2 * h + j) * SS_A_PLANE, m, &full[s], a.bcol0[h] + 64 * j, node, zt, 0);
              }
            }
            SS_TR(3, it);
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = ss_idesc(128, (uint32_t)a.afmt, (uint32_t)a.bfmt, 1, 1);
    const uint32_t idesc1 = ss_idesc(16, (uint32_t)a.afmt, (uint32_t)a.bfmt, 1, 1);
    const uint32_t idesc2 = ss_idesc(256, (uint32_t)a.afmt, (uint32_t)a.bfmt, 1, 1);
    const uint64_t odesc = ss_desc(smem_u32(ones), 128, 256, 0);
    int it = 0, lt = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < total && ok; tile += gridDim.x, ++lt) {
      const int split = (tile / a.mt) / a.G;
      const int b0 = split * a.bps, b1 = min(blocks_g, b0 + a.bps);
      if (!mbar_wait(&dempty, (lt & 1) ^ 1)) { if (lane == 0) atomicExch(a.err, 62); ok = false; break; }
      tc_fence_after();
      uint32_t acc[2] = {0u, 0u}, acc1 = 0u;
      for (int b = b0; b < b1 && ok; ++b) {
        const int t = (b / a.tpw) % a.T;
        for (int kb = 0; kb < kb_n; ++kb, ++it) {
          const int s = it % NST, ph = (it / NST) & 1;
          if (!mbar_wait(&full[s], ph)) { if (lane == 0) atomicExch(a.err, 63); ok = false; break; }
          tc_fence_after();
          if (ss_elect()) {
            SS_TR(1, it);
            const uint32_t as = smem_u32(smem + s * STAGE), bs = as + SS_A_STAGE;
            const int nk16 = min(4, (a.rpt - kb * 64 + 15) / 16);
            // both halves in ONE 256-column instruction per product when both are present (5 instructions per K step with
            // the bias products instead of 8): an MN-major operand costs ~145 clk per instruction whatever its width
            const bool both = a.nh == 2 && !((a.bshift[0] || a.bshift[1]) && t == 0);
            const int hsel = (a.nh == 2 && a.bshift[0] && t == 0) ? 1 : 0;   // the half that is present when only one is
            const bool any = both || !(a.bshift[hsel] && t == 0);
            for (int k16 = 0; k16 < nk16; ++k16) {
              // MN-major, no swizzle: [16 channel groups][64 rows][16 B] -> groups 1024 B apart (SBO), 8-row groups 128 B (LBO)
              const uint64_t ahi = ss_desc(as + k16 * 256, 128, 1024, 0), alo = ss_desc(as + SS_A_PLANE + k16 * 256, 128, 1024, 0);
              uint64_t bhi, blo;
              if (a.bvar[0] == 0) {   // TB8: hi planes of the halves back to back, then the lo planes
                bhi = ss_desc(bs + (both ? 0 : hsel) * SS_A_PLANE + k16 * 256, 128, 1024, 0);
                blo = ss_desc(bs + (a.nh + (both ? 0 : hsel)) * SS_A_PLANE + k16 * 256, 128, 1024, 0);
              } else {                // row-major source: 64-column boxes [plane][64 rows][128 B] 16 KB apart, SWIZZLE_128B
                bhi = ss_desc(bs + (both ? 0 : 2 * hsel) * SS_A_PLANE + k16 * 2048, 16384, 1024, 2);
                blo = ss_desc(bs + (both ? 0 : 2 * hsel) * SS_A_PLANE + 8192 + k16 * 2048, 16384, 1024, 2);
              }
              if (any) {
                const uint32_t d = tbase + (both ? 0 : hsel * 128), id = both ? idesc2 : idesc;
                if (both && acc[0] == acc[1]) {
                  ss_mma(d, ahi, bhi, id, acc[0]);
                  acc[0] = acc[1] = 1u;
                } else if (both) {
                  // one half has products already, the other (skipped at step 0) must be overwritten: two instructions, once
                  const uint64_t bh1 = a.bvar[0] == 0 ? ss_desc(bs + SS_A_PLANE + k16 * 256, 128, 1024, 0)
                                                      : ss_desc(bs + 2 * SS_A_PLANE + k16 * 2048, 16384, 1024, 2);
                  ss_mma(d, ahi, bhi, idesc, acc[0]);
                  ss_mma(d + 128, ahi, bh1, idesc, acc[1]);
                  acc[0] = acc[1] = 1u;
                } else {
                  ss_mma(d, ahi, bhi, id, acc[hsel]);
                  acc[hsel] = 1u;
                }
                ss_mma(d, alo, bhi, id, 1u);
                ss_mma(d, ahi, blo, id, 1u);
              }
              if (a.bias_part != nullptr) {   // row sums of dG^T: the bias gradients
                ss_mma(tbase + 256, ahi, odesc, idesc1, acc1);
                ss_mma(tbase + 256, alo, odesc, idesc1, 1u);
                acc1 = 1u;
              }
            }
            SS_TR(2, it);
            umma_commit(&empty[s]);
          }
          __syncwarp();
        }
      }
      if (ss_elect() && ok) umma_commit(&dfull);
      __syncwarp();
    }
  } else {
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
    const int ncol = a.nh * 128;
    int lt = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++lt) {
      const int mtile = tile % a.mt, g = (tile / a.mt) % a.G, split = (tile / a.mt) / a.G;
      if (!mbar_wait(&dfull, lt & 1)) { if (lane == 0) atomicExch(a.err, 64); break; }
      tc_fence_after();
      // partial tile = [ncol / 4 column groups][128 rows][4 floats]: a warp's float4 store covers 512 contiguous bytes
      float4* cblk = reinterpret_cast<float4*>(a.part) + ((((long long)split * a.G + g) * a.mt + mtile) * (ncol >> 2)) * 128 + row;
      // a half whose every block of this split was skipped (step 0 of a shifted operand) never touched its accumulator
      bool live[2] = {false, false};
      {
        const int b0 = split * a.bps, b1 = min(blocks_g, b0 + a.bps);
        for (int h = 0; h < a.nh; ++h)
          for (int b = b0; b < b1 && !live[h]; ++b) live[h] = !(a.bshift[h] && (b / a.tpw) % a.T == 0);
      }
      auto emit = [&](uint32_t (&v)[32], int cc) {
        if (!live[cc >> 7]) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          cblk[(long long)((cc + j) >> 2) * 128] = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      };
#pragma unroll 1
      for (int cc = 0; cc < ncol; cc += 32) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld32(tlane + cc, v);
        tmem_wait_ld();
        emit(v, cc);
      }
      if (a.bias_part != nullptr) {
        uint32_t v[8];
        __syncwarp();
        tmem_ld8(tlane + 256, v);
        tmem_wait_ld();
        a.bias_part[((long long)split * a.G + g) * (a.mt * 128) + mtile * 128 + row] = __uint_as_float(v[0]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dempty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, 512);
}

// ---- the same product on CTA PAIRS (tcgen05 cta_group::2, M = 256): the two CTAs of a cluster own two neighbouring
// 128-row tiles of dG's channels and HALF of the B columns each -- CTA r loads its own A tile (32 KB per stage) and columns
// [r N/2, (r + 1) N/2) of [X | H_prev] (32 KB), and every instruction, issued by the leader CTA alone, reads A from each
// CTA's own shared memory and the two B halves from both.  The one-CTA kernel ingests 96 KB per stage and SM (the four M
// tiles re-read the activation columns: it is bound by the ~40 B/clk an SM pulls from L2); a pair ingests 64 KB per SM for
// the same work, and three 64 KB stages fit where two 96 KB ones did.
//   full[s]  (leader's): expect_tx of both CTAs' bytes; both producers' TMA loads complete on it (.cta_group::2 loads may
//            signal an mbarrier of the peer CTA)
//   empty[s], dfull: tcgen05.commit multicast to both CTAs
//   dempty   (leader's): 8 arrivals, the peer's epilogue warps arrive through the cluster window
// A B half that has no term at step 0 (the shifted H operand) is loaded from a block coordinate past the tensor: the TMA
// unit fills zeros, so every instruction covers both halves.
__device__ __forceinline__ void ss_mma2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit2(uint64_t* bar) {   // arrives on `bar` of BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tma2_load_5d(void* dst, const CUtensorMap* m, uint32_t cbar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n"
               ::"r"(smem_u32(dst)), "l"(m), "r"(cbar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* m, uint32_t cbar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n"
               ::"r"(smem_u32(dst)), "l"(m), "r"(cbar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ uint32_t ss_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t ss_mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void ss_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void ss_arrive_remote(uint32_t cbar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cbar) : "memory");
}
// no data of the arriving thread is published: no fence (a cluster-scope release in the MMA-issuing thread stalls it for
// ~1.5k clk per stage)
__device__ __forceinline__ void ss_arrive_remote_relaxed(uint32_t cbar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cbar) : "memory");
}
// bounded polling wait with a cluster-scope acquire (arrivals come from the peer CTA)
__device__ __forceinline__ bool ss_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}

// The bias gradient (column sums of dG over the rows) costs the tensor pipe as much as the products when it is a
// 16-column instruction pair per K step: an instruction with an MN-major no-swizzle operand takes ~150 clk whatever its
// width, 8 of the 20 per stage.  Four extra warps (6-9) add the dG tile up straight from shared memory instead (16
// LDS.128 + 128 converts / adds per thread and stage, fp32 accumulators, one shuffle reduction per tile); they hold the
// stage (empty[s] counts them) and learn that it has landed from the leader's MMA warp (landed[s], also across the pair).
constexpr int WG2_STAGE = 65536;   // [A hi 16K][A lo 16K][B: this CTA's N/2 columns, 32K]
constexpr int WG2_NST = 3;
constexpr int WG2_THREADS = 320;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WG2_THREADS, 1)
wf_wg2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
              const __grid_constant__ CUtensorMap tmB1, const WgArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ones = smem + WG2_NST * WG2_STAGE;   // [16 rows][8 columns] of 1.0: this CTA's half of the 16-column bias operand
  __shared__ uint64_t full[WG2_NST], empty[WG2_NST], landed[WG2_NST], dfull[2], dempty[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ss_cluster_rank();
  const bool sum_warps = a.bias_part != nullptr && a.afmt == 1 && a.sumw != 0;   // bias gradient by warps 6-9 (bf16 dG), else by the tensor core
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int blocks_g = a.Bw * a.T * a.tpw;
  const int mp = a.mt >> 1;                      // M tile pairs
  const int total = a.splits * a.G * mp;
  const int kb_n = (a.rpt + 63) / 64;
  const int ncta = a.nh * 64;                    // B columns held by one CTA
  const uint32_t b_bytes = (uint32_t)ncta * 64u * 2u * 2u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < WG2_NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], sum_warps ? 5 : 1); mbar_init(&landed[s], 1); }
    for (int d = 0; d < 2; ++d) { mbar_init(&dfull[d], 1); mbar_init(&dempty[d], 8); }
    mbar_fence_init();
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB0); tma_prefetch_desc(&tmB1);
  }
  // with the bias gradient off the tensor core the accumulator is 256 columns: two of them, so that the epilogue of a
  // tile (128 KB of partials per CTA, ~10k clk) runs under the next tile's instructions
  const int nbuf = sum_warps ? 2 : 1;
  if (threadIdx.x < 64) reinterpret_cast<uint32_t*>(ones)[threadIdx.x] = a.bfmt == 0 ? 0x3C003C00u : 0x3F803F80u;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  ss_cluster_sync();   // both CTAs' barriers exist before anybody signals across
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int tile = pair; tile < total; tile += npairs) {
        const int mtile = 2 * (tile % mp) + (int)rank, g = (tile / mp) % a.G, split = (tile / mp) / a.G;
        const int b0 = split * a.bps, b1 = min(blocks_g, b0 + a.bps);
        for (int b = b0; b < b1; ++b) {
          const int t = (b / a.tpw) % a.T, nt = b % a.tpw;
          const int blk = g * blocks_g + b;
          for (int kb = 0; kb < kb_n; ++kb, ++it) {
            const int s = it % WG2_NST, ph = (it / WG2_NST) & 1;
            if (!mbar_wait(&empty[s], ph ^ 1)) { atomicExch(a.err, 71); tile = total; b = b1; break; }
            SS_TR(0, it);
            uint8_t* st = smem + s * WG2_STAGE;
            const uint32_t fbar = ss_mapa(smem_u32(&full[s]), 0);   // the leader's barrier collects both CTAs' bytes
            if (rank == 0) mbar_expect_tx(&full[s], 2u * ((uint32_t)SS_A_STAGE + b_bytes));
            tma2_load_5d(st, &tmA, fbar, 0, 2 * kb, mtile * 16, blk, 0);
            uint8_t* sb = st + SS_A_STAGE;
            // my columns of [half 0 | half 1]: with two halves CTA r holds half r, with one it holds 64 columns of it
            const int h = a.nh == 2 ? (int)rank : 0;
            const CUtensorMap* m = h == 0 ? &tmB0 : &tmB1;
            const bool absent = a.bshift[h] && t == 0;   // no such term: zeros from past the end of the tensor
            if (a.bvar[h] == 0) {
              const int bb = absent ? a.G * blocks_g : blk - a.bshift[h] * a.tpw;
              const int grp = a.bcol0[h] + (a.nh == 2 ? 0 : 8 * (int)rank);
              tma2_load_5d(sb, m, fbar, 0, 2 * kb, grp, bb, 0);
              tma2_load_5d(sb + SS_A_PLANE, m, fbar, 0, 2 * kb, grp, bb, 1);
            } else {
              const int zt = absent ? a.G * a.Bw * a.T : blk / a.tpw, node = nt * a.rpt + kb * 64;
              const int col = a.bcol0[h] + (a.nh == 2 ? 0 : 64 * (int)rank);
              for (int j = 0; j < ncta / 64; ++j) tma2_load_4d(sb + j * SS_A_PLANE, m, fbar, col + 64 * j, node, zt, 0);
            }
            SS_TR(3, it);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      const uint32_t idesc = ss_idesc(a.nh * 128, (uint32_t)a.afmt, (uint32_t)a.bfmt, 1, 1, 256);
      const uint32_t idesc1 = ss_idesc(16, (uint32_t)a.afmt, (uint32_t)a.bfmt, 1, 1, 256);
      const uint64_t odesc = ss_desc(smem_u32(ones), 128, 256, 0);
      int it = 0, lt = 0;
      bool ok = true;
      for (int tile = pair; tile < total && ok; tile += npairs, ++lt) {
        const int split = (tile / mp) / a.G;
        const int b0 = split * a.bps, b1 = min(blocks_g, b0 + a.bps);
        const int ds = lt % nbuf;
        const uint32_t dcol = tbase + (uint32_t)ds * 256u;
        if (!ss_wait_cluster(&dempty[ds], ((lt / nbuf) & 1) ^ 1)) { if (lane == 0) atomicExch(a.err, 72); ok = false; break; }
        tc_fence_after();
        uint32_t acc = 0u;
        for (int b = b0; b < b1 && ok; ++b) {
          for (int kb = 0; kb < kb_n; ++kb, ++it) {
            const int s = it % WG2_NST, ph = (it / WG2_NST) & 1;
            if (!mbar_wait(&full[s], ph)) { if (lane == 0) atomicExch(a.err, 73); ok = false; break; }
            tc_fence_after();
            if (ss_elect()) {
              SS_TR(1, it);
              if (sum_warps) {   // the stage is complete in BOTH CTAs: tell the summing warps of each
                // (the bytes were put into the peer's shared memory by its own TMA loads, complete before the transaction
                // count reached this CTA's full[s]; this thread publishes nothing of its own)
                mbar_arrive(&landed[s]);
                ss_arrive_remote_relaxed(ss_mapa(smem_u32(&landed[s]), 1));
              }
              const uint32_t as = smem_u32(smem + s * WG2_STAGE), bs = as + SS_A_STAGE;
              const int nk16 = min(4, (a.rpt - kb * 64 + 15) / 16);
              for (int k16 = 0; k16 < nk16; ++k16) {
                const uint64_t ahi = ss_desc(as + k16 * 256, 128, 1024, 0), alo = ss_desc(as + SS_A_PLANE + k16 * 256, 128, 1024, 0);
                uint64_t bhi, blo;
                if (a.bvar[0] == 0) {   // TB8: [groups of 8 columns][64 rows][16 B], hi plane then lo plane 16 KB apart
                  bhi = ss_desc(bs + k16 * 256, 128, 1024, 0);
                  blo = ss_desc(bs + SS_A_PLANE + k16 * 256, 128, 1024, 0);
                } else {                // row-major source: 64-column boxes [plane][64 rows][128 B], SWIZZLE_128B
                  bhi = ss_desc(bs + k16 * 2048, 16384, 1024, 2);
                  blo = ss_desc(bs + 8192 + k16 * 2048, 16384, 1024, 2);
                }
                ss_mma2(dcol, ahi, bhi, idesc, acc);
                ss_mma2(dcol, alo, bhi, idesc, 1u);
                ss_mma2(dcol, ahi, blo, idesc, 1u);
                if (a.bias_part != nullptr && !sum_warps) {
                  ss_mma2(tbase + 256, ahi, odesc, idesc1, acc);
                  ss_mma2(tbase + 256, alo, odesc, idesc1, 1u);
                }
                acc = 1u;
              }
              umma_commit2(&empty[s]);
              SS_TR(2, it);
            }
            __syncwarp();
          }
        }
        if (ss_elect() && ok) umma_commit2(&dfull[ds]);
        __syncwarp();
      }
    }
  } else if (warp >= 6) {
    if (sum_warps) {
      // column sums of my dG tile: thread (grp, rs) owns the 8 channels of group grp and rows rs, rs + 8, ... of a stage
      const int tid2 = threadIdx.x - 192, grp = tid2 >> 3, rs = tid2 & 7;
      int it = 0;
      bool ok = true;
      for (int tile = pair; tile < total && ok; tile += npairs) {
        const int mtile = 2 * (tile % mp) + (int)rank, g = (tile / mp) % a.G, split = (tile / mp) / a.G;
        const int b0 = split * a.bps, b1 = min(blocks_g, b0 + a.bps);
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int b = b0; b < b1 && ok; ++b) {
          for (int kb = 0; kb < kb_n; ++kb, ++it) {
            const int s = it % WG2_NST, ph = (it / WG2_NST) & 1;
            if (!ss_wait_cluster(&landed[s], ph)) { if (lane == 0) atomicExch(a.err, 75); ok = false; break; }
            const uint32_t as = smem_u32(smem + s * WG2_STAGE) + (uint32_t)grp * 1024u + (uint32_t)rs * 16u;
#pragma unroll
            for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                uint4 w;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w)
                             : "r"(as + (uint32_t)pl * (uint32_t)SS_A_PLANE + (uint32_t)i * 128u));
                acc[0] += __uint_as_float(w.x << 16); acc[1] += __uint_as_float(w.x & 0xFFFF0000u);
                acc[2] += __uint_as_float(w.y << 16); acc[3] += __uint_as_float(w.y & 0xFFFF0000u);
                acc[4] += __uint_as_float(w.z << 16); acc[5] += __uint_as_float(w.z & 0xFFFF0000u);
                acc[6] += __uint_as_float(w.w << 16); acc[7] += __uint_as_float(w.w & 0xFFFF0000u);
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);   // my reads of the stage are done
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
          acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 2);
          acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 4);
        }
        if (rs == 0 && ok) {
          float* dstb = a.bias_part + ((long long)split * a.G + g) * (a.mt * 128) + mtile * 128 + grp * 8;
          *reinterpret_cast<float4*>(dstb) = make_float4(acc[0], acc[1], acc[2], acc[3]);
          *reinterpret_cast<float4*>(dstb + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
      }
    }
  } else {
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
    const int ncol = a.nh * 128;
    int lt = 0;
    for (int tile = pair; tile < total; tile += npairs, ++lt) {
      const int mtile = 2 * (tile % mp) + (int)rank, g = (tile / mp) % a.G, split = (tile / mp) / a.G;
      const int ds = lt % nbuf;
      const uint32_t tl = tlane + (uint32_t)ds * 256u;
      if (!mbar_wait(&dfull[ds], (lt / nbuf) & 1)) { if (lane == 0) atomicExch(a.err, 74); break; }
      tc_fence_after();
      float4* cblk = reinterpret_cast<float4*>(a.part) + ((((long long)split * a.G + g) * a.mt + mtile) * (ncol >> 2)) * 128 + row;
#pragma unroll 1
      for (int cc = 0; cc < ncol; cc += 32) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld32(tl + cc, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          cblk[(long long)((cc + j) >> 2) * 128] = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
      if (a.bias_part != nullptr && !sum_warps) {
        uint32_t v[8];
        __syncwarp();
        tmem_ld8(tlane + 256, v);
        tmem_wait_ld();
        a.bias_part[((long long)split * a.G + g) * (a.mt * 128) + mtile * 128 + row] = __uint_as_float(v[0]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&dempty[ds]);
        else ss_arrive_remote(ss_mapa(smem_u32(&dempty[ds]), 0));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  ss_cluster_sync();   // nobody leaves (or frees tensor memory) while the pair's instructions may still touch this CTA
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tbase), "r"(512u) : "memory");
}

// ---- the resident-weight GEMM on CTA pairs (NODES row tiles, TB4 epilogue): the pair owns two consecutive row tiles
// (M = 256) and BN2 output columns; each CTA keeps HALF of the weight slice resident (BN2 / 2 rows x K, hi + lo <= 128 KB),
// streams its OWN row tile through a 3-stage ring and drains its own 128 x BN2 accumulator.  Against the one-CTA kernel:
// twice the columns per resident byte, so A is read half as often (dX: once instead of twice; K = 256 projection: twice
// instead of four times), and every instruction does four times the work of a 128 x 64 one for the same ~150 clk.
template <int BN2, bool DROP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SS_THREADS, 1)
wf_ss2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
              const __grid_constant__ CUtensorMap tmBlo, const SsArgs a) {
  constexpr int BNH = BN2 / 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;                           // [plane][k-block][BNH rows][128 B]
  uint8_t* sA = smem + SS_B_BYTES;              // ring of A stages: [plane][128 rows x 64 k]
  __shared__ uint64_t full[3], empty[3], dfull[2], dempty[2], bfull, bempty;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NST = 3;
  const uint32_t rank = ss_cluster_rank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int pidx = pair % a.n_parts, slot = pair / a.n_parts, slots = npairs / a.n_parts;
  const int total_pt = (a.m_tiles_g * a.G) >> 1;            // pair tiles: row tiles (2 pt, 2 pt + 1), same group
  const int per = (total_pt + slots - 1) / slots;
  const int pt0 = slot * per, pt1 = min(total_pt, pt0 + per);
  const int n0 = pidx * BN2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&dfull[s], 1); mbar_init(&dempty[s], 8); }
    mbar_init(&bfull, 1); mbar_init(&bempty, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmBhi); tma_prefetch_desc(&tmBlo);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  ss_cluster_sync();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  const uint32_t b_plane = (uint32_t)a.nkb * BNH * 128u;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0, bload = 0, gprev = -1;
      const uint32_t bfull0 = ss_mapa(smem_u32(&bfull), 0);
      for (int pt = pt0; pt < pt1; ++pt) {
        const SsTile t = ss_decode(a, 2 * pt + (int)rank);
        const int gb = a.b_per_group ? t.g : 0;
        if (gb != gprev) {  // (re)load my half of the resident weight slice of this group
          if (bload > 0 && !mbar_wait(&bempty, (bload - 1) & 1)) { atomicExch(a.err, 81); break; }
          if (rank == 0) mbar_expect_tx(&bfull, 4u * b_plane);   // both CTAs' halves, both planes
          for (int kb = 0; kb < a.nkb; ++kb) {
            asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                         ::"r"(smem_u32(sB + kb * BNH * 128)), "l"(&tmBhi), "r"(bfull0), "r"(kb * SS_BK), "r"(n0 + (int)rank * BNH), "r"(gb) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                         ::"r"(smem_u32(sB + b_plane + kb * BNH * 128)), "l"(&tmBlo), "r"(bfull0), "r"(kb * SS_BK), "r"(n0 + (int)rank * BNH), "r"(gb) : "memory");
          }
          gprev = gb; ++bload;
        }
        for (int kb = 0; kb < a.nkb; ++kb, ++it) {
          const int s = it % NST, ph = (it / NST) & 1;
          if (!mbar_wait(&empty[s], ph ^ 1)) { atomicExch(a.err, 82); pt = pt1; break; }
          SS_TR(0, it);
          uint8_t* st = sA + s * SS_A_STAGE;
          const uint32_t fbar = ss_mapa(smem_u32(&full[s]), 0);
          if (rank == 0) mbar_expect_tx(&full[s], 2u * (uint32_t)SS_A_STAGE);
          if (a.avar == SS_A_KT) tma2_load_5d(st, &tmA, fbar, 0, 0, kb * 8, t.blk, 0);
          else tma2_load_4d(st, &tmA, fbar, kb * SS_BK, t.node0, t.zt, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      const uint32_t idesc = ss_idesc(BN2, (uint32_t)a.afmt, (uint32_t)a.bfmt, 0, 0, 256);
      int it = 0, lt = 0, bload = 0, gprev = -1;
      bool ok = true;
      for (int pt = pt0; pt < pt1 && ok; ++pt, ++lt) {
        const int g = (2 * pt) / a.m_tiles_g;
        const int gb = a.b_per_group ? g : 0;
        const int ds = lt & 1;
        if (!ss_wait_cluster(&dempty[ds], ((lt >> 1) & 1) ^ 1)) { if (lane == 0) atomicExch(a.err, 83); ok = false; break; }
        if (gb != gprev) {
          if (!mbar_wait(&bfull, bload & 1)) { if (lane == 0) atomicExch(a.err, 84); ok = false; break; }
          gprev = gb; ++bload;
        }
        tc_fence_after();
        uint32_t acc = 0u;
        for (int kb = 0; kb < a.nkb; ++kb, ++it) {
          const int s = it % NST, ph = (it / NST) & 1;
          if (!mbar_wait(&full[s], ph)) { if (lane == 0) atomicExch(a.err, 85); ok = false; break; }
          tc_fence_after();
          if (ss_elect()) {
            SS_TR(1, it);
            const uint32_t ahi = smem_u32(sA + s * SS_A_STAGE), bhi = smem_u32(sB + kb * BNH * 128);
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16) {
#pragma unroll
              for (int p = 0; p < 3; ++p) {  // A_hi B_hi, A_lo B_hi, A_hi B_lo
                const uint32_t as = ahi + (p == 1 ? SS_A_PLANE : 0), bs = bhi + (p == 2 ? b_plane : 0);
                const uint64_t ad = a.avar == SS_A_KT ? ss_desc(as + k16 * 4096, 2048, 128, 0) : ss_desc(as + k16 * 32, 16, 1024, 2);
                ss_mma2(tbase + ds * BN2, ad, ss_desc(bs + k16 * 32, 16, 1024, 2), idesc, acc);
                acc = 1u;
              }
            }
            umma_commit2(&empty[s]);
            SS_TR(2, it);
            if (kb == a.nkb - 1) {
              umma_commit2(&dfull[ds]);
              const int gnext = pt + 1 < pt1 ? (a.b_per_group ? (2 * (pt + 1)) / a.m_tiles_g : 0) : -2;
              if (gnext != gb && gnext != -2) umma_commit2(&bempty);
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
    int lt = 0;
    bool ok = true;
    const bool has_bias = a.bias != nullptr || a.bias2 != nullptr;
    for (int pt = pt0; pt < pt1 && ok; ++pt, ++lt) {
      const SsTile t = ss_decode(a, 2 * pt + (int)rank);
      const int ds = lt & 1;
      const float* b1 = has_bias && a.bias ? a.bias + t.g * a.bias_gstride + n0 : nullptr;
      const float* b2 = has_bias && a.bias2 ? a.bias2 + t.g * a.bias_gstride + n0 : nullptr;
      if (!mbar_wait(&dfull[ds], (lt >> 1) & 1)) { if (lane == 0) atomicExch(a.err, 86); ok = false; break; }
      tc_fence_after();
      if (warp == 2 && lane == 0) SS_TR(4, lt);
      DropState dst;
      if (DROP) dst = wf_drop_state(a.drop);
      float4* cblk = reinterpret_cast<float4*>(a.C) + ((long long)t.blk * (a.c_cols >> 2) + (n0 >> 2)) * 128 + row;
      unsigned long long e4row = 0;
      if (DROP) e4row = (((unsigned long long)t.zt * a.Nn + (unsigned)(t.node0 + row)) * (unsigned)a.c_cols + (unsigned)n0) >> 2;
#pragma unroll 1
      for (int cc = 0; cc < BN2; cc += 32) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld32(tlane + ds * BN2 + cc, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          if (b1) { const float4 b = __ldg(reinterpret_cast<const float4*>(b1 + cc + j)); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
          if (b2) { const float4 b = __ldg(reinterpret_cast<const float4*>(b2 + cc + j)); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
          if (DROP) {
            float m[4];
            wf_drop4(dst, e4row + (unsigned)((cc + j) >> 2), m);
            o.x *= m[0]; o.y *= m[1]; o.z *= m[2]; o.w *= m[3];
          }
          if (row < a.rpt) cblk[(long long)((cc + j) >> 2) * 128] = o;   // rows >= rpt of a node tile are padding
        }
      }
      tc_fence_before();
      __syncwarp();
      if (warp == 2 && lane == 0) SS_TR(5, lt);
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&dempty[ds]);
        else ss_arrive_remote(ss_mapa(smem_u32(&dempty[ds]), 0));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  ss_cluster_sync();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tbase), "r"(512u) : "memory");
}

// dst_h[g][m][c] = sum_s part[s][g][m][h*128 + c] (fixed order); bias -> both LSTM bias gradients
__global__ void wf_wg_reduce_kernel(const float* __restrict__ part, const float* __restrict__ bias_part, int splits, int G, int nh, int M,
                                    float* dst0, int ld0, int w0, float* dst1, int ld1, int w1, float* db1, float* db2,
                                    long long gstride) {
  // columns [0, w0) of a tile row go to dst0 (row pitch ld0), columns [w0, w0 + w1) to dst1 (row pitch ld1)
  const int ncol = nh * 128, g = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // over M * ncol / 4 float4 (rows fastest) + M bias rows
  const int quads = M * ncol / 4, mt = M >> 7;
  if (i < quads) {
    const int m = i % M, c = (i / M) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {   // partial tiles: [split][g][m tile][ncol / 4][128 rows][4]
      const float4 v = reinterpret_cast<const float4*>(part)[((((long long)s * G + g) * mt + (m >> 7)) * (ncol >> 2) + (c >> 2)) * 128 + (m & 127)];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (c < w0) *reinterpret_cast<float4*>(dst0 + g * gstride + (long long)m * ld0 + c) = acc;
    else if (dst1 != nullptr && c - w0 < w1) *reinterpret_cast<float4*>(dst1 + g * gstride + (long long)m * ld1 + (c - w0)) = acc;
  } else if (i < quads + M && bias_part != nullptr && db1 != nullptr) {
    const int m = i - quads;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += bias_part[((long long)s * G + g) * M + m];
    db1[g * gstride + m] = acc;
    if (db2) db2[g * gstride + m] = acc;
  }
}

}  // namespace

// ================================================================================= host side
namespace {

int ss_dtype(int fmt) { return fmt == 0 ? 1 : 2; }   // wf_encode_tensor_map: 1 = f16, 2 = bf16

// row-major hl16 [2 planes][Z][rows][C]: K-major operand boxes {64 k, box_rows, both planes} (SWIZZLE_128B)
int map_rows_k(CUtensorMap* m, const void* base, uint64_t C, uint64_t rows, uint64_t Z, uint64_t plane, uint32_t box_rows, int fmt) {
  uint64_t dims[4] = {C, rows, Z, 2};
  uint64_t str[3] = {C * 2, rows * C * 2, plane * 2};
  uint32_t box[4] = {SS_BK, box_rows, 1, 2};
  return wf_encode_tensor_map(m, base, 4, dims, str, box, 1, ss_dtype(fmt));
}
// the same tensor as an MN-major operand (rows = K): boxes {64 columns, box_rows, box_planes} (SWIZZLE_128B);
// {64, 32, 1, 1} is also the GCN epilogue's store box
int map_rows_mn(CUtensorMap* m, const void* base, uint64_t C, uint64_t rows, uint64_t Z, uint64_t plane, int fmt,
                uint32_t box_rows = 32, uint32_t box_planes = 1) {
  uint64_t dims[4] = {C, rows, Z, 2};
  uint64_t str[3] = {C * 2, rows * C * 2, plane * 2};
  uint32_t box[4] = {64, box_rows, 1, box_planes};
  return wf_encode_tensor_map(m, base, 4, dims, str, box, 1, ss_dtype(fmt));
}
// the hl16 epilogue's store box: {32 columns, 32 rows} of one plane, SWIZZLE_64B
int map_rows_store(CUtensorMap* m, const void* base, uint64_t C, uint64_t rows, uint64_t Z, uint64_t plane, int fmt) {
  uint64_t dims[4] = {C, rows, Z, 2};
  uint64_t str[3] = {C * 2, rows * C * 2, plane * 2};
  uint32_t box[4] = {32, 32, 1, 1};
  return wf_encode_tensor_map(m, base, 4, dims, str, box, 2, ss_dtype(fmt));
}
// TB8 [2 planes][blocks][C/8][128 rows][8]: the 128 x 8 elements of a channel group are folded as {256, 4}.
// fold = 4, groups = 8: a K-major [128 rows][64 k] box; fold = 2, groups = 16: an MN-major [128 channels][64 rows] box;
// both planes in one box
int map_tb8(CUtensorMap* m, const void* base, uint64_t C, uint64_t blocks, uint64_t plane, uint32_t fold, uint32_t groups, int fmt,
            uint32_t box_planes = 2) {
  uint64_t dims[5] = {256, 4, C / 8, blocks, 2};
  uint64_t str[4] = {512, 2048, C * 256, plane * 2};
  uint32_t box[5] = {256, fold, groups, 1, box_planes};
  return wf_encode_tensor_map(m, base, 5, dims, str, box, 0, ss_dtype(fmt));
}
// weights [G][N][K] (one plane): resident-B boxes {64 k, bn rows} (SWIZZLE_128B)
int map_w(CUtensorMap* m, const void* base, uint64_t K, uint64_t N, uint64_t G, uint64_t ld, uint64_t gstride, uint32_t bn, int fmt) {
  uint64_t dims[3] = {K, N, G};
  uint64_t str[2] = {ld * 2, gstride * 2};
  uint32_t box[3] = {SS_BK, bn, 1};
  return wf_encode_tensor_map(m, base, 3, dims, str, box, 1, ss_dtype(fmt));
}

int ss_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

template <int BN>
int ss_launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmBhi, const CUtensorMap& tmBlo,
                 const CUtensorMap& tmOut, const CUtensorMap& tmOut2, SsArgs& a, cudaStream_t st) {
  // ring depth: 3 stages of 32 KB next to the resident weights, 2 where the epilogue needs the staging buffer -- and 2 for
  // K <= 128 (two k-blocks per tile): those GEMMs are bound by their stores, and a third load in flight costs 13 %
  // (h -> gates projection: 107.8 us with 3 stages, 93.7 us with 2; K >= 256: 141 vs 170 us the other way round)
  a.nst = (a.epi == SS_E_HL || a.nkb <= 2) ? 2 : 3;
  static const int dbg = getenv("WF_SS_DEBUG") ? atoi(getenv("WF_SS_DEBUG")) : 0;
  a.debug = dbg;
  static const int pf_env = getenv("WF_SS_PF") ? atoi(getenv("WF_SS_PF")) : 0;
  a.pf = pf_env;
  static const int nst_env = getenv("WF_SS_NST") ? atoi(getenv("WF_SS_NST")) : 0;
  if (nst_env > 0 && nst_env < a.nst) a.nst = nst_env;
  const int smem = SS_B_BYTES + a.nst * SS_A_STAGE + (a.epi == SS_E_HL ? SS_STG_BYTES : 0) + 1024;
  static bool configured = false;
  // 256-column parts always run with two stages (K <= 128): their static shared memory (bias copies) is 2 KB
  const int mx = SS_B_BYTES + (BN == 256 ? 2 : 3) * SS_A_STAGE + 1024;
  WF_REQUIRE(smem <= mx, "ss kernel: %d bytes of shared memory for %d-column parts (limit %d)", smem, BN, mx);
  if (!configured) {
    if (cudaFuncSetAttribute(wf_ss_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx) != cudaSuccess ||
        cudaFuncSetAttribute(wf_ss_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx) != cudaSuccess)
      return wf_fail(WF_ECUDA, "ss kernel: cannot raise dynamic shared memory to %d", mx);
    configured = true;
  }
  const int total_mt = a.m_tiles_g * a.G;
  if (a.k_parts < 1) a.k_parts = 1;
  int slots = ss_sms() / (a.n_parts * a.k_parts);
  if (slots > total_mt) slots = total_mt;
  if (slots < 1) slots = 1;
  const int grid = slots * a.n_parts * a.k_parts;
  const int threads = a.epi == SS_E_HL ? SS_THREADS_HL : SS_THREADS;
  if (a.drop.rng != nullptr) wf_ss_kernel<BN, true><<<grid, threads, smem, st>>>(tmA, tmA2, tmBhi, tmBlo, tmOut, tmOut2, a);
  else wf_ss_kernel<BN, false><<<grid, threads, smem, st>>>(tmA, tmA2, tmBhi, tmBlo, tmOut, tmOut2, a);
  WF_CHECK_LAUNCH("ss_kernel");
  return WF_OK;
}

template <int BN2>
int ss2_launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmBhi, const CUtensorMap& tmBlo, SsArgs& a, cudaStream_t st) {
  const int smem = SS_B_BYTES + 3 * SS_A_STAGE + 1024;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(wf_ss2_kernel<BN2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
        cudaFuncSetAttribute(wf_ss2_kernel<BN2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return wf_fail(WF_ECUDA, "ss pair kernel: cannot raise dynamic shared memory to %d", smem);
    configured = true;
  }
  const int total_pt = a.m_tiles_g * a.G / 2;
  int slots = (ss_sms() / 2) / a.n_parts;
  if (slots > total_pt) slots = total_pt;
  if (slots < 1) slots = 1;
  const int grid = 2 * slots * a.n_parts;
  if (a.drop.rng != nullptr) wf_ss2_kernel<BN2, true><<<grid, SS_THREADS, smem, st>>>(tmA, tmBhi, tmBlo, a);
  else wf_ss2_kernel<BN2, false><<<grid, SS_THREADS, smem, st>>>(tmA, tmBhi, tmBlo, a);
  WF_CHECK_LAUNCH("ss2_kernel");
  return WF_OK;
}

int ss_launch(int bn, const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmBhi, const CUtensorMap& tmBlo,
              const CUtensorMap& tmOut, const CUtensorMap& tmOut2, SsArgs& a, cudaStream_t st) {
  WF_REQUIRE((long long)bn * a.nkb * SS_BK * 4 <= SS_B_BYTES, "ss gemm: the weight slice %d x %d does not fit shared memory", bn, a.nkb * SS_BK);
  if (bn == 64) return ss_launch_bn<64>(tmA, tmA2, tmBhi, tmBlo, tmOut, tmOut2, a, st);
  if (bn == 128) return ss_launch_bn<128>(tmA, tmA2, tmBhi, tmBlo, tmOut, tmOut2, a, st);
  if (bn == 256) return ss_launch_bn<256>(tmA, tmA2, tmBhi, tmBlo, tmOut, tmOut2, a, st);
  return wf_fail(WF_EINVAL, "ss gemm: unsupported n-part width %d", bn);
}

}  // namespace

// C (TB4 fp32, c_cols = Ntot channels per block) = A W^T (+ bias + bias2) per (window, step, node tile).
// avar 0: A row-major hl16 [2][G*Bw*T][Nn][K]; 1: A TB8 hl16 [2][blocks][K/8][128][8].  W hi/lo: [G][Ntot][K] (ldb, gstride).
int wf_ss_launch_nodes(int bn, int avar, const void* A16, long long a_plane, int K, int afmt, const void* Bhi, const void* Blo,
                       int ldb, long long b_gstride, int Ntot, int bfmt, const float* bias, const float* bias2,
                       long long bias_gstride, float* C, int T, int Nn, int Bw, int G, const DropCfg* drop, int* err,
                       cudaStream_t st, int k_parts) {
  WF_REQUIRE(k_parts >= 1 && K % (SS_BK * k_parts) == 0 || k_parts == 1, "ss_nodes: K=%d does not split into %d parts of 64-wide blocks", K, k_parts);
  WF_REQUIRE(K % 8 == 0 && (avar == SS_A_KS || K % SS_BK == 0) && Ntot % bn == 0 && ldb % 8 == 0,
             "ss_nodes: K=%d must be a multiple of 8 (64 for TB8 operands), N=%d of %d", K, Ntot, bn);
  WF_REQUIRE(((uintptr_t)A16 | (uintptr_t)Bhi | (uintptr_t)Blo | (uintptr_t)C) % 16 == 0, "ss_nodes: pointers must be 16-byte aligned");
  const int tpw = wf_cdiv(Nn, 128);
  const long long ZT = (long long)G * Bw * T;
  CUtensorMap tmA, tmBhi, tmBlo;
  int rc;
  if (avar == SS_A_KT) rc = map_tb8(&tmA, A16, K, ZT * tpw, a_plane, 4, 8, afmt);
  else rc = map_rows_k(&tmA, A16, K, Nn, ZT, a_plane, 128, afmt);
  if (rc) return rc;
  if ((rc = map_w(&tmBhi, Bhi, K, Ntot, G, ldb, G > 1 ? b_gstride : (long long)Ntot * ldb, bn, bfmt))) return rc;
  if ((rc = map_w(&tmBlo, Blo, K, Ntot, G, ldb, G > 1 ? b_gstride : (long long)Ntot * ldb, bn, bfmt))) return rc;
  SsArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = SS_NODES; a.avar = avar; a.epi = SS_E_TB4; a.n_parts = Ntot / bn; a.m_tiles_g = Bw * T * tpw; a.G = G;
  a.nkb = (K + SS_BK - 1) / SS_BK / k_parts; a.k_parts = k_parts; a.b_per_group = G > 1 ? 1 : 0; a.afmt = afmt; a.bfmt = bfmt;
  a.T = T; a.Nn = Nn; a.Bw = Bw; a.tpw = tpw; a.rpt = wf_tile_rows(Nn);
  if (k_parts > 1 &&   // the parts add into C
      cudaMemsetAsync(C, 0, (size_t)ZT * tpw * Ntot * 128 * sizeof(float), st) != cudaSuccess)
    return wf_fail(WF_ECUDA, "ss_nodes: clearing the output failed");
  a.C = C; a.c_cols = Ntot; a.bias = bias; a.bias2 = bias2; a.bias_gstride = bias_gstride; a.err = err;
  if (drop != nullptr) a.drop = *drop;
  // CTA pairs (cta_group::2): K >= 256 (the store-bound K = 128 projection gains nothing), an even number of row tiles per
  // group, and a column count the pair's resident halves can cover
  static const int use_pairs = getenv("WF_SS_PAIRS") ? atoi(getenv("WF_SS_PAIRS")) : 1;
  const int bn2 = K >= 512 ? 128 : 256;
  if (use_pairs && k_parts == 1 && K >= 256 && K % SS_BK == 0 && (a.m_tiles_g % 2) == 0 && Ntot % bn2 == 0 &&
      (long long)(bn2 / 2) * K * 4 <= SS_B_BYTES) {
    CUtensorMap tmBhi2, tmBlo2;
    if ((rc = map_w(&tmBhi2, Bhi, K, Ntot, G, ldb, G > 1 ? b_gstride : (long long)Ntot * ldb, bn2 / 2, bfmt))) return rc;
    if ((rc = map_w(&tmBlo2, Blo, K, Ntot, G, ldb, G > 1 ? b_gstride : (long long)Ntot * ldb, bn2 / 2, bfmt))) return rc;
    a.n_parts = Ntot / bn2;
    return bn2 == 128 ? ss2_launch_bn<128>(tmA, tmBhi2, tmBlo2, a, st) : ss2_launch_bn<256>(tmA, tmBhi2, tmBlo2, a, st);
  }
  return ss_launch(bn, tmA, tmA, tmBhi, tmBlo, tmA, tmA, a, st);
}

// Weight gradients of one LSTM layer from dG (TB8 bf16) and up to two 128-column operand halves (see wf_wg_kernel).
// bsrc_h: base of the half's source; bvar 0: TB8 with bC channels per block, 1: row-major [2][G*Bw*T][Nn][bC].
// Columns [0, w0) of the result go to dst0 (row pitch ld0), [w0, w0 + w1) to dst1; row sums to db1 / db2 (optional).
int wf_ss_launch_wgrad(const void* dg16, long long dg_plane, int nh, const void* const* bsrc, const long long* bplane,
                       const int* bvar, const int* bshift, const int* bcol0, const int* bC, int T, int Nn, int Bw, int G,
                       float* part, size_t part_floats, float* dst0, int ld0, int w0, float* dst1, int ld1, int w1, float* db1,
                       float* db2, long long gstride, int* err, cudaStream_t st, int M) {
  WF_REQUIRE(nh == 1 || nh == 2, "ss_wgrad: one or two operand halves");
  WF_REQUIRE(M >= 128 && M % 128 == 0, "ss_wgrad: M=%d must be a multiple of 128", M);
  const int tpw = wf_cdiv(Nn, 128), blocks_g = Bw * T * tpw;
  const long long blocks = (long long)G * blocks_g, ZT = (long long)G * Bw * T;
  CUtensorMap tmA, tmB[2];
  int rc;
  if ((rc = map_tb8(&tmA, dg16, M, blocks, dg_plane, 2, 16, 1))) return rc;
  for (int h = 0; h < nh; ++h) {
    if (bvar[h] == 0) rc = map_tb8(&tmB[h], bsrc[h], bC[h], blocks, bplane[h], 2, 16, 1, 1);   // one plane per load
    else rc = map_rows_mn(&tmB[h], bsrc[h], bC[h], Nn, ZT, bplane[h], 1, 64, 2);
    if (rc) return rc;
  }
  if (nh == 1) tmB[1] = tmB[0];
  // split-K over blocks: the split count that minimises (rounds of tiles over the SMs) x (blocks per tile)
  const int ncol = nh * 128, sms = ss_sms();
  int best = 1;
  long long best_cost = -1;
  for (int s = 1; s <= 40 && s <= blocks_g; ++s) {
    if ((size_t)s * G * M * (ncol + 1) > part_floats) break;
    const long long cost = (long long)wf_cdiv((long long)(M / 128) * G * s, sms) * wf_cdiv(blocks_g, s) * 64 + s;  // + s: prefer fewer partials on ties
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
  }
  WF_REQUIRE((size_t)best * G * M * (ncol + 1) <= part_floats, "ss_wgrad: partial buffer too small");
  WgArgs a;
  memset(&a, 0, sizeof(a));
  a.G = G; a.Bw = Bw; a.T = T; a.tpw = tpw; a.rpt = wf_tile_rows(Nn); a.Nn = Nn;
  a.bps = wf_cdiv(blocks_g, best); a.splits = wf_cdiv(blocks_g, a.bps); a.nh = nh;
  for (int h = 0; h < nh; ++h) { a.bvar[h] = bvar[h]; a.bshift[h] = bshift[h]; a.bcol0[h] = bvar[h] == 0 ? bcol0[h] / 8 : bcol0[h]; }
  a.afmt = 1; a.bfmt = 1; a.nst = nh == 2 ? 2 : 3;   // kind::f16 takes ONE format for both operands: bf16 (dG's range)
  a.mt = M / 128;
  a.part = part; a.bias_part = db1 != nullptr ? part + (size_t)a.splits * G * M * ncol : nullptr; a.err = err;
  static const int use_pairs = getenv("WF_WG_PAIRS") ? atoi(getenv("WF_WG_PAIRS")) : 1;
  if (use_pairs && a.mt % 2 == 0 && (nh == 1 || bvar[0] == bvar[1])) {
    // CTA pairs (cta_group::2): the split count is chosen for pair tiles over sms / 2 pairs
    const int pairs = sms / 2, mp = a.mt / 2;
    int bestp = 1;
    long long bestc = -1;
    for (int s2 = 1; s2 <= 40 && s2 <= blocks_g; ++s2) {
      if ((size_t)s2 * G * M * (ncol + 1) > part_floats) break;
      // rounds of pair tiles x (blocks per tile + half a block of hand-over), fewer partial tiles on ties
      const long long cost = (long long)wf_cdiv((long long)mp * G * s2, pairs) * (wf_cdiv(blocks_g, s2) * 64 + 32) + s2;
      if (bestc < 0 || cost < bestc) { bestc = cost; bestp = s2; }
    }
    a.bps = wf_cdiv(blocks_g, bestp); a.splits = wf_cdiv(blocks_g, a.bps);
    a.bias_part = db1 != nullptr ? part + (size_t)a.splits * G * M * ncol : nullptr;
    CUtensorMap tmBp[2] = {tmB[0], tmB[1]};
    if (nh == 1 && bvar[0] == 0 && (rc = map_tb8(&tmBp[0], bsrc[0], bC[0], blocks, bplane[0], 2, 8, 1, 1))) return rc;  // 64-column boxes
    if (nh == 1) tmBp[1] = tmBp[0];
    const int smem2 = WG2_NST * WG2_STAGE + 1024 + 1024;
    static bool configured2 = false;
    if (!configured2) {
      if (cudaFuncSetAttribute(wf_wg2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2) != cudaSuccess)
        return wf_fail(WF_ECUDA, "wg pair kernel: cannot raise dynamic shared memory");
      configured2 = true;
    }
    static const int sumw = getenv("WF_WG_SUMWARPS") ? atoi(getenv("WF_WG_SUMWARPS")) : 1;
    a.sumw = sumw;
    const int total2 = a.splits * G * mp;
    wf_wg2_kernel<<<2 * (total2 < pairs ? total2 : pairs), WG2_THREADS, smem2, st>>>(tmA, tmBp[0], tmBp[1], a);
    WF_CHECK_LAUNCH("wg2_kernel");
    const int items2 = M * ncol / 4 + M;
    wf_wg_reduce_kernel<<<dim3(wf_cdiv(items2, 256), G), 256, 0, st>>>(part, a.bias_part, a.splits, G, nh, M, dst0, ld0, w0, dst1, ld1, w1,
                                                                      db1, db2, gstride);
    WF_CHECK_LAUNCH("wg_reduce");
    return WF_OK;
  }
  const int smem = a.nst * SS_A_STAGE * (1 + nh) + 1024 + 1024;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(wf_wg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * SS_A_STAGE * 3 + 2048) != cudaSuccess)
      return wf_fail(WF_ECUDA, "wg kernel: cannot raise dynamic shared memory");
    configured = true;
  }
  const int total = a.splits * G * a.mt;
  wf_wg_kernel<<<total < sms ? total : sms, WG_THREADS, smem, st>>>(tmA, tmB[0], tmB[1], a);
  WF_CHECK_LAUNCH("wg_kernel");
  const int items = M * ncol / 4 + M;
  wf_wg_reduce_kernel<<<dim3(wf_cdiv(items, 256), G), 256, 0, st>>>(part, a.bias_part, a.splits, G, nh, M, dst0, ld0, w0, dst1, ld1, w1,
                                                                   db1, db2, gstride);
  WF_CHECK_LAUNCH("wg_reduce");
  return WF_OK;
}

// ================================================================================= GCN layer on pre-split operands
namespace {

// fp32 windows [Z][R][C] (contiguous, or window z at element x_win_off[z] of a resident features tensor) -> fp16 hi / lo
// planes [2][Z][R][C].  One thread per 4 channels.
__global__ void wf_ss_split_windows_kernel(const float* __restrict__ X, const long long* __restrict__ x_win_off, int C, int R,
                                           uint16_t* __restrict__ out, long long plane) {
  const int z = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // float4 index inside the window
  const long long n4 = (long long)R * C / 4;
  if (i >= n4) return;
  const float4 v = *reinterpret_cast<const float4*>(X + (x_win_off ? x_win_off[z] : (long long)z * R * C) + 4 * i);
  uint2 hi, lo;
  ss_split_f16(v.x, v.y, hi.x, lo.x);
  ss_split_f16(v.z, v.w, hi.y, lo.y);
  const long long o = (long long)z * R * C + 4 * i;
  *reinterpret_cast<uint2*>(out + o) = hi;
  *reinterpret_cast<uint2*>(out + plane + o) = lo;
}

// Side buffer S[z][rr][:] = sum_p val[p] * X[z][col[p]][:] for the leading agg_rows rows of every window (all rows whose
// aggregation is not the unit self loop live there: the t = 0 slice for the reference's graphs).  Rows that
// only have their self loop come out as copies.  X as fp32 windows (first layer) or fp16 hi/lo planes; S is fp16 hi/lo
// [2][Z][agg_rows][C].  One warp per (row, window), 8 channels per lane and trip.
__global__ void __launch_bounds__(256) wf_ss_agg_rows_kernel(const float* __restrict__ X32, const long long* __restrict__ x_win_off,
                                                             const uint16_t* __restrict__ X16, long long x_plane, int C, int R,
                                                             int Bw, int agg_rows, const int* __restrict__ rowptr,
                                                             const int* __restrict__ col, const float* __restrict__ val,
                                                             long long g_rowptr, long long g_csr, uint16_t* __restrict__ S,
                                                             long long s_plane) {
  const int z = blockIdx.y, g = z / Bw, rr = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (rr >= agg_rows || rr >= R) return;
  const int* rp = rowptr + g * g_rowptr;
  const int p0 = rp[rr], p1 = rp[rr + 1];
  const long long xbase = X32 ? (x_win_off ? x_win_off[z] : (long long)z * R * C) : (long long)z * R * C;
  for (int c8 = lane; c8 < (C >> 3); c8 += 32) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int p = p0; p < p1; ++p) {
      const float v = __ldg(val + g * g_csr + p);
      const long long src = xbase + (long long)__ldg(col + g * g_csr + p) * C + 8 * c8;
      float x[8];
      if (X32) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(X32 + src)), b = __ldg(reinterpret_cast<const float4*>(X32 + src + 4));
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
      } else {
        const uint4 h = __ldg(reinterpret_cast<const uint4*>(X16 + src)), l = __ldg(reinterpret_cast<const uint4*>(X16 + x_plane + src));
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&hw[j]));
          const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&lw[j]));
          x[2 * j] = fh.x + fl.x; x[2 * j + 1] = fh.y + fl.y;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, x[j], acc[j]);
    }
    uint4 hi, lo;
    ss_split_f16(acc[0], acc[1], hi.x, lo.x); ss_split_f16(acc[2], acc[3], hi.y, lo.y);
    ss_split_f16(acc[4], acc[5], hi.z, lo.z); ss_split_f16(acc[6], acc[7], hi.w, lo.w);
    const long long o = ((long long)z * agg_rows + rr) * C + 8 * c8;
    *reinterpret_cast<uint4*>(S + o) = hi;
    *reinterpret_cast<uint4*>(S + s_plane + o) = lo;
  }
}

}  // namespace

// GCNConv + ReLU (+ train-mode dropout) on pre-split fp16 hi/lo activations (model.py:31-42, hybrid_model.py:65-75):
// Y16 = split(dropout(relu((A_hat X) W^T + b))), Y16 / X16 as planes [2][G*Bw][R][C]; Yb16 (optional): the same values
// as bf16 hi / lo planes (the LSTM's layer-0 weight gradient contracts them against bf16 dG).  The first layer passes the fp32
// windows instead (X32 + x_win_off, or dense when x_win_off == NULL) and a scratch `xsplit16` [2][G*Bw][R][Cin] that
// receives their split.  Rows whose aggregation is not the unit self loop must all lie in the leading agg_rows rows of a
// window (a multiple of 128, or R rounded up); they are aggregated into `side16` [2][G*Bw][agg_rows][Cin] first and the
// GEMM reads those row tiles from there.  Cin % 8 == 0, Cout % 128 == 0, W16 hi/lo = wf_split16(W, 0) shared by all groups.
extern "C" int wf_gcn_layer_fwd_ss(const float* X32, const long long* x_win_off, const void* X16, void* xsplit16,
                                   const void* W16_hi, const void* W16_lo, const float* bias, const int* rowptr,
                                   const int* col, const float* val, long long rowptr_group_stride, long long csr_group_stride,
                                   int agg_rows, void* side16, int R, int Cin, int Cout, int G, int Bw, int relu, void* Y16,
                                   void* Yb16, float p_drop, const unsigned long long* rng, int site, int* err, void* stream) {
  WF_REQUIRE(G > 0 && Bw > 0 && R > 0, "gcn_layer_fwd_ss: bad batch");
  WF_REQUIRE(Cin % 8 == 0 && Cout % 128 == 0, "gcn_layer_fwd_ss: Cin=%d must be a multiple of 8, Cout=%d of 128", Cin, Cout);
  WF_REQUIRE((X32 != nullptr) != (X16 != nullptr), "gcn_layer_fwd_ss: pass the input either as fp32 windows or as fp16 planes");
  WF_REQUIRE(X32 == nullptr || xsplit16 != nullptr, "gcn_layer_fwd_ss: fp32 input needs the split scratch");
  WF_REQUIRE(agg_rows >= 0 && agg_rows % 128 == 0 && (agg_rows == 0 || (rowptr != nullptr && side16 != nullptr)),
             "gcn_layer_fwd_ss: agg_rows=%d must be a multiple of 128 and comes with the CSR and the side buffer", agg_rows);
  WF_REQUIRE(p_drop >= 0.f && p_drop < 1.f && (p_drop == 0.f || rng != nullptr), "gcn_layer_fwd_ss: bad dropout arguments");
  WF_REQUIRE((long long)128 * ((Cin + 63) / 64 * 64) * 4 <= SS_B_BYTES, "gcn_layer_fwd_ss: Cin=%d too wide for a resident weight slice", Cin);
  cudaStream_t st = (cudaStream_t)stream;
  const long long Z = (long long)G * Bw, plane = Z * R * Cin;
  const uint16_t* A16 = (const uint16_t*)X16;
  if (X32 != nullptr) {
    wf_ss_split_windows_kernel<<<dim3(wf_cdiv((long long)R * Cin / 4, 256), (unsigned)Z), 256, 0, st>>>(X32, x_win_off, Cin, R,
                                                                                                      (uint16_t*)xsplit16, plane);
    WF_CHECK_LAUNCH("ss_split_windows");
    A16 = (const uint16_t*)xsplit16;
  }
  const int win_tiles = wf_cdiv(R, 128);
  if (agg_rows > win_tiles * 128) agg_rows = win_tiles * 128;
  const long long s_plane = Z * agg_rows * Cin;
  if (agg_rows > 0) {
    wf_ss_agg_rows_kernel<<<dim3(wf_cdiv(agg_rows, 8), (unsigned)Z), 256, 0, st>>>(X32, x_win_off, X32 ? nullptr : A16, plane, Cin, R, Bw,
                                                                                  agg_rows, rowptr, col, val, rowptr_group_stride,
                                                                                  csr_group_stride, (uint16_t*)side16, s_plane);
    WF_CHECK_LAUNCH("ss_agg_rows");
  }
  CUtensorMap tmA, tmA2, tmBhi, tmBlo, tmOut, tmOut2;
  int rc;
  if ((rc = map_rows_k(&tmA, A16, Cin, R, Z, plane, 128, 0))) return rc;
  if (agg_rows > 0) { if ((rc = map_rows_k(&tmA2, side16, Cin, agg_rows, Z, s_plane, 128, 0))) return rc; }
  else tmA2 = tmA;
  if ((rc = map_w(&tmBhi, W16_hi, Cin, Cout, 1, Cin, (long long)Cout * Cin, 128, 0))) return rc;
  if ((rc = map_w(&tmBlo, W16_lo, Cin, Cout, 1, Cin, (long long)Cout * Cin, 128, 0))) return rc;
  if ((rc = map_rows_store(&tmOut, Y16, Cout, R, Z, Z * R * Cout, 0))) return rc;
  if (Yb16 != nullptr) { if ((rc = map_rows_store(&tmOut2, Yb16, Cout, R, Z, Z * R * Cout, 1))) return rc; }
  else tmOut2 = tmOut;
  SsArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = SS_ROWS; a.avar = SS_A_KS; a.epi = SS_E_HL; a.n_parts = Cout / 128; a.m_tiles_g = Bw * win_tiles; a.G = G;
  a.nkb = (Cin + SS_BK - 1) / SS_BK; a.b_per_group = 0; a.afmt = 0; a.bfmt = 0;
  a.R = R; a.Bw = Bw; a.win_tiles = win_tiles; a.agg_tiles = agg_rows / 128;
  a.c_cols = Cout; a.bias = bias; a.relu = relu; a.err = err;
  a.drop = wf_drop_cfg(p_drop, rng, WF_SITE_GCN + site);
  a.range_limit = 32768.0f;
  a.out2 = Yb16 != nullptr ? 1 : 0;
  return ss_launch(128, tmA, tmA2, tmBhi, tmBlo, tmOut, tmOut2, a, st);
}

// hi + lo planes -> fp32 (tests, and the drop-in module API, whose tensors are fp32): out[i] = float(hi[i]) + float(lo[i])
static __global__ void wf_ss_join_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo, long long n, int fmt,
                                         float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a, b;
  if (fmt == 0) { a = __half2float(__ushort_as_half(hi[i])); b = __half2float(__ushort_as_half(lo[i])); }
  else { a = __uint_as_float((uint32_t)hi[i] << 16); b = __uint_as_float((uint32_t)lo[i] << 16); }
  out[i] = a + b;
}
extern "C" int wf_join16(const void* hi, const void* lo, long long n, int fmt, float* out, void* stream) {
  WF_REQUIRE(n > 0 && (fmt == 0 || fmt == 1), "join16: bad arguments");
  wf_ss_join_kernel<<<wf_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>((const uint16_t*)hi, (const uint16_t*)lo, n, fmt, out);
  WF_CHECK_LAUNCH("join16");
  return WF_OK;
}

// ================================================================================= GCNConv backward on the tensor cores
// loss.backward() through Y = dropout(relu((A_hat X) W^T + b)) (model.py:31-42 under autograd -- STGCN.forward, the only
// differentiable use of GCNConv):
//   dZ = dY * mask * (Y > 0)                  -> TB8 bf16 hi/lo planes, block = (window, 128-row tile)     wf_ss_dz_tb8_kernel
//   AX = A_hat X                               -> row-major bf16 hi/lo planes                               wf_ss_ax_rows_kernel
//   dW = dZ^T AX, db = dZ^T 1                  -> wf_wg_kernel (both operands MN-major, M = Cout)
//   P  = dZ W                                  -> wf_ss_kernel (A = the same dZ planes K-major, B = W^T bf16 hi/lo), TB4 fp32
//   dX = A_hat^T P                             -> wf_ss_dx_finish_kernel (transposed CSR; TB4 -> row-major fp32)
namespace {

// one CTA per TB8 block, one thread per row of it; rows past the window (or past rpt) are written as zeros
__global__ void __launch_bounds__(128) wf_ss_dz_tb8_kernel(const float* __restrict__ dY, const float* __restrict__ Y, int Cout, int R,
                                                           int tpw, int rpt, int relu, const DropCfg drop,
                                                           uint16_t* __restrict__ out, long long plane) {
  const long long blk = blockIdx.x;
  const int z = (int)(blk / tpw), nt = (int)(blk % tpw), rl = threadIdx.x;
  const int r = nt * rpt + rl;
  const bool valid = rl < rpt && r < R;
  const long long grow = (long long)z * R + r;
  DropState ds;
  if (drop.rng != nullptr) ds = wf_drop_state(drop);
  for (int c8 = 0; c8 < (Cout >> 3); ++c8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (valid) {
      const float* src = dY + grow * Cout + 8 * c8;
      const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src + 4));
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      if (drop.rng != nullptr) {
        float m[8];
        const unsigned long long e4 = (unsigned long long)(grow * Cout + 8 * c8) >> 2;
        wf_drop4(ds, e4, m);
        wf_drop4(ds, e4 + 1, m + 4);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= m[j];
      }
      if (relu) {
        const float* ys = Y + grow * Cout + 8 * c8;
        const float4 ya = __ldg(reinterpret_cast<const float4*>(ys)), yb = __ldg(reinterpret_cast<const float4*>(ys + 4));
        const float y[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = y[j] > 0.f ? v[j] : 0.f;
      }
    }
    uint4 hi, lo;
    ss_split_bf16(v[0], v[1], hi.x, lo.x); ss_split_bf16(v[2], v[3], hi.y, lo.y);
    ss_split_bf16(v[4], v[5], hi.z, lo.z); ss_split_bf16(v[6], v[7], hi.w, lo.w);
    const long long o = ((blk * (Cout >> 3) + c8) * 128 + rl) * 8;
    *reinterpret_cast<uint4*>(out + o) = hi;
    *reinterpret_cast<uint4*>(out + plane + o) = lo;
  }
}

// AX[z][r][:] = sum_p val[p] X[z][col[p]][:] as bf16 hi/lo planes with row pitch CinP >= Cin (columns past Cin zero):
// one warp per row, 8 channels per lane
__global__ void __launch_bounds__(256) wf_ss_ax_rows_kernel(const float* __restrict__ X, int Cin, int CinP, int R, long long rows,
                                                            const int* __restrict__ rowptr, const int* __restrict__ col,
                                                            const float* __restrict__ val, uint16_t* __restrict__ out,
                                                            long long plane) {
  const long long gr = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (gr >= rows) return;
  const long long z = gr / R;
  const int r = (int)(gr - z * R);
  const int p0 = __ldg(rowptr + r), p1 = __ldg(rowptr + r + 1);
  for (int c8 = lane; c8 < (CinP >> 3); c8 += 32) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (8 * c8 < Cin) {
      for (int p = p0; p < p1; ++p) {
        const float w = __ldg(val + p);
        const float* src = X + (z * R + __ldg(col + p)) * Cin + 8 * c8;
        const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src + 4));
        acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]); acc[2] = fmaf(w, a.z, acc[2]); acc[3] = fmaf(w, a.w, acc[3]);
        acc[4] = fmaf(w, b.x, acc[4]); acc[5] = fmaf(w, b.y, acc[5]); acc[6] = fmaf(w, b.z, acc[6]); acc[7] = fmaf(w, b.w, acc[7]);
      }
    }
    uint4 hi, lo;
    ss_split_bf16(acc[0], acc[1], hi.x, lo.x); ss_split_bf16(acc[2], acc[3], hi.y, lo.y);
    ss_split_bf16(acc[4], acc[5], hi.z, lo.z); ss_split_bf16(acc[6], acc[7], hi.w, lo.w);
    const long long o = gr * CinP + 8 * c8;
    *reinterpret_cast<uint4*>(out + o) = hi;
    *reinterpret_cast<uint4*>(out + plane + o) = lo;
  }
}

// dX[z][r][:] = sum_p val_t[p] P[z][col_t[p]][:], P in TB4 blocks [Cin / 4][128 rows][4]; one CTA per block, one thread per row
__global__ void __launch_bounds__(128) wf_ss_dx_finish_kernel(const float* __restrict__ P, int Cin, int R, int tpw, int rpt,
                                                              const int* __restrict__ rowptr_t, const int* __restrict__ col_t,
                                                              const float* __restrict__ val_t, float* __restrict__ dX) {
  const long long blk = blockIdx.x;
  const int z = (int)(blk / tpw), nt = (int)(blk % tpw), rl = threadIdx.x;
  const int r = nt * rpt + rl;
  if (rl >= rpt || r >= R) return;
  const int p0 = __ldg(rowptr_t + r), p1 = __ldg(rowptr_t + r + 1);
  const float4* P4 = reinterpret_cast<const float4*>(P);
  float* dst = dX + ((long long)z * R + r) * Cin;
  const int q = Cin >> 2;
  if (p1 - p0 == 1 && __ldg(col_t + p0) == r) {  // the common row: its own (unit or not) self loop only
    const float w = __ldg(val_t + p0);
    for (int c4 = 0; c4 < q; c4 += 2) {
      float4 a = P4[(blk * q + c4) * 128 + rl], b = P4[(blk * q + c4 + 1) * 128 + rl];
      a.x *= w; a.y *= w; a.z *= w; a.w *= w; b.x *= w; b.y *= w; b.z *= w; b.w *= w;
      *reinterpret_cast<float4*>(dst + 4 * c4) = a;
      *reinterpret_cast<float4*>(dst + 4 * c4 + 4) = b;
    }
    return;
  }
  for (int c4 = 0; c4 < q; c4 += 2) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    for (int p = p0; p < p1; ++p) {
      const float w = __ldg(val_t + p);
      const int c = __ldg(col_t + p);
      const long long sb = (long long)z * tpw + c / rpt;
      const int sl = c % rpt;
      const float4 u = P4[(sb * q + c4) * 128 + sl], v = P4[(sb * q + c4 + 1) * 128 + sl];
      a.x = fmaf(w, u.x, a.x); a.y = fmaf(w, u.y, a.y); a.z = fmaf(w, u.z, a.z); a.w = fmaf(w, u.w, a.w);
      b.x = fmaf(w, v.x, b.x); b.y = fmaf(w, v.y, b.y); b.z = fmaf(w, v.z, b.z); b.w = fmaf(w, v.w, b.w);
    }
    *reinterpret_cast<float4*>(dst + 4 * c4) = a;
    *reinterpret_cast<float4*>(dst + 4 * c4 + 4) = b;
  }
}

// WT[n][k] = W[k][n] as bf16 hi/lo planes [2][Cin][Cout]
__global__ void wf_ss_wt_split_kernel(const float* __restrict__ W, int Cout, int Cin, uint16_t* __restrict__ out, long long plane) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * Cin) return;
  const int n = i / Cout, k = i - n * Cout;
  const float w = W[(long long)k * Cin + n];
  const __nv_bfloat16 h = __float2bfloat16_rn(w);
  const __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
  out[i] = *reinterpret_cast<const uint16_t*>(&h);
  out[plane + i] = *reinterpret_cast<const uint16_t*>(&l);
}

struct GcnBwdWs { size_t dz, ax, p, wt, part, total; int cinp, splits_max; };
GcnBwdWs gcn_bwd_ws(int R, int Cin, int Cout, int Bw, bool need_dx) {
  GcnBwdWs w;
  const long long tpw = wf_cdiv(R, 128), blocks = (long long)Bw * tpw;
  w.cinp = Cin <= 128 ? 128 : 256;
  w.splits_max = 40;
  auto up = [](size_t b) { return (b + 1023) & ~(size_t)1023; };
  w.dz = 0;
  w.ax = w.dz + up((size_t)blocks * Cout * 128 * 2 * 2);
  w.p = w.ax + up((size_t)Bw * R * w.cinp * 2 * 2);
  w.wt = w.p + (need_dx ? up((size_t)blocks * Cin * 128 * 4) : 0);
  w.part = w.wt + up((size_t)Cin * Cout * 2 * 2);
  w.total = w.part + up((size_t)w.splits_max * Cout * (w.cinp + 1) * 4);
  return w;
}

}  // namespace

extern "C" size_t wf_gcn_layer_bwd_ss_workspace_bytes(int R, int Cin, int Cout, int Bw) {
  return gcn_bwd_ws(R, Cin, Cout, Bw, true).total;
}

// X [Bw*R][Cin], Y / dY [Bw*R][Cout], W [Cout][Cin] fp32 row-major; CSR by target (A_hat) and its transpose over the R rows of
// one window (shared by all windows).  dX may be NULL (input without gradient).  Cout a multiple of 128, Cin a multiple
// of 8 up to 256 (dX additionally needs Cin = 128 or 256).  Y is read only when relu != 0.
extern "C" int wf_gcn_layer_bwd_ss(const float* X, const float* Y, const float* dY, const float* W, const int* rowptr,
                                   const int* col, const float* val, const int* rowptr_t, const int* col_t, const float* val_t,
                                   int R, int Cin, int Cout, int Bw, int relu, float p_drop, const unsigned long long* rng,
                                   int site, float* dX, float* dW, float* db, void* workspace, size_t workspace_bytes, int* err,
                                   void* stream) {
  WF_REQUIRE(R > 0 && Bw > 0, "gcn_layer_bwd_ss: bad batch");
  WF_REQUIRE(Cout % 128 == 0 && Cin % 8 == 0 && Cin <= 256, "gcn_layer_bwd_ss: Cout=%d must be a multiple of 128, Cin=%d of 8 and <= 256", Cout, Cin);
  WF_REQUIRE(dX == nullptr || Cin % 128 == 0, "gcn_layer_bwd_ss: dX needs Cin=%d to be 128 or 256", Cin);
  WF_REQUIRE(!relu || Y != nullptr, "gcn_layer_bwd_ss: the ReLU gate needs Y");
  WF_REQUIRE(p_drop >= 0.f && p_drop < 1.f && (p_drop == 0.f || rng != nullptr), "gcn_layer_bwd_ss: bad dropout arguments");
  const GcnBwdWs w = gcn_bwd_ws(R, Cin, Cout, Bw, dX != nullptr);
  WF_REQUIRE(workspace != nullptr && workspace_bytes >= w.total && (uintptr_t)workspace % 256 == 0,
             "gcn_layer_bwd_ss: workspace too small or not 256-byte aligned (%zu < %zu)", workspace_bytes, w.total);
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* base = (uint8_t*)workspace;
  const int tpw = wf_cdiv(R, 128), rpt = wf_tile_rows(R);
  const long long blocks = (long long)Bw * tpw, rows = (long long)Bw * R;
  uint16_t* dz16 = (uint16_t*)(base + w.dz);
  uint16_t* ax16 = (uint16_t*)(base + w.ax);
  float* P = (float*)(base + w.p);
  uint16_t* wt16 = (uint16_t*)(base + w.wt);
  float* part = (float*)(base + w.part);
  const long long dz_plane = blocks * Cout * 128, ax_plane = rows * w.cinp;
  const DropCfg dc = wf_drop_cfg(p_drop, rng, site);
  wf_ss_dz_tb8_kernel<<<(unsigned)blocks, 128, 0, st>>>(dY, Y, Cout, R, tpw, rpt, relu, dc, dz16, dz_plane);
  WF_CHECK_LAUNCH("ss_dz_tb8");
  int rc;
  if (dW != nullptr || db != nullptr) {
    WF_REQUIRE(dW != nullptr, "gcn_layer_bwd_ss: db comes with dW");
    wf_ss_ax_rows_kernel<<<(unsigned)wf_cdiv(rows, 8), 256, 0, st>>>(X, Cin, w.cinp, R, rows, rowptr, col, val, ax16, ax_plane);
    WF_CHECK_LAUNCH("ss_ax_rows");
    const int nh = w.cinp / 128;
    const void* src[2] = {ax16, ax16};
    const long long plane[2] = {ax_plane, ax_plane};
    const int var[2] = {1, 1}, shift[2] = {0, 0}, col0[2] = {0, 128}, ch[2] = {w.cinp, w.cinp};
    const int w0 = Cin < 128 ? Cin : 128, w1 = Cin > 128 ? Cin - 128 : 0;
    rc = wf_ss_launch_wgrad(dz16, dz_plane, nh, src, plane, var, shift, col0, ch, 1, R, Bw, 1, part,
                            (size_t)w.splits_max * Cout * (w.cinp + 1), dW, Cin, w0, w1 > 0 ? dW + 128 : nullptr, Cin, w1, db, nullptr, 0,
                            err, st, Cout);
    if (rc) return rc;
  }
  if (dX != nullptr) {
    wf_ss_wt_split_kernel<<<wf_cdiv((long long)Cin * Cout, 256), 256, 0, st>>>(W, Cout, Cin, wt16, (long long)Cin * Cout);
    WF_CHECK_LAUNCH("ss_wt_split");
    rc = wf_ss_launch_nodes(128, SS_A_KT, dz16, dz_plane, Cout, 1, wt16, wt16 + (long long)Cin * Cout, Cout, 0, Cin, 1, nullptr, nullptr,
                            0, P, 1, R, Bw, 1, nullptr, err, st, 1);
    if (rc) return rc;
    wf_ss_dx_finish_kernel<<<(unsigned)blocks, 128, 0, st>>>(P, Cin, R, tpw, rpt, rowptr_t, col_t, val_t, dX);
    WF_CHECK_LAUNCH("ss_dx_finish");
  }
  return WF_OK;
}

// ---- test entry points: the two kernels on operands given as plain 16-bit planes, so that every operand layout
// (K-major SWIZZLE_64B / TB8 K-major / TB8 MN-major / row-major MN-major SWIZZLE_128B, fp16 and bf16) can be checked
// against a float64 product in isolation (tests/test_gpu_ss.py).
// C (TB4 fp32 [blocks][Ntot/4][128][4]) = A W^T: avar 0: A16 row-major planes [2][G*Bw*T][Nn][K]; 1: TB8 planes.
extern "C" int wf_ss_nodes_gemm(int bn, int avar, const void* A16, long long a_plane, int K, int afmt, const void* W16_hi,
                                const void* W16_lo, long long w_group_stride, int Ntot, int bfmt, const float* bias,
                                const float* bias2, long long bias_group_stride, float* C, int T, int Nn, int Bw, int G,
                                int k_parts, int* err, void* stream) {
  return wf_ss_launch_nodes(bn, avar, A16, a_plane, K, afmt, W16_hi, W16_lo, K, w_group_stride, Ntot, bfmt, bias, bias2,
                            bias_group_stride, C, T, Nn, Bw, G, nullptr, err, (cudaStream_t)stream, k_parts);
}
// dst0 [G][512][w0] (+ dst1 [G][512][w1], db [G][512]) = dG^T [B0 | B1] over all blocks: dg16 TB8 bf16 planes (512 channels);
// half h: bvar 0 TB8 fp16 planes with bC channels (bcol0: first channel), 1 row-major fp16 planes [2][G*Bw*T][Nn][bC];
// bshift 1: the block of the previous step.  part: scratch of part_floats floats.
extern "C" int wf_ss_wgrad(const void* dg16, long long dg_plane, int nh, const void* b0, long long b0_plane, int b0var,
                           int b0shift, int b0col0, int b0C, const void* b1, long long b1_plane, int b1var, int b1shift,
                           int b1col0, int b1C, int T, int Nn, int Bw, int G, float* part, long long part_floats, float* dst0,
                           int ld0, int w0, float* dst1, int ld1, int w1, float* db, long long gstride, int* err,
                           void* stream) {
  const void* src[2] = {b0, b1};
  const long long plane[2] = {b0_plane, b1_plane};
  const int var[2] = {b0var, b1var}, shift[2] = {b0shift, b1shift}, col0[2] = {b0col0, b1col0}, ch[2] = {b0C, b1C};
  return wf_ss_launch_wgrad(dg16, dg_plane, nh, src, plane, var, shift, col0, ch, T, Nn, Bw, G, part, (size_t)part_floats, dst0,
                            ld0, w0, dst1, ld1, w1, db, nullptr, gstride, err, (cudaStream_t)stream, 512);
}

#ifdef WF_SS_TRACE
extern "C" int wf_ss_trace_read(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, wf_ss_trace_buf, sizeof(long long) * 6 * 256);
}
#endif
