// Region-graph kernels: grid kNN builder, GCN symmetric normalisation -> CSR (+ transpose),
// and the stand-alone CSR SpMM.
//
// Replaces, for the hot path:
//   * graphBuilder.py:9-47  build_spatial_graph  (scipy cKDTree on the host + a Python loop)
//   * PyG gcn_norm, re-run inside every GCNConv.forward (model.py:31-40,
//     hybrid_model.py:65-74) -> here computed once per region and kept as CSR values.
#include "wf_common.cuh"

#define WF_KNN_MAXK 32

// ------------------------------------------------------------------ kNN
// One thread per node.  Candidates are ranked by (squared Euclidean distance in raw
// degrees [float64, no FMA contraction], node index) -- the deterministic rule documented in
// DESIGN.md.  stencil > 0: only the (2*stencil+1)^2 index box around the node is scanned,
// which is exhaustive for strictly monotonic coordinate axes when stencil >= k (every point
// outside the box is farther than k points of the node's own row/column).  stencil == 0:
// brute force over all nodes.
__global__ void wf_knn_kernel(const double* __restrict__ lats, int nlat, const double* __restrict__ lons,
                              int nlon, int k, int stencil, long long* __restrict__ edge_index) {
  const int n = nlat * nlon;
  int node = blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= n) return;
  const int il = node / nlon, io = node - il * nlon;
  const double la = lats[il], lo = lons[io];
  double bd[WF_KNN_MAXK];
  int bi[WF_KNN_MAXK];
  int cnt = 0;
  int l0 = 0, l1 = nlat - 1, o0 = 0, o1 = nlon - 1;
  if (stencil > 0) {
    l0 = max(0, il - stencil); l1 = min(nlat - 1, il + stencil);
    o0 = max(0, io - stencil); o1 = min(nlon - 1, io + stencil);
  }
  for (int a = l0; a <= l1; ++a) {
    const double dla = __dsub_rn(lats[a], la);
    const double dla2 = __dmul_rn(dla, dla);
    for (int b = o0; b <= o1; ++b) {
      const int j = a * nlon + b;
      if (j == node) continue;
      const double dlo = __dsub_rn(lons[b], lo);
      const double d2 = __dadd_rn(dla2, __dmul_rn(dlo, dlo));
      // candidates arrive in ascending index order, so "strictly smaller" keeps the
      // lower index first among equal distances
      if (cnt == k && !(d2 < bd[k - 1])) continue;
      int pos = cnt < k ? cnt : k - 1;
      while (pos > 0 && d2 < bd[pos - 1]) {
        bd[pos] = bd[pos - 1];
        bi[pos] = bi[pos - 1];
        --pos;
      }
      bd[pos] = d2;
      bi[pos] = j;
      if (cnt < k) ++cnt;
    }
  }
  const long long E = (long long)n * k;
  for (int q = 0; q < k; ++q) {
    edge_index[(long long)node * k + q] = node;                       // row 0: source = node
    edge_index[E + (long long)node * k + q] = q < cnt ? bi[q] : node;  // row 1: neighbour
  }
}

// ------------------------------------------------------------------ gcn_norm -> CSR
// Pass 1: in-degree by target over the non-self-loop edges (+1 self loop added later) and
// out-degree by source for the transposed structure.
__global__ void wf_csr_count_kernel(const long long* __restrict__ ei, long long E, int R, int* cnt_in,
                                    int* cnt_out, int* err) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  long long s = ei[e], d = ei[E + e];
  if (s < 0 || s >= R || d < 0 || d >= R) { atomicExch(err, 1); return; }
  if (s == d) return;
  atomicAdd(cnt_in + d, 1);
  atomicAdd(cnt_out + s, 1);
}

// Single-block exclusive scan of (cnt[r] + 1) -> rowptr[0..R]; also dis[r] = 1/sqrt(cnt_in+1)
// computed as 1.0f / sqrtf(deg) with two IEEE roundings, which is bitwise what
// torch's deg.pow(-0.5) yields on the CPU (SURVEY.md "Hard parts").
__global__ void wf_csr_scan_kernel(const int* __restrict__ cnt, int R, int* __restrict__ rowptr,
                                   float* __restrict__ dis /* may be null */) {
  __shared__ int sh[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < R; base += 1024) {
    int r = base + threadIdx.x;
    int v = r < R ? cnt[r] + 1 : 0;
    if (dis && r < R) dis[r] = __fdiv_rn(1.0f, __fsqrt_rn((float)v));
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (r < R) rowptr[r] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) rowptr[R] = carry;
}

// Pass 2: drop each non-self edge id into its row segment (arbitrary order, sorted in pass 3).
__global__ void wf_csr_fill_kernel(const long long* __restrict__ ei, long long E, const int* __restrict__ rp_in,
                                   const int* __restrict__ rp_out, int* cur_in, int* cur_out, int* eid_in,
                                   int* eid_out) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int s = (int)ei[e], d = (int)ei[E + e];
  if (s == d) return;
  eid_in[rp_in[d] + atomicAdd(cur_in + d, 1)] = (int)e;
  eid_out[rp_out[s] + atomicAdd(cur_out + s, 1)] = (int)e;
}

// Pass 3: per row, order the edge ids ascending (PyG scatters messages in edge order with the
// self loops appended last, so this reproduces the reference's summation order), then emit
// (column, weight).  transposed == 0: row = target, column = source (A_hat, forward).
// transposed == 1: row = source, column = target (A_hat^T, backward).  Weight of edge
// (s -> d) is dis[s] * dis[d] either way; the self loop (r, r) gets dis[r]^2 and comes last.
__global__ void wf_csr_emit_kernel(const long long* __restrict__ ei, long long E, int R,
                                   const int* __restrict__ rowptr, int* __restrict__ eid,
                                   const float* __restrict__ dis, int transposed, int* __restrict__ col,
                                   float* __restrict__ val) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  int p0 = rowptr[r], p1 = rowptr[r + 1] - 1;  // last slot is the self loop
  for (int i = p0 + 1; i < p1; ++i) {
    int key = eid[i], j = i - 1;
    while (j >= p0 && eid[j] > key) { eid[j + 1] = eid[j]; --j; }
    eid[j + 1] = key;
  }
  for (int p = p0; p < p1; ++p) {
    int e = eid[p];
    int s = (int)ei[e], d = (int)ei[E + e];
    col[p] = transposed ? d : s;
    val[p] = __fmul_rn(dis[s], dis[d]);
  }
  col[p1] = r;
  val[p1] = __fmul_rn(dis[r], dis[r]);
}

// ------------------------------------------------------------------ SpMM  Z = A_hat X
// One warp per output row; lanes stride the channel quads with 128-bit loads.  Rows whose
// only entry is the unit self loop (every row >= N on the reference path)
// reduce to a straight copy.
__global__ void wf_spmm_kernel(const float* __restrict__ X, RowMap xm, long long gX, const int* __restrict__ rowptr,
                               const int* __restrict__ col, const float* __restrict__ val, long long gRowptr,
                               long long gCsr, int R, int rows /* per group = Bw*R */, int C,
                               float* __restrict__ Z, long long gZ) {
  const int g = blockIdx.y;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int w = warp / R, rr = warp - w * R;
  const int* rp = rowptr + g * gRowptr;
  const int* cl = col + g * gCsr;
  const float* vl = val + g * gCsr;
  const float* Xg = X + g * gX;
  const long long wbase = row_off(xm, w * R);
  const int p0 = rp[rr], p1 = rp[rr + 1];
  float* zrow = Z + g * gZ + (long long)warp * C;
  for (int c = lane * 4; c < C; c += 128) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = p0; p < p1; ++p) {
      float v = __ldg(vl + p);
      float4 x = __ldg(reinterpret_cast<const float4*>(Xg + wbase + (long long)__ldg(cl + p) * xm.ld + c));
      a.x = fmaf(v, x.x, a.x); a.y = fmaf(v, x.y, a.y);
      a.z = fmaf(v, x.z, a.z); a.w = fmaf(v, x.w, a.w);
    }
    *reinterpret_cast<float4*>(zrow + c) = a;
  }
}

// ------------------------------------------------------------------ C-ABI
extern "C" int wf_knn_grid_build(const double* lats, int nlat, const double* lons, int nlon, int k,
                                 int monotonic, long long* edge_index, void* stream) {
  WF_REQUIRE(nlat > 0 && nlon > 0, "knn: empty grid");
  WF_REQUIRE(k >= 1 && k <= WF_KNN_MAXK, "knn: k=%d out of range [1,%d]", k, WF_KNN_MAXK);
  WF_REQUIRE((long long)nlat * nlon > k, "knn: need more than k nodes");
  int n = nlat * nlon;
  int stencil = monotonic ? k : 0;
  wf_knn_kernel<<<wf_cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(lats, nlat, lons, nlon, k, stencil, edge_index);
  WF_CHECK_LAUNCH("knn");
  return WF_OK;
}

extern "C" size_t wf_gcn_norm_workspace_bytes(long long E, int R) {
  // cnt_in, cnt_out, cur_in, cur_out [R] ; rp_out [R+1]; eid_in, eid_out [E+R]; dis [R]; err [1]
  return sizeof(int) * (size_t)(4LL * R + (R + 1) + 2 * (E + R) + 1) + sizeof(float) * (size_t)R + 64;
}

// Builds A_hat (rowptr/col/val, rows = targets) and A_hat^T (rowptr_t/col_t/val_t, rows =
// sources) over R = window * N rows from edge_index i64[2, E].  Capacity of col/val: E + R.
// The true nnz is rowptr[R] (E minus self loops, plus R).
extern "C" int wf_gcn_norm_csr(const long long* edge_index, long long E, int R, int* rowptr, int* col, float* val,
                               int* rowptr_t, int* col_t, float* val_t, void* workspace, size_t workspace_bytes,
                               void* stream) {
  WF_REQUIRE(R > 0 && E >= 0, "gcn_norm: bad sizes");
  if (workspace_bytes < wf_gcn_norm_workspace_bytes(E, R))
    return wf_fail(WF_EWORKSPACE, "gcn_norm: workspace %zu < %zu", workspace_bytes, wf_gcn_norm_workspace_bytes(E, R));
  cudaStream_t st = (cudaStream_t)stream;
  int* cnt_in = (int*)workspace;
  int* cnt_out = cnt_in + R;
  int* cur_in = cnt_out + R;
  int* cur_out = cur_in + R;
  int* err = cur_out + R;
  int* eid_in = err + 1;
  int* eid_out = eid_in + (E + R);
  float* dis = (float*)(eid_out + (E + R));
  if (cudaMemsetAsync(workspace, 0, sizeof(int) * (4 * (size_t)R + 1), st) != cudaSuccess)
    return wf_fail(WF_ECUDA, "gcn_norm: memset failed");
  if (E > 0) {
    wf_csr_count_kernel<<<wf_cdiv(E, 256), 256, 0, st>>>(edge_index, E, R, cnt_in, cnt_out, err);
    WF_CHECK_LAUNCH("csr_count");
  }
  wf_csr_scan_kernel<<<1, 1024, 0, st>>>(cnt_in, R, rowptr, dis);
  wf_csr_scan_kernel<<<1, 1024, 0, st>>>(cnt_out, R, rowptr_t, nullptr);
  WF_CHECK_LAUNCH("csr_scan");
  if (E > 0) {
    wf_csr_fill_kernel<<<wf_cdiv(E, 256), 256, 0, st>>>(edge_index, E, rowptr, rowptr_t, cur_in, cur_out, eid_in, eid_out);
    WF_CHECK_LAUNCH("csr_fill");
  }
  wf_csr_emit_kernel<<<wf_cdiv(R, 128), 128, 0, st>>>(edge_index, E, R, rowptr, eid_in, dis, 0, col, val);
  wf_csr_emit_kernel<<<wf_cdiv(R, 128), 128, 0, st>>>(edge_index, E, R, rowptr_t, eid_out, dis, 1, col_t, val_t);
  WF_CHECK_LAUNCH("csr_emit");
  return WF_OK;
}

int wf_launch_spmm(const float* X, RowMap xm, long long gX, const int* rowptr, const int* col, const float* val,
                   long long gRowptr, long long gCsr, int R, int rows, int C, float* Z, long long gZ, int groups,
                   cudaStream_t st) {
  WF_REQUIRE(C % 4 == 0 && xm.ld % 4 == 0, "spmm: channels must be a multiple of 4");
  WF_REQUIRE(xm.rows_per_blk == R, "spmm: rows_per_blk must equal R");
  dim3 grid(wf_cdiv((long long)rows * 32, 256), groups);
  wf_spmm_kernel<<<grid, 256, 0, st>>>(X, xm, gX, rowptr, col, val, gRowptr, gCsr, R, rows, C, Z, gZ);
  WF_CHECK_LAUNCH("spmm");
  return WF_OK;
}
