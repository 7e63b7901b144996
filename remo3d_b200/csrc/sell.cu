// Sliced-ELLPACK copy of the stiffness matrix for the multi-right-hand-side PCG SpMM (Q = A P + fused p.q).
//
// Why a second layout.  The CSR SpMM (solver.cu k_spmm_p) sits at 90 % of the SM's L1 data pipe (profiles/r01_notes.md):
// per matrix entry one wavefront for the gathered row of P -- 1.4 when the 48-byte rows of a 6-column block straddle
// 128-byte lines -- plus the shuffles that hand the (col, val) pairs of a row to the lanes of its group.  Here:
//   * slices of 8 rows (one warp = 8 groups of 4 lanes at 5..8 right-hand sides); inside a slice the entries are stored
//     in chunks of 4 per row, rows interleaved: col[chunk][row 0..7][4] (int32), val[chunk][row 0..7][4] (fp64), so the
//     (col, val) data of a warp step are 384 contiguous bytes and need no shuffles;
//   * rows are taken in Morton order of their dof location and sorted by length inside windows of 2048 rows
//     (SELL-C-sigma): a slice is padded only to the longest of 8 similar rows (2 % padding on the bench meshes) and the
//     8 groups of a warp run the same trip count; padding entries are (own row, 0.0);
//   * the search directions P are kept in their own block with a power-of-two row stride (64 bytes for 5..8 columns),
//     so a gathered row never straddles a line.
// Two kernels read it: k_spmm_stream8 (P stride 8, i.e. 5..8 right-hand sides: per-warp contiguous chunk ranges streamed
// through a cp.async ring in shared memory) and the generic k_spmm_sell<KS> (KS/2 lanes per row, register prefetch of
// the next chunk) for the other strides.  Measured on B200 at 4.8 M dofs (136.9 M non-zeros): 5 RHS 1.04 -> 0.865 ms
// (36.5 % of the measured HBM peak by algorithmic bytes), 8 RHS 0.856 ms (41 %), 2 RHS 0.39 ms (72 %).
// The CSR arrays stay the assembly target and the parity export (remo_matrix_get); this copy is made once per matrix
// by remo_precond_setup (two radix sorts of ndof keys + one streaming pass over the matrix, ~5 ms at 4.8 M dofs).
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdlib>

#include "space_view.cuh"

namespace {

constexpr int TB = 256;
constexpr int KMAX = REMO_MAX_RHS;
int sell_env(const char* name, int def) {
  const char* e = getenv(name);
  return e ? atoi(e) : def;
}

int sigma_rows() {  // sorting window (rows); a multiple of the slice height
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("REMO_SELL_SIGMA");
    v = e ? atoi(e) : 2048;
    if (v < 8) v = 8;
    v &= ~7;
  }
  return v;
}

// position i of the cluster order holds row order0[i] (identity when order0 == nullptr)
__global__ void k_sell_keys(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ order0, int64_t n, int SIGMA,
                            uint32_t* __restrict__ key, int32_t* __restrict__ idx) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t r = order0 ? order0[i] : i;
  const int64_t len = rowptr[r + 1] - rowptr[r];
  const uint32_t inv = 4095u - (uint32_t)(len > 4095 ? 4095 : len);  // longest rows first inside a window
  key[i] = ((uint32_t)(i / SIGMA) << 12) | inv;
  idx[i] = (int32_t)r;
}

// ---- spatial cluster order of the rows: Morton code of the dof's location (vertex, edge midpoint, face / cell centroid)
// bounding box of the vertices, two stages: per-block min / max, then one block over the block results
__global__ void k_bbox_stage(const double* __restrict__ in, int64_t count, int dim, int from_blocks, double* __restrict__ out) {
  // from_blocks == 0: `in` = xyz (count vertices);  1: `in` = count block results [6] -> out[0..2] = min, out[3..5] = max
  __shared__ double smin[3][TB], smax[3][TB];
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int64_t i = (int64_t)blockIdx.x * TB + threadIdx.x; i < count; i += (int64_t)gridDim.x * TB)
    for (int d = 0; d < dim; d++) {
      const double lo = from_blocks ? in[i * 6 + d] : in[i * dim + d];
      const double hi = from_blocks ? in[i * 6 + 3 + d] : lo;
      mn[d] = fmin(mn[d], lo);
      mx[d] = fmax(mx[d], hi);
    }
  for (int d = 0; d < 3; d++) { smin[d][threadIdx.x] = mn[d]; smax[d][threadIdx.x] = mx[d]; }
  __syncthreads();
  for (int w = TB / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w)
      for (int d = 0; d < 3; d++) {
        smin[d][threadIdx.x] = fmin(smin[d][threadIdx.x], smin[d][threadIdx.x + w]);
        smax[d][threadIdx.x] = fmax(smax[d][threadIdx.x], smax[d][threadIdx.x + w]);
      }
    __syncthreads();
  }
  if (threadIdx.x < 3) {
    out[(int64_t)blockIdx.x * 6 + threadIdx.x] = smin[threadIdx.x][0];
    out[(int64_t)blockIdx.x * 6 + 3 + threadIdx.x] = smax[threadIdx.x][0];
  }
}

__global__ void k_sell_morton(SpaceView s, const double* __restrict__ xyz, const double* __restrict__ lohi,
                              uint64_t* __restrict__ code, int32_t* __restrict__ idx) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= s.ndof) return;
  int32_t v[3];
  int m = 1;
  if (i < s.nv) {
    v[0] = (int32_t)i;
  } else if (i < s.face_base) {
    const uint64_t k = s.edge_keys[(i - s.edge_base) / (s.order - 1)];
    v[0] = (int32_t)(k >> 32); v[1] = (int32_t)(k & 0xffffffffu); m = 2;
  } else if (s.dim == 3) {
    const uint64_t f = s.face_keys[i - s.face_base];
    const uint64_t k = s.edge_keys[f >> 32];
    v[0] = (int32_t)(k >> 32); v[1] = (int32_t)(k & 0xffffffffu); v[2] = (int32_t)(f & 0xffffffffu); m = 3;
  } else {
    const int64_t t = i - s.face_base;
    v[0] = s.sv[t * 3]; v[1] = s.sv[t * 3 + 1]; v[2] = s.sv[t * 3 + 2]; m = 3;
  }
  uint64_t c = 0;
  for (int d = 0; d < s.dim; d++) {
    double x = 0.0;
    for (int q = 0; q < m; q++) x += xyz[(int64_t)v[q] * s.dim + d];
    x /= m;
    const double ext = lohi[3 + d] - lohi[d];
    const double t = ext > 0 ? (x - lohi[d]) / ext : 0.0;
    const uint64_t q = (uint64_t)fmin(fmax(t * 2097151.0, 0.0), 2097151.0);
    c |= spread21s(q) << d;
  }
  code[i] = c;
  idx[i] = (int32_t)i;
}

// chunks (of 4 entries) of every slice = ceil(longest of its 8 rows / 4); rows beyond n are padding (-1)
__global__ void k_sell_chunks(const int64_t* __restrict__ rowptr, int32_t* __restrict__ srow, int64_t n, int64_t nslices,
                              int32_t* __restrict__ nch) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= nslices) return;
  int64_t mx = 0;
  for (int g = 0; g < 8; g++) {
    const int64_t rs = s * 8 + g;
    if (rs < n) {
      const int32_t r = srow[rs];
      const int64_t len = rowptr[r + 1] - rowptr[r];
      mx = len > mx ? len : mx;
    } else {
      srow[rs] = -1;
    }
  }
  nch[s] = (int32_t)((mx + 3) >> 2);
}

__global__ void k_sell_ptr(const int32_t* __restrict__ nch, const int64_t* __restrict__ incl, int64_t nslices, int64_t* __restrict__ sptr) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s > nslices) return;
  sptr[s] = (s == 0) ? 0 : incl[s - 1];
  (void)nch;
}

__global__ void k_i32_to_i64(const int32_t* __restrict__ a, int64_t n, int64_t* __restrict__ b) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) b[i] = a[i];
}

// CTA b of the blocked SpMM owns the slices [part[b], part[b+1]) holding an equal share of the chunks
__global__ void k_sell_partition(const int64_t* __restrict__ sptr, int64_t nslices, int nparts, int64_t* __restrict__ part) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nparts) return;
  if (b == nparts) { part[b] = nslices; return; }
  const int64_t target = (int64_t)(((__int128)sptr[nslices] * b) / nparts);
  int64_t lo = 0, hi = nslices;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (sptr[mid] < target) lo = mid + 1; else hi = mid;
  }
  part[b] = lo;
}

// a 4-lane group per row slot: lane l copies entry l of every chunk (16 B of columns / 32 B of values per group and chunk)
__global__ void __launch_bounds__(TB) k_sell_fill(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                  const double* __restrict__ val, const int32_t* __restrict__ srow,
                                                  const int64_t* __restrict__ sptr, int64_t nslots, int32_t* __restrict__ scol,
                                                  double* __restrict__ sval) {
  const int l = threadIdx.x & 3;
  for (int64_t rs = ((int64_t)blockIdx.x * TB + threadIdx.x) >> 2; rs < nslots; rs += ((int64_t)gridDim.x * TB) >> 2) {
    const int64_t s = rs >> 3;
    const int g = (int)(rs & 7);
    const int32_t row = srow[rs];
    const int64_t st = row >= 0 ? rowptr[row] : 0;
    const int64_t len = row >= 0 ? rowptr[row + 1] - st : 0;
    const int32_t self = row >= 0 ? row : 0;
    const int64_t c0 = sptr[s], c1 = sptr[s + 1];
    for (int64_t ch = c0; ch < c1; ch++) {
      const int64_t j = (ch - c0) * 4 + l;
      const bool in = j < len;
      const int64_t o = (ch * 8 + g) * 4 + l;
      scol[o] = in ? __ldcs(col + st + j) : self;
      sval[o] = in ? __ldcs(val + st + j) : 0.0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Q = A P (constrained rows -> 0) and the per-column partial dots p.q.
// KS = row stride of P in doubles (2, 4, 8, 16, 32); KS/2 lanes own a row, every lane two adjacent right-hand sides
// (one 16-byte gather per entry); Q has the PCG's own stride ks (even, >= the number of right-hand sides).
// ------------------------------------------------------------------------------------------------
template <int KS, int TBK, int MINB, int U>
__global__ void __launch_bounds__(TBK, MINB) k_spmm_sell(const int64_t* __restrict__ sptr, const int32_t* __restrict__ scol,
                                                  const double* __restrict__ sval, const int32_t* __restrict__ srow,
                                                  const uint8_t* __restrict__ constrained, const double* __restrict__ P,
                                                  double* __restrict__ Q, int ks, int64_t nslots, double* __restrict__ partial,
                                                  const int64_t* __restrict__ part) {
  constexpr int LPR = KS / 2;     // lanes per row
  constexpr int RPC = TBK / LPR;  // row slots per CTA and pass
  const int l = threadIdx.x % LPR;
  const int slot = threadIdx.x / LPR;
  const bool on = 2 * l < ks;
  const double2* __restrict__ P2 = reinterpret_cast<const double2*>(P) + l;
  double dot0 = 0.0, dot1 = 0.0;
  // row slots of this CTA: a contiguous range of slices with an equal share of the chunks (part != nullptr: consecutive
  // row blocks of one CTA reuse the rows of P that its L1 already holds), or grid-strided
  int64_t rs = (int64_t)blockIdx.x * RPC + slot, rs_end = nslots, rs_step = (int64_t)gridDim.x * RPC;
  if (part) { rs = part[blockIdx.x] * 8 + slot; rs_end = part[blockIdx.x + 1] * 8; rs_step = RPC; }
  for (; rs < rs_end; rs += rs_step) {
    const int64_t s = rs >> 3;
    const int g = (int)(rs & 7);
    const int64_t c0 = sptr[s], c1 = sptr[s + 1];
    const int4* pc = reinterpret_cast<const int4*>(scol) + c0 * 8 + g;
    const double2* pv = reinterpret_cast<const double2*>(sval) + (c0 * 8 + g) * 2;
    double acc0 = 0.0, acc1 = 0.0;
    int4 nc = make_int4(0, 0, 0, 0);
    double2 nv0 = make_double2(0.0, 0.0), nv1 = nv0;
    if (c0 < c1) { nc = __ldcs(pc); nv0 = __ldcs(pv); nv1 = __ldcs(pv + 1); }
#pragma unroll U
    for (int64_t ch = c0; ch < c1; ch++) {
      const int4 cc = nc;
      const double2 v0 = nv0, v1 = nv1;
      pc += 8;
      pv += 16;
      if (ch + 1 < c1) { nc = __ldcs(pc); nv0 = __ldcs(pv); nv1 = __ldcs(pv + 1); }  // next chunk in flight during the gathers
      const double2 x0 = P2[(int64_t)cc.x * LPR];
      const double2 x1 = P2[(int64_t)cc.y * LPR];
      const double2 x2 = P2[(int64_t)cc.z * LPR];
      const double2 x3 = P2[(int64_t)cc.w * LPR];
      acc0 = fma(v0.x, x0.x, acc0); acc1 = fma(v0.x, x0.y, acc1);
      acc0 = fma(v0.y, x1.x, acc0); acc1 = fma(v0.y, x1.y, acc1);
      acc0 = fma(v1.x, x2.x, acc0); acc1 = fma(v1.x, x2.y, acc1);
      acc0 = fma(v1.y, x3.x, acc0); acc1 = fma(v1.y, x3.y, acc1);
    }
    const int32_t row = srow[rs];
    if (row >= 0 && on) {
      if (constrained[row]) { acc0 = 0.0; acc1 = 0.0; }
      const double2 p = P2[(int64_t)row * LPR];
      *reinterpret_cast<double2*>(Q + (int64_t)row * ks + 2 * l) = make_double2(acc0, acc1);
      dot0 = fma(acc0, p.x, dot0);
      dot1 = fma(acc1, p.y, dot1);
    }
  }
  __shared__ double sh[2][TBK];
  sh[0][threadIdx.x] = dot0;
  sh[1][threadIdx.x] = dot1;
  __syncthreads();
  if (threadIdx.x < LPR && 2 * threadIdx.x < ks) {
    double t0 = 0.0, t1 = 0.0;
    for (int i = threadIdx.x; i < TBK; i += LPR) { t0 += sh[0][i]; t1 += sh[1][i]; }
    partial[(int64_t)blockIdx.x * KMAX + 2 * threadIdx.x] = t0;
    partial[(int64_t)blockIdx.x * KMAX + 2 * threadIdx.x + 1] = t1;
  }
}

// ------------------------------------------------------------------------------------------------
// Streaming variant for 5..8 right-hand sides (P stride 8: a warp = one slice, 4 lanes per row).
// The generic kernel above is LATENCY bound (ncu: 80 % of the warp time on the long scoreboard, L1 data pipe 55-60 %,
// DRAM 30 %): every row start costs a dependent chain sptr -> first (col,val) chunk -> gathers, and the matrix stream is
// only one chunk ahead.  Here every warp owns a CONTIGUOUS range of slices, i.e. one contiguous run of chunks of the
// SELL arrays, and streams it through a private ring in shared memory with cp.async (one 16-byte copy per lane moves a
// whole 384-byte chunk: 128 B of columns + 256 B of values), D-1 chunks ahead and straight across row boundaries.
// The loop body then contains only the four gathers of P per lane; the row epilogue (own row of P, Q store, p.q) is
// deferred by one slice so that its loads are in flight during the next slice.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int D, int MINB>
__global__ void __launch_bounds__(256, MINB) k_spmm_stream8(const int64_t* __restrict__ sptr, const int32_t* __restrict__ scol,
                                                            const double* __restrict__ sval, const int32_t* __restrict__ srow,
                                                            const uint8_t* __restrict__ constrained, const double* __restrict__ P,
                                                            double* __restrict__ Q, int ks, double* __restrict__ partial,
                                                            const int64_t* __restrict__ wpart) {
  constexpr int CHB = 384;  // bytes of one chunk: 8 rows x 4 columns (int32) + 8 rows x 4 values (fp64)
  __shared__ __align__(16) unsigned char ring[8][D][CHB];
  __shared__ double sh[2][256];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int g = lane >> 2, l = lane & 3;
  const bool on = 2 * l < ks;
  const double2* __restrict__ P2 = reinterpret_cast<const double2*>(P) + l;
  double dot0 = 0.0, dot1 = 0.0;
  const int64_t gw = (int64_t)blockIdx.x * 8 + w;
  int64_t s = wpart[gw];
  const int64_t s_end = wpart[gw + 1];
  if (s < s_end) {
    int64_t ch = sptr[s];
    const int64_t ch_end = sptr[s_end];
    // this lane's 16 bytes of every chunk: lanes 0..7 the columns, 8..23 the values, 24..31 idle
    const unsigned char* src = lane < 8 ? reinterpret_cast<const unsigned char*>(scol) + lane * 16
                                        : reinterpret_cast<const unsigned char*>(sval) + (lane - 8) * 16;
    const int64_t src_step = lane < 8 ? 128 : 256;
    const uint32_t dst0 = smem_addr(&ring[w][0][0]) + (lane < 8 ? lane * 16 : 128 + (lane - 8) * 16);
    auto fetch = [&](int64_t c) {
      if (lane < 24 && c < ch_end)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (uint32_t)(c % D) * CHB), "l"(src + c * src_step) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int i = 0; i < D - 1; i++) fetch(ch + i);
    int64_t slice_end = sptr[s + 1];
    int32_t row = srow[s * 8 + g];
    double acc0 = 0.0, acc1 = 0.0;
    // deferred epilogue of the previous slice
    int32_t prow = -1;
    double pa0 = 0.0, pa1 = 0.0;
    double2 pp = make_double2(0.0, 0.0);
    uint8_t pcon = 0;
    for (; ch < ch_end; ch++) {
      asm volatile("cp.async.wait_group %0;" ::"n"(D - 2) : "memory");
      __syncwarp();
      fetch(ch + D - 1);  // refills the stage read in the previous iteration
      const unsigned char* stg = &ring[w][ch % D][0];
      const int4 cc = *reinterpret_cast<const int4*>(stg + g * 16);
      const double2 v0 = *reinterpret_cast<const double2*>(stg + 128 + g * 32);
      const double2 v1 = *reinterpret_cast<const double2*>(stg + 128 + g * 32 + 16);
      const double2 x0 = P2[(int64_t)cc.x * 4];
      const double2 x1 = P2[(int64_t)cc.y * 4];
      const double2 x2 = P2[(int64_t)cc.z * 4];
      const double2 x3 = P2[(int64_t)cc.w * 4];
      acc0 = fma(v0.x, x0.x, acc0); acc1 = fma(v0.x, x0.y, acc1);
      acc0 = fma(v0.y, x1.x, acc0); acc1 = fma(v0.y, x1.y, acc1);
      acc0 = fma(v1.x, x2.x, acc0); acc1 = fma(v1.x, x2.y, acc1);
      acc0 = fma(v1.y, x3.x, acc0); acc1 = fma(v1.y, x3.y, acc1);
      if (ch + 1 == slice_end) {  // warp-uniform: the slice is complete
        if (prow >= 0 && on) {
          if (pcon) { pa0 = 0.0; pa1 = 0.0; }
          *reinterpret_cast<double2*>(Q + (int64_t)prow * ks + 2 * l) = make_double2(pa0, pa1);
          dot0 = fma(pa0, pp.x, dot0);
          dot1 = fma(pa1, pp.y, dot1);
        }
        prow = row; pa0 = acc0; pa1 = acc1;
        if (row >= 0) { pp = P2[(int64_t)row * 4]; pcon = constrained[row]; }
        acc0 = 0.0; acc1 = 0.0;
        s++;
        if (s < s_end) { slice_end = sptr[s + 1]; row = srow[s * 8 + g]; }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (prow >= 0 && on) {
      if (pcon) { pa0 = 0.0; pa1 = 0.0; }
      *reinterpret_cast<double2*>(Q + (int64_t)prow * ks + 2 * l) = make_double2(pa0, pa1);
      dot0 = fma(pa0, pp.x, dot0);
      dot1 = fma(pa1, pp.y, dot1);
    }
  }
  sh[0][threadIdx.x] = dot0;
  sh[1][threadIdx.x] = dot1;
  __syncthreads();
  if (threadIdx.x < 4 && 2 * threadIdx.x < ks) {
    double t0 = 0.0, t1 = 0.0;
    for (int i = threadIdx.x; i < 256; i += 4) { t0 += sh[0][i]; t1 += sh[1][i]; }
    partial[(int64_t)blockIdx.x * KMAX + 2 * threadIdx.x] = t0;
    partial[(int64_t)blockIdx.x * KMAX + 2 * threadIdx.x + 1] = t1;
  }
}

}  // namespace

// min (lohi[0..2]) and max (lohi[3..5]) of the vertex coordinates, computed once per mesh
const double* mesh_bbox(Ctx* c) {
  if (!c->have_bbox) {
    const int nb = c->num_sms * 4;
    c->bbox.ensure((size_t)(nb + 1) * 6, c->stream);
    LAUNCH(c, k_bbox_stage, nb, TB, 0, c->xyz.p, c->nv, c->dim, 0, c->bbox.p + 6);
    LAUNCH(c, k_bbox_stage, 1, TB, 0, c->bbox.p + 6, (int64_t)nb, c->dim, 1, c->bbox.p);
    c->have_bbox = true;
  }
  return c->bbox.p;
}

int sell_pstride(int ks) {
  int p = 2;
  while (p < ks) p <<= 1;
  return p;
}

void sell_build(Ctx* c) {
  cudaStream_t st = c->stream;
  const int64_t n = c->ndof;
  const int64_t nslices = (n + 7) / 8, nslots = nslices * 8;
  size_t bytes = 0;
  uint32_t* key = scratch<uint32_t>(c, 0, n);
  uint32_t* keys = scratch<uint32_t>(c, 1, n);
  int32_t* idx = scratch<int32_t>(c, 2, n);
  c->sell_row.ensure(nslots, st);
  // rows in spatial cluster order (Morton code of the dof location): the rows a CTA works on at any time share most of
  // their columns, so the gathered rows of P are served by the SM's L1 instead of being re-fetched from L2
  const int32_t* order0 = nullptr;
  if (sell_env("REMO_SELL_ORDER", 1)) {
    const double* lohi = mesh_bbox(c);
    uint64_t* code = scratch<uint64_t>(c, 6, n);
    uint64_t* codes = scratch<uint64_t>(c, 7, n);
    int32_t* ord = scratch<int32_t>(c, 8, n);
    LAUNCH(c, k_sell_morton, grid_for(n, TB), TB, 0, make_view(c), c->xyz.p, lohi, code, idx);
    CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, code, codes, idx, ord, n, 0, 63, st));
    c->tmp.ensure(bytes, st);
    CK(cub::DeviceRadixSort::SortPairs(c->tmp.p, bytes, code, codes, idx, ord, n, 0, 63, st));
    c->launches += 4;
    order0 = ord;
  }
  LAUNCH(c, k_sell_keys, grid_for(n, TB), TB, 0, c->rowptr.p, order0, n, sigma_rows(), key, idx);
  CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, key, keys, idx, c->sell_row.p, n, 0, 32, st));
  c->tmp.ensure(bytes, st);
  CK(cub::DeviceRadixSort::SortPairs(c->tmp.p, bytes, key, keys, idx, c->sell_row.p, n, 0, 32, st));
  c->launches += 4;
  int32_t* nch = scratch<int32_t>(c, 3, nslices);
  int64_t* nch64 = scratch<int64_t>(c, 4, nslices);
  int64_t* incl = scratch<int64_t>(c, 5, nslices);
  LAUNCH(c, k_sell_chunks, grid_for(nslices, TB), TB, 0, c->rowptr.p, c->sell_row.p, n, nslices, nch);
  LAUNCH(c, k_i32_to_i64, grid_for(nslices, TB), TB, 0, nch, nslices, nch64);
  CK(cub::DeviceScan::InclusiveSum(nullptr, bytes, nch64, incl, nslices, st));
  c->tmp.ensure(bytes, st);
  CK(cub::DeviceScan::InclusiveSum(c->tmp.p, bytes, nch64, incl, nslices, st));
  c->launches += 2;
  c->sell_ptr.ensure(nslices + 1, st);
  LAUNCH(c, k_sell_ptr, grid_for(nslices + 1, TB), TB, 0, nch, incl, nslices, c->sell_ptr.p);
  int64_t total = 0;
  CK(cudaMemcpyAsync(&total, incl + (nslices - 1), sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  c->sell_chunks = total;
  c->sell_slots = nslots;
  c->sell_col.ensure((size_t)total * 32, st);
  c->sell_val.ensure((size_t)total * 32, st);
  LAUNCH(c, k_sell_fill, c->num_sms * 8, TB, 0, c->rowptr.p, c->col.p, c->val.p, c->sell_row.p, c->sell_ptr.p, nslots, c->sell_col.p, c->sell_val.p);
  {
    // blocked distribution: `waves` CTAs per resident CTA slot; 0 = grid-strided rows
    const int waves = sell_env("REMO_SELL_WAVES", 4);
    c->sell_nparts = waves > 0 ? std::min(c->num_sms * 4 * waves, c->num_sms * 64) : 0;
    if (c->sell_nparts > 0) {
      c->sell_part.ensure(c->sell_nparts + 1, st);
      LAUNCH(c, k_sell_partition, grid_for(c->sell_nparts + 1, TB), TB, 0, c->sell_ptr.p, nslices, c->sell_nparts, c->sell_part.p);
    }
  }
  {
    // streaming kernel (P stride 8): one contiguous range of slices per WARP, equal chunk counts.  Ranges of ~14 slices:
    // long enough to amortise the pipeline start of a warp, short enough that the CTAs in flight (dispatched in row
    // order) sweep the matrix as a narrow front, so the gathered rows of P stay L2-resident even when P (308 MB at
    // 4.8 M dofs) is larger than the L2 -- measured 1.01 ms with 42 slices per warp, 0.84 ms with 11..16.
    const int spw = std::max(1, sell_env("REMO_SELL_SPW", 14));
    const int64_t want = (nslices + 8 * spw - 1) / (8 * spw);
    c->sell_sgrid = (int)std::min<int64_t>(std::max<int64_t>(want, c->num_sms * 4), (int64_t)c->num_sms * 192);
    const int nw = c->sell_sgrid * 8;
    c->sell_wpart.ensure(nw + 1, st);
    LAUNCH(c, k_sell_partition, grid_for(nw + 1, TB), TB, 0, c->sell_ptr.p, nslices, nw, c->sell_wpart.p);
  }
  c->have_sell = true;
}

static bool sell_stream(const Ctx* c) { return c->pstride == 8 && sell_env("REMO_SELL_STREAM", 1) != 0; }

int sell_grid(const Ctx* c) {
  if (sell_stream(c)) return c->sell_sgrid;
  return c->sell_nparts > 0 ? c->sell_nparts : c->num_sms * 8;
}

void launch_spmm_sell(Ctx* c, const double* P, double* Q, int ks, int pstride) {
  cudaStream_t st = c->stream;
  auto* sp = c->sell_ptr.p; auto* sc = c->sell_col.p; auto* sv = c->sell_val.p; auto* sr = c->sell_row.p;
  auto* cs = c->constrained.p;
  double* pt = c->partial.p;
  const int64_t ns = c->sell_slots;
  const int64_t* part = c->sell_nparts > 0 ? c->sell_part.p : nullptr;
  const int grid = sell_grid(c);
  if (sell_stream(c)) {
    // ring depth 4: three chunks ahead cover the DRAM latency of the matrix stream (depth 8 measured equal); 64 registers,
    // 4 CTAs per SM -- tighter register caps spill and run 1.5x slower, a software-pipelined variant with 8 gathers in
    // flight per lane needs 80-94 registers and is 1.3-1.8x slower (fewer warps): profiles/r01_notes.md
    k_spmm_stream8<4, 4><<<grid, 256, 0, st>>>(sp, sc, sv, sr, cs, P, Q, ks, pt, c->sell_wpart.p);
    c->launches++;
    CK(cudaGetLastError());
    return;
  }
#define SELL_CASE(KS_) \
  case KS_: k_spmm_sell<KS_, 256, 4, 1><<<grid, 256, 0, st>>>(sp, sc, sv, sr, cs, P, Q, ks, ns, pt, part); break;
  switch (pstride) {
    SELL_CASE(2) SELL_CASE(4) SELL_CASE(8) SELL_CASE(16) SELL_CASE(32)
    default: FAIL(REMO_ERR_ARG, "launch_spmm_sell: unsupported stride %d", pstride);
  }
#undef SELL_CASE
  c->launches++;
  CK(cudaGetLastError());
}
