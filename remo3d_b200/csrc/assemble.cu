// Numeric assembly of the H1 stiffness matrix into the precomputed CSR pattern.
// Replaces `a += grad(u)*grad(v)*sigma*dx` / `2*pi*grad(u)*grad(v)*x*sigma*dx` + `a.Assemble()`
// (/root/reference/remo3d/ngsolve_functions.py:31-36, 47).
//
// Two kernels:
//   k_geom_tet / k_geom_tri : one thread per element, coalesced loads of the 4 (3) sorted vertex ids and
//       their coordinates -> the metric  gm[t][m] = sigma |K| grad l_i . grad l_j  (10 / 6 numbers per element)
//   k_assemble_rows<NLD, LANES> : GATHER form.  One LANES-wide sub-warp owns one matrix row (= one dof).  It
//       walks the elements adjacent to that dof in ascending element order (adjacency from symbolic.cu); lane b
//       evaluates entry (a,b) of the element matrix  K_ab = sum_m gm[t][m] T[m][a][b]  from the exact reference
//       tensors staged in shared memory, finds its column by binary search in the row's sorted column list and
//       accumulates in a shared-memory row buffer.  No atomics, fixed summation order -> bit-reproducible.
#include "space_view.cuh"

namespace {
#include "ref_tensors.inc"

constexpr int TB = 256;

__global__ void k_geom_tet(const int32_t* __restrict__ sv, const double* __restrict__ xyz, const int32_t* __restrict__ mat,
                           const double* __restrict__ sigma, int nmat, double* __restrict__ gm, int64_t nt,
                           int* __restrict__ bad) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nt) return;
  const int4 v = reinterpret_cast<const int4*>(sv)[t];
  const int32_t vv[4] = {v.x, v.y, v.z, v.w};
  double x[4][3];
#pragma unroll
  for (int i = 0; i < 4; i++) {
#pragma unroll
    for (int d = 0; d < 3; d++) x[i][d] = xyz[3 * (int64_t)vv[i] + d];
  }
  double e1[3], e2[3], e3[3];
#pragma unroll
  for (int d = 0; d < 3; d++) { e1[d] = x[1][d] - x[0][d]; e2[d] = x[2][d] - x[0][d]; e3[d] = x[3][d] - x[0][d]; }
  double g[4][3];
  // cofactors: grad l_1 = (e2 x e3)/det, grad l_2 = (e3 x e1)/det, grad l_3 = (e1 x e2)/det
  g[1][0] = e2[1] * e3[2] - e2[2] * e3[1]; g[1][1] = e2[2] * e3[0] - e2[0] * e3[2]; g[1][2] = e2[0] * e3[1] - e2[1] * e3[0];
  g[2][0] = e3[1] * e1[2] - e3[2] * e1[1]; g[2][1] = e3[2] * e1[0] - e3[0] * e1[2]; g[2][2] = e3[0] * e1[1] - e3[1] * e1[0];
  g[3][0] = e1[1] * e2[2] - e1[2] * e2[1]; g[3][1] = e1[2] * e2[0] - e1[0] * e2[2]; g[3][2] = e1[0] * e2[1] - e1[1] * e2[0];
  const double det = e1[0] * g[1][0] + e1[1] * g[1][1] + e1[2] * g[1][2];
  const int m_id = mat[t];
  if (m_id < 0 || m_id >= nmat) { atomicExch(bad, 1); return; }
  if (det == 0.0) atomicExch(bad, 2);
  const double inv = 1.0 / det;
#pragma unroll
  for (int i = 1; i < 4; i++)
#pragma unroll
    for (int d = 0; d < 3; d++) g[i][d] *= inv;
#pragma unroll
  for (int d = 0; d < 3; d++) g[0][d] = -(g[1][d] + g[2][d] + g[3][d]);
  const double w = sigma[m_id] * fabs(det) * (1.0 / 6.0);
  int m = 0;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = i; j < 4; j++) {
      gm[t * 10 + m] = w * (g[i][0] * g[j][0] + g[i][1] * g[j][1] + g[i][2] * g[j][2]);
      m++;
    }
}

// 2D axisymmetric: gm[t][k*6+m] = 2 pi sigma |K| r_k grad l_i . grad l_j   (18 numbers per triangle)
__global__ void k_geom_tri(const int32_t* __restrict__ sv, const double* __restrict__ xy, const int32_t* __restrict__ mat,
                           const double* __restrict__ sigma, int nmat, double* __restrict__ gm, int64_t nt,
                           int* __restrict__ bad) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nt) return;
  double x[3][2];
  for (int i = 0; i < 3; i++) {
    int64_t v = sv[3 * t + i];
    x[i][0] = xy[2 * v];
    x[i][1] = xy[2 * v + 1];
  }
  const double e1[2] = {x[1][0] - x[0][0], x[1][1] - x[0][1]}, e2[2] = {x[2][0] - x[0][0], x[2][1] - x[0][1]};
  const double det = e1[0] * e2[1] - e1[1] * e2[0];
  const int m_id = mat[t];
  if (m_id < 0 || m_id >= nmat) { atomicExch(bad, 1); return; }
  if (det == 0.0) atomicExch(bad, 2);
  const double inv = 1.0 / det;
  double g[3][2];
  g[1][0] = e2[1] * inv; g[1][1] = -e2[0] * inv;
  g[2][0] = -e1[1] * inv; g[2][1] = e1[0] * inv;
  g[0][0] = -(g[1][0] + g[2][0]); g[0][1] = -(g[1][1] + g[2][1]);
  const double w = 2.0 * 3.14159265358979323846 * sigma[m_id] * fabs(det) * 0.5;
  for (int k = 0; k < 3; k++) {
    int m = 0;
    for (int i = 0; i < 3; i++)
      for (int j = i; j < 3; j++) {
        gm[t * 18 + k * 6 + m] = w * x[k][0] * (g[i][0] * g[j][0] + g[i][1] * g[j][1]);
        m++;
      }
  }
}

// NM = number of metric numbers per element (10 in 3D, 18 in 2D); tensor layout T[m][a][b]
template <int NLD, int LANES, int NM>
__global__ void __launch_bounds__(TB) k_assemble_rows(SpaceView s, const double* __restrict__ tensors,
                                                      const double* __restrict__ gm, const int64_t* __restrict__ adj_ptr,
                                                      const uint32_t* __restrict__ adj, const int64_t* __restrict__ rowptr,
                                                      const int32_t* __restrict__ col, double* __restrict__ val,
                                                      int maxrow, int64_t nrows, int ncl) {
  extern __shared__ double smem[];
  double* T = smem;                                 // NM*NLD*NLD
  double* rowbuf = smem + NM * NLD * NLD;           // (TB/LANES) x maxrow
  for (int i = threadIdx.x; i < NM * NLD * NLD; i += blockDim.x) T[i] = tensors[i];
  __syncthreads();
  const int sub = threadIdx.x / LANES, lane = threadIdx.x % LANES;
  const int subs_per_block = TB / LANES;
  double* acc = rowbuf + (size_t)sub * maxrow;
  // sub-warps of one warp work on different rows with different trip counts: synchronise the sub-warp only
  const unsigned full = (LANES == 32) ? 0xffffffffu : (((1u << (LANES & 31)) - 1u) << (((threadIdx.x & 31) / LANES) * LANES));
  // nrows = ndof, ncl = NLD: the whole matrix.  nrows = nv, ncl = vertices per element: its leading vertex block alone
  // (same element order, same arithmetic: bit-identical to the corresponding entries of the whole matrix)
  for (int64_t row = (int64_t)blockIdx.x * subs_per_block + sub; row < nrows; row += (int64_t)gridDim.x * subs_per_block) {
    const int64_t rs = rowptr[row];
    const int len = (int)(rowptr[row + 1] - rs);
    const int32_t* rc = col + rs;
    for (int j = lane; j < len; j += LANES) acc[j] = 0.0;
    __syncwarp(full);
    const int64_t a0 = adj_ptr[row], a1 = adj_ptr[row + 1];
    for (int64_t ai = a0; ai < a1; ai++) {
      const uint32_t pay = adj[ai];
      const int64_t t = pay / NLD;
      const int a = (int)(pay - t * NLD);
      if (lane < ncl) {
        const double* g = gm + t * NM;
        double k = 0.0;
#pragma unroll
        for (int m = 0; m < NM; m++) k = fma(g[m], T[(m * NLD + a) * NLD + lane], k);
        const int32_t cd = (int32_t)elem_dof(s, t, lane);
        int lo = 0, hi = len;
        while (lo < hi) {
          int mid = (lo + hi) >> 1;
          if (rc[mid] < cd) lo = mid + 1; else hi = mid;
        }
        acc[lo] += k;  // distinct lanes of one element hit distinct columns
      }
      __syncwarp(full);
    }
    for (int j = lane; j < len; j += LANES) val[rs + j] = acc[j];
    __syncwarp(full);
  }
}

__global__ void k_max_rowlen(const int64_t* __restrict__ rowptr, int64_t n, int* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int len = (i < n) ? (int)(rowptr[i + 1] - rowptr[i]) : 0;
  for (int o = 16; o; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
  if ((threadIdx.x & 31) == 0 && len > 0) atomicMax(out, len);
}

// diag(A) straight from the element metrics: one thread per dof walks its adjacent elements in ascending order (the order
// and the arithmetic of k_assemble_rows: bit-identical to the diagonal of the assembled matrix).  dinv = 1 / a_ii on free
// dofs, 0 on constrained ones.
template <int NLD, int NM>
__global__ void __launch_bounds__(TB) k_diag_rows(const double* __restrict__ tensors, const double* __restrict__ gm,
                                                  const int64_t* __restrict__ adj_ptr, const uint32_t* __restrict__ adj,
                                                  const uint8_t* __restrict__ constrained, int64_t ndof, double* __restrict__ dinv) {
  __shared__ double Td[NM * NLD];  // T[m][a][a]
  for (int i = threadIdx.x; i < NM * NLD; i += blockDim.x) {
    const int m = i / NLD, a = i - m * NLD;
    Td[i] = tensors[(m * NLD + a) * NLD + a];
  }
  __syncthreads();
  const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= ndof) return;
  double d = 0.0;
  if (!constrained[row]) {
    for (int64_t ai = adj_ptr[row]; ai < adj_ptr[row + 1]; ai++) {
      const uint32_t pay = adj[ai];
      const int64_t t = pay / NLD;
      const int a = (int)(pay - t * NLD);
      const double* g = gm + t * NM;
      double k = 0.0;
#pragma unroll
      for (int m = 0; m < NM; m++) k = fma(g[m], Td[m * NLD + a], k);
      d += k;
    }
  }
  dinv[row] = d > 0.0 ? 1.0 / d : 0.0;
}

template <int NLD, int LANES, int NM>
void launch_rows(Ctx* c, const double* tensors_dev, int maxrow, const int64_t* rowptr, const int32_t* col, double* val,
                 int64_t nrows, int ncl) {
  const size_t smem = (size_t)(NM * NLD * NLD + (TB / LANES) * maxrow) * sizeof(double);
  auto kern = k_assemble_rows<NLD, LANES, NM>;
  const int dev_max = allow_max_smem(kern, c->device);
  if (smem > (size_t)dev_max || smem > 200 * 1024) FAIL(REMO_ERR_MESH, "remo_assemble: a matrix row has %d entries, too many for the row buffer", maxrow);
  int per_sm = 1;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TB, smem));
  if (per_sm < 1) per_sm = 1;
  const int64_t want = (nrows + (TB / LANES) - 1) / (TB / LANES);
  const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)c->num_sms * per_sm);
  SpaceView sview = make_view(c);
  kern<<<grid, TB, smem, c->stream>>>(sview, tensors_dev, c->gm.p, c->adj_ptr.p, c->adj.p, rowptr, col, val, maxrow, nrows, ncl);
  c->launches++;
  CK(cudaGetLastError());
}

struct TensorCache {
  double* dev[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
};
TensorCache g_tensors[16];  // per device

const double* tensors_for(Ctx* c) {
  TensorCache& tc = g_tensors[c->device & 15];
  const int di = c->dim - 2, pi = c->order - 1;
  if (tc.dev[di][pi]) return tc.dev[di][pi];
  const double* src = nullptr;
  size_t n = 0;
  if (c->dim == 3) {
    if (pi == 0) { src = REF_T3_P1; n = sizeof(REF_T3_P1); }
    if (pi == 1) { src = REF_T3_P2; n = sizeof(REF_T3_P2); }
    if (pi == 2) { src = REF_T3_P3; n = sizeof(REF_T3_P3); }
  } else {
    if (pi == 0) { src = REF_T2_P1; n = sizeof(REF_T2_P1); }
    if (pi == 1) { src = REF_T2_P2; n = sizeof(REF_T2_P2); }
    if (pi == 2) { src = REF_T2_P3; n = sizeof(REF_T2_P3); }
  }
  double* d = nullptr;
  CK(cudaMalloc((void**)&d, n));
  CK(cudaMemcpyAsync(d, src, n, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  tc.dev[di][pi] = d;
  return d;
}

void geometry(Ctx* c) {
  cudaStream_t st = c->stream;
  const int nm = (c->dim == 3) ? 10 : 18;
  c->gm.ensure(c->nt * nm, st);
  DBuf<int> bad;
  bad.ensure(2, st);
  CK(cudaMemsetAsync(bad.p, 0, 2 * sizeof(int), st));
  if (c->dim == 3)
    LAUNCH(c, k_geom_tet, grid_for(c->nt, TB), TB, 0, c->sv.p, c->xyz.p, c->mat.p, c->sigma.p, (int)c->sigma.n, c->gm.p, c->nt, bad.p);
  else
    LAUNCH(c, k_geom_tri, grid_for(c->nt, TB), TB, 0, c->sv.p, c->xyz.p, c->mat.p, c->sigma.p, (int)c->sigma.n, c->gm.p, c->nt, bad.p);
  int hb = 0;
  CK(cudaMemcpyAsync(&hb, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  bad.release(st);
  if (hb == 1) FAIL(REMO_ERR_ARG, "remo_assemble: a material index is outside the sigma list (nmat=%d)", (int)c->sigma.n);
  if (hb == 2) FAIL(REMO_ERR_MESH, "remo_assemble: degenerate (zero volume) element");
}

// row-gather assembly into a CSR pattern: the whole matrix (nrows = ndof, ncl = nld) or its vertex block (nrows = nv, ncl = dim + 1)
void assemble_rows(Ctx* c, const int64_t* rowptr, const int32_t* col, double* val, int64_t nrows, int ncl) {
  cudaStream_t st = c->stream;
  int* mx = scratch<int>(c, 8, 4);
  CK(cudaMemsetAsync(mx, 0, sizeof(int), st));
  LAUNCH(c, k_max_rowlen, grid_for(nrows, TB), TB, 0, rowptr, nrows, mx);
  int hm = 0;
  CK(cudaMemcpyAsync(&hm, mx, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const int maxrow = (hm + 7) & ~7;
  const double* T = tensors_for(c);
  if (c->dim == 3) {
    if (c->order == 1) launch_rows<4, 4, 10>(c, T, maxrow, rowptr, col, val, nrows, ncl);
    if (c->order == 2) launch_rows<10, 16, 10>(c, T, maxrow, rowptr, col, val, nrows, ncl);
    if (c->order == 3) launch_rows<20, 32, 10>(c, T, maxrow, rowptr, col, val, nrows, ncl);
  } else {
    if (c->order == 1) launch_rows<3, 4, 18>(c, T, maxrow, rowptr, col, val, nrows, ncl);
    if (c->order == 2) launch_rows<6, 8, 18>(c, T, maxrow, rowptr, col, val, nrows, ncl);
    if (c->order == 3) launch_rows<10, 16, 18>(c, T, maxrow, rowptr, col, val, nrows, ncl);
  }
}

}  // namespace

// geometry + row-gather kernels of the whole matrix (also timed alone by remo_kernel_time which=1)
void assemble_kernels_only(Ctx* c) {
  pattern_build(c);
  geometry(c);
  assemble_rows(c, c->rowptr.p, c->col.p, c->val.p, c->ndof, c->nld);
  c->have_values = true;
}

// The CSR values exist only when somebody needs them (remo_matrix_get, the SELL / CSR SpMM kernels): the element-wise PCG
// product, diag(A) and the vertex block of the V-cycle all come straight from the element metrics gm.
void ensure_values(Ctx* c) {
  if (!c->have_matrix) FAIL(REMO_ERR_STATE, "matrix not assembled (call remo_assemble first)");
  if (c->have_values) return;
  pattern_build(c);
  assemble_rows(c, c->rowptr.p, c->col.p, c->val.p, c->ndof, c->nld);
  c->have_values = true;
}

// dinv = 1 / diag(A) on free dofs, 0 on constrained ones, from the element metrics
void diag_from_elements(Ctx* c, double* dinv) {
  const double* T = tensors_for(c);
  const unsigned g = grid_for(c->ndof, TB);
#define DIAG(NLD_, NM_) LAUNCH(c, (k_diag_rows<NLD_, NM_>), g, TB, 0, T, c->gm.p, c->adj_ptr.p, c->adj.p, c->constrained.p, c->ndof, dinv)
  if (c->dim == 3) {
    if (c->order == 1) DIAG(4, 10);
    if (c->order == 2) DIAG(10, 10);
    if (c->order == 3) DIAG(20, 10);
  } else {
    if (c->order == 1) DIAG(3, 18);
    if (c->order == 2) DIAG(6, 18);
    if (c->order == 3) DIAG(10, 18);
  }
#undef DIAG
}

// values of the vertex block (P1 stiffness matrix in the hierarchical basis) into the pattern (rowptr, col) of vertex_block_pattern
void assemble_vertex_block(Ctx* c, const int64_t* rowptr, const int32_t* col, double* val) {
  assemble_rows(c, rowptr, col, val, c->nv, c->dim + 1);
}

void assemble(Ctx* c, int nmat, const double* sigma) {
  if (!c->have_space) FAIL(REMO_ERR_STATE, "remo_assemble: no space (call remo_space_build first)");
  if (nmat < 1 || !sigma) FAIL(REMO_ERR_ARG, "remo_assemble: need at least one material conductivity");
  StageTimer timer(c, ST_ASM);
  c->sigma.ensure(nmat, c->stream);
  CK(cudaMemcpyAsync(c->sigma.p, sigma, nmat * sizeof(double), cudaMemcpyDefault, c->stream));
  c->have_matrix = false;
  c->have_values = false;
  geometry(c);
  c->have_matrix = true;
  if (!c->lazy_matrix) ensure_values(c);
  c->have_sell = false;
  c->have_ebe = false;
  c->pkind = -1;
  c->have_solution = false;
}
