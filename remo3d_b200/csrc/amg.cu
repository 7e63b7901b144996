// "multigrid" preconditioner (REMO_PRECOND_MULTIGRID): two-level hierarchical-basis splitting with an
// aggregation-AMG V-cycle on the P1 block.
//
// Reference behaviour being replaced: `c = ngs.Preconditioner(a, "multigrid")`
// (/root/reference/remo3d/ngsolve_functions.py:46, default of remo3d.py:82).  On a single-level mesh NGSolve's
// multigrid is a two-level method: exact solve of the lowest-order (P1) problem + smoothing of the high-order
// dofs [NGS, SURVEY 8a-a4].  Here:
//   * the dofs are [vertices | edges | faces] in a HIERARCHICAL basis, so the P1 stiffness matrix is literally the
//     leading nv x nv block of A (no second assembly, prolongation = injection);
//   * z = M r :  z_high = D^-1 r_high  (Jacobi on the edge / face dofs, whose block is well conditioned), and
//                z_vert = V-cycle(A_vv) r_vert;
//   * the V-cycle hierarchy is built on the GPU by STRENGTH-BASED PAIRWISE aggregation: per level `amg_passes` (3)
//     passes of "every unmatched row points at its strongest unmatched neighbour (strength -a_ij / sqrt(a_ii a_jj),
//     ties to the smaller index); mutual pointers become a pair" (`amg_rounds` handshake rounds per pass, rows left over
//     stay single), each pass followed by the Galerkin product of the pair map (sort + reduce-by-key), so an aggregate
//     has at most 2^passes rows and follows the strong couplings -- across the 1:100 conductivity jumps and along the
//     anisotropy of the graded mesh -- instead of cutting them the way the geometric (Morton-rank) aggregates of round 1
//     did: 155 -> 119 PCG iterations on the 200 k-dof bench mesh, 105 with an exact P1 solve
//     (tools/precond_study/amg_variants.py).  `remo_set_option("amg_agg", 0)` restores the Morton aggregates.
//     Prolongation piecewise constant, the coarsest level (<= 256 rows) is inverted densely.  Smoother: l1-Jacobi
//     (always a contraction), symmetric cycle, coarse correction over-weighted by 1.5 (plain aggregation
//     under-corrects) -> M is SPD and plain PCG applies.
//   * round 2: the cycle itself runs in fp32 (values, l1 weights and work blocks of every level; level 0 reads the PCG's fp64
//     residual and writes its fp64 z inside the sweeps) around the fp64 PCG -- remo_set_option("amg_fp32", 0) restores fp64.
// Everything works on row-major n x nrhs blocks, like the PCG.
#include <cooperative_groups.h>
#include <cub/cub.cuh>

#include <cstdlib>
#include <cstring>

#include "space_view.cuh"

namespace {

constexpr int TB = 256;
constexpr int AGG = 8;
constexpr int COARSEST = 512;  // dense inverse below this many rows (one launch per pivot step, k_gj_step)
constexpr int MAXLEV = 12;

double env_d(const char* name, double def) {
  const char* e = getenv(name);
  return e ? atof(e) : def;
}

// smoother diagonal: l1-Jacobi, dinv_i = 1 / sum_j |a_ij|.  D_l1 - A is diagonally dominant with a non-negative
// diagonal, hence positive semi-definite: the sweep x += dinv (b - A x) is a contraction in the A-norm for any SPD A
// (no global eigenvalue estimate needed, robust on sliver elements), and the symmetric V-cycle is SPD.
// dinv = 0 marks rows outside the coarse space (constrained vertices, empty aggregates).
__global__ void k_level_dinv(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ val,
                             const uint8_t* __restrict__ con, int64_t n, double* __restrict__ dinv, double* __restrict__ diag) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double d = 0.0, l1 = 0.0;
  if (!(con && con[i]))
    for (int64_t j = rowptr[i]; j < rowptr[i + 1]; j++) {
      const int32_t cj = col[j];
      if (con && con[cj] && cj != (int32_t)i) continue;  // couplings to constrained vertices (their x is always 0)
      l1 += fabs(val[j]);
      if (cj == (int32_t)i) d = val[j];
    }
  dinv[i] = d > 0.0 ? 1.0 / l1 : 0.0;
  if (diag) diag[i] = d > 0.0 ? d : 0.0;
}

// ---------------------------------------------------------------- pairwise aggregation (handshake matching)
// match[i]: -2 = not part of the coarse space (constrained / empty row), -1 = unmatched, >= 0 = partner row
__global__ void k_pair_init(const double* __restrict__ diag, int64_t n, int32_t* __restrict__ match) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) match[i] = diag[i] > 0.0 ? -1 : -2;
}

// 8 lanes per row: strongest unmatched neighbour (only negative off-diagonals couple; ties -> smaller column)
__global__ void __launch_bounds__(TB) k_pair_pick(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                  const double* __restrict__ val, const double* __restrict__ diag,
                                                  const int32_t* __restrict__ match, int64_t n, int32_t* __restrict__ pick) {
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3;
  const int l8 = threadIdx.x & 7;
  double bw = 0.0;
  int32_t bj = -1;
  const bool live = row < n && match[row] == -1;
  if (live) {
    const double di = diag[row];
    for (int64_t j = rowptr[row] + l8; j < rowptr[row + 1]; j += 8) {
      const int32_t cj = col[j];
      const double a = val[j];
      if (cj == (int32_t)row || !(a < 0.0) || match[cj] != -1) continue;
      const double w = -a / sqrt(di * diag[cj]);
      if (w > bw || (w == bw && bj >= 0 && cj < bj)) { bw = w; bj = cj; }
    }
  }
#pragma unroll
  for (int o = 4; o; o >>= 1) {
    const double ow = __shfl_xor_sync(0xffffffffu, bw, o);
    const int32_t oj = __shfl_xor_sync(0xffffffffu, bj, o);
    if (oj >= 0 && (ow > bw || (ow == bw && (bj < 0 || oj < bj)))) { bw = ow; bj = oj; }
  }
  if (row < n && l8 == 0) pick[row] = live ? bj : -1;
}

__global__ void k_pair_match(const int32_t* __restrict__ pick, int64_t n, int32_t* __restrict__ match) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n || match[i] != -1) return;
  const int32_t j = pick[i];
  if (j >= 0 && pick[j] == (int32_t)i) match[i] = j;
}

// leaders: single rows and the smaller row of every pair
__global__ void k_pair_flags(const int32_t* __restrict__ match, int64_t n, int32_t* __restrict__ flag) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t m = match[i];
  flag[i] = (m == -1 || (m >= 0 && (int32_t)i < m)) ? 1 : 0;
}

__global__ void k_pair_ids(const int32_t* __restrict__ match, const int32_t* __restrict__ incl, int64_t n, int32_t* __restrict__ id) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t m = match[i];
  id[i] = (m == -2) ? -1 : ((m == -1 || (int32_t)i < m) ? incl[i] - 1 : incl[m] - 1);
}

// fallback when a pass hardly coarsens: groups of `g` consecutive live rows
__global__ void k_live_flags(const double* __restrict__ diag, int64_t n, int32_t* __restrict__ flag) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) flag[i] = diag[i] > 0.0 ? 1 : 0;
}
__global__ void k_block_ids(const int32_t* __restrict__ flag, const int32_t* __restrict__ incl, int64_t n, int g, int32_t* __restrict__ id) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) id[i] = flag[i] ? (incl[i] - 1) / g : -1;
}

// tot[i] = id[tot[i]]  (composition of the pass maps; -1 stays -1)
__global__ void k_compose(int32_t* __restrict__ tot, const int32_t* __restrict__ id, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) { const int32_t t = tot[i]; tot[i] = t >= 0 ? id[t] : -1; }
}

__global__ void k_diag_of(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ val, int64_t n,
                          double* __restrict__ diag) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double d = 0.0;
  for (int64_t j = rowptr[i]; j < rowptr[i + 1]; j++)
    if (col[j] == (int32_t)i) d = val[j];
  diag[i] = d > 0.0 ? d : 0.0;
}

__global__ void k_agg_keys(const int32_t* __restrict__ agg, int64_t n, uint32_t* __restrict__ key, int32_t* __restrict__ row) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) { key[i] = (uint32_t)agg[i]; row[i] = (int32_t)i; }  // -1 -> 0xffffffff sorts behind every aggregate
}

__global__ void k_agg_ptr(const uint32_t* __restrict__ skey, int64_t n, int64_t nc, int32_t* __restrict__ ptr) {
  int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (a > nc) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (skey[mid] < (uint32_t)a) lo = mid + 1; else hi = mid;
  }
  ptr[a] = (int32_t)lo;
}

// ---------------------------------------------------------------- Morton ranks of the vertices
__device__ __forceinline__ uint64_t spread21(uint64_t v) {
  v &= 0x1fffffull;
  v = (v | (v << 32)) & 0x1f00000000ffffull;
  v = (v | (v << 16)) & 0x1f0000ff0000ffull;
  v = (v | (v << 8)) & 0x100f00f00f00f00full;
  v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}

__global__ void k_morton(const double* __restrict__ xyz, int64_t nv, int dim, const double* __restrict__ lohi,
                         uint64_t* __restrict__ code, uint32_t* __restrict__ idx) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nv) return;
  uint64_t c = 0;
  for (int d = 0; d < dim; d++) {
    const double ext = lohi[3 + d] - lohi[d];
    const double t = ext > 0 ? (xyz[i * dim + d] - lohi[d]) / ext : 0.0;
    const uint64_t q = (uint64_t)fmin(fmax(t * 2097151.0, 0.0), 2097151.0);
    c |= spread21(q) << d;
  }
  code[i] = c;
  idx[i] = (uint32_t)i;
}

// perm = vertices in Morton order -> aggregate of vertex perm[p] is p / AGG  (amg_agg = 0: the round-1 aggregates)
__global__ void k_agg_from_perm(const uint32_t* __restrict__ perm, const double* __restrict__ diag, int64_t n, int32_t* __restrict__ agg) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p < n) agg[perm[p]] = diag[perm[p]] > 0.0 ? (int32_t)(p / AGG) : -1;  // constrained vertices join no aggregate
}
__global__ void k_agg_consecutive(int64_t n, int32_t* __restrict__ agg) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) agg[i] = (int32_t)(i / AGG);
}

// ---------------------------------------------------------------- Galerkin coarse operator, piecewise-constant P
__global__ void k_galerkin_pairs(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ val,
                                 const int32_t* __restrict__ agg, const double* __restrict__ dinv, int64_t n, uint64_t nc,
                                 uint64_t* __restrict__ keys, double* __restrict__ vals) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t ai = agg[i];
  const bool dead_i = ai < 0 || dinv[i] == 0.0;  // constrained or empty row: not part of the coarse space
  for (int64_t j = rowptr[i]; j < rowptr[i + 1]; j++) {
    const int32_t c = col[j];
    const int32_t ac = agg[c];
    const bool dead = dead_i || ac < 0 || dinv[c] == 0.0;
    keys[j] = dead ? (nc << 32) : (((uint64_t)(uint32_t)ai << 32) | (uint32_t)ac);  // dropped entries sort behind every row
    vals[j] = dead ? 0.0 : val[j];
  }
}

__global__ void k_coarse_rowptr(const uint64_t* __restrict__ ukeys, int64_t nruns, int64_t nc, int64_t* __restrict__ rowptr) {
  int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (a > nc) return;
  rowptr[a] = lower_bound_u64(ukeys, nruns, (uint64_t)a << 32);
}

__global__ void k_coarse_cols(const uint64_t* __restrict__ ukeys, int64_t nnz, int32_t* __restrict__ col) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < nnz) col[i] = (int32_t)(ukeys[i] & 0xffffffffu);
}

// ---------------------------------------------------------------- V-cycle kernels (row-major n x k blocks)
// mode 0: OUT = X + omega * dinv * (B - A X)    (l1-Jacobi sweep)      mode 1: OUT = B - A X   (residual)
// Same lane layout as the PCG SpMM (solver.cu k_spmm_p): a G-lane group owns a row, every lane owns two adjacent
// right-hand sides (16-byte gathers; ks even) -- or one when ks == 1.
template <typename T> struct Pair2;
template <> struct Pair2<double> { using type = double2; };
template <> struct Pair2<float> { using type = float2; };

// T = storage type of the level (matrix values, dinv, X): double, or float for the mixed-precision cycle
// (remo_set_option("amg_fp32", 1), the default): TBt / TOt = types of B and OUT -- on level 0 those are the PCG's own fp64
// blocks R / Z, so the conversion happens inside the sweeps and costs no extra pass.  Sums are taken in T.
template <int G, int KP2, int W, typename T, typename TBt, typename TOt>
__global__ void __launch_bounds__(TB) k_smooth(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                               const T* __restrict__ val, const T* __restrict__ dinv,
                                               const TBt* __restrict__ B, const T* __restrict__ X,
                                               TOt* __restrict__ OUT, int ks, int64_t n, T omega, int mode) {
  using P2 = typename Pair2<T>::type;
  constexpr int J = G / KP2;
  constexpr int GROUPS = TB / G;
  const int lane = threadIdx.x & 31;
  const int gl = threadIdx.x % G, grp = threadIdx.x / G;
  const int jsub = gl / KP2, r2 = gl % KP2;
  const bool on = W * r2 < ks;
  const int cc = on ? W * r2 : 0;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << ((lane / G) * G));
  for (int64_t row = (int64_t)blockIdx.x * GROUPS + grp; row < n; row += (int64_t)gridDim.x * GROUPS) {
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    T acc0 = 0, acc1 = 0;
    for (int64_t base = s; base < e; base += G) {
      int32_t myc = (int32_t)row;
      T myv = 0;
      if (base + gl < e) { myc = col[base + gl]; myv = val[base + gl]; }
#pragma unroll
      for (int i = 0; i < KP2; i++) {
        const int j = i * J + jsub;
        const int32_t c = __shfl_sync(gmask, myc, j, G);
        const T v = __shfl_sync(gmask, myv, j, G);
        if (W == 2) {
          const P2 x = *reinterpret_cast<const P2*>(X + (int64_t)c * ks + cc);
          acc0 = fma(v, x.x, acc0);
          acc1 = fma(v, x.y, acc1);
        } else {
          acc0 = fma(v, X[(int64_t)c * ks + cc], acc0);
        }
      }
    }
#pragma unroll
    for (int o = G / 2; o >= KP2; o >>= 1) {
      acc0 += __shfl_xor_sync(gmask, acc0, o, G);
      if (W == 2) acc1 += __shfl_xor_sync(gmask, acc1, o, G);
    }
    if (jsub == 0 && on) {
      const int64_t idx = row * ks + cc;
      const T w = omega * dinv[row];
      const T res0 = (T)B[idx] - acc0;
      OUT[idx] = (TOt)((mode == 1) ? res0 : fma(w, res0, X[idx]));
      if (W == 2) {
        const T res1 = (T)B[idx + 1] - acc1;
        OUT[idx + 1] = (TOt)((mode == 1) ? res1 : fma(w, res1, X[idx + 1]));
      }
    }
  }
}

template <typename T, typename TBt>
__global__ void k_jacobi0(const T* __restrict__ dinv, const TBt* __restrict__ B, T* __restrict__ X, int k, int64_t n, T omega) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n * k) return;
  X[e] = omega * dinv[e / k] * (T)B[e];
}

// rows of an aggregate in ascending order (members / aggptr): fixed summation order -> bit-reproducible
template <typename T>
__global__ void k_restrict(const int32_t* __restrict__ aggptr, const int32_t* __restrict__ members, const T* __restrict__ R,
                           T* __restrict__ BC, int k, int64_t nc) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nc * k) return;
  const int64_t a = e / k;
  const int r = (int)(e - a * k);
  T s = 0;
  for (int32_t p = aggptr[a]; p < aggptr[a + 1]; p++) s += R[(int64_t)members[p] * k + r];
  BC[e] = s;
}

template <typename T>
__global__ void k_prolong(const int32_t* __restrict__ agg, const T* __restrict__ dinv, const T* __restrict__ XC,
                          T* __restrict__ X, int k, int64_t n, T alpha) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n * k) return;
  const int64_t i = e / k;
  const int r = (int)(e - i * k);
  const int32_t a = agg[i];
  if (a < 0 || dinv[i] == (T)0) return;  // constrained / empty rows stay zero
  X[e] = fma(alpha, XC[(int64_t)a * k + r], X[e]);
}

// x = Ainv b on the coarsest level (the inverse stays fp64; vectors in T): 8 lanes per output entry split the dot product
template <typename T>
__global__ void k_dense_apply_t(int n, const double* __restrict__ Ainv, const T* __restrict__ B, T* __restrict__ X, int k) {
  const int t = threadIdx.x + blockIdx.x * blockDim.x;
  const int e = t >> 3, part = t & 7;
  double s = 0.0;
  const bool ok = e < n * k;
  const int i = ok ? e / k : 0, r = ok ? e - i * k : 0;
  if (ok)
    for (int j = part; j < n; j += 8) s = fma(Ainv[(int64_t)i * n + j], (double)B[(int64_t)j * k + r], s);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  if (ok && part == 0) X[e] = (T)s;
}

__global__ void k_to_float(const double* __restrict__ in, float* __restrict__ out, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

// ---------------------------------------------------------------- coarsest level: dense inverse
__global__ void k_dense_fill(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ val,
                             int n, double* __restrict__ M) {  // dense n x n copy of the coarsest matrix
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool diag = false;
  for (int64_t j = rowptr[i]; j < rowptr[i + 1]; j++) {
    M[(int64_t)i * n + col[j]] = val[j];
    if (col[j] == i && val[j] > 0.0) diag = true;
  }
  if (!diag) M[(int64_t)i * n + i] = 1.0;  // empty aggregate: identity row keeps the matrix regular
}

// One pivot step of the in-place Gauss-Jordan inversion (no pivoting: the matrix is SPD), out-of-place between two
// buffers so that no thread reads what another one writes.  One launch per pivot: n launches of n^2 threads -- 1 ms at
// n = 400, where the single-CTA version of round 1 took 9.4 ms at n = 250 (ncu, profiles/r02_notes.md).
__global__ void k_gj_step(int n, int p, const double* __restrict__ in, double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * n) return;
  const int i = e / n, j = e - i * n;
  const double ip = 1.0 / in[(int64_t)p * n + p];
  double v;
  if (i == p) v = (j == p) ? ip : in[e] * ip;
  else if (j == p) v = -in[e] * ip;
  else v = fma(-in[(int64_t)i * n + p] * ip, in[(int64_t)p * n + j], in[e]);
  out[e] = v;
}

// ---------------------------------------------------------------- the small levels of the V-cycle in ONE launch
// Below ~16 k rows a level's kernels are pure launch latency (5 launches of 3-10 us per level and V-cycle, ~25 launches
// = ~120 us of a 1.3 ms PCG iteration at 4.8 M dofs).  k_vcycle_tail runs the whole sub-cycle of those levels -- first
// sweep, residual, restriction down to the dense coarsest solve, then prolongation and smoothing back up -- as one
// thread-block CLUSTER of 8 CTAs x 1024 threads: the phases are separated by the hardware cluster barrier
// (barrier.cluster arrive.release / wait.acquire orders the global-memory traffic between the CTAs), the data stays in
// L2.  Same arithmetic as the per-level kernels (k_jacobi0, k_smooth, k_restrict, k_prolong, k_dense_apply), and for the
// strides 6 and 8 (5..8 right-hand sides: k_smooth<4,4,2>) also the same summation order, i.e. bit-identical results
// (remo_set_option("amg_fused_tail", 0) selects the per-level launches).  Column pairs: ks even, ks <= 8.
template <typename T>
struct TailLevel {
  const int64_t* rowptr;
  const int32_t* col;
  const T* val;
  const T* dinv;
  const int32_t* agg;
  const int32_t* aggptr;
  const int32_t* members;
  const T* b;
  T* x;
  T* t;
  int64_t n;
  T omega;
};
template <typename T>
struct TailArgs {
  TailLevel<T> L[MAXLEV];
  int nl;               // levels handled here; the last one is the coarsest (dense inverse)
  const double* dense;  // inverse of the coarsest matrix (fp64 in both precisions)
  int k;                // row stride of the blocks (even, <= 8)
  int sweeps;
  T alpha;
};
constexpr int TAIL_CTAS = 8, TAIL_THREADS = 1024;

// OUT = X + omega dinv (B - A X) (mode 0) or B - A X (mode 1): the loop of k_smooth<4, 4, 2> over the cluster's threads
template <typename T>
__device__ __forceinline__ void tail_smooth(const TailLevel<T>& L, const T* __restrict__ X, T* __restrict__ OUT, int ks, int mode,
                                            int ctid, int nthreads) {
  using P2 = typename Pair2<T>::type;
  constexpr int G = 4;
  const int lane = ctid & 31;
  const int gl = ctid % G;
  const int r2 = gl;
  const bool on = 2 * r2 < ks;
  const int cc = on ? 2 * r2 : 0;
  const unsigned gmask = ((1u << G) - 1u) << ((lane / G) * G);
  for (int64_t row = ctid / G; row < L.n; row += nthreads / G) {
    const int64_t s = L.rowptr[row], e = L.rowptr[row + 1];
    T acc0 = 0, acc1 = 0;
    for (int64_t base = s; base < e; base += G) {
      int32_t myc = (int32_t)row;
      T myv = 0;
      if (base + gl < e) { myc = L.col[base + gl]; myv = L.val[base + gl]; }
#pragma unroll
      for (int i = 0; i < G; i++) {
        const int32_t c = __shfl_sync(gmask, myc, i, G);
        const T v = __shfl_sync(gmask, myv, i, G);
        const P2 x = *reinterpret_cast<const P2*>(X + (int64_t)c * ks + cc);
        acc0 = fma(v, x.x, acc0);
        acc1 = fma(v, x.y, acc1);
      }
    }
    if (on) {
      const int64_t idx = row * ks + cc;
      const T w = L.omega * L.dinv[row];
      const T res0 = L.b[idx] - acc0, res1 = L.b[idx + 1] - acc1;
      OUT[idx] = (mode == 1) ? res0 : fma(w, res0, X[idx]);
      OUT[idx + 1] = (mode == 1) ? res1 : fma(w, res1, X[idx + 1]);
    }
  }
}

template <typename T>
__global__ void __cluster_dims__(TAIL_CTAS, 1, 1) __launch_bounds__(TAIL_THREADS) k_vcycle_tail(const __grid_constant__ TailArgs<T> a) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int ctid = (int)cluster.block_rank() * TAIL_THREADS + threadIdx.x;
  const int nthreads = TAIL_CTAS * TAIL_THREADS;
  const int k = a.k, nl = a.nl;
  T* xs[MAXLEV];
  T* ts[MAXLEV];
#pragma unroll
  for (int l = 0; l < MAXLEV; l++) { xs[l] = a.L[l].x; ts[l] = a.L[l].t; }
  // ---- down
  for (int l = 0; l < nl - 1; l++) {
    const TailLevel<T>& L = a.L[l];
    for (int64_t e = ctid; e < L.n * k; e += nthreads) xs[l][e] = L.omega * L.dinv[e / k] * L.b[e];
    cluster.sync();
    for (int sw = 1; sw < a.sweeps; sw++) {
      tail_smooth<T>(L, xs[l], ts[l], k, 0, ctid, nthreads);
      T* tmp = xs[l]; xs[l] = ts[l]; ts[l] = tmp;
      cluster.sync();
    }
    tail_smooth<T>(L, xs[l], ts[l], k, 1, ctid, nthreads);  // t = b - A x
    cluster.sync();
    const TailLevel<T>& C = a.L[l + 1];
    T* bc = const_cast<T*>(C.b);
    for (int64_t e = ctid; e < C.n * k; e += nthreads) {
      const int64_t ag = e / k;
      const int r = (int)(e - ag * k);
      T sum = 0;
      for (int32_t p = L.aggptr[ag]; p < L.aggptr[ag + 1]; p++) sum += ts[l][(int64_t)L.members[p] * k + r];
      bc[e] = sum;
    }
    cluster.sync();
  }
  // ---- coarsest: x = Ainv b, 8 lanes per entry as k_dense_apply_t
  {
    const TailLevel<T>& L = a.L[nl - 1];
    const int n = (int)L.n;
    for (int base = 0; base < n * k * 8; base += nthreads) {
      const int t = base + ctid;
      const int e = t >> 3, part = t & 7;
      double sum = 0.0;
      const bool ok = e < n * k;
      const int i = ok ? e / k : 0, r = ok ? e - i * k : 0;
      if (ok)
        for (int j = part; j < n; j += 8) sum = fma(a.dense[(int64_t)i * n + j], (double)L.b[(int64_t)j * k + r], sum);
      sum += __shfl_xor_sync(0xffffffffu, sum, 4);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      if (ok && part == 0) xs[nl - 1][e] = (T)sum;
    }
    cluster.sync();
  }
  // ---- up
  for (int l = nl - 2; l >= 0; l--) {
    const TailLevel<T>& L = a.L[l];
    const T* xc = xs[l + 1];
    for (int64_t e = ctid; e < L.n * k; e += nthreads) {
      const int64_t i = e / k;
      const int r = (int)(e - i * k);
      const int32_t ag = L.agg[i];
      if (ag >= 0 && L.dinv[i] != (T)0) xs[l][e] = fma(a.alpha, xc[(int64_t)ag * k + r], xs[l][e]);
    }
    cluster.sync();
    for (int sw = 0; sw < a.sweeps; sw++) {
      tail_smooth<T>(L, xs[l], ts[l], k, 0, ctid, nthreads);
      T* tmp = xs[l]; xs[l] = ts[l]; ts[l] = tmp;
      cluster.sync();
    }
  }
}

int kp_of(int k) {
  int kp = 1;
  while (kp < k) kp <<= 1;
  return kp;
}

template <typename T, typename TBt, typename TOt>
void smooth_t(Ctx* c, const int64_t* rowptr, const int32_t* col, const T* val, const T* dinv, const TBt* B, const T* X, TOt* OUT, int k,
              int64_t n, double omega_d, int mode) {
  cudaStream_t st = c->stream;
  const T omega = (T)omega_d;
  auto grid_of = [&](int G) {
    const int64_t want = (n + (TB / G) - 1) / (TB / G);
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)c->num_sms * 8));
  };
  if (k & 1) {  // odd stride (only k == 1 in practice): one right-hand side per lane
    const int kp = kp_of(k);
    if (kp == 1) k_smooth<8, 1, 1, T, TBt, TOt><<<grid_of(8), TB, 0, st>>>(rowptr, col, val, dinv, B, X, OUT, k, n, omega, mode);
    else if (kp <= 8) k_smooth<8, 8, 1, T, TBt, TOt><<<grid_of(8), TB, 0, st>>>(rowptr, col, val, dinv, B, X, OUT, k, n, omega, mode);
    else k_smooth<32, 32, 1, T, TBt, TOt><<<grid_of(32), TB, 0, st>>>(rowptr, col, val, dinv, B, X, OUT, k, n, omega, mode);
  } else if (k <= 2) k_smooth<4, 1, 2, T, TBt, TOt><<<grid_of(4), TB, 0, st>>>(rowptr, col, val, dinv, B, X, OUT, k, n, omega, mode);
  else if (k <= 4) k_smooth<4, 2, 2, T, TBt, TOt><<<grid_of(4), TB, 0, st>>>(rowptr, col, val, dinv, B, X, OUT, k, n, omega, mode);
  else if (k <= 8) {
    // 8 lanes per row = two entries of the row in flight per step (rows of the vertex block have ~15 entries: 2 steps instead
    // of 4 dependent load -> gather rounds; the sweeps are latency-bound, profiles/r02_notes.md)
    if (c->amg_lanes8) k_smooth<8, 4, 2, T, TBt, TOt><<<grid_of(8), TB, 0, st>>>(rowptr, col, val, dinv, B, X, OUT, k, n, omega, mode);
    else k_smooth<4, 4, 2, T, TBt, TOt><<<grid_of(4), TB, 0, st>>>(rowptr, col, val, dinv, B, X, OUT, k, n, omega, mode);
  }
  else if (k <= 16) k_smooth<8, 8, 2, T, TBt, TOt><<<grid_of(8), TB, 0, st>>>(rowptr, col, val, dinv, B, X, OUT, k, n, omega, mode);
  else k_smooth<16, 16, 2, T, TBt, TOt><<<grid_of(16), TB, 0, st>>>(rowptr, col, val, dinv, B, X, OUT, k, n, omega, mode);
  c->launches++;
  CK(cudaGetLastError());
}

}  // namespace

void spmm_smooth(Ctx* c, const int64_t* rowptr, const int32_t* col, const double* val, const double* dinv, const double* B,
                 const double* X, double* OUT, int k, int64_t n, double omega, int mode) {
  smooth_t<double, double, double>(c, rowptr, col, val, dinv, B, X, OUT, k, n, omega, mode);
}

void amg_release(Ctx* c) {
  cudaStream_t s = c->stream;
  for (auto& L : c->amg) {
    L.rowptr.release(s); L.col.release(s); L.val.release(s); L.dinv.release(s); L.diag.release(s); L.agg.release(s);
    L.members.release(s); L.aggptr.release(s); L.b.release(s); L.x.release(s); L.t.release(s);
    L.valf.release(s); L.dinvf.release(s); L.bf.release(s); L.xf.release(s); L.tf.release(s);
  }
  c->amg.clear();
  for (auto& T : c->amg_tmp) { T.rowptr.release(s); T.col.release(s); T.val.release(s); T.diag.release(s); }
  for (auto& w : c->amg_w) w.release(s);
  c->amg_dense.release(s);
  c->amg_nrhs = 0;
  c->amg_nlev = 0;
}

namespace {

struct CsrRef {
  int64_t n, nnz;
  const int64_t* rowptr;
  const int32_t* col;
  const double* val;
  const double* diag;  // > 0 on the rows that belong to the coarse space
};

int bits_of(uint64_t v) {
  int b = 1;
  while (b < 64 && (v >> b)) b++;
  return b;
}

int32_t inclusive_scan_total(Ctx* c, const int32_t* in, int32_t* out, int64_t n) {
  size_t bytes = 0;
  cudaStream_t st = c->stream;
  CK(cub::DeviceScan::InclusiveSum(nullptr, bytes, in, out, n, st));
  c->tmp.ensure(bytes, st);
  CK(cub::DeviceScan::InclusiveSum(c->tmp.p, bytes, in, out, n, st));
  c->launches += 2;
  int32_t total = 0;
  CK(cudaMemcpyAsync(&total, out + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return total;
}

// Galerkin product P^T A P with the piecewise-constant prolongation of the row -> coarse row map `agg` (-1 = dropped):
// (coarse row, coarse column) keys of every entry, radix sort, reduce-by-key -> coarse CSR.  Returns its nnz.
int64_t galerkin(Ctx* c, const CsrRef& F, const int32_t* agg, int64_t nc, DBuf<int64_t>& rowptr, DBuf<int32_t>& col, DBuf<double>& val) {
  cudaStream_t st = c->stream;
  size_t bytes = 0;
  uint64_t* keys = scratch<uint64_t>(c, 0, F.nnz);
  uint64_t* keys2 = scratch<uint64_t>(c, 1, F.nnz);
  uint64_t* ukeys = scratch<uint64_t>(c, 6, F.nnz);
  double* vals = scratch<double>(c, 2, F.nnz);
  double* vals2 = scratch<double>(c, 3, F.nnz);
  double* uvals = scratch<double>(c, 7, F.nnz);
  int64_t* nruns = scratch<int64_t>(c, 9, 2);
  LAUNCH(c, k_galerkin_pairs, grid_for(F.n, TB), TB, 0, F.rowptr, F.col, F.val, agg, F.diag, F.n, (uint64_t)nc, keys, vals);
  const int end_bit = std::min(64, 32 + bits_of((uint64_t)nc));  // dropped entries carry the key nc << 32
  CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys, keys2, vals, vals2, F.nnz, 0, end_bit, st));
  c->tmp.ensure(bytes, st);
  CK(cub::DeviceRadixSort::SortPairs(c->tmp.p, bytes, keys, keys2, vals, vals2, F.nnz, 0, end_bit, st));
  CK(cub::DeviceReduce::ReduceByKey(nullptr, bytes, keys2, ukeys, vals2, uvals, nruns, cub::Sum(), F.nnz, st));
  c->tmp.ensure(bytes, st);
  CK(cub::DeviceReduce::ReduceByKey(c->tmp.p, bytes, keys2, ukeys, vals2, uvals, nruns, cub::Sum(), F.nnz, st));
  c->launches += 8;
  int64_t hr = 0;
  uint64_t lastkey = 0;
  CK(cudaMemcpyAsync(&hr, nruns, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (hr > 0) {
    CK(cudaMemcpyAsync(&lastkey, ukeys + (hr - 1), sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if ((lastkey >> 32) >= (uint64_t)nc) hr--;  // the run of dropped (constrained) entries
  }
  rowptr.ensure(nc + 1, st); col.ensure(std::max<int64_t>(hr, 1), st); val.ensure(std::max<int64_t>(hr, 1), st);
  LAUNCH(c, k_coarse_rowptr, grid_for(nc + 1, TB), TB, 0, ukeys, hr, nc, rowptr.p);
  if (hr) {
    LAUNCH(c, k_coarse_cols, grid_for(hr, TB), TB, 0, ukeys, hr, col.p);
    CK(cudaMemcpyAsync(val.p, uvals, hr * sizeof(double), cudaMemcpyDeviceToDevice, st));
  }
  return hr;
}

// Pairwise aggregation of level F: fills F.agg (row -> coarse row, -1 = not in the coarse space) and the coarse level's
// CSR (the Galerkin product of the last pass IS the coarse operator).  Returns the number of coarse rows.
int64_t aggregate_pairwise(Ctx* c, Ctx::AmgLevel& F, Ctx::AmgLevel& C) {
  cudaStream_t st = c->stream;
  CsrRef cur{F.n, F.nnz, F.rowptr.p, F.col.p, F.val.p, F.diag.p};
  const int passes = std::max(1, c->amg_passes), rounds = std::max(1, c->amg_rounds);
  int64_t nc = 0;
  for (int p = 0; p < passes; p++) {
    const int64_t n = cur.n;
    for (auto& w : c->amg_w) w.ensure(n, st);
    int32_t *match = c->amg_w[0].p, *pick = c->amg_w[1].p, *flag = c->amg_w[2].p, *incl = c->amg_w[3].p, *id = c->amg_w[4].p;
    LAUNCH(c, k_pair_init, grid_for(n, TB), TB, 0, cur.diag, n, match);
    for (int r = 0; r < rounds; r++) {
      LAUNCH(c, k_pair_pick, grid_for(n * 8, TB), TB, 0, cur.rowptr, cur.col, cur.val, cur.diag, match, n, pick);
      LAUNCH(c, k_pair_match, grid_for(n, TB), TB, 0, pick, n, match);
    }
    LAUNCH(c, k_pair_flags, grid_for(n, TB), TB, 0, match, n, flag);
    nc = inclusive_scan_total(c, flag, incl, n);
    if (nc > (int64_t)(0.7 * (double)n) && n > COARSEST) {
      // the strength graph hardly matched anything (small coarse levels are nearly dense, with few negative couplings
      // left): pairs of consecutive live rows instead -- the numbering keeps the locality of the level above
      LAUNCH(c, k_live_flags, grid_for(n, TB), TB, 0, cur.diag, n, flag);
      const int64_t live = inclusive_scan_total(c, flag, incl, n);
      nc = (live + 1) / 2;
      LAUNCH(c, k_block_ids, grid_for(n, TB), TB, 0, flag, incl, n, 2, id);
    } else {
      LAUNCH(c, k_pair_ids, grid_for(n, TB), TB, 0, match, incl, n, id);
    }
    if (nc < 1) FAIL(REMO_ERR_MESH, "amg_setup: no unconstrained vertex is left on level with %lld rows", (long long)n);
    if (p == 0) CK(cudaMemcpyAsync(F.agg.p, id, n * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    else LAUNCH(c, k_compose, grid_for(F.n, TB), TB, 0, F.agg.p, id, F.n);
    const bool last = (p == passes - 1) || nc <= COARSEST;
    if (last) {
      C.nnz = galerkin(c, cur, id, nc, C.rowptr, C.col, C.val);
      break;
    }
    Ctx::AmgTmp& T = c->amg_tmp[p & 1];
    T.n = nc;
    T.nnz = galerkin(c, cur, id, nc, T.rowptr, T.col, T.val);
    T.diag.ensure(nc, st);
    LAUNCH(c, k_diag_of, grid_for(nc, TB), TB, 0, T.rowptr.p, T.col.p, T.val.p, nc, T.diag.p);
    cur = CsrRef{T.n, T.nnz, T.rowptr.p, T.col.p, T.val.p, T.diag.p};
  }
  return nc;
}

}  // namespace

// Level 0 (rowptr / col / val / dinv / diag of c->amg[0], n = nv) must be in place: builds the coarser levels.
// The level objects (and all temporaries, scratch slots of the context) persist across meshes: setup only grows buffers.
void amg_build_hierarchy(Ctx* c) {
  cudaStream_t st = c->stream;
  size_t bytes = 0;
  int nlev = 1;
  // ---- coarser levels by Galerkin products until the level is small enough for a dense inverse
  while (c->amg[nlev - 1].n > COARSEST && nlev < MAXLEV) {
    if ((int)c->amg.size() <= nlev) c->amg.emplace_back();
    Ctx::AmgLevel& F = c->amg[nlev - 1];
    Ctx::AmgLevel& C = c->amg[nlev];
    F.agg.ensure(F.n, st);
    int64_t nc = 0;
    if (c->amg_agg != 0) {
      nc = aggregate_pairwise(c, F, C);
    } else {
      // round-1 aggregates: 8 consecutive Morton ranks of the vertex coordinates on level 0, 8 consecutive rows below
      nc = (F.n + AGG - 1) / AGG;
      if (nlev == 1) {
        const double* lohi = mesh_bbox(c);
        uint64_t* code = scratch<uint64_t>(c, 0, F.n);
        uint64_t* codes = scratch<uint64_t>(c, 1, F.n);
        uint32_t* idx = scratch<uint32_t>(c, 2, F.n);
        uint32_t* perm = scratch<uint32_t>(c, 3, F.n);
        LAUNCH(c, k_morton, grid_for(F.n, TB), TB, 0, c->xyz.p, F.n, c->dim, lohi, code, idx);
        CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, code, codes, idx, perm, F.n, 0, 63, st));
        c->tmp.ensure(bytes, st);
        CK(cub::DeviceRadixSort::SortPairs(c->tmp.p, bytes, code, codes, idx, perm, F.n, 0, 63, st));
        c->launches += 4;
        LAUNCH(c, k_agg_from_perm, grid_for(F.n, TB), TB, 0, perm, F.diag.p, F.n, F.agg.p);
      } else {
        LAUNCH(c, k_agg_consecutive, grid_for(F.n, TB), TB, 0, F.n, F.agg.p);
      }
      CsrRef ref{F.n, F.nnz, F.rowptr.p, F.col.p, F.val.p, F.diag.p};
      C.nnz = galerkin(c, ref, F.agg.p, nc, C.rowptr, C.col, C.val);
    }
    C.n = nc;
    nlev++;
    // rows of every aggregate, ascending (deterministic restriction): stable sort of (aggregate, row)
    {
      uint32_t* key = scratch<uint32_t>(c, 0, F.n);
      uint32_t* keys = scratch<uint32_t>(c, 1, F.n);
      int32_t* row = scratch<int32_t>(c, 2, F.n);
      F.members.ensure(F.n, st);
      F.aggptr.ensure(nc + 1, st);
      LAUNCH(c, k_agg_keys, grid_for(F.n, TB), TB, 0, F.agg.p, F.n, key, row);
      CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, key, keys, row, F.members.p, F.n, 0, 32, st));
      c->tmp.ensure(bytes, st);
      CK(cub::DeviceRadixSort::SortPairs(c->tmp.p, bytes, key, keys, row, F.members.p, F.n, 0, 32, st));
      c->launches += 4;
      LAUNCH(c, k_agg_ptr, grid_for(nc + 1, TB), TB, 0, keys, F.n, nc, F.aggptr.p);
    }
    C.dinv.ensure(nc, st); C.diag.ensure(nc, st);
    LAUNCH(c, k_level_dinv, grid_for(nc, TB), TB, 0, C.rowptr.p, C.col.p, C.val.p, (const uint8_t*)nullptr, nc, C.dinv.p, C.diag.p);
  }
  c->amg_nlev = nlev;
  for (int l = 0; l < nlev; l++) c->amg[l].omega = c->amg_omega_scale;
  if (c->amg_fp32)
    for (int l = 0; l < nlev; l++) {  // fp32 copies for the mixed-precision cycle
      Ctx::AmgLevel& L = c->amg[l];
      L.valf.ensure(L.nnz, st); L.dinvf.ensure(L.n, st);
      LAUNCH(c, k_to_float, grid_for(L.nnz, TB), TB, 0, L.val.p, L.valf.p, L.nnz);
      LAUNCH(c, k_to_float, grid_for(L.n, TB), TB, 0, L.dinv.p, L.dinvf.p, L.n);
    }  // l1-Jacobi: any weight <= 1 keeps the cycle SPD
  // ---- dense inverse of the coarsest level
  {
    Ctx::AmgLevel& L = c->amg[nlev - 1];
    const int n = (int)L.n;
    if (n > 2048) FAIL(REMO_ERR_ARG, "amg_setup: coarsest level still has %d rows", n);
    double* M = scratch<double>(c, 6, (size_t)n * n);
    c->amg_dense.ensure((size_t)n * n, st);
    double* buf[2] = {(n & 1) ? M : c->amg_dense.p, (n & 1) ? c->amg_dense.p : M};  // n swaps later the result sits in amg_dense
    CK(cudaMemsetAsync(buf[0], 0, (size_t)n * n * sizeof(double), st));
    LAUNCH(c, k_dense_fill, grid_for(n, TB), TB, 0, L.rowptr.p, L.col.p, L.val.p, n, buf[0]);
    for (int p = 0; p < n; p++) {
      k_gj_step<<<grid_for((int64_t)n * n, TB), TB, 0, st>>>(n, p, buf[p & 1], buf[(p + 1) & 1]);
    }
    c->launches += n;
    CK(cudaGetLastError());
  }
}

void amg_setup(Ctx* c) {
  cudaStream_t st = c->stream;
  const int64_t nv = c->nv;
  if (c->amg.capacity() < MAXLEV) c->amg.reserve(MAXLEV);
  if (c->amg.empty()) c->amg.emplace_back();
  {
    // ---- level 0: the leading nv x nv block of A = the P1 stiffness matrix, built without the assembled matrix: pattern
    // from the edge list, values by the row-gather assembly restricted to the vertex rows / columns (bit-identical to the
    // entries of the whole matrix).  Constrained vertices keep their entries but have dinv = diag = 0: their x stays 0 in
    // every sweep, they join no aggregate and the Galerkin products drop them.
    Ctx::AmgLevel& L = c->amg[0];
    L.n = nv;
    L.nnz = vertex_block_pattern(c, L.rowptr, L.col);
    L.val.ensure(L.nnz, st); L.dinv.ensure(nv, st); L.diag.ensure(nv, st);
    assemble_vertex_block(c, L.rowptr.p, L.col.p, L.val.p);
    LAUNCH(c, k_level_dinv, grid_for(nv, TB), TB, 0, L.rowptr.p, L.col.p, L.val.p, c->constrained.p, nv, L.dinv.p, L.diag.p);
  }
  amg_build_hierarchy(c);
}
// work blocks of every level for k right-hand sides (grow-only; must be called before a CUDA-graph capture of amg_apply)
void amg_prepare(Ctx* c, int k) {
  cudaStream_t st = c->stream;
  for (int l = 0; l < c->amg_nlev; l++) {
    Ctx::AmgLevel& L = c->amg[l];
    if (c->amg_fp32) {
      L.bf.ensure(L.n * k, st); L.xf.ensure(L.n * k, st); L.tf.ensure(L.n * k, st);
    } else {
      L.b.ensure(L.n * k, st); L.x.ensure(L.n * k, st); L.t.ensure(L.n * k, st);
    }
  }
  c->amg_nrhs = k;
}

namespace {

// per-level storage of the cycle in precision T
template <typename T> struct Lv;
template <> struct Lv<double> {
  static const double* val(Ctx::AmgLevel& L) { return L.val.p; }
  static const double* dinv(Ctx::AmgLevel& L) { return L.dinv.p; }
  static double*& b(Ctx::AmgLevel& L) { return L.b.p; }
  static double*& x(Ctx::AmgLevel& L) { return L.x.p; }
  static double*& t(Ctx::AmgLevel& L) { return L.t.p; }
};
template <> struct Lv<float> {
  static const float* val(Ctx::AmgLevel& L) { return L.valf.p; }
  static const float* dinv(Ctx::AmgLevel& L) { return L.dinvf.p; }
  static float*& b(Ctx::AmgLevel& L) { return L.bf.p; }
  static float*& x(Ctx::AmgLevel& L) { return L.xf.p; }
  static float*& t(Ctx::AmgLevel& L) { return L.tf.p; }
};

// The V-cycle with the level storage in T.  Level 0 reads its right-hand side from the PCG's fp64 block R and its last sweep
// writes the fp64 block Z: with T = float every conversion is fused into a sweep.  Levels [lt, nl) run inside the fused
// tail kernel (lt == nl: per-level launches all the way down and the dense coarsest solve as its own launch).
template <typename T>
void vcycle(Ctx* c, const double* R, double* Z, int k, int lt) {
  cudaStream_t st = c->stream;
  const T alpha = (T)c->amg_alpha;
  const int sweeps = c->amg_sweeps;
  const int nl = c->amg_nlev;
  auto down = [&](int l) {
    Ctx::AmgLevel& L = c->amg[l];
    const double omega = L.omega;
    if (l == 0) {
      LAUNCH(c, (k_jacobi0<T, double>), grid_for(L.n * k, TB), TB, 0, Lv<T>::dinv(L), R, Lv<T>::x(L), k, L.n, (T)omega);
    } else {
      LAUNCH(c, (k_jacobi0<T, T>), grid_for(L.n * k, TB), TB, 0, Lv<T>::dinv(L), (const T*)Lv<T>::b(L), Lv<T>::x(L), k, L.n, (T)omega);
    }
    for (int s = 1; s < sweeps; s++) {
      if (l == 0) smooth_t<T, double, T>(c, L.rowptr.p, L.col.p, Lv<T>::val(L), Lv<T>::dinv(L), R, Lv<T>::x(L), Lv<T>::t(L), k, L.n, omega, 0);
      else smooth_t<T, T, T>(c, L.rowptr.p, L.col.p, Lv<T>::val(L), Lv<T>::dinv(L), (const T*)Lv<T>::b(L), Lv<T>::x(L), Lv<T>::t(L), k, L.n, omega, 0);
      std::swap(Lv<T>::x(L), Lv<T>::t(L));
    }
    // t = b - A x
    if (l == 0) smooth_t<T, double, T>(c, L.rowptr.p, L.col.p, Lv<T>::val(L), Lv<T>::dinv(L), R, Lv<T>::x(L), Lv<T>::t(L), k, L.n, omega, 1);
    else smooth_t<T, T, T>(c, L.rowptr.p, L.col.p, Lv<T>::val(L), Lv<T>::dinv(L), (const T*)Lv<T>::b(L), Lv<T>::x(L), Lv<T>::t(L), k, L.n, omega, 1);
    Ctx::AmgLevel& C = c->amg[l + 1];
    LAUNCH(c, k_restrict<T>, grid_for(C.n * k, TB), TB, 0, L.aggptr.p, L.members.p, (const T*)Lv<T>::t(L), Lv<T>::b(C), k, C.n);
  };
  auto up = [&](int l) {
    Ctx::AmgLevel& L = c->amg[l];
    Ctx::AmgLevel& C = c->amg[l + 1];
    const double omega = L.omega;
    LAUNCH(c, k_prolong<T>, grid_for(L.n * k, TB), TB, 0, L.agg.p, Lv<T>::dinv(L), (const T*)Lv<T>::x(C), Lv<T>::x(L), k, L.n, alpha);
    for (int s = 0; s < sweeps; s++) {
      if (l == 0 && s == sweeps - 1) {
        smooth_t<T, double, double>(c, L.rowptr.p, L.col.p, Lv<T>::val(L), Lv<T>::dinv(L), R, Lv<T>::x(L), Z, k, L.n, omega, 0);
      } else {
        if (l == 0) smooth_t<T, double, T>(c, L.rowptr.p, L.col.p, Lv<T>::val(L), Lv<T>::dinv(L), R, Lv<T>::x(L), Lv<T>::t(L), k, L.n, omega, 0);
        else smooth_t<T, T, T>(c, L.rowptr.p, L.col.p, Lv<T>::val(L), Lv<T>::dinv(L), (const T*)Lv<T>::b(L), Lv<T>::x(L), Lv<T>::t(L), k, L.n, omega, 0);
        std::swap(Lv<T>::x(L), Lv<T>::t(L));
      }
    }
  };
  const int top = std::min(lt, nl - 1);  // per-level launches for the levels [0, top)
  for (int l = 0; l < top; l++) down(l);
  if (lt < nl) {
    TailArgs<T> a;
    memset(&a, 0, sizeof a);
    a.nl = nl - lt;
    a.dense = c->amg_dense.p;
    a.k = k;
    a.sweeps = sweeps;
    a.alpha = alpha;
    for (int l = lt; l < nl; l++) {
      Ctx::AmgLevel& L = c->amg[l];
      TailLevel<T>& Tl = a.L[l - lt];
      Tl.rowptr = L.rowptr.p; Tl.col = L.col.p; Tl.val = Lv<T>::val(L); Tl.dinv = Lv<T>::dinv(L);
      Tl.agg = L.agg.p; Tl.aggptr = L.aggptr.p; Tl.members = L.members.p;
      Tl.b = Lv<T>::b(L); Tl.x = Lv<T>::x(L); Tl.t = Lv<T>::t(L); Tl.n = L.n; Tl.omega = (T)L.omega;
    }
    k_vcycle_tail<T><<<TAIL_CTAS, TAIL_THREADS, 0, st>>>(a);
    c->launches++;
    CK(cudaGetLastError());
    // mirror the kernel's pointer swaps: the result of a non-coarsest tail level ends up where its t block was after an
    // odd number of swaps
    for (int l = lt; l < nl - 1; l++)
      if (((sweeps - 1) + sweeps) & 1) std::swap(Lv<T>::x(c->amg[l]), Lv<T>::t(c->amg[l]));
  } else {  // coarsest: dense inverse
    Ctx::AmgLevel& L = c->amg[nl - 1];
    LAUNCH(c, k_dense_apply_t<T>, grid_for(L.n * k * 8, 128), 128, 0, (int)L.n, c->amg_dense.p, (const T*)Lv<T>::b(L), Lv<T>::x(L), k);
  }
  for (int l = top - 1; l >= 0; l--) up(l);
}

}  // namespace

// z_vert = V-cycle(r_vert) on the leading nv rows of the ndof x k blocks R, Z of the PCG.
void amg_apply(Ctx* c, const double* R, double* Z, int k) {
  const int nl = c->amg_nlev;
  amg_prepare(c, k);
  if (nl == 1) {  // the vertex block itself is small enough for the dense inverse
    Ctx::AmgLevel& L = c->amg[0];
    LAUNCH(c, k_dense_apply_t<double>, grid_for(L.n * k * 8, 128), 128, 0, (int)L.n, c->amg_dense.p, R, Z, k);
    return;
  }
  // levels [lt, nl) go into the fused tail kernel: every level below amg_tail_rows rows, but never level 0 (its right-hand
  // side / result are the PCG's blocks) and only for the column-pair strides the kernel is written for
  int lt = nl;
  if (c->amg_fused_tail && (k & 1) == 0 && k <= 8 && nl >= 2) {
    lt = nl - 1;
    while (lt > 1 && c->amg[lt - 1].n <= c->amg_tail_rows) lt--;
    if (nl - lt < 2) lt = nl;  // the coarsest level alone: nothing to fuse
  }
  if (c->amg_fp32) vcycle<float>(c, R, Z, k, lt);
  else vcycle<double>(c, R, Z, k, lt);
  // the high-order rows (z = D^-1 r) are handled inside the PCG vector kernels (k_init / k_update_r / k_update_px, tail_from = nv)
}
