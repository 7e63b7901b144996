// Element-wise multi-right-hand-side product Q = A P (+ fused p.q) for order-2 and order-3 tetrahedra: the PCG "SpMM" without
// the assembled matrix.  (Written for order 2 first -- the text below; order 3 is the same kernel on batches of 128 tets x 20
// local dofs with the element product generated from the exact reference tensors, see Batch<NLD> and ebe_p3_apply.inc:
// 0.685 ms for 5 right-hand sides at 4.75 M dofs / 227.6 M nnz = 70 % of the HBM roofline by the bytes of the SpMM it
// replaces, against 1.353 ms of the SELL kernel.  Order-3 axisymmetric triangles -- the reference's default 2D configuration --
// use the 256 x 10 shape with 18 metric numbers per element and p3tri_apply.)
//
// Why.  The CSR / SELL SpMM gathers one 64-byte row of P per matrix entry (136.9 M gathers at 4.8 M dofs) and is bound
// by the latency x concurrency of those gathers at ~36 % of the HBM roofline (profiles/r01_notes.md).  The same product
// taken element by element needs 10 dofs per tet -- 38 M (tet, dof) incidences instead of 137 M entries -- and the
// element matrix need not be read at all: for the hierarchical P2 basis (vertex l_i, edge l_a l_b; oracle
// local_basis, assemble.cu tensors) it is a closed form of the 10 metric numbers  S_ij = sigma |K| grad l_i . grad l_j
// that k_geom_tet already leaves in c->gm:
//     grad u = sum_j c_j(l) grad l_j,   c_j = x_j + sum_{b != j} x_{jb} l_b         (linear in the barycentrics)
//     y_i    = sum_j S_ij mean(c_j)                    mean(c_j)     = x_j + s_j / 4,          s_j = sum_b x_{jb}
//     y_ab   = sum_j S_bj mean(l_a c_j) + S_aj mean(l_b c_j),   mean(l_a c_j) = x_j / 4 + (s_j + x_{ja}) / 20
// (~150 flops per tet and right-hand side; checked against the exact tensors in tests/test_oracle.py).
//
// Layout.  Tets are taken in Morton order of their centroid and cut into batches of 256 (one CTA pass, one tet per
// thread).  Per batch, built once per matrix by k_ebe_batch (block radix sort of the 2560 (dof, slot) pairs):
//   udof   [U]        the distinct dofs of the batch (U ~ 620), bit 31 = constrained
//   lidx   [10][256]  slot -> position in udof (16 bit), coalesced per slot
//   lpos   [10][256]  slot -> position of its result in the scratch; the scratch is dof-major in jagged-diagonal
//                     form (dofs ranked by entry count, entry i of dof u at jd[i] + u), ucnt [U] = entries per dof
//   gmb    [10][256]  the metric numbers in batch order, coalesced per number
// The kernel stages the U rows of P in shared memory with cp.async (one gather per DISTINCT dof of a batch: 8.2 M
// instead of 137 M), then per right-hand side: every thread applies its element to the staged values and parks the 10
// results at their dof-major positions of a scratch; one thread per dof adds its entries (fixed order, no
// shared-memory atomics), takes its p.q share and overwrites the staged value in place.  After the last pass one fp64
// RED per (dof, right-hand side) adds the batch's share to Q, which is zeroed first.  Sums across batches are
// RED-ordered, so Q is reproducible to rounding only.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdlib>

#include "space_view.cuh"

namespace {

// One CTA pass = one batch of tets, one tet per thread.  Order 2: 256 tets x 10 local dofs; order 3: 128 tets x 20 local dofs --
// 2560 (tet, local dof) entries per batch either way, so every table and shared-memory array has the same size for both.
constexpr int ENTRIES = 2560;
// metric numbers per element: NG = 10 on tets (pairs (i <= j) of grad l_i . grad l_j), 18 on axisymmetric triangles (6 pairs x
// the 3 weights r_k); shapes in use: <NLD, NG> = <10, 10> order-2 tets, <20, 10> order-3 tets, <10, 18> order-3 triangles
template <int NLD> struct Batch {
  static constexpr int TPB = ENTRIES / NLD;            // tets per batch = threads per CTA
  static constexpr int TSH = (TPB == 256) ? 8 : 7;     // log2(TPB)
  static constexpr int NJ = 1024 / TPB;                // dofs per thread of the per-dof sums (fast8: at most 1024 dofs per batch)
  // resident CTAs per SM the product kernel is compiled for: order 2 fits 80 registers x 768 threads; the order-3 apply holds
  // x[20], y[20] and the metric numbers (~200 registers): 2 x 128 threads, no spills (at 3 CTAs: 1.3 KB of spills per pass)
  static constexpr int MINB = 3;
  static_assert(TPB == 256 || TPB == 128, "batch shapes: 256 x 10 (order 2), 128 x 20 (order 3)");
};
constexpr int KMAX = REMO_MAX_RHS;
constexpr int EBE_MAX_RHS = 8;  // array bound
constexpr int EBE_USE_RHS = 6;  // measured at 4.8 M dofs: 0.69 / 0.83 / 1.19 ms for 5 / 6 / 8 columns against 0.87 ms of the SELL kernel
constexpr uint32_t SENT = 0xffffffffu;
constexpr int EBE_JD = 264;  // jagged-diagonal offsets kept per batch (a dof has at most 256 entries in a batch)

template <int DIM>
__global__ void k_tet_morton(const int32_t* __restrict__ sv, const double* __restrict__ xyz, const double* __restrict__ lohi,
                             int64_t nt, uint64_t* __restrict__ code, int32_t* __restrict__ idx) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nt) return;
  uint64_t c = 0;
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    double x = 0.0;
#pragma unroll
    for (int i = 0; i <= DIM; i++) x += xyz[DIM * (int64_t)sv[t * (DIM + 1) + i] + d];
    x *= 1.0 / (DIM + 1);
    const double ext = lohi[3 + d] - lohi[d];
    const double u = ext > 0 ? (x - lohi[d]) / ext : 0.0;
    const uint64_t q = (uint64_t)fmin(fmax(u * 2097151.0, 0.0), 2097151.0);
    c |= spread21s(q) << d;
  }
  code[t] = c;
  idx[t] = (int32_t)t;
}

// One CTA per batch.  FILL = false: count the distinct dofs; true: write the batch tables at the offsets uoff[].
//
// Conflict-aware layout (color != 0, pieces of at most 8 entries).  The product kernel makes two random 8-byte
// shared-memory accesses per (tet, slot, right-hand side): it GATHERS the staged row of the slot's dof (address ~ row)
// and SCATTERS the element's result to entry i of that dof in the jagged-diagonal scratch (address jd[i] + row).  The 16
// lanes of a half-warp executing one slot form a GROUP; a group's request costs one wavefront per distinct address in its
// busiest 8-byte bank pair, and with rows / entry orders taken as they come that is 2.4 (gathers) and 3.0 (scatters)
// wavefronts instead of 1 (tools/ebe_layout_sim.py; ncu: 40 M of the 120 M wavefronts of a product were conflicts).
// Both layouts have slack: a piece may take ANY row of its entry-count class, and the entries of a piece may take its
// indices 0..n-1 in ANY order.  Warp 0 deals them greedily, pieces with the most groups first: the row whose bank pair
// (row mod 16) is least used in the piece's groups, then for every entry the free index whose scratch bank pair is least
// used in the entry's group.  Model: 1.2 / 2.3 wavefronts per request.  Fixed visiting order -> the tables, and with
// them the summation order of the product, are reproducible.
template <bool FILL, int NLD>
__global__ void __launch_bounds__(Batch<NLD>::TPB) k_ebe_batch(SpaceView s, const int32_t* __restrict__ tperm,
                                                   int split, int color, const uint8_t* __restrict__ constrained, int64_t* __restrict__ ucount,
                                                   int* __restrict__ umax, const int64_t* __restrict__ uoff,
                                                   int32_t* __restrict__ udof, uint16_t* __restrict__ lidx,
                                                   uint16_t* __restrict__ lpos, uint16_t* __restrict__ ucnt, uint16_t* __restrict__ jdp) {
  constexpr int TPB = Batch<NLD>::TPB, TSH = Batch<NLD>::TSH;
  using Sort = cub::BlockRadixSort<uint32_t, TPB, NLD, uint16_t>;
  using Scan = cub::BlockScan<int, TPB>;
  constexpr int NGRP = (TPB / 16) * NLD;  // gather / scatter groups of a batch: (half-warp, slot)
  struct Greedy {
    uint8_t gocc[NGRP * 16];   // rows already dealt per (group, bank pair)
    uint8_t socc[NGRP * 16];   // scratch entries already dealt per (group, bank pair)
    uint8_t sidx[TPB * NLD];   // sorted entry -> its index inside its piece
    uint16_t nxt[9 * 16];      // next free row per (entry-count class, bank pair)
  };
  __shared__ union Tmp {
    typename Sort::TempStorage sort;
    typename Scan::TempStorage scan;
    Greedy g;
  } tmp;
  __shared__ uint32_t skey[TPB * NLD];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int64_t ti = (int64_t)b * TPB + tid;
  uint32_t key[NLD];
  uint16_t val[NLD];
  if (ti < s.nt) {
    const int64_t t = tperm[ti];
#pragma unroll
    for (int k = 0; k < NLD; k++) key[k] = (uint32_t)elem_dof(s, t, k);  // vertices, edge dofs, (order 3) face dofs
  } else {
#pragma unroll
    for (int k = 0; k < NLD; k++) key[k] = SENT;
  }
#pragma unroll
  for (int k = 0; k < NLD; k++) val[k] = (uint16_t)(k * TPB + tid);
  Sort(tmp.sort).Sort(key, val);  // blocked arrangement: thread i holds the sorted positions 10 i .. 10 i + 9
#pragma unroll
  for (int k = 0; k < NLD; k++) skey[tid * NLD + k] = key[k];
  __syncthreads();
  bool head[NLD];
  int heads = 0;
#pragma unroll
  for (int k = 0; k < NLD; k++) {
    const int pos = tid * NLD + k;
    head[k] = key[k] != SENT && (pos == 0 || skey[pos - 1] != key[k]);
    if (split > 0 && key[k] != SENT && !head[k]) {
      // a dof with many entries in the batch (a vertex: up to 40) is cut into pieces of `split` entries; every piece is
      // a dof of its own for the product kernel (staged, summed and added to Q separately), so no thread of the per-dof
      // sums runs much longer than the others
      int lo = pos > TPB ? pos - TPB : 0, hi = pos;  // first entry of this dof: a dof has at most 256 entries
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (skey[mid] < key[k]) lo = mid + 1; else hi = mid;
      }
      head[k] = ((pos - lo) % split) == 0;
    }
    heads += head[k] ? 1 : 0;
  }
  int base = 0, total = 0;
  Scan(tmp.scan).ExclusiveSum(heads, base, total);
  if (!FILL) {
    if (tid == 0) {
      ucount[b] = total;
      atomicMax(umax, total);
    }
    return;
  }
  // The scratch of the product kernel is laid out as JAGGED DIAGONALS: the dofs of the batch are ranked by their number
  // of entries (descending) and entry i of the dof with rank u lives at jd[i] + u, jd[i] = number of entries with a
  // smaller index i.  One thread per dof then walks i = 0 .. count-1: neighbouring lanes read neighbouring words (no
  // bank conflicts) and the lanes of a warp have (almost) the same trip count -- a vertex has 20-40 entries in a
  // batch, an edge ~3, and in dof order they would share warps.
  __shared__ uint16_t spos[TPB * NLD + 1];  // first sorted position of every piece (by its rank in dof order)
  __shared__ uint16_t srow[TPB * NLD];      // rank in dof order -> row (rank by entry count, permuted inside its class)
  __shared__ uint16_t sord[TPB * NLD];      // rank by entry count -> rank in dof order
  __shared__ uint16_t sval[TPB * NLD];      // sorted position -> slot * 256 + tet
  __shared__ uint16_t sjd[EBE_JD];
  const bool colored = color != 0 && split > 0 && split <= 8;
  const int nvalid = (int)((s.nt - (int64_t)b * TPB < TPB ? s.nt - (int64_t)b * TPB : TPB) * NLD);
  {
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < NLD; k++) {
      if (head[k]) spos[base + cnt++] = (uint16_t)(tid * NLD + k);
      sval[tid * NLD + k] = val[k];
    }
    if (tid == 0) spos[total] = (uint16_t)nvalid;  // padding tets sort behind every real entry
  }
  __syncthreads();
  uint32_t ckey[NLD];
  uint16_t cval[NLD];
#pragma unroll
  for (int k = 0; k < NLD; k++) {
    const int rho = tid * NLD + k;
    cval[k] = (uint16_t)rho;
    if (rho >= total) { ckey[k] = SENT; continue; }
    const int p0 = spos[rho], n = spos[rho + 1] - p0;
    if (!colored) { ckey[k] = 0xffffu - (uint32_t)n; continue; }
    // entries descending, then pieces that touch the most groups first (they are the hardest to place)
    int ngr = 0;
    uint32_t seen[8];
    for (int e = 0; e < n; e++) {
      const uint32_t v = sval[p0 + e];
      const uint32_t g = ((v & (TPB - 1)) >> 4) * NLD + (v >> TSH);
      bool dup = false;
      for (int f = 0; f < ngr; f++) dup |= seen[f] == g;
      if (!dup) seen[ngr++] = g;
    }
    ckey[k] = ((uint32_t)(8 - n) << 4) | (uint32_t)(8 - ngr);
  }
  Sort(tmp.sort).Sort(ckey, cval);
  uint32_t* scount = skey;  // entry counts by new rank (descending); the sorted dof keys are no longer needed
  __syncthreads();          // every thread has read skey (head flags) and the sort is done with its temp storage
#pragma unroll
  for (int k = 0; k < NLD; k++) {
    const int j = tid * NLD + k;
    if (j < total) {
      sord[j] = cval[k];
      srow[cval[k]] = (uint16_t)j;
      scount[j] = colored ? 8u - (ckey[k] >> 4) : 0xffffu - ckey[k];
    }
  }
  __syncthreads();
  int ng[NLD], jd[NLD];
#pragma unroll
  for (int k = 0; k < NLD; k++) {  // dofs with more than i entries = first new rank whose count is <= i
    const uint32_t i = (uint32_t)(tid * NLD + k);
    int lo = 0, hi = total;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (scount[mid] > i) lo = mid + 1; else hi = mid;
    }
    ng[k] = lo;
  }
  Scan(tmp.scan).ExclusiveSum(ng, jd);
  const int64_t u0 = uoff[b];
  const int64_t bo = (int64_t)b * (TPB * NLD);
#pragma unroll
  for (int k = 0; k < NLD; k++) {
    const int i = tid * NLD + k;
    if (i < EBE_JD) {
      sjd[i] = (uint16_t)jd[k];
      jdp[(int64_t)b * EBE_JD + i] = (uint16_t)jd[k];
    }
  }
  __syncthreads();  // sjd complete; the scan is done with its temp storage
  if (colored) {
    Greedy& G = tmp.g;
    for (int i = tid; i < NGRP * 16; i += TPB) { G.gocc[i] = 0; G.socc[i] = 0; }
    if (tid < 9 * 16) {
      // class c (pieces of c entries) owns the rows [lo, hi): lo = pieces with more than c entries.  First row of the
      // class in bank pair r: lo + ((r - lo) mod 16)
      const int c = tid >> 4, r = tid & 15;
      int lo = 0;
      if (c >= 1) {
        int l = 0, h = total;
        while (l < h) {
          const int mid = (l + h) >> 1;
          if (scount[mid] > (uint32_t)c) l = mid + 1; else h = mid;
        }
        lo = l;
      }
      G.nxt[tid] = (uint16_t)(lo + ((r - lo) & 15));
    }
    __syncthreads();
    if (tid < 32) {
      const int lane = tid;
      for (int j = 0; j < total; j++) {
        const int rho = sord[j], n = (int)scount[j];
        const int p0 = spos[rho];
        // class range end: first rank with a smaller count = pieces with at least n entries
        // (ranks are count-descending, so it is the first j' with scount[j'] < n; found once per class below)
        // ---- row: least used bank pair of the piece's groups, among the bank pairs that still have a row in the class
        uint32_t cost = 0xffffu;
        if (lane < 16) {
          const int row = G.nxt[n * 16 + lane];
          // the row is free iff it is still inside the class: rows of the class are [lo, hi) and hi = lo + size; a row
          // index >= total or with another count is outside
          if (row < total && (int)scount[row] == n) {
            cost = 0;
            for (int e = 0; e < n; e++) {
              const uint32_t v = sval[p0 + e];
              cost += G.gocc[(((v & (TPB - 1)) >> 4) * NLD + (v >> TSH)) * 16 + lane];
            }
          }
        }
        uint32_t best = (cost << 8) | (uint32_t)(lane & 15);
#pragma unroll
        for (int o = 8; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        best = __shfl_sync(0xffffffffu, best, 0);  // lanes 16..31 computed on dummies: take the half-warp's result
        const int r = (int)(best & 15u);
        const int row = G.nxt[n * 16 + r];
        __syncwarp();
        if (lane == 0) {
          G.nxt[n * 16 + r] = (uint16_t)(row + 16);
          srow[rho] = (uint16_t)row;
          for (int e = 0; e < n; e++) {
            const uint32_t v = sval[p0 + e];
            G.gocc[(((v & (TPB - 1)) >> 4) * NLD + (v >> TSH)) * 16 + r]++;
          }
        }
        // ---- entry order: every entry takes the free index whose scratch bank pair is least used in its group
        uint32_t avail = (1u << n) - 1u;
        for (int e = 0; e < n; e++) {
          const uint32_t v = sval[p0 + e];
          const int g = (int)(((v & (TPB - 1)) >> 4) * NLD + (v >> TSH));
          uint32_t c2 = 0xffffu;
          if (lane < n && ((avail >> lane) & 1u)) c2 = G.socc[g * 16 + ((sjd[lane] + row) & 15)];
          uint32_t b2 = (c2 << 8) | (uint32_t)(lane & 7);
#pragma unroll
          for (int o = 4; o; o >>= 1) b2 = min(b2, __shfl_xor_sync(0xffffffffu, b2, o));
          b2 = __shfl_sync(0xffffffffu, b2, 0);
          const int i = (int)(b2 & 7u);
          avail &= ~(1u << i);
          if (lane == 0) {
            G.sidx[p0 + e] = (uint8_t)i;
            G.socc[g * 16 + ((sjd[i] + row) & 15)]++;
          }
          __syncwarp();
        }
      }
    }
    __syncthreads();
  }
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < NLD; k++) {
    const int pos = tid * NLD + k;
    if (head[k]) cnt++;
    const int rho = base + cnt - 1;  // rank (dof order) of the piece this entry belongs to
    const bool real = key[k] != SENT;
    const int nr = real ? srow[rho] : 0;
    if (head[k]) {
      udof[u0 + nr] = (int32_t)(key[k] | (constrained[key[k]] ? 0x80000000u : 0u));
      ucnt[u0 + nr] = (uint16_t)(spos[rho + 1] - spos[rho]);
    }
    lidx[bo + val[k]] = (uint16_t)nr;
    // a dof appears at most once per tet: at most 256 entries, so i < EBE_JD.  Padding tets keep a slot behind the
    // nvalid real entries
    const int i = real ? (colored ? (int)tmp.g.sidx[pos] : pos - spos[rho]) : 0;
    lpos[bo + val[k]] = (uint16_t)(real ? sjd[i] + nr : pos);
  }
}

template <int TPB, int NG>
__global__ void k_ebe_gm(const double* __restrict__ gm, const int32_t* __restrict__ tperm, int64_t nt, int64_t nb,
                         double* __restrict__ gmb) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // i = b * TPB + tet
  if (i >= nb * TPB) return;
  const int64_t b = i / TPB, l = i - b * TPB;
  const bool in = i < nt;
  const double* g = gm + (in ? (int64_t)tperm[i] : 0) * NG;
#pragma unroll
  for (int m = 0; m < NG; m++) gmb[(b * NG + m) * TPB + l] = in ? g[m] : 0.0;
}

__global__ void k_ebe_offsets_in(const int64_t* __restrict__ ucount, int64_t nb, int64_t* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i <= nb) out[i] = i < nb ? ucount[i] : 0;
}

// y = K_e x for one P2 tet from its 10 metric numbers (pairs (0,0) (0,1) (0,2) (0,3) (1,1) (1,2) (1,3) (2,2) (2,3) (3,3);
// local dofs: vertices 0..3, then the edges (0,1) (0,2) (0,3) (1,2) (1,3) (2,3) of the sorted tet)
__device__ __forceinline__ void p2_apply(const double (&g)[10], const double (&x)[10], double (&y)[10]) {
  const double S[4][4] = {{g[0], g[1], g[2], g[3]}, {g[1], g[4], g[5], g[6]}, {g[2], g[5], g[7], g[8]}, {g[3], g[6], g[8], g[9]}};
  // xe[j][a] = edge value between the local vertices j and a (0 on the diagonal)
  const double xe[4][4] = {{0.0, x[4], x[5], x[6]}, {x[4], 0.0, x[7], x[8]}, {x[5], x[7], 0.0, x[9]}, {x[6], x[8], x[9], 0.0}};
  double d0[4], bs[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const double sj = (xe[j][0] + xe[j][1]) + (xe[j][2] + xe[j][3]);
    d0[j] = fma(0.25, sj, x[j]);
    bs[j] = fma(0.05, sj, 0.25 * x[j]);
  }
  double B[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    y[i] = fma(S[i][0], d0[0], fma(S[i][1], d0[1], fma(S[i][2], d0[2], S[i][3] * d0[3])));
    B[i] = fma(S[i][0], bs[0], fma(S[i][1], bs[1], fma(S[i][2], bs[2], S[i][3] * bs[3])));
  }
  // V[b][a] = sum_j S_bj xe[j][a]
  double V[4][4];
#pragma unroll
  for (int bb = 0; bb < 4; bb++)
#pragma unroll
    for (int a = 0; a < 4; a++) {
      if (a == bb) { V[bb][a] = 0.0; continue; }
      double v = 0.0;
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (j != a) v = fma(S[bb][j], xe[j][a], v);
      V[bb][a] = v;
    }
  constexpr int EA[6] = {0, 0, 0, 1, 1, 2}, EB[6] = {1, 2, 3, 2, 3, 3};
#pragma unroll
  for (int e = 0; e < 6; e++) y[4 + e] = (B[EA[e]] + B[EB[e]]) + 0.05 * (V[EB[e]][EA[e]] + V[EA[e]][EB[e]]);
}

// order 3: the same product straight from the exact reference tensors, symmetric pairs only (tools/gen_ebe_apply.py)
#include "ebe_p3_apply.inc"

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NLD, int NG, int MINB = (NG == 18 ? 2 : Batch<NLD>::MINB)>  // triangles: 18 metric numbers live, 2 CTAs
__global__ void __launch_bounds__(Batch<NLD>::TPB, MINB) k_spmm_ebe(int nb, const int64_t* __restrict__ uoff, const int32_t* __restrict__ udof,
                                                     const uint16_t* __restrict__ lidx, const uint16_t* __restrict__ lpos,
                                                     const uint16_t* __restrict__ ucnt, const uint16_t* __restrict__ jdp,
                                                     const double* __restrict__ gmb,
                                                     const double* __restrict__ P, int pstride, double* __restrict__ Q, int ks,
                                                     int nr, int xst, int umax, int fast8, int pf, double* __restrict__ partial) {
  constexpr int TPB = Batch<NLD>::TPB, NJ = Batch<NLD>::NJ;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // xs: umax x xst staged rows of P (xst odd: random rows spread over the banks); column r is overwritten in place by
  // the batch's share of Q once pass r no longer needs it.  scr: the element results of the current pass in SORTED
  // (dof-major) order.  Offsets into xs / scr are BYTE offsets (< 64 K, kept as 16-bit pairs).
  unsigned char* xs = smem_raw;
  double* scr = reinterpret_cast<double*>(xs + (size_t)umax * xst * 8);
  int32_t* sdof = reinterpret_cast<int32_t*>(scr + NLD * TPB);
  uint16_t* scnt = reinterpret_cast<uint16_t*>(sdof + umax);
  uint16_t* sjd = scnt + umax;
  __shared__ double sdot[TPB / 32][EBE_MAX_RHS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fr = tid & 7;  // staging / flush: 8 lanes per dof row, lane = right-hand side
  const int xstep = xst * 8;
  if (tid < (TPB / 32) * EBE_MAX_RHS) (&sdot[0][0])[tid] = 0.0;
  for (int b = blockIdx.x; b < nb; b += gridDim.x) {
    const int64_t u0 = uoff[b];
    const int U = (int)(uoff[b + 1] - u0);
    if (pf && b + (int)gridDim.x < nb) {
      // the tables of this CTA's next batch are first-touch DRAM reads: pull their lines into L2 now
      const int64_t nb1 = b + gridDim.x;
      constexpr int L_IDX = ENTRIES / 64, L_GM = NG * TPB / 16, L_JD = (EBE_JD + 63) / 64;  // 128-byte lines of each table
      for (int i = tid; i < 2 * L_IDX + L_GM + L_JD; i += TPB) {
        const void* a;
        if (i < L_IDX) a = lidx + nb1 * ENTRIES + i * 64;
        else if (i < 2 * L_IDX) a = lpos + nb1 * ENTRIES + (i - L_IDX) * 64;
        else if (i < 2 * L_IDX + L_GM) a = gmb + nb1 * (NG * TPB) + (i - 2 * L_IDX) * 16;
        else a = jdp + nb1 * EBE_JD + (i - 2 * L_IDX - L_GM) * 64;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
      }
      if (tid < 36) {
        const int64_t un = uoff[nb1];
        const void* a2 = tid < 24 ? (const void*)(udof + un + tid * 32) : (const void*)(ucnt + un + (tid - 24) * 64);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a2));
      }
    }
    for (int i = tid; i < U; i += TPB) sdof[i] = udof[u0 + i];
    for (int i = tid; i < U; i += TPB) scnt[i] = ucnt[u0 + i];
    for (int i = tid; i < EBE_JD; i += TPB) sjd[i] = jdp[(int64_t)b * EBE_JD + i];
    __syncthreads();
    // stage the U rows of P: every copy of the batch is issued before anything waits (cp.async, 8 bytes per lane)
    if (fr < nr) {
      const uint32_t dst = smem_u32(xs) + fr * 8;
      const double* src = P + fr;
      for (int row = tid >> 3; row < U; row += TPB / 8) {
        const int64_t dof = sdof[row] & 0x7fffffff;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + row * xstep), "l"(src + dof * pstride) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    uint32_t lxo[NLD / 2], lso[NLD / 2];  // per slot: byte offset of its staged row in xs / of its sorted position in scr
    double g[NG];
#pragma unroll
    for (int k = 0; k < NLD / 2; k++) {
      const int64_t o0 = ((int64_t)b * NLD + 2 * k) * TPB + tid, o1 = o0 + TPB;
      lxo[k] = ((uint32_t)__ldcs(lidx + o0) * xstep) | (((uint32_t)__ldcs(lidx + o1) * xstep) << 16);
      lso[k] = ((uint32_t)__ldcs(lpos + o0) * 8u) | (((uint32_t)__ldcs(lpos + o1) * 8u) << 16);
    }
#pragma unroll
    for (int k = 0; k < NG; k++) g[k] = __ldcs(gmb + ((int64_t)b * NG + k) * TPB + tid);
    // fast8: entry counts (+ constrained flag in bit 7) of this thread's up to NJ dofs, and the first 8 diagonal offsets
    uint64_t info = 0;
    uint32_t jdr[4] = {0, 0, 0, 0};
    if (fast8) {
#pragma unroll
      for (int j = 0; j < NJ; j++) {
        const int u = tid + j * TPB;
        if (u < U) info |= (uint64_t)((uint32_t)scnt[u] | (sdof[u] < 0 ? 0x80u : 0u)) << (8 * j);
      }
#pragma unroll
      for (int i = 0; i < 4; i++) jdr[i] = (uint32_t)sjd[2 * i] | ((uint32_t)sjd[2 * i + 1] << 16);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int r = 0; r < nr; r++) {
      {
        double x[NLD];
        const unsigned char* xr = xs + r * 8;
#pragma unroll
        for (int k = 0; k < NLD; k++) x[k] = *reinterpret_cast<const double*>(xr + ((lxo[k >> 1] >> ((k & 1) * 16)) & 0xffffu));
        unsigned char* sb = reinterpret_cast<unsigned char*>(scr);
        double y[NLD];
        if constexpr (NLD == 10 && NG == 10) p2_apply(g, x, y);
        else if constexpr (NLD == 20) p3_apply(g, x, y);
        else p3tri_apply(g, x, y);
#pragma unroll
        for (int k = 0; k < NLD; k++) *reinterpret_cast<double*>(sb + ((lso[k >> 1] >> ((k & 1) * 16)) & 0xffffu)) = y[k];
      }
      __syncthreads();
      // one thread per dof (ranked by entry count): entry i of dof u sits at jd[i] + u -- conflict-free, equal trip
      // counts inside a warp, fixed order, no atomics
      double d = 0.0;
      if (fast8) {
        // every dof has at most 8 entries and the batch at most 1024 dofs: the diagonal offsets and this thread's entry
        // counts stay in registers for all passes (no table loads in the loop)
#pragma unroll
        for (int j = 0; j < NJ; j++) {
          const int n = (int)((info >> (8 * j)) & 15u);
          if (n) {
            const int u = tid + j * TPB;
            const double* e = scr + u;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
              if (i < n) s0 += e[(jdr[i >> 1] & 0xffffu)];
              if (i + 1 < n) s1 += e[(jdr[i >> 1] >> 16)];
            }
            s0 += s1;
            double* px = reinterpret_cast<double*>(xs + u * xstep + r * 8);
            if (!((info >> (8 * j + 7)) & 1u)) d = fma(s0, *px, d);  // p.q: constrained rows do not count
            *px = s0;
          }
        }
      } else {
        for (int u = tid; u < U; u += TPB) {
          const double* e = scr + u;
          const int n = scnt[u];
          double s0 = 0.0, s1 = 0.0;
          int i = 0;
          for (; i + 2 <= n; i += 2) {
            s0 += e[sjd[i]];
            s1 += e[sjd[i + 1]];
          }
          if (i < n) s0 += e[sjd[i]];
          s0 += s1;
          double* px = reinterpret_cast<double*>(xs + u * xstep + r * 8);
          if (sdof[u] >= 0) d = fma(s0, *px, d);  // p.q: constrained rows do not count (and are not written to Q)
          *px = s0;
        }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (lane == 0) sdot[warp][r] += d;
      __syncthreads();  // everybody has read scr: the next pass may overwrite it
    }
    // flush the batch's share of Q: one RED per (dof, right-hand side), 8 lanes per dof row
    if (fr < nr) {
      for (int row = tid >> 3; row < U; row += TPB / 8) {
        const int32_t dc = sdof[row];
        if (dc >= 0) atomicAdd(Q + (int64_t)dc * ks + fr, *reinterpret_cast<const double*>(xs + row * xstep + fr * 8));
      }
    }
    __syncthreads();  // the next batch restages xs / sdof
  }
  if (tid < ks) {
    double t = 0.0;
    if (tid < nr)
      for (int w = 0; w < TPB / 32; w++) t += sdot[w][tid];
    partial[(int64_t)blockIdx.x * KMAX + tid] = t;
  }
}

// Table validator (remo_set_option("ebe_check", 1); compute-sanitizer is not available on every pool, so the bounds the
// product kernel relies on are asserted here, once per matrix): per batch U <= umax, every dof number < ndof, every slot's
// staged row < U, the 2560 scratch positions are a permutation of 0..2559 (no two results of a pass land on the same word),
// every real entry sits on one of the first ucnt[row] diagonals of its row, and the entry counts add up.
template <int NLD>
__global__ void __launch_bounds__(Batch<NLD>::TPB) k_ebe_check(int64_t nt, int64_t ndof, int umax, const int64_t* __restrict__ uoff,
                                                   const int32_t* __restrict__ udof, const uint16_t* __restrict__ lidx,
                                                   const uint16_t* __restrict__ lpos, const uint16_t* __restrict__ ucnt,
                                                   const uint16_t* __restrict__ jdp, int* __restrict__ err) {
  constexpr int TPB = Batch<NLD>::TPB;
  __shared__ uint32_t seen[(TPB * NLD + 31) / 32];
  __shared__ int cnt_sum;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int64_t u0 = uoff[b];
  const int U = (int)(uoff[b + 1] - u0);
  for (int i = tid; i < (TPB * NLD + 31) / 32; i += TPB) seen[i] = 0;
  if (tid == 0) cnt_sum = 0;
  __syncthreads();
  int bad = 0;
  if (tid == 0 && (U > umax || U < 1)) bad++;
  int mine = 0;
  for (int u = tid; u < U; u += TPB) {
    if ((int64_t)(udof[u0 + u] & 0x7fffffff) >= ndof) bad++;
    const int n = ucnt[u0 + u];
    if (n < 1 || n > TPB) bad++;
    mine += n;
  }
  atomicAdd(&cnt_sum, mine);
  const int64_t left = nt - (int64_t)b * TPB;
  const bool real = tid < left;
  const uint16_t* jd = jdp + (int64_t)b * EBE_JD;
  for (int k = 0; k < NLD; k++) {
    const int64_t o = ((int64_t)b * NLD + k) * TPB + tid;
    const int row = lidx[o], pos = lpos[o];
    if (pos >= TPB * NLD) { bad++; continue; }
    if (atomicOr(&seen[pos >> 5], 1u << (pos & 31)) & (1u << (pos & 31))) bad++;
    if (real) {
      if (row >= U) { bad++; continue; }
      const int n = ucnt[u0 + row];
      bool found = false;
      for (int i = 0; i < n && i < EBE_JD; i++) found |= (int)jd[i] + row == pos;
      if (!found) bad++;
    }
  }
  __syncthreads();
  const int nvalid = (int)((left < TPB ? left : TPB) * NLD);
  if (tid == 0 && cnt_sum != nvalid) bad++;
  if (bad) atomicAdd(err, bad);
}

size_t ebe_smem(int umax, int nr) {
  const int xst = nr | 1;
  return (size_t)umax * xst * 8 + (size_t)ENTRIES * 8 + (size_t)umax * 4 + (size_t)(umax + EBE_JD) * 2;
}

}  // namespace

bool ebe_eligible(const Ctx* c) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("REMO_SPMM_EBE");
    on = e ? atoi(e) : 1;
  }
  if (c->ebe_on >= 0 ? c->ebe_on == 0 : on == 0) return false;
  if (c->ndof >= 0x7fffffff) return false;
  return (c->dim == 3 && (c->order == 2 || c->order == 3)) || (c->dim == 2 && c->order == 3);
}

int ebe_max_rhs() { return EBE_USE_RHS; }

// Will the element-wise product take a block of nr right-hand sides?  Decides the row stride of the vector blocks
// (solver.cu), so it must agree with ebe_usable once the tables exist: a batch may not fit (the staged rows are addressed
// with 16-bit byte offsets: order-3 triangles have ~1200 distinct dofs per batch, which rules out 6 columns), so the tables
// are built here if the element metrics are there.
bool ebe_serves(Ctx* c, int nr) {
  if (!ebe_eligible(c) || nr < 1 || nr > EBE_USE_RHS) return false;
  if (!c->have_ebe && c->have_matrix) ebe_build(c);
  return c->have_ebe ? c->ebe_occ[nr] > 0 : true;
}

bool ebe_usable(const Ctx* c, int nr) { return c->have_ebe && nr >= 1 && nr <= EBE_USE_RHS && c->ebe_occ[nr] > 0; }

int ebe_grid(const Ctx* c, int nr) { return (int)std::min<int64_t>(c->ebe_nb, (int64_t)c->num_sms * c->ebe_occ[nr]); }

namespace {

template <int NLD, int NG>
void ebe_build_t(Ctx* c) {
  constexpr int TPB = Batch<NLD>::TPB;
  constexpr int DIM = (NG == 10) ? 3 : 2;
  cudaStream_t st = c->stream;
  const int64_t nt = c->nt, nb = (nt + TPB - 1) / TPB;
  size_t bytes = 0;
  // tets in Morton order of their centroid: a batch is a compact blob, consecutive batches are neighbours
  uint64_t* code = scratch<uint64_t>(c, 0, nt);
  uint64_t* codes = scratch<uint64_t>(c, 1, nt);
  int32_t* idx = scratch<int32_t>(c, 2, nt);
  int32_t* tperm = scratch<int32_t>(c, 3, nt);
  LAUNCH(c, k_tet_morton<DIM>, grid_for(nt, 256), 256, 0, c->sv.p, c->xyz.p, mesh_bbox(c), nt, code, idx);
  CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, code, codes, idx, tperm, nt, 0, 63, st));
  c->tmp.ensure(bytes, st);
  CK(cub::DeviceRadixSort::SortPairs(c->tmp.p, bytes, code, codes, idx, tperm, nt, 0, 63, st));
  c->launches += 4;
  static const int split = [] { const char* e = getenv("REMO_EBE_SPLIT"); return e ? atoi(e) : 8; }();
  // pass 1: distinct dofs per batch -> offsets
  int64_t* ucount = scratch<int64_t>(c, 4, nb + 1);
  int64_t* uin = scratch<int64_t>(c, 5, nb + 1);
  int* umax_d = scratch<int>(c, 6, 1);
  CK(cudaMemsetAsync(umax_d, 0, sizeof(int), st));
  SpaceView sview = make_view(c);
  static const int color = [] { const char* e = getenv("REMO_EBE_COLOR"); return e ? atoi(e) : 0; }();  // off until the greedy build is cheap (53 ms at 4.8 M dofs for 0.05 ms per product)
  LAUNCH(c, (k_ebe_batch<false, NLD>), (unsigned)nb, TPB, 0, sview, tperm, split, color, c->constrained.p, ucount, umax_d, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  LAUNCH(c, k_ebe_offsets_in, grid_for(nb + 1, 256), 256, 0, ucount, nb, uin);
  c->ebe_uoff.ensure(nb + 1, st);
  CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, uin, c->ebe_uoff.p, nb + 1, st));
  c->tmp.ensure(bytes, st);
  CK(cub::DeviceScan::ExclusiveSum(c->tmp.p, bytes, uin, c->ebe_uoff.p, nb + 1, st));
  c->launches++;
  int64_t total = 0;
  int umax = 0;
  CK(cudaMemcpyAsync(&total, c->ebe_uoff.p + nb, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&umax, umax_d, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  umax = (umax + 1) & ~1;  // even: the 16-byte alignment of the shared-memory arrays behind xs
  // pass 2: tables
  c->ebe_udof.ensure(total, st);
  c->ebe_ucnt.ensure(total, st);
  c->ebe_jd.ensure((size_t)nb * EBE_JD, st);
  c->ebe_lidx.ensure((size_t)nb * ENTRIES, st);
  c->ebe_lpos.ensure((size_t)nb * ENTRIES, st);
  c->ebe_gm.ensure((size_t)nb * TPB * NG, st);
  LAUNCH(c, (k_ebe_batch<true, NLD>), (unsigned)nb, TPB, 0, sview, tperm, split, color, c->constrained.p, nullptr, nullptr, c->ebe_uoff.p, c->ebe_udof.p,
         c->ebe_lidx.p, c->ebe_lpos.p, c->ebe_ucnt.p, c->ebe_jd.p);
  LAUNCH(c, (k_ebe_gm<TPB, NG>), grid_for(nb * TPB, 256), 256, 0, c->gm.p, tperm, nt, nb, c->ebe_gm.p);
  c->ebe_nb = nb;
  c->ebe_nld = NLD;
  c->ebe_ng = NG;
  c->ebe_fast8 = (split > 0 && split <= 8 && umax <= Batch<NLD>::NJ * TPB) ? 1 : 0;
  c->ebe_umax = umax;
  // resident CTAs per SM for every right-hand-side count (the shared-memory row stride of xs depends on it)
  // raised once per (kernel, device) to the opt-in maximum, never to a per-mesh value: other contexts of this GPU launch the
  // same kernel for other meshes (ctx.cuh allow_max_smem)
  // order 3 also has a 128-register build (4 resident CTAs, 240 B of spills): remo_set_option("ebe_p3_ctas", 4)
  bool four = false;
  if constexpr (NLD == 20) four = c->ebe_p3_ctas >= 4;
  int dev_max = 0;
  if constexpr (NLD == 20) {
    dev_max = four ? allow_max_smem(k_spmm_ebe<NLD, NG, 4>, c->device) : allow_max_smem(k_spmm_ebe<NLD, NG>, c->device);
  } else {
    dev_max = allow_max_smem(k_spmm_ebe<NLD, NG>, c->device);
  }
  for (int nr = 1; nr <= EBE_MAX_RHS; nr++) {
    const size_t sm = ebe_smem(umax, nr);
    int occ = 0;
    // the kernel keeps byte offsets into xs as 16-bit numbers
    if (sm <= (size_t)dev_max && (size_t)umax * (nr | 1) * 8 < 65536) {
      if constexpr (NLD == 20) {
        if (four) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmm_ebe<NLD, NG, 4>, TPB, sm));
        else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmm_ebe<NLD, NG>, TPB, sm));
      } else {
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmm_ebe<NLD, NG>, TPB, sm));
      }
    }
    c->ebe_occ[nr] = occ;  // 0: a batch does not fit (degenerate mesh) -> the SELL / CSR kernels take over
  }
  if (c->ebe_check) {
    int* err_d = scratch<int>(c, 6, 1);
    int err = 0;
    CK(cudaMemsetAsync(err_d, 0, sizeof(int), st));
    LAUNCH(c, k_ebe_check<NLD>, (unsigned)nb, TPB, 0, nt, c->ndof, umax, c->ebe_uoff.p, c->ebe_udof.p, c->ebe_lidx.p, c->ebe_lpos.p, c->ebe_ucnt.p, c->ebe_jd.p, err_d);
    CK(cudaMemcpyAsync(&err, err_d, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (err) FAIL(REMO_ERR_STATE, "ebe_check: %d violations in the batch tables", err);
  }
  c->have_ebe = true;
}

template <int NLD, int NG>
void launch_t(Ctx* c, const double* P, int pstride, double* Q, int ks, int nr) {
  cudaStream_t st = c->stream;
  const int xst = nr | 1;
  const int grid = ebe_grid(c, nr);
  static const int pf = [] { const char* e = getenv("REMO_EBE_PREFETCH"); return e ? atoi(e) : 1; }();
  const size_t sm = ebe_smem(c->ebe_umax, nr);
  if constexpr (NLD == 20) {
    if (c->ebe_p3_ctas >= 4) {
      k_spmm_ebe<NLD, NG, 4><<<grid, Batch<NLD>::TPB, sm, st>>>((int)c->ebe_nb, c->ebe_uoff.p, c->ebe_udof.p, c->ebe_lidx.p, c->ebe_lpos.p, c->ebe_ucnt.p,
                                                              c->ebe_jd.p, c->ebe_gm.p, P, pstride, Q, ks, nr, xst, c->ebe_umax, c->ebe_fast8, pf, c->partial.p);
      return;
    }
  }
  k_spmm_ebe<NLD, NG><<<grid, Batch<NLD>::TPB, sm, st>>>((int)c->ebe_nb, c->ebe_uoff.p, c->ebe_udof.p, c->ebe_lidx.p, c->ebe_lpos.p, c->ebe_ucnt.p, c->ebe_jd.p,
                                                       c->ebe_gm.p, P, pstride, Q, ks, nr, xst, c->ebe_umax, c->ebe_fast8, pf, c->partial.p);
}

}  // namespace

void ebe_build(Ctx* c) {
  if (c->dim == 2) ebe_build_t<10, 18>(c);  // order-3 triangles
  else if (c->order == 3) ebe_build_t<20, 10>(c);
  else ebe_build_t<10, 10>(c);
}

void launch_spmm_ebe(Ctx* c, const double* P, int pstride, double* Q, int ks, int nr) {
  CK(cudaMemsetAsync(Q, 0, (size_t)c->ndof * ks * sizeof(double), c->stream));
  if (c->ebe_ng == 18) launch_t<10, 18>(c, P, pstride, Q, ks, nr);
  else if (c->ebe_nld == 20) launch_t<20, 10>(c, P, pstride, Q, ks, nr);
  else launch_t<10, 10>(c, P, pstride, Q, ks, nr);
  c->launches += 2;
  CK(cudaGetLastError());
}
