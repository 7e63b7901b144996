// Symbolic phase on the GPU: topology (edges, faces), dof numbering, Dirichlet dofs, dof -> element
// adjacency and the CSR sparsity pattern.  Replaces the space construction
// `fes = ngs.H1(mesh, order=3, dirichlet=...)` (/root/reference/remo3d/ngsolve_functions.py:27) and the
// sparsity-graph part of `a.Assemble()` (:47).  Numbering contract: SURVEY.md section 10.2 /
// oracle/fem_oracle.py `Space` (edges and faces numbered lexicographically by sorted vertex tuples).
//
// Everything is integer work; sorting / scanning uses CUB device primitives, the rest are small
// hand-written kernels.  The only host synchronisations are the reads of ne, nf and nnz (needed to
// size the allocations).
#include <cub/cub.cuh>

#include "space_view.cuh"

namespace {

constexpr int TB = 256;

__device__ __forceinline__ void cswap(int32_t& a, int32_t& b) {
  if (a > b) { int32_t t = a; a = b; b = t; }
}

__global__ void k_sort_verts(const int32_t* __restrict__ elems, int32_t* __restrict__ sv, int64_t nt, int nvl) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nt) return;
  if (nvl == 4) {
    int4 e = reinterpret_cast<const int4*>(elems)[t];
    int32_t a = e.x, b = e.y, c = e.z, d = e.w;
    cswap(a, b); cswap(c, d); cswap(a, c); cswap(b, d); cswap(b, c);
    reinterpret_cast<int4*>(sv)[t] = make_int4(a, b, c, d);
  } else {
    int32_t a = elems[3 * t], b = elems[3 * t + 1], c = elems[3 * t + 2];
    cswap(a, b); cswap(b, c); cswap(a, b);
    sv[3 * t] = a; sv[3 * t + 1] = b; sv[3 * t + 2] = c;
  }
}

// one key per (element, local edge): (a<<32)|b, payload = flat index into elem_edges
__global__ void k_edge_keys(const int32_t* __restrict__ sv, uint64_t* __restrict__ keys, uint32_t* __restrict__ pay,
                            int64_t nt, int nvl, int nle) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nt * nle) return;
  int64_t t = i / nle;
  int le = (int)(i - t * nle);
  int a = (nvl == 4) ? LE3[le][0] : LE2[le][0];
  int b = (nvl == 4) ? LE3[le][1] : LE2[le][1];
  keys[i] = ((uint64_t)(uint32_t)sv[t * nvl + a] << 32) | (uint32_t)sv[t * nvl + b];
  pay[i] = (uint32_t)i;
}

// faces of tets: key = (edge(i,j) << 32) | k ; lexicographic in (i,j,k) because edge numbers are lexicographic in (i,j)
__global__ void k_face_keys(const int32_t* __restrict__ sv, const int32_t* __restrict__ elem_edges,
                            uint64_t* __restrict__ keys, uint32_t* __restrict__ pay, int64_t nt) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nt * 4) return;
  int64_t t = i >> 2;
  int lf = (int)(i & 3);
  uint32_t e = (uint32_t)elem_edges[t * 6 + LF3_EDGE[lf]];
  keys[i] = ((uint64_t)e << 32) | (uint32_t)sv[t * 4 + LF3_V[lf][2]];
  pay[i] = (uint32_t)i;
}

__global__ void k_flag_first(const uint64_t* __restrict__ keys, int32_t* __restrict__ flag, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// id = inclusive_scan(flag) - 1 ; scatter ids back to the (element, local) slots and keep the unique keys
__global__ void k_assign_ids(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ pay,
                             const int32_t* __restrict__ scan, int32_t* __restrict__ elem_ids,
                             uint64_t* __restrict__ uniq, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t id = scan[i] - 1;
  elem_ids[pay[i]] = id;
  if (i == 0 || keys[i] != keys[i - 1]) uniq[id] = keys[i];
}

__global__ void k_mark_dirichlet(SpaceView s, const int32_t* __restrict__ bfacets, const uint8_t* __restrict__ bdir,
                                 int64_t nb, uint8_t* __restrict__ constrained, int* __restrict__ bad) {
  int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= nb || !bdir[f]) return;
  const int d = s.dim;
  int32_t v[3];
  for (int i = 0; i < d; i++) v[i] = bfacets[f * d + i];
  if (d == 3) { cswap(v[0], v[1]); cswap(v[1], v[2]); cswap(v[0], v[1]); } else { cswap(v[0], v[1]); }
  for (int i = 0; i < d; i++) constrained[v[i]] = 1;
  if (s.order == 1) return;
  const int pe = s.order - 1;
  int64_t e01 = -1;
  for (int i = 0; i < d; i++)
    for (int j = i + 1; j < d; j++) {
      int64_t e = find_edge(s, v[i], v[j]);
      if (e < 0) { atomicExch(bad, 1); continue; }
      if (i == 0 && j == 1) e01 = e;
      for (int k = 0; k < pe; k++) constrained[s.edge_base + pe * e + k] = 1;
    }
  if (s.order == 3 && d == 3 && e01 >= 0) {
    uint64_t key = ((uint64_t)e01 << 32) | (uint32_t)v[2];
    int64_t pos = lower_bound_u64(s.face_keys, s.nf, key);
    if (pos < s.nf && s.face_keys[pos] == key) constrained[s.face_base + pos] = 1; else atomicExch(bad, 1);
  }
}

// adjacency entries: key = global dof, payload = element*nld + local dof
__global__ void k_adj_pairs(SpaceView s, uint32_t* __restrict__ keys, uint32_t* __restrict__ pay, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t t = i / s.nld;
  int b = (int)(i - t * s.nld);
  keys[i] = (uint32_t)elem_dof(s, t, b);
  pay[i] = (uint32_t)i;
}

// adj_ptr[d] = first position whose key >= d (keys sorted); handles dofs without elements
__global__ void k_adj_ptr(const uint32_t* __restrict__ keys, int64_t n, int64_t ndof, int64_t* __restrict__ ptr) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > n) return;
  int64_t prev = (i == 0) ? -1 : (int64_t)keys[i - 1];
  int64_t cur = (i == n) ? ndof : (int64_t)keys[i];
  for (int64_t d = prev + 1; d <= cur; d++) ptr[d] = i;
}

// ---- CSR pattern.  Row d = sorted distinct dofs of the elements adjacent to d.  A group (a warp; a whole CTA for the
// rare rows with more than WARP_CAP candidates) builds the candidate list of its row in shared memory straight from the
// adjacency (nld dofs per adjacent element: 73 candidates per row on average at order 2 for 28 distinct columns), flags
// the repeats (bit 31), ranks the distinct values by counting the smaller ones and writes them, sorted, to the head of
// the row's segment of `out`; `count[d]` = distinct columns.  O(n^2) compares per row from broadcast 16-byte
// shared-memory reads: ~10 ms at 4.8 M dofs where the segmented radix sort of the 354 M candidates took 43 ms, and
// no candidate buffers in global memory (2 x 1.4 GB) are needed any more.
constexpr int WARP_CAP = 2048;    // candidates per warp-row (8 KB of shared memory per warp)
constexpr int CTA_CAP = 49152;    // candidates of a row handled by a whole CTA (192 KB)
constexpr uint32_t DUP = 0x80000000u;

template <bool CTA>
__device__ __forceinline__ void group_sync() {
  if (CTA) __syncthreads(); else __syncwarp();
}

// buf: n4 = n rounded up to 4 entries (the tail holds 0xffffffff).  Returns this thread's number of distinct values.
template <bool CTA>
__device__ __forceinline__ int row_unique(uint32_t* buf, int n, int tid, int nth, int32_t* __restrict__ out) {
  const int n4 = (n + 3) & ~3;
  const uint4* b4 = reinterpret_cast<const uint4*>(buf);
  // 1. flag every value that already occurs at a smaller position (the first occurrence is never flagged, so reading
  //    entries that other threads are flagging meanwhile is harmless: the flag bit is masked out of the comparison)
  for (int base = 0; base < n; base += nth) {
    const int i = base + tid;
    const uint32_t v = i < n ? (buf[i] & ~DUP) : 0xfffffffeu;
    bool dup = false;
    const int jend = min(n4, (base + nth + 3) & ~3);
    for (int j = 0; j < jend; j += 4) {
      const uint4 q = b4[j >> 2];
      dup |= (j < i && (q.x & ~DUP) == v) | (j + 1 < i && (q.y & ~DUP) == v) | (j + 2 < i && (q.z & ~DUP) == v) | (j + 3 < i && (q.w & ~DUP) == v);
    }
    if (i < n && dup) buf[i] = v | DUP;
  }
  group_sync<CTA>();
  // 2. rank of a distinct value = number of distinct values below it (flagged entries and the tail compare as huge)
  int mine = 0;
  for (int i = tid; i < n; i += nth) {
    const uint32_t v = buf[i];
    if (v & DUP) continue;
    int rank = 0;
    for (int j = 0; j < n4; j += 4) {
      const uint4 q = b4[j >> 2];
      rank += (q.x < v) + (q.y < v) + (q.z < v) + (q.w < v);
    }
    out[rank] = (int32_t)v;
    mine++;
  }
  return mine;
}

template <bool CTA>
__device__ __forceinline__ void row_candidates(const SpaceView& s, const uint32_t* __restrict__ adj, int64_t a0, int n, int tid,
                                               int nth, uint32_t* buf) {
  for (int i = tid; i < ((n + 3) & ~3); i += nth) {
    uint32_t v = 0xffffffffu;
    if (i < n) {
      const int a = i / s.nld, b = i - a * s.nld;
      v = (uint32_t)elem_dof(s, adj[a0 + a] / s.nld, b);
    }
    buf[i] = v;
  }
  group_sync<CTA>();
}

__global__ void __launch_bounds__(256) k_pattern_warp(SpaceView s, const int64_t* __restrict__ adj_ptr, const uint32_t* __restrict__ adj,
                                                      int32_t* __restrict__ out, int32_t* __restrict__ count, int* __restrict__ nbig) {
  extern __shared__ __align__(16) uint32_t smw[];  // 8 warps x WARP_CAP
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t* mybuf = smw + w * WARP_CAP;
  for (int64_t row = (int64_t)blockIdx.x * 8 + w; row < s.ndof; row += (int64_t)gridDim.x * 8) {
    const int64_t a0 = adj_ptr[row];
    const int64_t nn = (adj_ptr[row + 1] - a0) * s.nld;
    if (nn > WARP_CAP) {  // left to k_pattern_cta
      if (lane == 0) atomicAdd(nbig, 1);
      continue;
    }
    const int n = (int)nn;
    row_candidates<false>(s, adj, a0, n, lane, 32, mybuf);
    int mine = row_unique<false>(mybuf, n, lane, 32, out + a0 * s.nld);
    for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if (lane == 0) count[row] = mine;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) k_pattern_cta(SpaceView s, const int64_t* __restrict__ adj_ptr, const uint32_t* __restrict__ adj,
                                                     int32_t* __restrict__ out, int32_t* __restrict__ count, int* __restrict__ bad) {
  extern __shared__ __align__(16) uint32_t big[];
  __shared__ int total;
  for (int64_t row = blockIdx.x; row < s.ndof; row += gridDim.x) {
    const int64_t a0 = adj_ptr[row];
    const int64_t nn = (adj_ptr[row + 1] - a0) * s.nld;
    if (nn <= WARP_CAP) continue;
    if (nn > CTA_CAP) {
      if (threadIdx.x == 0) { atomicExch(bad, 1); count[row] = 0; }
      continue;
    }
    if (threadIdx.x == 0) total = 0;
    const int n = (int)nn;
    row_candidates<true>(s, adj, a0, n, threadIdx.x, 256, big);
    const int mine = row_unique<true>(big, n, threadIdx.x, 256, out + a0 * s.nld);
    atomicAdd(&total, mine);
    __syncthreads();
    if (threadIdx.x == 0) count[row] = total;
    __syncthreads();
  }
}

// col[rowptr[d] ..] = head of the row's segment
__global__ void k_pattern_copy(const int64_t* __restrict__ adj_ptr, int nld, const int32_t* __restrict__ out,
                               const int64_t* __restrict__ rowptr, int64_t ndof, int32_t* __restrict__ col) {
  const int lane = threadIdx.x & 31;
  for (int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; row < ndof; row += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const int64_t r0 = rowptr[row], n = rowptr[row + 1] - r0;
    const int32_t* src = out + adj_ptr[row] * nld;
    for (int64_t i = lane; i < n; i += 32) col[r0 + i] = src[i];
  }
}

__global__ void k_count_to_ptr(const int64_t* __restrict__ incl, int64_t n, int64_t* __restrict__ ptr) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i <= n) ptr[i] = i == 0 ? 0 : incl[i - 1];
}

__global__ void k_widen(const int32_t* __restrict__ a, int64_t n, int64_t* __restrict__ b) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) b[i] = a[i];
}

__global__ void k_unpack_edges(const uint64_t* __restrict__ keys, int64_t ne, int32_t* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= ne) return;
  out[2 * i] = (int32_t)(keys[i] >> 32);
  out[2 * i + 1] = (int32_t)(keys[i] & 0xffffffffu);
}

__global__ void k_unpack_faces(const uint64_t* __restrict__ fkeys, const uint64_t* __restrict__ ekeys, int64_t nf,
                               int32_t* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nf) return;
  uint64_t ek = ekeys[fkeys[i] >> 32];
  out[3 * i] = (int32_t)(ek >> 32);
  out[3 * i + 1] = (int32_t)(ek & 0xffffffffu);
  out[3 * i + 2] = (int32_t)(fkeys[i] & 0xffffffffu);
}

int bits_for(uint64_t n) {
  int b = 1;
  while (b < 64 && (n >> b)) b++;
  return b;
}

// sort (key,payload) pairs, number the distinct keys 0.. in ascending order; returns the count
template <typename KeyT>
void sort_pairs(Ctx* c, const KeyT* keys, const uint32_t* pay, int64_t n, int end_bit, KeyT* keys_sorted, uint32_t* pay_sorted) {
  size_t bytes = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys, keys_sorted, pay, pay_sorted, n, 0, end_bit, c->stream));
  c->tmp.ensure(bytes, c->stream);
  CK(cub::DeviceRadixSort::SortPairs(c->tmp.p, bytes, keys, keys_sorted, pay, pay_sorted, n, 0, end_bit, c->stream));
  c->launches += 4;
}

int32_t scan_flags(Ctx* c, const int32_t* flag, int32_t* scan, int64_t n) {
  size_t bytes = 0;
  CK(cub::DeviceScan::InclusiveSum(nullptr, bytes, flag, scan, n, c->stream));
  c->tmp.ensure(bytes, c->stream);
  CK(cub::DeviceScan::InclusiveSum(c->tmp.p, bytes, flag, scan, n, c->stream));
  c->launches += 2;
  int32_t total = 0;
  CK(cudaMemcpyAsync(&total, scan + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return total;
}

struct ScaleOffsets {
  const int64_t* ptr;
  int nld;
  __host__ __device__ int64_t operator()(int64_t i) const { return ptr[i] * nld; }
};

}  // namespace

void space_build(Ctx* c, int order) {
  if (!c->have_mesh) FAIL(REMO_ERR_STATE, "remo_space_build: no mesh (call remo_mesh_set first)");
  if (order < 1 || order > 3) FAIL(REMO_ERR_ARG, "remo_space_build: order must be 1, 2 or 3 (got %d)", order);
  StageTimer timer(c, ST_SPACE);
  cudaStream_t st = c->stream;
  const int dim = c->dim, nvl = dim + 1;
  const int64_t nt = c->nt;
  c->order = order;
  c->nle = (dim == 3) ? 6 : 3;
  c->nlf = (dim == 3) ? 4 : 1;
  c->npair = nvl * (nvl + 1) / 2;
  c->nld = nvl + c->nle * (order - 1) + (order == 3 ? c->nlf : 0);
  c->have_space = c->have_matrix = c->have_sell = c->have_ebe = c->have_pattern = c->have_values = false;
  c->pkind = -1;

  // 1. sorted element vertices
  c->sv.ensure(nt * nvl, st);
  LAUNCH(c, k_sort_verts, grid_for(nt, TB), TB, 0, c->elems.p, c->sv.p, nt, nvl);

  // scratch slots: 0 keys, 1 sorted keys, 2 payload, 3 sorted payload, 4 flags, 5 scan, 6/7 candidate columns
  const int64_t nadj = nt * c->nld;
  if (nadj >= (int64_t)1 << 32) FAIL(REMO_ERR_ARG, "remo_space_build: mesh too large (nt*nld >= 2^32)");
  const int64_t nkeys = std::max<int64_t>(nt * c->nle, nt * 4);
  uint64_t* k64 = scratch<uint64_t>(c, 0, std::max<int64_t>(nkeys, (nadj + 1) / 2));
  uint64_t* k64s = scratch<uint64_t>(c, 1, std::max<int64_t>(nkeys, (nadj + 1) / 2));
  uint32_t* pay = scratch<uint32_t>(c, 2, std::max(nkeys, nadj));
  uint32_t* pays = scratch<uint32_t>(c, 3, std::max(nkeys, nadj));
  int32_t* flag = scratch<int32_t>(c, 4, nkeys);
  int32_t* scan = scratch<int32_t>(c, 5, nkeys);

  // 2. edges (always built: order >= 2 needs the dofs, order 1 needs nothing but the cost is small and
  //    remo_topology_get exports them)
  {
    const int64_t n = nt * c->nle;
    LAUNCH(c, k_edge_keys, grid_for(n, TB), TB, 0, c->sv.p, k64, pay, nt, nvl, c->nle);
    sort_pairs(c, k64, pay, n, 32 + bits_for((uint64_t)c->nv), k64s, pays);
    LAUNCH(c, k_flag_first, grid_for(n, TB), TB, 0, k64s, flag, n);
    c->ne = scan_flags(c, flag, scan, n);
    c->edge_keys.ensure(c->ne, st);
    c->elem_edges.ensure(n, st);
    LAUNCH(c, k_assign_ids, grid_for(n, TB), TB, 0, k64s, pays, scan, c->elem_edges.p, c->edge_keys.p, n);
  }
  // 3. faces (3D order 3 only); in 2D the order-3 cell bubbles are numbered by element
  c->nf = 0;
  if (order == 3 && dim == 3) {
    const int64_t n = nt * 4;
    LAUNCH(c, k_face_keys, grid_for(n, TB), TB, 0, c->sv.p, c->elem_edges.p, k64, pay, nt);
    sort_pairs(c, k64, pay, n, 32 + bits_for((uint64_t)c->ne), k64s, pays);
    LAUNCH(c, k_flag_first, grid_for(n, TB), TB, 0, k64s, flag, n);
    c->nf = scan_flags(c, flag, scan, n);
    c->face_keys.ensure(c->nf, st);
    c->elem_faces.ensure(n, st);
    LAUNCH(c, k_assign_ids, grid_for(n, TB), TB, 0, k64s, pays, scan, c->elem_faces.p, c->face_keys.p, n);
  } else if (order == 3 && dim == 2) {
    c->nf = nt;
  }
  c->edge_base = c->nv;
  c->face_base = c->nv + (int64_t)(order - 1) * c->ne;
  c->ndof = c->face_base + (order == 3 ? c->nf : 0);
  if (c->ndof >= (int64_t)1 << 31) FAIL(REMO_ERR_ARG, "remo_space_build: %lld dofs exceed the int32 column index", (long long)c->ndof);

  SpaceView sview = make_view(c);
  int* bad = scratch<int>(c, 8, 4);

  // 4. Dirichlet dofs
  c->constrained.ensure(c->ndof, st);
  CK(cudaMemsetAsync(c->constrained.p, 0, c->ndof, st));
  if (c->nb > 0) {
    CK(cudaMemsetAsync(bad, 0, sizeof(int), st));
    LAUNCH(c, k_mark_dirichlet, grid_for(c->nb, TB), TB, 0, sview, c->bfacets.p, c->bdir.p, c->nb, c->constrained.p, bad);
    int hbad = 0;
    CK(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (hbad) FAIL(REMO_ERR_MESH, "remo_space_build: a Dirichlet boundary facet is not a face of the mesh");
  }

  // 5. dof -> (element, local dof) adjacency, elements ascending within a dof (stable radix sort)
  c->nadj = nadj;
  uint32_t* akey = reinterpret_cast<uint32_t*>(k64);
  uint32_t* akeys = reinterpret_cast<uint32_t*>(k64s);
  LAUNCH(c, k_adj_pairs, grid_for(nadj, TB), TB, 0, sview, akey, pay, nadj);
  c->adj.ensure(nadj, st);
  sort_pairs(c, akey, pay, nadj, bits_for((uint64_t)c->ndof), akeys, c->adj.p);
  c->adj_ptr.ensure(c->ndof + 1, st);
  LAUNCH(c, k_adj_ptr, grid_for(nadj + 1, TB), TB, 0, akeys, nadj, c->ndof, c->adj_ptr.p);

  c->have_space = true;
  c->have_pattern = false;
  c->nnz = 0;
  // 6. CSR pattern: only when somebody needs the assembled matrix (pattern_build); the element-wise PCG path never does
  if (!c->lazy_matrix) pattern_build(c);
}

// CSR pattern of the whole matrix: per row the sorted distinct dofs of the adjacent elements (k_pattern_*).  Idempotent.
void pattern_build(Ctx* c) {
  if (!c->have_space) FAIL(REMO_ERR_STATE, "no space (call remo_space_build first)");
  if (c->have_pattern) return;
  cudaStream_t st = c->stream;
  SpaceView sview = make_view(c);
  const int64_t ncand = c->nadj * c->nld;
  if (ncand >= (int64_t)1 << 31) FAIL(REMO_ERR_ARG, "CSR pattern: %lld candidate entries exceed the 2^31 limit of the pattern builder", (long long)ncand);
  {
    int32_t* seg = scratch<int32_t>(c, 6, ncand);          // row d's columns at the head of [adj_ptr[d] * nld, ...)
    int32_t* count = scratch<int32_t>(c, 4, c->ndof + 1);
    int64_t* count64 = scratch<int64_t>(c, 5, c->ndof + 1);
    int64_t* incl = scratch<int64_t>(c, 7, c->ndof + 1);
    int* flags = scratch<int>(c, 8, 4);
    CK(cudaMemsetAsync(flags, 0, 4 * sizeof(int), st));
    allow_max_smem(k_pattern_warp, c->device);  // per (kernel, device), see ctx.cuh
    LAUNCH(c, k_pattern_warp, c->num_sms * 6, 256, 8 * WARP_CAP * sizeof(uint32_t), sview, c->adj_ptr.p, c->adj.p, seg, count, flags);
    int hf[2] = {0, 0};
    CK(cudaMemcpyAsync(hf, flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (hf[0] > 0) {  // rows with more than WARP_CAP candidates: one CTA per row
      allow_max_smem(k_pattern_cta, c->device);
      LAUNCH(c, k_pattern_cta, c->num_sms, 256, CTA_CAP * sizeof(uint32_t), sview, c->adj_ptr.p, c->adj.p, seg, count, flags + 1);
      CK(cudaMemcpyAsync(hf, flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (hf[1]) FAIL(REMO_ERR_MESH, "remo_space_build: a dof is shared by more than %d elements", CTA_CAP / c->nld);
    }
    LAUNCH(c, k_widen, grid_for(c->ndof, TB), TB, 0, count, c->ndof, count64);
    size_t bytes = 0;
    CK(cub::DeviceScan::InclusiveSum(nullptr, bytes, count64, incl, c->ndof, st));
    c->tmp.ensure(bytes, st);
    CK(cub::DeviceScan::InclusiveSum(c->tmp.p, bytes, count64, incl, c->ndof, st));
    c->launches += 2;
    int64_t nnz = 0;
    CK(cudaMemcpyAsync(&nnz, incl + (c->ndof - 1), sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (nnz >= (int64_t)1 << 31) FAIL(REMO_ERR_ARG, "remo_space_build: %lld non-zeros exceed the 2^31 limit", (long long)nnz);
    c->nnz = nnz;
    c->rowptr.ensure(c->ndof + 1, st);
    c->col.ensure(c->nnz, st);
    c->val.ensure(c->nnz, st);
    LAUNCH(c, k_count_to_ptr, grid_for(c->ndof + 1, TB), TB, 0, incl, c->ndof, c->rowptr.p);
    LAUNCH(c, k_pattern_copy, c->num_sms * 8, 256, 0, c->adj_ptr.p, c->nld, seg, c->rowptr.p, c->ndof, c->col.p);
  }
  c->have_pattern = true;
}

namespace {

__global__ void k_vv_keys(const uint64_t* __restrict__ edge_keys, int64_t ne, uint64_t* __restrict__ rev) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e < ne) rev[e] = (edge_keys[e] << 32) | (edge_keys[e] >> 32);  // (b, a): neighbours with a smaller number, by row b
}

// row a of the vertex graph: [neighbours c < a, ascending | a | neighbours b > a, ascending]; its first entry sits at
// a + (edges whose smaller vertex is < a) + (edges whose larger vertex is < a) -- no scan needed
__global__ void k_vv_pattern(const uint64_t* __restrict__ edge_keys, const uint64_t* __restrict__ rev, int64_t ne, int64_t nv,
                             int64_t* __restrict__ rowptr, int32_t* __restrict__ col) {
  int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (a > nv) return;
  const int64_t u0 = lower_bound_u64(edge_keys, ne, (uint64_t)a << 32), l0 = lower_bound_u64(rev, ne, (uint64_t)a << 32);
  const int64_t o = a + u0 + l0;
  rowptr[a] = o;
  if (a == nv) return;
  const int64_t u1 = lower_bound_u64(edge_keys, ne, (uint64_t)(a + 1) << 32), l1 = lower_bound_u64(rev, ne, (uint64_t)(a + 1) << 32);
  int64_t w = o;
  for (int64_t j = l0; j < l1; j++) col[w++] = (int32_t)(rev[j] & 0xffffffffu);
  col[w++] = (int32_t)a;
  for (int64_t j = u0; j < u1; j++) col[w++] = (int32_t)(edge_keys[j] & 0xffffffffu);
}

}  // namespace

// CSR pattern of the vertex block (= the vertex graph of the mesh + the diagonal) from the edge list.  Returns its nnz.
int64_t vertex_block_pattern(Ctx* c, DBuf<int64_t>& rowptr, DBuf<int32_t>& col) {
  cudaStream_t st = c->stream;
  const int64_t ne = c->ne, nv = c->nv, nnz = nv + 2 * ne;
  uint64_t* rev = scratch<uint64_t>(c, 0, ne);
  uint64_t* revs = scratch<uint64_t>(c, 1, ne);
  LAUNCH(c, k_vv_keys, grid_for(ne, TB), TB, 0, c->edge_keys.p, ne, rev);
  size_t bytes = 0;
  const int end_bit = std::min(64, 32 + bits_for((uint64_t)nv));
  CK(cub::DeviceRadixSort::SortKeys(nullptr, bytes, rev, revs, ne, 0, end_bit, st));
  c->tmp.ensure(bytes, st);
  CK(cub::DeviceRadixSort::SortKeys(c->tmp.p, bytes, rev, revs, ne, 0, end_bit, st));
  c->launches += 2;
  rowptr.ensure(nv + 1, st);
  col.ensure(nnz, st);
  LAUNCH(c, k_vv_pattern, grid_for(nv + 1, TB), TB, 0, c->edge_keys.p, revs, ne, nv, rowptr.p, col.p);
  return nnz;
}


void topology_get(Ctx* c, int32_t* edges, int32_t* faces, int32_t* elem_edges, int32_t* elem_faces) {
  if (!c->have_space) FAIL(REMO_ERR_STATE, "remo_topology_get: no space (call remo_space_build first)");
  cudaStream_t st = c->stream;
  if (edges && c->ne) {
    DBuf<int32_t> t;
    t.ensure(2 * c->ne, st);
    LAUNCH(c, k_unpack_edges, grid_for(c->ne, TB), TB, 0, c->edge_keys.p, c->ne, t.p);
    CK(cudaMemcpyAsync(edges, t.p, 2 * c->ne * sizeof(int32_t), cudaMemcpyDefault, st));
    CK(cudaStreamSynchronize(st));
    t.release(st);
  }
  if (faces && c->nf && c->dim == 3) {
    DBuf<int32_t> t;
    t.ensure(3 * c->nf, st);
    LAUNCH(c, k_unpack_faces, grid_for(c->nf, TB), TB, 0, c->face_keys.p, c->edge_keys.p, c->nf, t.p);
    CK(cudaMemcpyAsync(faces, t.p, 3 * c->nf * sizeof(int32_t), cudaMemcpyDefault, st));
    CK(cudaStreamSynchronize(st));
    t.release(st);
  }
  if (elem_edges) CK(cudaMemcpyAsync(elem_edges, c->elem_edges.p, c->nt * c->nle * sizeof(int32_t), cudaMemcpyDefault, st));
  if (elem_faces && c->order == 3 && c->dim == 3)
    CK(cudaMemcpyAsync(elem_faces, c->elem_faces.p, c->nt * 4 * sizeof(int32_t), cudaMemcpyDefault, st));
  CK(cudaStreamSynchronize(st));
}
