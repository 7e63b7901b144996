// Device-side view of the discrete space: local -> global dof map and axis evaluation.
#pragma once
#include <stdint.h>

#include "ctx.cuh"

struct SpaceView {
  int dim, order, nvl, nle, nlf, nld;
  int64_t nv, nt, ne, nf, ndof, edge_base, face_base;
  const int32_t* sv;          // nt x nvl sorted vertices
  const int32_t* elem_edges;  // nt x nle
  const int32_t* elem_faces;  // nt x nlf (3D only)
  const uint64_t* edge_keys;  // ne
  const uint64_t* face_keys;  // nf (3D, order 3)
  // axis
  int64_t naxis;
  const int32_t* axis_v;
  const double* axis_z;
};

// local edges / faces of the sorted simplex (same order as oracle/fem_oracle.py local_edges/local_faces)
static __device__ __constant__ const int8_t LE3[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
static __device__ __constant__ const int8_t LE2[3][2] = {{0, 1}, {0, 2}, {1, 2}};
// face (i,j,k): the local edge (i,j) it extends and the third vertex k
static __device__ __constant__ const int8_t LF3_EDGE[4] = {0, 0, 1, 3};
static __device__ __constant__ const int8_t LF3_V[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};

__device__ __forceinline__ int64_t elem_dof(const SpaceView& s, int64_t t, int b) {
  if (b < s.nvl) return s.sv[t * s.nvl + b];
  b -= s.nvl;
  const int pe = s.order - 1;
  if (b < s.nle * pe) {
    const int le = b / pe, k = b - le * pe;
    return s.edge_base + (int64_t)pe * s.elem_edges[t * s.nle + le] + k;
  }
  const int lf = b - s.nle * pe;
  if (s.dim == 2) return s.face_base + t;  // cell bubble, numbered by element
  return s.face_base + s.elem_faces[t * s.nlf + lf];
}

__device__ __forceinline__ int64_t lower_bound_u64(const uint64_t* a, int64_t n, uint64_t key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// global edge number of the vertex pair (a<b), -1 if the mesh has no such edge
__device__ __forceinline__ int64_t find_edge(const SpaceView& s, int32_t a, int32_t b) {
  const uint64_t key = ((uint64_t)(uint32_t)a << 32) | (uint32_t)b;
  const int64_t pos = lower_bound_u64(s.edge_keys, s.ne, key);
  return (pos < s.ne && s.edge_keys[pos] == key) ? pos : -1;
}

// 21 bits -> every third bit (Morton codes of dof / element locations)
__device__ __forceinline__ uint64_t spread21s(uint64_t v) {
  v &= 0x1fffffull;
  v = (v | (v << 32)) & 0x1f00000000ffffull;
  v = (v | (v << 16)) & 0x1f0000ff0000ffull;
  v = (v | (v << 8)) & 0x100f00f00f00f00full;
  v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}

#define REMO_SNAP_TOL 1e-9

// Non-zero basis functions at the axis point z (oracle Axis.shape).
// Returns the number of (dof, value) pairs (1..4), 0 if z is outside the axis, -1 if the two axis
// vertices around z are not joined by a mesh edge.
__device__ __forceinline__ int axis_shape(const SpaceView& s, double z, int64_t dof[4], double val[4]) {
  int64_t lo = 0, hi = s.naxis;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (s.axis_z[mid] < z) lo = mid + 1; else hi = mid;
  }
  const int64_t i = lo;  // first axis vertex with z_i >= z
  if (i > 0 && fabs(s.axis_z[i - 1] - z) <= REMO_SNAP_TOL) { dof[0] = s.axis_v[i - 1]; val[0] = 1.0; return 1; }
  if (i < s.naxis && fabs(s.axis_z[i] - z) <= REMO_SNAP_TOL) { dof[0] = s.axis_v[i]; val[0] = 1.0; return 1; }
  if (i == 0 || i == s.naxis) return 0;
  const int32_t v0 = s.axis_v[i - 1], v1 = s.axis_v[i];
  const double t = (z - s.axis_z[i - 1]) / (s.axis_z[i] - s.axis_z[i - 1]);
  int32_t a, b;
  double la, lb;
  if (v0 < v1) { a = v0; b = v1; la = 1.0 - t; lb = t; } else { a = v1; b = v0; la = t; lb = 1.0 - t; }
  dof[0] = a; val[0] = la;
  dof[1] = b; val[1] = lb;
  if (s.order == 1) return 2;
  const int64_t e = find_edge(s, a, b);
  if (e < 0) return -1;
  const int pe = s.order - 1;
  dof[2] = s.edge_base + pe * e; val[2] = la * lb;
  if (s.order == 2) return 3;
  dof[3] = s.edge_base + pe * e + 1; val[3] = la * lb * (lb - la);
  return 4;
}

static inline SpaceView make_view(const Ctx* c) {
  SpaceView s;
  s.dim = c->dim; s.order = c->order; s.nvl = c->dim + 1; s.nle = c->nle; s.nlf = c->nlf; s.nld = c->nld;
  s.nv = c->nv; s.nt = c->nt; s.ne = c->ne; s.nf = c->nf; s.ndof = c->ndof;
  s.edge_base = c->edge_base; s.face_base = c->face_base;
  s.sv = c->sv.p; s.elem_edges = c->elem_edges.p; s.elem_faces = c->elem_faces.p;
  s.edge_keys = c->edge_keys.p; s.face_keys = c->face_keys.p;
  s.naxis = c->naxis; s.axis_v = c->axis_v.p; s.axis_z = c->axis_z.p;
  return s;
}
