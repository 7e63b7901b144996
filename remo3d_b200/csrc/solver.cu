// Preconditioned conjugate gradients for all right-hand sides of one mesh at once, point sources,
// axis sampling and apparent resistivity.  Replaces
//   c = ngs.Preconditioner(a, preconditioner); inv = ngs.CGSolver(a.mat, c.mat, maxsteps=1000);
//   gfu.vec.data = inv * f.vec                       (/root/reference/remo3d/ngsolve_functions.py:46-51,
//                                                      device form ngsolve_functions_gpu.py:41-47)
//   AddPointSource                                    (ngsolve_functions.py:10-21, 39-44)
//   gfu(mesh(0.0, 0.0, z)), Ra = |K dU| / 2           (workers/worker.py:113-134)
//
// Data layout: every vector block (F, X, R, Z, P, Q) is row-major ndof x nrhs, so the nrhs values of one dof
// are contiguous: the SpMM gathers one contiguous nrhs*8-byte segment per matrix entry and the matrix
// (12 B per non-zero) is streamed once per iteration for ALL right-hand sides.
// The columns run in lockstep with their own alpha/beta (device-resident scalars, no host round trip per
// iteration); a converged column is frozen (alpha = beta = 0).  Dot products are reduced block-wise into a
// partial array and finished in a fixed order by a one-block kernel -> bit-reproducible runs.
#include <cstdlib>

#include "space_view.cuh"

namespace {

constexpr int TB = 256;
constexpr int KMAX = REMO_MAX_RHS;  // 32
// scalar slots (each KMAX doubles)
enum { S_RZ = 0, S_ALPHA, S_BETA, S_RR, S_BB, S_ACTIVE, S_TOL2, S_PQ, S_NSLOT };

__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// SpMM  Q = A P  (constrained rows -> 0) with the fused per-column dot  p.q
// one warp per row; lane = (jsub, r): KP lanes cover the right-hand sides, 32/KP matrix entries in flight
// ------------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(TB) k_spmm(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                             const double* __restrict__ val, const uint8_t* __restrict__ constrained,
                                             const double* __restrict__ P, double* __restrict__ Q, int k, int64_t n,
                                             double* __restrict__ partial) {
  constexpr int J = 32 / KP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int jsub = lane / KP, r = lane % KP;
  const bool on = r < k;
  const int64_t nwarps = (int64_t)gridDim.x * (TB / 32);
  double dot = 0.0;
  for (int64_t row = (int64_t)blockIdx.x * (TB / 32) + warp; row < n; row += nwarps) {
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    double acc = 0.0;
    for (int64_t j = s + jsub; j < e; j += J) {
      const int32_t c = __ldg(col + j);
      const double v = __ldg(val + j);
      if (on) acc = fma(v, P[(int64_t)c * k + r], acc);
    }
#pragma unroll
    for (int o = 16; o >= KP; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (constrained[row]) acc = 0.0;
    if (jsub == 0 && on) {
      Q[row * k + r] = acc;
      dot = fma(acc, P[row * k + r], dot);
    }
  }
  // block reduction of the per-column dots: lanes with equal r across the warps
  __shared__ double sh[TB / 32][32];
  sh[warp][lane] = (jsub == 0 && on) ? dot : 0.0;
  __syncthreads();
  if (warp == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < TB / 32; w++) t += sh[w][lane];
    if (lane < KP) partial[(int64_t)blockIdx.x * KMAX + lane] = t;
  }
}

// ------------------------------------------------------------------------------------------------
// Variant 3: G-lane group per row.  The plain warp-per-row kernel is LATENCY bound (one dependent
// rowptr -> col/val -> P chain per warp, 64 rows in flight per SM).  Here a group of G lanes (8 for nrhs <= 8)
// owns a row, so 32/G rows are in flight per warp; the group loads G consecutive (col,val) pairs with one
// coalesced request each, broadcasts them by shuffle to its (jsub, r) lanes and issues all G/J gathers of P
// back to back before the FMAs.  No cross-lane reduction when G == KP.
// ------------------------------------------------------------------------------------------------
template <int G, int KP>
__global__ void __launch_bounds__(TB) k_spmm_g(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                               const double* __restrict__ val, const uint8_t* __restrict__ constrained,
                                               const double* __restrict__ P, double* __restrict__ Q, int k, int64_t n,
                                               double* __restrict__ partial) {
  constexpr int J = G / KP;           // matrix entries processed per step by one group
  constexpr int GROUPS = TB / G;      // rows in flight per CTA
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = threadIdx.x % G;     // lane within the group
  const int grp = threadIdx.x / G;
  const int jsub = gl / KP, r = gl % KP;
  const bool on = r < k;
  const int rr = on ? r : 0;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << ((lane / G) * G));
  double dot = 0.0;
  for (int64_t row = (int64_t)blockIdx.x * GROUPS + grp; row < n; row += (int64_t)gridDim.x * GROUPS) {
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    double acc = 0.0;
    // software pipeline: the next chunk's (col,val) are requested before the current chunk's gathers
    int32_t nc = (int32_t)row;
    double nv = 0.0;
    if (s + gl < e) { nc = __ldcs(col + s + gl); nv = __ldcs(val + s + gl); }
    for (int64_t base = s; base < e; base += G) {
      const int32_t myc = nc;
      const double myv = nv;
      nc = (int32_t)row; nv = 0.0;
      if (base + G + gl < e) { nc = __ldcs(col + base + G + gl); nv = __ldcs(val + base + G + gl); }
      double x[KP];
      double vv[KP];
#pragma unroll
      for (int i = 0; i < KP; i++) {
        const int j = i * J + jsub;
        const int32_t c = __shfl_sync(gmask, myc, j, G);
        vv[i] = __shfl_sync(gmask, myv, j, G);
        x[i] = P[(int64_t)c * k + rr];
      }
#pragma unroll
      for (int i = 0; i < KP; i++) acc = fma(vv[i], x[i], acc);
    }
#pragma unroll
    for (int o = G / 2; o >= KP; o >>= 1) acc += __shfl_xor_sync(gmask, acc, o, G);
    if (constrained[row]) acc = 0.0;
    if (jsub == 0 && on) {
      __stcs(Q + row * k + r, acc);
      dot = fma(acc, P[row * k + r], dot);
    }
  }
  // per-column block reduction: threads with jsub == 0 hold column r = gl
  __shared__ double sh[TB];
  sh[threadIdx.x] = (jsub == 0 && on) ? dot : 0.0;
  __syncthreads();
  if (threadIdx.x < KP) {
    double t = 0.0;
    for (int i = threadIdx.x; i < TB; i += G) t += sh[i];
    partial[(int64_t)blockIdx.x * KMAX + threadIdx.x] = t;
  }
  (void)warp;
}

// ------------------------------------------------------------------------------------------------
// Variant 4 (default for nrhs >= 2): as variant 3, but every lane owns TWO adjacent right-hand sides and gathers them
// with one 16-byte load (the row stride of the vector blocks is kept even for this).  An entry then needs KP2 = ks/2
// lanes instead of ks, a 4-lane group owns a row (8 rows in flight per warp) and one shuffle serves twice as many
// matrix entries: the L1 data-pipe cost per entry (profiles/r01_notes.md) drops from ~2.45 to ~2.0 wavefronts.
// ------------------------------------------------------------------------------------------------
template <int G, int KP2>
__global__ void __launch_bounds__(TB) k_spmm_p(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                               const double* __restrict__ val, const uint8_t* __restrict__ constrained,
                                               const double* __restrict__ P, double* __restrict__ Q, int ks, int64_t n,
                                               double* __restrict__ partial) {
  constexpr int J = G / KP2;
  constexpr int GROUPS = TB / G;
  const int lane = threadIdx.x & 31;
  const int gl = threadIdx.x % G, grp = threadIdx.x / G;
  const int jsub = gl / KP2, r2 = gl % KP2;
  const bool on = 2 * r2 < ks;
  const int cc = on ? 2 * r2 : 0;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << ((lane / G) * G));
  double dot0 = 0.0, dot1 = 0.0;
  for (int64_t row = (int64_t)blockIdx.x * GROUPS + grp; row < n; row += (int64_t)gridDim.x * GROUPS) {
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    double acc0 = 0.0, acc1 = 0.0;
    int32_t nc = (int32_t)row;
    double nv = 0.0;
    if (s + gl < e) { nc = __ldg(col + s + gl); nv = __ldg(val + s + gl); }
    for (int64_t base = s; base < e; base += G) {
      const int32_t myc = nc;
      const double myv = nv;
      nc = (int32_t)row; nv = 0.0;
      if (base + G + gl < e) { nc = __ldg(col + base + G + gl); nv = __ldg(val + base + G + gl); }
      double2 x[KP2];
      double vv[KP2];
#pragma unroll
      for (int i = 0; i < KP2; i++) {
        const int j = i * J + jsub;
        const int32_t c = __shfl_sync(gmask, myc, j, G);
        vv[i] = __shfl_sync(gmask, myv, j, G);
        x[i] = *reinterpret_cast<const double2*>(P + (int64_t)c * ks + cc);
      }
#pragma unroll
      for (int i = 0; i < KP2; i++) { acc0 = fma(vv[i], x[i].x, acc0); acc1 = fma(vv[i], x[i].y, acc1); }
    }
#pragma unroll
    for (int o = G / 2; o >= KP2; o >>= 1) {
      acc0 += __shfl_xor_sync(gmask, acc0, o, G);
      acc1 += __shfl_xor_sync(gmask, acc1, o, G);
    }
    if (constrained[row]) { acc0 = 0.0; acc1 = 0.0; }
    if (jsub == 0 && on) {
      const double2 p = *reinterpret_cast<const double2*>(P + row * ks + cc);
      *reinterpret_cast<double2*>(Q + row * ks + cc) = make_double2(acc0, acc1);
      dot0 = fma(acc0, p.x, dot0);
      dot1 = fma(acc1, p.y, dot1);
    }
  }
  __shared__ double sh[2][TB];
  sh[0][threadIdx.x] = (jsub == 0 && on) ? dot0 : 0.0;
  sh[1][threadIdx.x] = (jsub == 0 && on) ? dot1 : 0.0;
  __syncthreads();
  if (threadIdx.x < KP2 && 2 * threadIdx.x < ks) {
    double t0 = 0.0, t1 = 0.0;
    for (int i = threadIdx.x; i < TB; i += G) { t0 += sh[0][i]; t1 += sh[1][i]; }
    partial[(int64_t)blockIdx.x * KMAX + 2 * threadIdx.x] = t0;
    partial[(int64_t)blockIdx.x * KMAX + 2 * threadIdx.x + 1] = t1;
  }
}

// ------------------------------------------------------------------------------------------------
// vector kernels.  The blocks X, R, Z, Q (row stride ks) are walked as FLAT arrays of W-wide elements (W = 2: double2,
// ks even; W = 1 only for a single right-hand side), so every lane is active and every access is a coalesced 16-byte
// one whatever ks is (the former (row, column) thread layout left 2 of 8 lanes idle at ks = 6 and reached 3.9 TB/s; flat:
// ~6 TB/s).  The grid holds a multiple of hc = ks / W threads, so a thread keeps the same column pair for its whole
// stride loop (its alpha / beta and dot accumulators live in registers) and its row index advances by a constant.
// P has its own row stride kps (sell.cu).
// ------------------------------------------------------------------------------------------------
template <int W> struct VecT { typedef double2 T; };
template <> struct VecT<1> { typedef double T; };
__device__ __forceinline__ double lo(double v) { return v; }
__device__ __forceinline__ double hi(double) { return 0.0; }
__device__ __forceinline__ double lo(double2 v) { return v.x; }
__device__ __forceinline__ double hi(double2 v) { return v.y; }
template <int W> __device__ __forceinline__ typename VecT<W>::T mk(double a, double b);
template <> __device__ __forceinline__ double mk<1>(double a, double) { return a; }
template <> __device__ __forceinline__ double2 mk<2>(double a, double b) { return make_double2(a, b); }

// per-column block partials: thread t of block b holds column pair (b * TB + t) % hc; summed in ascending t (fixed order)
template <int W>
__device__ __forceinline__ void flat_partials(double a0, double a1, double b0, double b1, int hc, double* __restrict__ partial,
                                              int nslots) {
  __shared__ double sh[4][TB];
  sh[0][threadIdx.x] = a0; sh[1][threadIdx.x] = a1; sh[2][threadIdx.x] = b0; sh[3][threadIdx.x] = b1;
  __syncthreads();
  if ((int)threadIdx.x < hc) {
    const int first = (int)((hc - (int)(((int64_t)blockIdx.x * TB) % hc) + (int)threadIdx.x) % hc);  // first t with that pair
    double t0 = 0.0, t1 = 0.0, u0 = 0.0, u1 = 0.0;
    for (int i = first; i < TB; i += hc) { t0 += sh[0][i]; t1 += sh[1][i]; u0 += sh[2][i]; u1 += sh[3][i]; }
    const int64_t o = (int64_t)blockIdx.x * 2 * KMAX + W * threadIdx.x;
    partial[o] = t0;
    if (W == 2) partial[o + 1] = t1;
    if (nslots > 1) {
      partial[o + KMAX] = u0;
      if (W == 2) partial[o + KMAX + 1] = u1;
    }
  }
}

// X = 0, R = masked F; rows >= tail_from: P = z = dinv r; rows < tail_from (preconditioned by the V-cycle): r goes to the
// V-cycle's right-hand-side block B0 (row stride kz).  partials: [0] = r.r (= b.b), [1] = r.z over the rows >= tail_from
template <int W>
__global__ void __launch_bounds__(TB) k_init(const double* __restrict__ F, const uint8_t* __restrict__ constrained,
                                             const double* __restrict__ dinv, double* __restrict__ X, double* __restrict__ R,
                                             double* __restrict__ B0, double* __restrict__ P, int ks, int kps, int kz, int64_t n,
                                             int64_t tail_from, double* __restrict__ partial) {
  typedef typename VecT<W>::T V;
  const int hc = ks / W;
  const int64_t T = (int64_t)gridDim.x * TB, gt = (int64_t)blockIdx.x * TB + threadIdx.x;
  const int cp = (int)(gt % hc);
  const int64_t total = n * hc, di = T / hc;
  double rr0 = 0.0, rr1 = 0.0, rz0 = 0.0, rz1 = 0.0;
  int64_t i = gt / hc;
  for (int64_t e = gt; e < total; e += T, i += di) {
    V f = reinterpret_cast<const V*>(F)[e];
    if (constrained[i]) f = mk<W>(0.0, 0.0);
    reinterpret_cast<V*>(X)[e] = mk<W>(0.0, 0.0);
    reinterpret_cast<V*>(R)[e] = f;
    rr0 = fma(lo(f), lo(f), rr0); rr1 = fma(hi(f), hi(f), rr1);
    if (i >= tail_from) {  // rows preconditioned by their diagonal: all rows ("local"), the high-order rows ("multigrid")
      const double d = dinv[i];
      const V z = mk<W>(d * lo(f), d * hi(f));
      reinterpret_cast<V*>(P)[i * (kps / W) + cp] = z;
      rz0 = fma(lo(f), lo(z), rz0); rz1 = fma(hi(f), hi(z), rz1);
    } else {
      reinterpret_cast<V*>(B0)[i * (kz / W) + cp] = f;
    }
  }
  flat_partials<W>(rr0, rr1, rz0, rz1, hc, partial, 2);
}

// r -= alpha q ; partials [0] = r.r, [1] = r.z over the rows >= tail_from, where z = dinv r is NOT stored (k_update_px
// recomputes it from r); rows < tail_from: the new r also goes to the V-cycle's right-hand-side block B0 (row stride kz).
// x += alpha p moved to k_update_px, which reads p anyway: this kernel touches neither P nor X.
template <int W>
__global__ void __launch_bounds__(TB) k_update_r(double* __restrict__ R, const double* __restrict__ Q, const double* __restrict__ dinv,
                                                 double* __restrict__ B0, const double* __restrict__ scal, int ks, int kz,
                                                 int64_t n, int64_t tail_from, double* __restrict__ partial) {
  typedef typename VecT<W>::T V;
  const int hc = ks / W;
  const int64_t T = (int64_t)gridDim.x * TB, gt = (int64_t)blockIdx.x * TB + threadIdx.x;
  const int cp = (int)(gt % hc);
  const int64_t total = n * hc, di = T / hc;
  const double al0 = scal[S_ALPHA * KMAX + W * cp], al1 = (W == 2) ? scal[S_ALPHA * KMAX + W * cp + 1] : 0.0;
  double rr0 = 0.0, rr1 = 0.0, rz0 = 0.0, rz1 = 0.0;
  int64_t i = gt / hc;
  for (int64_t e = gt; e < total; e += T, i += di) {
    const V q = reinterpret_cast<const V*>(Q)[e];
    const V r = reinterpret_cast<const V*>(R)[e];
    const double s0 = fma(-al0, lo(q), lo(r)), s1 = fma(-al1, hi(q), hi(r));
    reinterpret_cast<V*>(R)[e] = mk<W>(s0, s1);
    rr0 = fma(s0, s0, rr0); rr1 = fma(s1, s1, rr1);
    if (i >= tail_from) {
      const double d = dinv[i];
      rz0 = fma(s0, d * s0, rz0); rz1 = fma(s1, d * s1, rz1);
    } else {
      reinterpret_cast<V*>(B0)[i * (kz / W) + cp] = mk<W>(s0, s1);
    }
  }
  flat_partials<W>(rr0, rr1, rz0, rz1, hc, partial, 2);
}

// partial [1] += r.z over rows [0, n): the rows preconditioned by the V-cycle (z in Zv, row stride kz; the others were
// summed by k_update_r / k_init)
template <int W>
__global__ void __launch_bounds__(TB) k_dot_rz(const double* __restrict__ R, const double* __restrict__ Zv, int ks, int kz, int64_t n,
                                               double* __restrict__ partial) {
  typedef typename VecT<W>::T V;
  const int hc = ks / W;
  const int64_t T = (int64_t)gridDim.x * TB, gt = (int64_t)blockIdx.x * TB + threadIdx.x;
  const int cp = (int)(gt % hc);
  const int64_t total = n * hc, di = T / hc;
  double rz0 = 0.0, rz1 = 0.0;
  int64_t i = gt / hc;
  for (int64_t e = gt; e < total; e += T, i += di) {
    const V r = reinterpret_cast<const V*>(R)[e], z = reinterpret_cast<const V*>(Zv)[i * (kz / W) + cp];
    rz0 = fma(lo(r), lo(z), rz0); rz1 = fma(hi(r), hi(z), rz1);
  }
  __shared__ double sh[2][TB];
  sh[0][threadIdx.x] = rz0; sh[1][threadIdx.x] = rz1;
  __syncthreads();
  if ((int)threadIdx.x < hc) {
    const int first = (int)((hc - (int)(((int64_t)blockIdx.x * TB) % hc) + (int)threadIdx.x) % hc);
    double t0 = 0.0, t1 = 0.0;
    for (int i2 = first; i2 < TB; i2 += hc) { t0 += sh[0][i2]; t1 += sh[1][i2]; }
    const int64_t o = ((int64_t)blockIdx.x * 2 + 1) * KMAX + W * threadIdx.x;
    partial[o] += t0;
    if (W == 2) partial[o + 1] += t1;
  }
}

// x += alpha p ; p = z + beta p, with z = dinv r recomputed on the rows >= tail_from and read from the V-cycle's result
// Zv (row stride kz) on the others: one pass over P serves both updates, and Z is never stored for the high-order rows.
template <int W>
__global__ void __launch_bounds__(TB) k_update_px(double* __restrict__ X, double* __restrict__ P, const double* __restrict__ R,
                                                  const double* __restrict__ dinv, const double* __restrict__ Zv,
                                                  const double* __restrict__ scal, int ks, int kps, int kz, int64_t n,
                                                  int64_t tail_from) {
  typedef typename VecT<W>::T V;
  const int hc = ks / W;
  const int64_t T = (int64_t)gridDim.x * TB, gt = (int64_t)blockIdx.x * TB + threadIdx.x;
  const int cp = (int)(gt % hc);
  const int64_t total = n * hc, di = T / hc;
  const double al0 = scal[S_ALPHA * KMAX + W * cp], al1 = (W == 2) ? scal[S_ALPHA * KMAX + W * cp + 1] : 0.0;
  const double be0 = scal[S_BETA * KMAX + W * cp], be1 = (W == 2) ? scal[S_BETA * KMAX + W * cp + 1] : 0.0;
  int64_t i = gt / hc;
  for (int64_t e = gt; e < total; e += T, i += di) {
    V* pp = reinterpret_cast<V*>(P) + i * (kps / W) + cp;
    const V p = *pp, x = reinterpret_cast<const V*>(X)[e];
    reinterpret_cast<V*>(X)[e] = mk<W>(fma(al0, lo(p), lo(x)), fma(al1, hi(p), hi(x)));
    V z;
    if (i >= tail_from) {
      const double d = dinv[i];
      const V r = reinterpret_cast<const V*>(R)[e];
      z = mk<W>(d * lo(r), d * hi(r));
    } else {
      z = reinterpret_cast<const V*>(Zv)[i * (kz / W) + cp];
    }
    *pp = mk<W>(fma(be0, lo(p), lo(z)), fma(be1, hi(p), hi(z)));
  }
}

// ---- one-block scalar kernels: finish the reductions in a fixed order (32 x 32 threads: column r, strip j)
// 1024 threads = ncol columns (power of two >= the column count) x 1024 / ncol strips: with few right-hand sides the
// strips are short even when the SpMM grid has thousands of blocks.  Fixed summation order -> bit-reproducible.
__device__ __forceinline__ double reduce_partials(const double* __restrict__ partial, int nblk, int stride, int slot, int ncol) {
  const int r = threadIdx.x & (ncol - 1), j = threadIdx.x / ncol, ns = 1024 / ncol;
  double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
  int b = j;
#pragma unroll 2
  for (; b + 3 * ns < nblk; b += 4 * ns) {
    t0 += partial[((int64_t)b * stride + slot) * KMAX + r];
    t1 += partial[((int64_t)(b + ns) * stride + slot) * KMAX + r];
    t2 += partial[((int64_t)(b + 2 * ns) * stride + slot) * KMAX + r];
    t3 += partial[((int64_t)(b + 3 * ns) * stride + slot) * KMAX + r];
  }
  for (; b < nblk; b += ns) t0 += partial[((int64_t)b * stride + slot) * KMAX + r];
  const double t = (t0 + t1) + (t2 + t3);
  __shared__ double sh[1024];
  __syncthreads();
  sh[threadIdx.x] = t;
  __syncthreads();
  double tot = 0.0;
  if (j == 0)
    for (int q = 0; q < ns; q++) tot += sh[q * ncol + r];
  return tot;  // valid for threads < ncol (columns >= the column count sum never-written zeros)
}

// after k_init: bb = rr, rz, active, tolerance
__global__ void k_scal_init(const double* __restrict__ partial, int nblk, double* __restrict__ scal, int* __restrict__ iters,
                            int k, double rtol, int ncol) {
  double rr = reduce_partials(partial, nblk, 2, 0, ncol);
  double rz = reduce_partials(partial, nblk, 2, 1, ncol);
  if (threadIdx.x >= ncol) { rr = 0.0; rz = 0.0; }
  if (threadIdx.x < KMAX) {
    const int r = threadIdx.x;
    const bool act = r < k && rr > 0.0;
    scal[S_BB * KMAX + r] = rr;
    scal[S_RR * KMAX + r] = rr;
    scal[S_RZ * KMAX + r] = rz;
    scal[S_TOL2 * KMAX + r] = rtol * rtol * rr;
    scal[S_ACTIVE * KMAX + r] = act ? 1.0 : 0.0;
    scal[S_ALPHA * KMAX + r] = 0.0;
    scal[S_BETA * KMAX + r] = 0.0;
    iters[r] = 0;
  }
}

// rz only (general preconditioner path, after the first apply)
__global__ void k_scal_rz0(const double* __restrict__ partial, int nblk, double* __restrict__ scal, int ncol) {
  const double rz = reduce_partials(partial, nblk, 2, 1, ncol);
  if (threadIdx.x < ncol) scal[S_RZ * KMAX + threadIdx.x] = rz;
}

// after the SpMM: alpha = rz / p.q
__global__ void k_scal_alpha(const double* __restrict__ partial, int nblk, double* __restrict__ scal, int ncol) {
  const double pq = reduce_partials(partial, nblk, 1, 0, ncol);
  if (threadIdx.x < ncol) {
    const int r = threadIdx.x;
    const bool act = scal[S_ACTIVE * KMAX + r] != 0.0;
    scal[S_PQ * KMAX + r] = pq;
    scal[S_ALPHA * KMAX + r] = (act && pq > 0.0) ? scal[S_RZ * KMAX + r] / pq : 0.0;
  }
}

// after the residual update (+ preconditioner): rr, convergence, beta = rz_new / rz
__global__ void k_scal_beta(const double* __restrict__ partial, int nblk, double* __restrict__ scal, int* __restrict__ iters,
                            int ncol) {
  const double rr = reduce_partials(partial, nblk, 2, 0, ncol);
  const double rzn = reduce_partials(partial, nblk, 2, 1, ncol);
  if (threadIdx.x < ncol) {
    const int r = threadIdx.x;
    bool act = scal[S_ACTIVE * KMAX + r] != 0.0;
    if (act) {
      iters[r] += 1;
      scal[S_RR * KMAX + r] = rr;
      const double rz = scal[S_RZ * KMAX + r];
      if (rr <= scal[S_TOL2 * KMAX + r] || !(rz > 0.0)) {
        act = false;
        scal[S_ACTIVE * KMAX + r] = 0.0;
      }
      scal[S_BETA * KMAX + r] = act ? rzn / rz : 0.0;
      scal[S_RZ * KMAX + r] = rzn;
    } else {
      scal[S_BETA * KMAX + r] = 0.0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// point sources, sampling, apparent resistivity
// ------------------------------------------------------------------------------------------------
__global__ void k_point_sources(SpaceView s, int nrhs, int stride, const int64_t* __restrict__ src_ptr, const double* __restrict__ src_z,
                                const double* __restrict__ src_fac, double* __restrict__ F, int* __restrict__ bad) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrhs) return;
  for (int64_t i = src_ptr[r]; i < src_ptr[r + 1]; i++) {
    const double fac = src_fac[i];
    if (fac == 0.0) continue;
    int64_t dof[4];
    double val[4];
    const int m = axis_shape(s, src_z[i], dof, val);
    if (m <= 0) { atomicExch(bad, m == 0 ? 1 : 2); continue; }
    for (int q = 0; q < m; q++) F[dof[q] * stride + r] += fac * val[q];  // one thread per column: no race
  }
}

__device__ __forceinline__ double eval_axis(const SpaceView& s, const double* __restrict__ X, int nrhs, int rhs, double z,
                                            int* bad) {
  int64_t dof[4];
  double val[4];
  const int m = axis_shape(s, z, dof, val);
  if (m <= 0) { atomicExch(bad, m == 0 ? 1 : 2); return nan(""); }
  double u = 0.0;
  for (int q = 0; q < m; q++) u = fma(val[q], X[dof[q] * nrhs + rhs], u);
  return u;
}

__global__ void k_sample(SpaceView s, const double* __restrict__ X, int nrhs, int nuser, int npts, const int32_t* __restrict__ pt_rhs,
                         const double* __restrict__ z, double* __restrict__ out, int* __restrict__ bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npts) return;
  const int rhs = pt_rhs ? pt_rhs[i] : 0;
  if (rhs < 0 || rhs >= nuser) { atomicExch(bad, 3); out[i] = nan(""); return; }
  out[i] = eval_axis(s, X, nrhs, rhs, z[i], bad);
}

__global__ void k_resistivity(SpaceView s, const double* __restrict__ X, int nrhs, int nuser, int npts, const int32_t* __restrict__ pt_rhs,
                              const double* __restrict__ z0, const double* __restrict__ z1, const double* __restrict__ kf,
                              double scale, double* __restrict__ ra, int* __restrict__ bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npts) return;
  const int rhs = pt_rhs[i];
  if (rhs < 0 || rhs >= nuser) { atomicExch(bad, 3); ra[i] = nan(""); return; }
  const double u0 = eval_axis(s, X, nrhs, rhs, z0[i], bad);
  double du = u0;
  if (z1[i] == z1[i]) du = eval_axis(s, X, nrhs, rhs, z1[i], bad) - u0;
  ra[i] = fabs(kf[i] * du) * scale;
}

int kp_for(int k) {
  int kp = 1;
  while (kp < k) kp <<= 1;
  return kp;
}

#define DISPATCH_KP(kp, CALL)        \
  switch (kp) {                      \
    case 1: { constexpr int KP = 1; CALL; } break;   \
    case 2: { constexpr int KP = 2; CALL; } break;   \
    case 4: { constexpr int KP = 4; CALL; } break;   \
    case 8: { constexpr int KP = 8; CALL; } break;   \
    case 16: { constexpr int KP = 16; CALL; } break; \
    default: { constexpr int KP = 32; CALL; } break; \
  }

// grid of the flat vector kernels: grid * TB threads must be a multiple of hc = ks / W (see above)
int vec_grid(Ctx* c, int ks) {
  int odd = (ks & 1) ? ks : ks / 2;
  while ((odd & 1) == 0) odd >>= 1;
  const int g = c->num_sms * 8;
  return (g + odd - 1) / odd * odd;
}
#define DISPATCH_W(ks, CALL)                      \
  if ((ks) & 1) { constexpr int W = 1; CALL; }    \
  else { constexpr int W = 2; CALL; }
// the SELL kernels want P at the stride sell_pstride picks (a power of two); any other even stride goes to the CSR kernels
bool use_sell(const Ctx* c, int nrhs, const double* P) {
  return c->have_sell && c->pstride >= 2 && (nrhs & 1) == 0 && P == c->P.p && c->pstride == sell_pstride(nrhs);
}
// element-wise product (ebe.cu): order-2 / order-3 tets and order-3 triangles, the PCG's own P block, up to 6 right-hand sides
bool use_ebe(const Ctx* c, const double* P) { return P == c->P.p && ebe_usable(c, c->nrhs_user); }
int spmm_grid(Ctx* c, int nrhs) {
  if (use_ebe(c, c->P.p)) return ebe_grid(c, c->nrhs_user);
  return use_sell(c, nrhs, c->P.p) ? sell_grid(c) : c->num_sms * 8;
}

void check_bad(Ctx* c, DBuf<int>& bad, const char* who) {
  int h = 0;
  CK(cudaMemcpyAsync(&h, bad.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  bad.release(c->stream);
  if (h == 1) FAIL(REMO_ERR_MESH, "%s: an axis point lies outside the mesh axis", who);
  if (h == 2) FAIL(REMO_ERR_MESH, "%s: consecutive axis vertices are not joined by a mesh edge", who);
  if (h == 3) FAIL(REMO_ERR_ARG, "%s: right-hand-side index out of range", who);
}

}  // namespace

void alloc_solver_state(Ctx* c, int nrhs) {
  cudaStream_t st = c->stream;
  const size_t n = (size_t)c->ndof * nrhs;
  c->F.ensure(n, st); c->X.ensure(n, st); c->R.ensure(n, st);
  c->Q.ensure(n, st);
  // the rows preconditioned by the V-cycle (the vertex block) have their own right-hand-side / result blocks with an even
  // row stride (amg.cu works on column pairs); the padding column of an odd count stays zero
  c->kz = (nrhs + 1) & ~1;
  c->B0.ensure((size_t)c->nv * c->kz, st); c->Zv.ensure((size_t)c->nv * c->kz, st);
  CK(cudaMemsetAsync(c->B0.p, 0, (size_t)c->nv * c->kz * sizeof(double), st));
  CK(cudaMemsetAsync(c->Zv.p, 0, (size_t)c->nv * c->kz * sizeof(double), st));
  // P alone gets a power-of-two row stride when the SELL SpMM gathers it (sell.cu): no gathered row straddles a line.
  // The element-wise product stages whole rows of P once per batch: it takes the plain stride.
  const bool ebe = ebe_serves(c, nrhs);
  c->pstride = (!ebe && spmm_variant() >= 5 && nrhs >= 2 && (nrhs & 1) == 0) ? sell_pstride(nrhs) : nrhs;
  c->P.ensure((size_t)c->ndof * c->pstride, st);
  if (c->pstride != nrhs) CK(cudaMemsetAsync(c->P.p, 0, (size_t)c->ndof * c->pstride * sizeof(double), st));
  c->partial.ensure((size_t)std::max(vec_grid(c, nrhs), c->num_sms * 192) * 2 * KMAX, st);  // room for any SpMM grid (sell.cu caps its own)
  CK(cudaMemsetAsync(c->partial.p, 0, c->partial.n * sizeof(double), st));
  c->scal.ensure(S_NSLOT * KMAX, st);
  c->iters_d.ensure(KMAX, st);
  c->nrhs = nrhs;
}

int spmm_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("REMO_SPMM_VARIANT");
    v = e ? atoi(e) : 5;
  }
  return v;
}

// CSR / SELL kernels: even (they gather two right-hand sides per 16-byte load); the extra column of an odd count has
// b = 0.  Element-wise product (ebe.cu): exactly nrhs -- no dead column in any vector pass.
int solver_stride(Ctx* c, int nrhs) {
  if (ebe_serves(c, nrhs)) return nrhs;
  return (nrhs > 1 && spmm_variant() >= 4) ? ((nrhs + 1) & ~1) : nrhs;
}
int spmm_blocks(Ctx* c, int ks) { return spmm_grid(c, ks); }

void launch_spmm(Ctx* c, const double* P, double* Q, int nrhs) {
  const int kp = kp_for(nrhs);
  const int grid = c->num_sms * 8;
  if (use_ebe(c, P)) {  // no assembled matrix at all: element by element from the metric numbers (ebe.cu)
    launch_spmm_ebe(c, P, c->pstride, Q, nrhs, c->nrhs_user);
    return;
  }
  if (use_sell(c, nrhs, P)) {  // SELL copy + power-of-two P stride (sell.cu)
    launch_spmm_sell(c, P, Q, nrhs, c->pstride);
    return;
  }
  if (c->pstride != nrhs && P == c->P.p) FAIL(REMO_ERR_STATE, "launch_spmm: P has stride %d but no SELL matrix exists", c->pstride);
  if (spmm_variant() >= 3 && (nrhs & 1) == 0) {  // even stride: paired-column kernel
    auto* rp = c->rowptr.p; auto* cl = c->col.p; auto* vl = c->val.p; auto* cs = c->constrained.p;
    double* pt = c->partial.p;
    const int64_t n = c->ndof;
    cudaStream_t st = c->stream;
    if (nrhs <= 2) k_spmm_p<4, 1><<<grid, TB, 0, st>>>(rp, cl, vl, cs, P, Q, nrhs, n, pt);
    else if (nrhs <= 4) k_spmm_p<4, 2><<<grid, TB, 0, st>>>(rp, cl, vl, cs, P, Q, nrhs, n, pt);
    else if (nrhs <= 8) k_spmm_p<4, 4><<<grid, TB, 0, st>>>(rp, cl, vl, cs, P, Q, nrhs, n, pt);
    else if (nrhs <= 16) k_spmm_p<8, 8><<<grid, TB, 0, st>>>(rp, cl, vl, cs, P, Q, nrhs, n, pt);
    else k_spmm_p<16, 16><<<grid, TB, 0, st>>>(rp, cl, vl, cs, P, Q, nrhs, n, pt);
    c->launches++;
    CK(cudaGetLastError());
    return;
  }
  if (spmm_variant() >= 3) {
    auto* rp = c->rowptr.p; auto* cl = c->col.p; auto* vl = c->val.p; auto* cs = c->constrained.p;
    double* pt = c->partial.p;
    const int64_t n = c->ndof;
    cudaStream_t st = c->stream;
    switch (kp) {
      case 1: k_spmm_g<8, 1><<<grid, TB, 0, st>>>(rp, cl, vl, cs, P, Q, nrhs, n, pt); break;
      case 2: k_spmm_g<8, 2><<<grid, TB, 0, st>>>(rp, cl, vl, cs, P, Q, nrhs, n, pt); break;
      case 4: k_spmm_g<8, 4><<<grid, TB, 0, st>>>(rp, cl, vl, cs, P, Q, nrhs, n, pt); break;
      case 8: k_spmm_g<8, 8><<<grid, TB, 0, st>>>(rp, cl, vl, cs, P, Q, nrhs, n, pt); break;
      case 16: k_spmm_g<16, 16><<<grid, TB, 0, st>>>(rp, cl, vl, cs, P, Q, nrhs, n, pt); break;
      default: k_spmm_g<32, 32><<<grid, TB, 0, st>>>(rp, cl, vl, cs, P, Q, nrhs, n, pt); break;
    }
    c->launches++;
    CK(cudaGetLastError());
    return;
  }
  DISPATCH_KP(kp, (k_spmm<KP><<<grid, TB, 0, c->stream>>>(c->rowptr.p, c->col.p, c->val.p, c->constrained.p, P, Q, nrhs, c->ndof, c->partial.p)));
  c->launches++;
  CK(cudaGetLastError());
}

// after alloc_solver_state / the right-hand sides are known: make sure the matrix copy the chosen SpMM kernel reads exists
void spmm_prepare(Ctx* c) {
  if (!c->have_matrix || use_ebe(c, c->P.p)) return;
  ensure_values(c);  // the CSR / SELL kernels read the assembled matrix
  if (spmm_variant() >= 5 && c->pstride >= 2 && !c->have_sell) sell_build(c);
}

// 0 = CSR kernels, 1 = SELL copy, 2 = element-wise (for remo_spmm_kind)
int spmm_kind(Ctx* c) {
  if (use_ebe(c, c->P.p)) return 2;
  return use_sell(c, c->nrhs, c->P.p) ? 1 : 0;
}

void launch_vector_updates(Ctx* c, int nrhs) {
  const int grid = vec_grid(c, nrhs);
  DISPATCH_W(nrhs, (k_update_r<W><<<grid, TB, 0, c->stream>>>(c->R.p, c->Q.p, c->dinv.p, c->B0.p, c->scal.p, nrhs, c->kz, c->ndof, (int64_t)0, c->partial.p)));
  DISPATCH_W(nrhs, (k_update_px<W><<<grid, TB, 0, c->stream>>>(c->X.p, c->P.p, c->R.p, c->dinv.p, c->Zv.p, c->scal.p, nrhs, c->pstride, c->kz, c->ndof, (int64_t)0)));
  c->launches += 2;
  CK(cudaGetLastError());
}

void precond_setup(Ctx* c, int kind) {
  if (!c->have_matrix) FAIL(REMO_ERR_STATE, "remo_precond_setup: no matrix (call remo_assemble first)");
  if (kind != REMO_PRECOND_LOCAL && kind != REMO_PRECOND_MULTIGRID) FAIL(REMO_ERR_ARG, "remo_precond_setup: unknown preconditioner kind %d", kind);
  StageTimer timer(c, ST_PRECOND);
  // diag(A) and the V-cycle's vertex block come straight from the element metrics (assemble.cu): no assembled matrix
  c->dinv.ensure(c->ndof, c->stream);
  diag_from_elements(c, c->dinv.p);
  if (kind == REMO_PRECOND_MULTIGRID) amg_setup(c);
  // element-wise path (ebe_eligible): the product needs no copy of the matrix at all; the CSR values and the SELL copy are made
  // on demand (spmm_prepare) if a block wider than ebe.cu takes shows up
  if (ebe_eligible(c)) { if (!c->have_ebe) ebe_build(c); }
  else {
    ensure_values(c);
    if (spmm_variant() >= 5 && !c->have_sell) sell_build(c);
  }
  c->pkind = kind;
}

void rhs_point_sources(Ctx* c, int nrhs, const int64_t* src_ptr, const double* src_z, const double* src_fac) {
  if (!c->have_space) FAIL(REMO_ERR_STATE, "remo_rhs_point_sources: no space (call remo_space_build first)");
  if (nrhs < 1 || nrhs > REMO_MAX_RHS) FAIL(REMO_ERR_ARG, "remo_rhs_point_sources: nrhs must be in 1..%d (got %d)", REMO_MAX_RHS, nrhs);
  StageTimer timer(c, ST_RHS);
  cudaStream_t st = c->stream;
  // row stride of the vector blocks: even (the SpMM gathers two right-hand sides per 16-byte load); the extra column
  // of an odd count has b = 0 and is frozen from the first iteration
  const int ks = solver_stride(c, nrhs);
  alloc_solver_state(c, ks);
  c->nrhs_user = nrhs;
  std::vector<int64_t> hp(nrhs + 1);
  CK(cudaMemcpyAsync(hp.data(), src_ptr, (nrhs + 1) * sizeof(int64_t), cudaMemcpyDefault, st));
  CK(cudaStreamSynchronize(st));
  const int64_t ns = hp[nrhs];
  if (hp[0] != 0 || ns < 0) FAIL(REMO_ERR_ARG, "remo_rhs_point_sources: malformed src_ptr");
  DBuf<int64_t> dp;
  DBuf<double> dz, df;
  DBuf<int> bad;
  dp.ensure(nrhs + 1, st); dz.ensure(ns + 1, st); df.ensure(ns + 1, st); bad.ensure(1, st);
  CK(cudaMemcpyAsync(dp.p, hp.data(), (nrhs + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  if (ns) {
    CK(cudaMemcpyAsync(dz.p, src_z, ns * sizeof(double), cudaMemcpyDefault, st));
    CK(cudaMemcpyAsync(df.p, src_fac, ns * sizeof(double), cudaMemcpyDefault, st));
  }
  CK(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
  CK(cudaMemsetAsync(c->F.p, 0, (size_t)c->ndof * ks * sizeof(double), st));
  LAUNCH(c, k_point_sources, 1, 32, 0, make_view(c), nrhs, ks, dp.p, dz.p, df.p, c->F.p, bad.p);
  check_bad(c, bad, "remo_rhs_point_sources");
  dp.release(st); dz.release(st); df.release(st);
  c->have_rhs = true;
  c->have_solution = false;
}

int solve(Ctx* c, double rtol, int maxit, int* iters, double* relres) {
  if (!c->have_matrix || c->pkind < 0) FAIL(REMO_ERR_STATE, "remo_solve: matrix / preconditioner missing (remo_assemble, remo_precond_setup)");
  if (!c->have_rhs) FAIL(REMO_ERR_STATE, "remo_solve: no right-hand side (call remo_rhs_point_sources first)");
  if (!(rtol > 0.0) || maxit < 1) FAIL(REMO_ERR_ARG, "remo_solve: rtol must be > 0 and maxit >= 1");
  StageTimer timer(c, ST_SOLVE);
  spmm_prepare(c);
  cudaStream_t st = c->stream;
  const int k = c->nrhs, kp = kp_for(k);
  const int64_t n = c->ndof;
  const int vg = vec_grid(c, k), sg = spmm_grid(c, k);
  const int jac = (c->pkind == REMO_PRECOND_LOCAL) ? 1 : 0;
  const int64_t tail = jac ? 0 : c->nv;  // rows >= tail: z = D^-1 r fused into the vector kernels; rows < tail: V-cycle

  const int kps = c->pstride, kz = c->kz;
  DISPATCH_W(k, (k_init<W><<<vg, TB, 0, st>>>(c->F.p, c->constrained.p, c->dinv.p, c->X.p, c->R.p, c->B0.p, c->P.p, k, kps, kz, n, tail, c->partial.p)));
  c->launches++;
  if (!jac) {
    amg_apply(c, c->B0.p, c->Zv.p, kz);
    DISPATCH_W(k, (k_dot_rz<W><<<vg, TB, 0, st>>>(c->R.p, c->Zv.p, k, kz, tail, c->partial.p)));
    CK(cudaMemcpy2DAsync(c->P.p, (size_t)kps * sizeof(double), c->Zv.p, (size_t)kz * sizeof(double), (size_t)k * sizeof(double), (size_t)tail, cudaMemcpyDeviceToDevice, st));
    c->launches++;
  }
  k_scal_init<<<1, 1024, 0, st>>>(c->partial.p, vg, c->scal.p, c->iters_d.p, k, rtol, kp);
  c->launches++;
  CK(cudaGetLastError());

  const int check_every = 16;
  std::vector<double> hs(S_NSLOT * KMAX);
  int it = 0;
  bool done = false;

  // one PCG iteration = ~8 (Jacobi) to ~40 (multigrid) launches, most of them tiny: the launch-bound inner loop is
  // captured once per solve as a CUDA graph and replayed; the host only polls the convergence flags every 16 iterations
  auto iteration = [&](int itno) {
    const size_t pe = (size_t)2 * itno;
    if (c->prof) {
      while (c->prof_ev.size() < pe + 2) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        c->prof_ev.push_back(e);
      }
      CK(cudaEventRecord(c->prof_ev[pe], st));
    }
    launch_spmm(c, c->P.p, c->Q.p, k);
    if (c->prof) CK(cudaEventRecord(c->prof_ev[pe + 1], st));
    k_scal_alpha<<<1, 1024, 0, st>>>(c->partial.p, sg, c->scal.p, kp);
    DISPATCH_W(k, (k_update_r<W><<<vg, TB, 0, st>>>(c->R.p, c->Q.p, c->dinv.p, c->B0.p, c->scal.p, k, kz, n, tail, c->partial.p)));
    if (!jac) {
      amg_apply(c, c->B0.p, c->Zv.p, kz);
      DISPATCH_W(k, (k_dot_rz<W><<<vg, TB, 0, st>>>(c->R.p, c->Zv.p, k, kz, tail, c->partial.p)));
      c->launches++;
    }
    k_scal_beta<<<1, 1024, 0, st>>>(c->partial.p, vg, c->scal.p, c->iters_d.p, kp);
    DISPATCH_W(k, (k_update_px<W><<<vg, TB, 0, st>>>(c->X.p, c->P.p, c->R.p, c->dinv.p, c->Zv.p, c->scal.p, k, kps, kz, n, tail)));
    c->launches += 4;
  };

  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  int64_t launches_per_iter = 0;
  const bool graphs = c->use_graph && !c->prof && getenv("REMO_NO_GRAPH") == nullptr;
  if (graphs) {
    if (!jac) amg_prepare(c, kz);  // no allocation may happen inside the capture
    const int64_t l0 = c->launches;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    try {
      iteration(0);
    } catch (...) {
      cudaStreamEndCapture(st, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    CK(cudaStreamEndCapture(st, &graph));
    launches_per_iter = c->launches - l0;
    c->launches = l0;
    CK(cudaGraphInstantiate(&gexec, graph, 0));
  }
  while (it < maxit && !done) {
    const int chunk = std::min(check_every, maxit - it);
    for (int q = 0; q < chunk; q++) {
      if (graphs) {
        CK(cudaGraphLaunch(gexec, st));
        c->launches += launches_per_iter;
      } else {
        iteration(it + q);
      }
    }
    CK(cudaGetLastError());
    it += chunk;
    CK(cudaMemcpyAsync(hs.data(), c->scal.p, hs.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    done = true;
    for (int r = 0; r < k; r++)
      if (hs[S_ACTIVE * KMAX + r] != 0.0) done = false;
  }
  if (gexec) cudaGraphExecDestroy(gexec);
  if (graph) cudaGraphDestroy(graph);
  std::vector<int> hit(KMAX);
  CK(cudaMemcpyAsync(hit.data(), c->iters_d.p, KMAX * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  bool conv = true;
  for (int r = 0; r < c->nrhs_user; r++) {
    const double bb = hs[S_BB * KMAX + r], rr = hs[S_RR * KMAX + r];
    const double rel = bb > 0.0 ? sqrt(rr / bb) : 0.0;
    if (iters) iters[r] = hit[r];
    if (relres) relres[r] = rel;
    // converged means the residual criterion is met -- a column frozen by a breakdown (r.z <= 0) is NOT converged
    if (bb > 0.0 && !(rr <= hs[S_TOL2 * KMAX + r])) conv = false;
  }
  if (c->prof) {
    for (int q = 0; q < it; q++) {
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, c->prof_ev[2 * q], c->prof_ev[2 * q + 1]));
      c->prof_spmm_ms += ms;
    }
    c->prof_spmm_n += it;
  }
  c->have_solution = true;
  c->host_scal = hs;
  return conv ? REMO_OK : REMO_ERR_NOCONV;
}

void sample_axis(Ctx* c, int npts, const int32_t* pt_rhs, const double* z, double* out) {
  if (!c->have_solution) FAIL(REMO_ERR_STATE, "remo_sample_axis: no solution (call remo_solve first)");
  if (npts < 1) return;
  StageTimer timer(c, ST_SAMPLE);
  cudaStream_t st = c->stream;
  DBuf<int32_t> dr;
  DBuf<double> dz, dout;
  DBuf<int> bad;
  dr.ensure(npts, st); dz.ensure(npts, st); dout.ensure(npts, st); bad.ensure(1, st);
  if (pt_rhs) CK(cudaMemcpyAsync(dr.p, pt_rhs, npts * sizeof(int32_t), cudaMemcpyDefault, st));
  CK(cudaMemcpyAsync(dz.p, z, npts * sizeof(double), cudaMemcpyDefault, st));
  CK(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
  LAUNCH(c, k_sample, grid_for(npts, 128), 128, 0, make_view(c), c->X.p, c->nrhs, c->nrhs_user, npts, pt_rhs ? dr.p : nullptr, dz.p, dout.p, bad.p);
  CK(cudaMemcpyAsync(out, dout.p, npts * sizeof(double), cudaMemcpyDefault, st));
  check_bad(c, bad, "remo_sample_axis");
  dr.release(st); dz.release(st); dout.release(st);
}

void apparent_resistivity(Ctx* c, int npts, const int32_t* pt_rhs, const double* z0, const double* z1, const double* kf,
                          double scale, double* ra) {
  if (!c->have_solution) FAIL(REMO_ERR_STATE, "remo_apparent_resistivity: no solution (call remo_solve first)");
  if (npts < 1) return;
  StageTimer timer(c, ST_SAMPLE);
  cudaStream_t st = c->stream;
  DBuf<int32_t> dr;
  DBuf<double> d0, d1, dk, dout;
  DBuf<int> bad;
  dr.ensure(npts, st); d0.ensure(npts, st); d1.ensure(npts, st); dk.ensure(npts, st); dout.ensure(npts, st); bad.ensure(1, st);
  CK(cudaMemcpyAsync(dr.p, pt_rhs, npts * sizeof(int32_t), cudaMemcpyDefault, st));
  CK(cudaMemcpyAsync(d0.p, z0, npts * sizeof(double), cudaMemcpyDefault, st));
  CK(cudaMemcpyAsync(d1.p, z1, npts * sizeof(double), cudaMemcpyDefault, st));
  CK(cudaMemcpyAsync(dk.p, kf, npts * sizeof(double), cudaMemcpyDefault, st));
  CK(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
  LAUNCH(c, k_resistivity, grid_for(npts, 128), 128, 0, make_view(c), c->X.p, c->nrhs, c->nrhs_user, npts, dr.p, d0.p, d1.p, dk.p, scale, dout.p, bad.p);
  CK(cudaMemcpyAsync(ra, dout.p, npts * sizeof(double), cudaMemcpyDefault, st));
  check_bad(c, bad, "remo_apparent_resistivity");
  dr.release(st); d0.release(st); d1.release(st); dk.release(st); dout.release(st);
}
