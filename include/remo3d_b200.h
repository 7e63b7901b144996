/*
 * remo3d_b200.h -- C ABI of libremo3d_b200.so: the B200-native (sm_100a) replacement for the
 * NGSolve calls on ReMo3D's per-measurement-point forward solve.
 *
 * The reference reaches its numerics through pybind11 into NGSolve (a third-party C++ library,
 * not vendored); this ABI is what a `ctypes` binding in the reference's
 * `remo3d/ngsolve_functions.py` / `ngsolve_functions_gpu.py` would call instead.  Every entry
 * point cites the reference lines it replaces (paths relative to /root/reference).
 *
 * Conventions
 *   - return 0 on success, <0 on error; `remo_last_error(ctx)` returns the message.  No C++
 *     exception crosses this boundary.
 *   - the context owns all device memory; the caller owns every buffer it passes for the
 *     duration of the call only.  Pointers may be host OR device pointers (unified virtual
 *     addressing decides the copy direction), so callers can hand over pinned host memory
 *     or `torch.Tensor.data_ptr()` of a CUDA tensor alike.
 *   - a context is bound to one device and one CUDA stream and is not thread-safe; use one
 *     context per worker thread / GPU.  Calls do not hold the Python GIL (ctypes releases it).
 *   - all indices are 0-based; vertex / edge / face / dof numbers are the reference numbering
 *     (SURVEY.md section 10.2): dofs = [vertices | edges x (p-1) | faces (p=3)], edges and
 *     faces numbered lexicographically by their sorted global vertex tuples.
 *   - right-hand sides and solutions are stored on the device as a row-major ndof x nrhs block.
 */
#ifndef REMO3D_B200_H
#define REMO3D_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REMO_OK 0
#define REMO_ERR_CUDA -1     /* a CUDA runtime call failed                                  */
#define REMO_ERR_ARG -2      /* invalid argument                                            */
#define REMO_ERR_STATE -3    /* call sequence violated (e.g. assemble before space_build)   */
#define REMO_ERR_MESH -4     /* inconsistent mesh (axis point outside mesh, missing edge..) */
#define REMO_ERR_NOCONV -5   /* PCG hit maxit before reaching rtol                          */

#define REMO_PRECOND_LOCAL 0      /* NGSolve "local"     = Jacobi on the free dofs           */
#define REMO_PRECOND_MULTIGRID 1  /* NGSolve "multigrid" = two-level: P1 (vertex block) coarse
                                     correction + Jacobi smoothing of the high-order dofs    */

#define REMO_MAX_RHS 32

/* Context = one GPU worker (replaces one MPI worker process, workers/worker.py:19-35). */
int remo_ctx_create(int device, void** ctx);
int remo_ctx_destroy(void* ctx);
const char* remo_last_error(void* ctx);
/* Run all further work of this context on a caller-owned CUDA stream (a `cudaStream_t`, e.g.
 * `torch.cuda.Stream().cuda_stream`), so the caller can bracket calls with its own CUDA events.
 * The stream must outlive the context and must not be the legacy default stream 0.            */
int remo_ctx_set_stream(void* ctx, void* stream);

/* Upload one mesh (replaces `ngs.Mesh(mesh)`, workers/worker.py:100).
 *   dim          2 (axisymmetric r,z triangles) or 3 (tets)
 *   xyz          nv x dim doubles; depth z is the last coordinate
 *   elems        nt x (dim+1) int32
 *   mat          nt int32, 0-based index into the sigma list of remo_assemble
 *   bfacets      nb x dim int32 boundary triangles / segments
 *   bdirichlet   nb uint8: 1 if the facet belongs to the `dirichlet=` selection of ngs.H1
 *                (ngsolve_functions.py:27; names/numbers resolved by the caller, worker.py:90,97)
 *   axis_vertices naxis int32: vertices on the electrode axis sorted by z (all electrodes lie on
 *                x=0[,y=0], netgen_functions.py:226-229, gmsh_functions.py:566-575)            */
int remo_mesh_set(void* ctx, int dim, int64_t nv, const double* xyz, int64_t nt, const int32_t* elems,
                  const int32_t* mat, int64_t nb, const int32_t* bfacets, const uint8_t* bdirichlet,
                  int64_t naxis, const int32_t* axis_vertices);

/* Symbolic phase: topology, dof numbering, Dirichlet dofs, dof -> element adjacency
 * (replaces `fes = ngs.H1(mesh, order=3, dirichlet=..)`, ngsolve_functions.py:27, and the
 * sparsity-graph part of `a.Assemble()`, :47).  order in {1,2,3}.
 * The CSR pattern itself is built on demand (remo_matrix_nnz / remo_matrix_get / a solve that
 * reads the assembled matrix): the element-wise PCG path of order-2 / order-3 tets never needs it.  *nnz is
 * 0 unless the pattern exists when the call returns (remo_set_option("lazy_matrix", 0): always). */
int remo_space_build(void* ctx, int order, int64_t* ndof, int64_t* nnz, int64_t* nedges, int64_t* nfaces);

/* Numbering export for parity tests: edges ne x 2, faces nf x 3 (sorted vertex tuples),
 * elem_edges nt x 6 (3 in 2D), elem_faces nt x 4 (1 in 2D); any pointer may be NULL.           */
int remo_topology_get(void* ctx, int32_t* edges, int32_t* faces, int32_t* elem_edges, int32_t* elem_faces);

/* Numeric assembly of  a += grad(u)*grad(v)*sigma*dx  (3D)  /  2*pi*grad(u)*grad(v)*x*sigma*dx  (2D)
 * (ngsolve_functions.py:31-36, 47).  sigma: nmat doubles, one per material (worker.py:101).    */
int remo_assemble(void* ctx, int nmat, const double* sigma);

/* Numeric assembly ... with "lazy_matrix" (default) remo_assemble evaluates the element metrics only; the CSR
 * values are gathered row by row (atomic-free, bit-reproducible) when first needed.
 * Export the assembled matrix, CSR with sorted columns, Dirichlet rows kept (reference numbering). */
int remo_matrix_get(void* ctx, int64_t* rowptr, int32_t* col, double* val);
/* Non-zeros of the CSR pattern (builds the pattern if it does not exist yet).                   */
int remo_matrix_nnz(void* ctx, int64_t* nnz);
int remo_dirichlet_get(void* ctx, uint8_t* constrained);

/* `c = ngs.Preconditioner(a, "local"|"multigrid")` (ngsolve_functions.py:46).                  */
int remo_precond_setup(void* ctx, int kind);

/* Parity export of what remo_precond_setup built (every pointer may be NULL): dinv = ndof doubles, 1 / a_ii on free dofs and
 * 0 on constrained ones; the vertex block of A (= P1 stiffness matrix, level 0 of the "multigrid" V-cycle) as CSR with
 * nv + 1 row pointers and nv + 2 * nedges entries; *nlevels in = capacity of level_rows / level_nnz, out = levels of the
 * aggregation hierarchy (0 for "local").                                                                              */
int remo_precond_get(void* ctx, double* dinv, int64_t* vv_rowptr, int32_t* vv_col, double* vv_val, int* nlevels,
                     int64_t* level_rows, int64_t* level_nnz);

/* Right-hand sides: AddPointSource for every (rhs, source) pair (ngsolve_functions.py:10-21, 39-44).
 * src_ptr: nrhs+1 offsets into src_z / src_fac.  1 <= nrhs <= REMO_MAX_RHS.                    */
int remo_rhs_point_sources(void* ctx, int nrhs, const int64_t* src_ptr, const double* src_z, const double* src_fac);
int remo_rhs_get(void* ctx, int rhs, double* f);

/* PCG for all right-hand sides at once (`inv = ngs.CGSolver(a.mat, c.mat, maxsteps=1000);
 * gfu.vec.data = inv * f.vec`, ngsolve_functions.py:50-51; device form ngsolve_functions_gpu.py:41-47).
 * x0 = 0; column r stops when ||r||_2 <= rtol ||b||_2.  iters / relres: nrhs entries (may be NULL).
 * Returns REMO_ERR_NOCONV if any column did not converge within maxit (results are still stored). */
int remo_solve(void* ctx, double rtol, int maxit, int* iters, double* relres);

/* Point evaluation on the axis, `gfu(mesh(0.0, z))` / `gfu(mesh(0.0, 0.0, z))` (worker.py:122-131). */
int remo_sample_axis(void* ctx, int npts, const int32_t* pt_rhs, const double* z, double* u_out);

/* Apparent resistivity of every log point (worker.py:113-134):
 *   ra = scale * |k * (u(z1) - u(z0))|   or   scale * |k * u(z0)|  when z1 is NaN;
 * scale = 0.5 on the 3D half-ball (worker.py:129,131), 1 in 2D.                               */
int remo_apparent_resistivity(void* ctx, int npts, const int32_t* pt_rhs, const double* z0, const double* z1,
                              const double* k, double scale, double* ra);

/* Full solution vector of one right-hand side (ndof doubles), `gfu.vec` (parity tests).        */
int remo_solution_get(void* ctx, int rhs, double* u);

/* Measurement hooks (bench.py): time `reps` launches of one kernel with CUDA events on the
 * context's stream; returns the average milliseconds per launch in *ms.
 *   which = 0: SpMM  Q = A P with the fused p.q dot (the PCG kernel), nrhs columns
 *   which = 1: numeric assembly (remo_assemble's kernels)
 *   which = 2: PCG vector update kernels for nrhs columns                                       */
int remo_kernel_time(void* ctx, int which, int nrhs, int reps, float* ms);

/* Parity hook (tests): one launch of the PCG's SpMM on caller-supplied search directions.
 *   p  : host, ndof x nrhs row-major (the nrhs values of a dof contiguous)
 *   q  : host, ndof x nrhs row-major, receives Q = A P with the rows of constrained dofs set to 0
 *   pq : host, nrhs, receives the fused dots p_r . q_r (block partials finished in a fixed order)
 * Runs exactly the kernel remo_solve would use for this nrhs (SELL streaming / SELL generic / CSR), including the
 * internal strides; needs remo_assemble (sets up the "local" preconditioner if none is set).  Invalidates the
 * right-hand sides and the solution of the context.                                                            */
int remo_spmm_apply(void* ctx, int nrhs, const double* p, double* q, double* pq);
/* Which SpMM the PCG uses for the right-hand sides currently set: 0 = CSR kernels, 1 = SELL-8 copy (sell.cu),
 * 2 = element-wise product from the metric numbers, no assembled matrix read (ebe.cu: order-2 / order-3 tets, order-3 triangles, <= 6 columns). */
int remo_spmm_kind(void* ctx);
/* Solver tunables (defaults in parentheses): "amg_sweeps" (1) pre = post damped-Jacobi sweeps per AMG level,
 * "amg_alpha" (1.5) scaling of the coarse-grid correction, "spmm_ebe" (1) 0 switches the element-wise product off (the SELL kernels take over), "amg_omega_scale" (1.0) weight of the l1-Jacobi sweeps,
 * must be <= 1 (changing it invalidates the preconditioner: call remo_precond_setup again).
 * Round 2: "amg_fp32" (1) the V-cycle runs in fp32 inside the fp64 PCG; "amg_lanes8" (1) 8 lanes per row in the sweeps for 5..8
 * columns; "amg_fused_tail" (0) / "amg_tail_rows" (20000) small levels in one cluster kernel; "amg_agg" (1) strength-based
 * pairwise aggregation (0: Morton-rank aggregates), "amg_passes" (3), "amg_rounds" (4); "ebe_check" (0) validate the
 * element-wise batch tables after every build (fails with REMO_ERR_STATE on a violation); "ebe_p3_ctas" (3) build of the
 * order-3 product kernel: 3 or 4 resident CTAs per SM; "lazy_matrix" (1) CSR pattern / values only on demand.          */
int remo_set_option(void* ctx, const char* name, double value);
/* Per-launch timing of the SpMM inside remo_solve: while on, CUDA events bracket every SpMM launch of the
 * PCG loop on the context's stream; remo_profile_get returns the summed milliseconds and the launch count
 * since remo_profile(ctx, 1).  (bench.py roofline: average launch duration inside the timed region.)    */
int remo_profile(void* ctx, int on);
int remo_profile_get(void* ctx, double* spmm_ms_total, int64_t* spmm_launches);
/* Number of kernel launches issued by this context since creation (bench.py gpu_launches).    */
int64_t remo_launch_count(void* ctx);
/* Per-stage device times of the last calls, milliseconds: [mesh_set, space_build, assemble,
 * precond_setup, rhs, solve, sample]                                                           */
int remo_stage_times(void* ctx, float* ms7);

#ifdef __cplusplus
}
#endif
#endif
