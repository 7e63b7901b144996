// Internal context of libremo3d_b200 (not part of the public ABI; see include/remo3d_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <mutex>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "../../include/remo3d_b200.h"

#define REMO_NSTAGE 7
enum { ST_MESH = 0, ST_SPACE, ST_ASM, ST_PRECOND, ST_RHS, ST_SOLVE, ST_SAMPLE };

struct Ctx;

// ---- error plumbing: every CUDA call goes through CK(), every entry point through the try/catch in cabi.cu
struct RemoError {
  int code;
  std::string msg;
};
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      char b__[512];                                                                               \
      snprintf(b__, sizeof b__, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      throw RemoError{REMO_ERR_CUDA, b__};                                                         \
    }                                                                                              \
  } while (0)
#define FAIL(code, ...)                      \
  do {                                       \
    char b__[512];                           \
    snprintf(b__, sizeof b__, __VA_ARGS__);  \
    throw RemoError{code, b__};              \
  } while (0)

// ---- stream-ordered device buffer (pool keeps freed blocks, so per-mesh re-allocation is cheap)
template <typename T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;    // elements in use
  size_t cap = 0;  // elements allocated
  void ensure(size_t count, cudaStream_t s) {
    if (count > cap) {
      if (p) CK(cudaFreeAsync(p, s));
      p = nullptr;
      cap = 0;
      size_t want = count + count / 8 + 64;
      CK(cudaMallocAsync((void**)&p, want * sizeof(T), s));
      cap = want;
    }
    n = count;
  }
  void release(cudaStream_t s) {
    if (p) cudaFreeAsync(p, s);
    p = nullptr;
    n = cap = 0;
  }
};

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  std::string err;
  int64_t launches = 0;
  int num_sms = 148;
  cudaEvent_t ev0[REMO_NSTAGE], ev1[REMO_NSTAGE];
  bool ev_set[REMO_NSTAGE] = {false};

  // ---- mesh (remo_mesh_set)
  int dim = 0;
  int64_t nv = 0, nt = 0, nb = 0, naxis = 0;
  DBuf<double> xyz;       // nv x dim
  DBuf<int32_t> elems;    // nt x (dim+1)
  DBuf<int32_t> mat;      // nt
  DBuf<int32_t> bfacets;  // nb x dim
  DBuf<uint8_t> bdir;     // nb
  DBuf<int32_t> axis_v;   // naxis
  DBuf<double> axis_z;    // naxis
  bool have_mesh = false;
  DBuf<double> bbox;  // [0..5] = min / max of the vertex coordinates (mesh_bbox), then per-block scratch
  bool have_bbox = false;

  // ---- space (remo_space_build)
  int order = 0, nld = 0, nle = 0, nlf = 0, npair = 0;
  int64_t ne = 0, nf = 0, ndof = 0, nnz = 0, nadj = 0;
  int64_t edge_base = 0, face_base = 0;
  DBuf<int32_t> sv;           // nt x (dim+1), vertices of every element sorted ascending
  DBuf<uint64_t> edge_keys;   // ne   (a<<32 | b), ascending  == lexicographic edge numbering
  DBuf<int32_t> elem_edges;   // nt x nle
  DBuf<uint64_t> face_keys;   // nf   (edge(a,b)<<32 | c), ascending == lexicographic face numbering
  DBuf<int32_t> elem_faces;   // nt x nlf
  DBuf<uint8_t> constrained;  // ndof
  DBuf<int64_t> adj_ptr;      // ndof+1: dof -> range in adj
  DBuf<uint32_t> adj;         // nadj = nt*nld: element*nld + local dof, elements ascending per dof
  DBuf<int64_t> rowptr;       // ndof+1
  DBuf<int32_t> col;          // nnz, ascending per row
  DBuf<double> val;           // nnz
  // have_space: topology, dofs, Dirichlet mask, adjacency.  have_pattern: rowptr / col.  have_matrix: the element metrics gm
  // of the current sigma (all the element-wise PCG path needs).  have_values: val filled for that gm.
  bool have_space = false, have_matrix = false, have_pattern = false, have_values = false;
  bool lazy_matrix = true;  // build the CSR pattern / values only on demand (remo_set_option("lazy_matrix", 0): eagerly)
  // sliced-ELLPACK copy of the matrix for the multi-RHS PCG SpMM (sell.cu): slices of 8 rows, chunks of 4 entries
  DBuf<int64_t> sell_ptr;   // nslices+1: first chunk of every slice
  DBuf<int32_t> sell_col;   // chunks x 8 rows x 4
  DBuf<double> sell_val;    // chunks x 8 rows x 4
  DBuf<int32_t> sell_row;   // row slot -> matrix row (sorted by length inside windows), -1 = padding
  DBuf<int64_t> sell_part;  // blocked distribution: first slice of every CTA (equal chunk counts)
  int sell_nparts = 0;
  DBuf<int64_t> sell_wpart;  // streaming kernel: first slice of every warp
  int sell_sgrid = 0;
  int64_t sell_chunks = 0, sell_slots = 0;
  bool have_sell = false;

  // ---- element-wise product (ebe.cu): batches of 256 (order-2 tets, order-3 triangles) or 128 (order-3 tets) Morton-ordered elements
  DBuf<int64_t> ebe_uoff;     // nb+1: first distinct dof of every batch
  DBuf<int32_t> ebe_udof;     // distinct dofs of every batch, bit 31 = constrained
  DBuf<uint16_t> ebe_lidx;    // nb x 10 x 256: (slot, tet) -> position in the batch's dof list
  DBuf<uint16_t> ebe_lpos;    // nb x 10 x 256: (slot, tet) -> position in the batch's dof-major scratch
  DBuf<uint16_t> ebe_ucnt;    // entries (incident tets) of every distinct dof inside its batch
  DBuf<uint16_t> ebe_jd;      // nb x 264: jagged-diagonal offsets of the batch's scratch
  DBuf<double> ebe_gm;        // nb x 10 x 256: metric numbers in batch order
  int64_t ebe_nb = 0;
  int ebe_umax = 0;
  int ebe_fast8 = 0;  // every dof piece has <= 8 entries and every batch <= 1024 of them: register tables in the kernel
  int ebe_occ[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // resident CTAs per SM by right-hand-side count
  bool have_ebe = false;
  int ebe_p3_ctas = 3;  // order-3 product kernel build: 3 (168 registers) or 4 (128 registers) resident CTAs per SM
  int ebe_ng = 10;   // metric numbers per element of the tables in use: 10 (tets) or 18 (axisymmetric triangles)
  int ebe_nld = 10;  // local dofs per tet of the tables in use: 10 (order 2, batches of 256 tets) or 20 (order 3, batches of 128)
  int ebe_check = 0;  // remo_set_option("ebe_check", 1): validate the batch tables after every build (ebe.cu k_ebe_check)
  int ebe_on = -1;  // remo_set_option("spmm_ebe"): 0 / 1, -1 = the REMO_SPMM_EBE environment default (on)

  // ---- numeric
  DBuf<double> gm;     // nt x npair: sigma |K| grad l_i . grad l_j
  DBuf<double> sigma;  // nmat
  DBuf<double> rvert;  // 2D: nt x 3 radii of the sorted vertices

  // ---- preconditioner
  int pkind = -1;
  DBuf<double> dinv;  // ndof: 1/diag on free dofs, 0 on constrained
  // "multigrid": aggregation AMG hierarchy on the vertex (P1) block, see amg.cu
  struct AmgLevel {
    int64_t n = 0, nnz = 0;
    DBuf<int64_t> rowptr;
    DBuf<int32_t> col;
    DBuf<double> val, dinv;   // dinv: l1-Jacobi weights 1 / sum_j |a_ij| (0 = row outside the coarse space)
    DBuf<double> diag;        // a_ii (0 on rows outside the coarse space): coupling strengths of the pairwise aggregation
    DBuf<int32_t> agg;        // fine row -> aggregate (coarse row) of the next level, -1 = none
    DBuf<int32_t> members;    // rows sorted by aggregate, ascending inside one: deterministic restriction
    DBuf<int32_t> aggptr;     // nc + 1: first entry of every aggregate in members
    double omega = 1.0;       // weight of the l1-Jacobi sweeps
    DBuf<double> b, x, t;     // n x nrhs work blocks (level 0 uses R / Z of the PCG directly for b / x)
    DBuf<float> valf, dinvf, bf, xf, tf;  // mixed-precision cycle (amg_fp32): fp32 copies of the matrix / weights, fp32 work blocks
  };
  std::vector<AmgLevel> amg;
  struct AmgTmp {  // Galerkin products between the passes of one level's pairwise aggregation
    int64_t n = 0, nnz = 0;
    DBuf<int64_t> rowptr;
    DBuf<int32_t> col;
    DBuf<double> val, diag;
  } amg_tmp[2];
  DBuf<int32_t> amg_w[5];    // match, pick, flag, scan, id of the current pass
  int amg_agg = 1;           // 1 = strength-based pairwise aggregation, 0 = Morton-rank aggregates of 8 (round 1)
  int amg_passes = 3;        // pairwise passes per level: aggregates of at most 2^passes rows
  int amg_rounds = 4;        // handshake rounds per pass
  int amg_fused_tail = 0;    // 1: the levels below amg_tail_rows rows run as one cluster kernel (amg.cu k_vcycle_tail); measured slower, off
  int64_t amg_tail_rows = 20000;
  double amg_alpha = 1.5, amg_omega_scale = 1.0;  // coarse-correction scaling, weight of the l1-Jacobi sweeps (<= 1)
  int amg_sweeps = 1;                              // pre = post smoothing sweeps
  int amg_lanes8 = 1;                              // k_smooth<8,4,2> instead of <4,4,2> for 5..8 right-hand sides
  int amg_fp32 = 1;                                // 1: the V-cycle runs in fp32 (fp64 PCG around it), amg.cu vcycle_levels<float>
  int amg_gamma = 1;                               // cycle index: 1 = V, 2 = W
  DBuf<double> amg_dense;     // inverse of the coarsest matrix (n x n)
  int amg_nrhs = 0;
  int amg_nlev = 0;           // levels in use (the vector keeps its buffers across meshes)

  // ---- right-hand sides / PCG state, row-major ndof x nrhs
  int nrhs = 0;       // internal column count = row stride of the vector blocks (user count rounded up to even)
  int nrhs_user = 0;  // right-hand sides the caller asked for
  int pstride = 0;    // row stride of the P block alone: = nrhs, or the next power of two when the SELL SpMM gathers it
  DBuf<double> F, X, R, P, Q;
  int kz = 0;            // row stride (even) of the V-cycle's blocks B0 / Zv
  DBuf<double> B0, Zv;   // nv x kz: residual of the vertex rows (right-hand side of the V-cycle) and its result
  DBuf<double> partial;  // per-block partial dot products
  DBuf<double> scal;     // device scalars, see solver.cu
  DBuf<int> iters_d;
  bool have_rhs = false, have_solution = false;

  // ---- per-launch SpMM timing inside remo_solve (remo_profile): CUDA events around every SpMM launch
  bool use_graph = true;  // replay one PCG iteration as a CUDA graph (off while profiling: events cannot sit inside a graph)
  bool prof = false;
  std::vector<cudaEvent_t> prof_ev;
  double prof_spmm_ms = 0.0;
  int64_t prof_spmm_n = 0;

  // ---- scratch: grow-only slots reused by every phase (no per-mesh cudaMallocAsync / cudaFreeAsync churn: freeing and
  //      re-allocating multi-GB temporaries through the pool costs 50-1500 ms at random, measured)
  DBuf<uint8_t> scr[10];
  DBuf<uint8_t> tmp;  // CUB temp storage
  std::vector<double> host_scal;
};

template <typename T>
static inline T* scratch(Ctx* c, int slot, size_t count) {
  c->scr[slot].ensure(count * sizeof(T) + 16, c->stream);
  return reinterpret_cast<T*>(c->scr[slot].p);
}

// ---- launch helper: counts launches (bench.py gpu_launches) and checks the launch
#define LAUNCH(ctx, kern, grid, block, smem, ...)                      \
  do {                                                                 \
    kern<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);     \
    (ctx)->launches++;                                                 \
    CK(cudaGetLastError());                                            \
  } while (0)

// Dynamic shared memory above 48 KB needs cudaFuncAttributeMaxDynamicSharedMemorySize, which is per-function,
// process-wide state: several contexts (host threads) solve different meshes on one GPU, so the limit is raised ONCE per
// (kernel, device) to the device's opt-in maximum and never lowered; callers only check their own size against it.
template <typename K>
static inline int allow_max_smem(K kern, int device) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> done;
  int dev_max = 0;
  CK(cudaDeviceGetAttribute(&dev_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  std::lock_guard<std::mutex> g(mu);
  const auto key = std::make_pair((const void*)kern, device);
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, kern));
  const int room = dev_max - (int)fa.sharedSizeBytes;  // static + dynamic shared memory share the opt-in limit
  if (!done.count(key)) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, room));
    done.insert(key);
  }
  return room;
}

static inline unsigned grid_for(int64_t n, int block, int64_t cap = (1 << 30)) {
  int64_t g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (unsigned)g;
}

struct StageTimer {
  Ctx* c;
  int st;
  StageTimer(Ctx* c_, int st_) : c(c_), st(st_) { cudaEventRecord(c->ev0[st], c->stream); }
  ~StageTimer() {
    cudaEventRecord(c->ev1[st], c->stream);
    c->ev_set[st] = true;
  }
};

// symbolic.cu
void space_build(Ctx* c, int order);
void pattern_build(Ctx* c);
int64_t vertex_block_pattern(Ctx* c, DBuf<int64_t>& rowptr, DBuf<int32_t>& col);
void topology_get(Ctx* c, int32_t* edges, int32_t* faces, int32_t* elem_edges, int32_t* elem_faces);
// assemble.cu
void assemble(Ctx* c, int nmat, const double* sigma);
void assemble_kernels_only(Ctx* c);
void ensure_values(Ctx* c);
void diag_from_elements(Ctx* c, double* dinv);
void assemble_vertex_block(Ctx* c, const int64_t* rowptr, const int32_t* col, double* val);
// solver.cu
void precond_setup(Ctx* c, int kind);
void rhs_point_sources(Ctx* c, int nrhs, const int64_t* src_ptr, const double* src_z, const double* src_fac);
int solve(Ctx* c, double rtol, int maxit, int* iters, double* relres);
void sample_axis(Ctx* c, int npts, const int32_t* pt_rhs, const double* z, double* out);
void apparent_resistivity(Ctx* c, int npts, const int32_t* pt_rhs, const double* z0, const double* z1, const double* k,
                          double scale, double* ra);
void launch_spmm(Ctx* c, const double* P, double* Q, int nrhs);
void launch_vector_updates(Ctx* c, int nrhs);
void alloc_solver_state(Ctx* c, int nrhs);
int spmm_variant();
void spmm_prepare(Ctx* c);
int spmm_kind(Ctx* c);
int solver_stride(Ctx* c, int nrhs);  // row stride of the PCG vector blocks for nrhs right-hand sides
int spmm_blocks(Ctx* c, int ks);   // CTAs (= partial-dot slots) of the SpMM launch for stride ks
// sell.cu
const double* mesh_bbox(Ctx* c);
int sell_pstride(int ks);
void sell_build(Ctx* c);
int sell_grid(const Ctx* c);
void launch_spmm_sell(Ctx* c, const double* P, double* Q, int ks, int pstride);
// ebe.cu
bool ebe_eligible(const Ctx* c);
bool ebe_serves(Ctx* c, int nr);
bool ebe_usable(const Ctx* c, int nr);
int ebe_max_rhs();
void ebe_build(Ctx* c);
int ebe_grid(const Ctx* c, int nr);
void launch_spmm_ebe(Ctx* c, const double* P, int pstride, double* Q, int ks, int nr);
// amg.cu
void amg_setup(Ctx* c);
void amg_build_hierarchy(Ctx* c);
void amg_apply(Ctx* c, const double* R, double* Z, int nrhs);
void amg_prepare(Ctx* c, int nrhs);
void amg_release(Ctx* c);
void spmm_smooth(Ctx* c, const int64_t* rowptr, const int32_t* col, const double* val, const double* dinv, const double* B,
                 const double* X, double* OUT, int k, int64_t n, double omega, int mode);
