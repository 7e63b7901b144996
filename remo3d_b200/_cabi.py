"""ctypes binding of libremo3d_b200.so (ABI: include/remo3d_b200.h).

The library is the ONLY compute path of this package: there is no CPU or PyTorch fallback.  If the
shared object is missing or no B200 is visible, construction fails loudly (`RemoError`)."""
import ctypes as C
import os

import numpy as np

_LIB = None
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libremo3d_b200.so")

OK, ERR_CUDA, ERR_ARG, ERR_STATE, ERR_MESH, ERR_NOCONV = 0, -1, -2, -3, -4, -5
PRECOND = {"local": 0, "multigrid": 1}
MAX_RHS = 32

_p = C.c_void_p
_i64p = C.POINTER(C.c_int64)

# name -> (restype, argtypes); every symbol of include/remo3d_b200.h
SIGNATURES = {
    "remo_ctx_create": (C.c_int, [C.c_int, C.POINTER(_p)]),
    "remo_ctx_destroy": (C.c_int, [_p]),
    "remo_last_error": (C.c_char_p, [_p]),
    "remo_ctx_set_stream": (C.c_int, [_p, _p]),
    "remo_mesh_set": (C.c_int, [_p, C.c_int, C.c_int64, _p, C.c_int64, _p, _p, C.c_int64, _p, _p, C.c_int64, _p]),
    "remo_space_build": (C.c_int, [_p, C.c_int, _i64p, _i64p, _i64p, _i64p]),
    "remo_topology_get": (C.c_int, [_p, _p, _p, _p, _p]),
    "remo_assemble": (C.c_int, [_p, C.c_int, _p]),
    "remo_matrix_get": (C.c_int, [_p, _p, _p, _p]),
    "remo_matrix_nnz": (C.c_int, [_p, _i64p]),
    "remo_dirichlet_get": (C.c_int, [_p, _p]),
    "remo_precond_setup": (C.c_int, [_p, C.c_int]),
    "remo_precond_get": (C.c_int, [_p, _p, _p, _p, _p, C.POINTER(C.c_int), _p, _p]),
    "remo_rhs_point_sources": (C.c_int, [_p, C.c_int, _p, _p, _p]),
    "remo_rhs_get": (C.c_int, [_p, C.c_int, _p]),
    "remo_solve": (C.c_int, [_p, C.c_double, C.c_int, _p, _p]),
    "remo_sample_axis": (C.c_int, [_p, C.c_int, _p, _p, _p]),
    "remo_apparent_resistivity": (C.c_int, [_p, C.c_int, _p, _p, _p, _p, C.c_double, _p]),
    "remo_solution_get": (C.c_int, [_p, C.c_int, _p]),
    "remo_kernel_time": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "remo_spmm_apply": (C.c_int, [_p, C.c_int, _p, _p, _p]),
    "remo_set_option": (C.c_int, [_p, C.c_char_p, C.c_double]),
    "remo_profile": (C.c_int, [_p, C.c_int]),
    "remo_profile_get": (C.c_int, [_p, C.POINTER(C.c_double), _i64p]),
    "remo_launch_count": (C.c_int64, [_p]),
    "remo_spmm_kind": (C.c_int, [_p]),
    "remo_stage_times": (C.c_int, [_p, _p]),
}


class RemoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libremo3d_b200 error %d: %s" % (code, msg))
        self.code = code


def load():
    """Load the shared library (once) and declare every prototype."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RemoError(ERR_STATE, "%s not found: build it with `python build.py` (there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _LIB = lib
    return _LIB


def _ptr(a):
    """Raw pointer of a numpy array / torch tensor / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()  # torch tensor (host pinned or CUDA)


def _np(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


class Context:
    """One GPU worker: owns the device memory of one mesh at a time (include/remo3d_b200.h)."""

    def __init__(self, device=0):
        self.lib = load()
        h = _p()
        rc = self.lib.remo_ctx_create(int(device), C.byref(h))
        if rc != OK:
            raise RemoError(rc, (self.lib.remo_last_error(None) or b"").decode())
        self.h = h
        self.device = int(device)
        self.ndof = self.ne = self.nf = 0
        self._nnz = 0
        self.nrhs = 0
        self.order = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.remo_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_handle):
        """Adopt a caller-owned stream (int handle, e.g. torch.cuda.Stream().cuda_stream)."""
        self._ck(self.lib.remo_ctx_set_stream(self.h, C.c_void_p(int(cuda_stream_handle))))

    def _ck(self, rc, allow=()):
        if rc != OK and rc not in allow:
            raise RemoError(rc, (self.lib.remo_last_error(self.h) or b"").decode())
        return rc

    # ---- mesh / space / matrix
    def mesh_set(self, dim, points, elems, mat, bfacets, bdirichlet, axis_vertices):
        """Arrays may be numpy (host) or torch tensors (pinned host or CUDA); dtypes as in the header."""
        nv = points.shape[0]
        nt = elems.shape[0]
        nb = 0 if bfacets is None else bfacets.shape[0]
        na = 0 if axis_vertices is None else axis_vertices.shape[0]
        self._keep = (points, elems, mat, bfacets, bdirichlet, axis_vertices)
        self._ck(self.lib.remo_mesh_set(self.h, dim, nv, _ptr(points), nt, _ptr(elems), _ptr(mat), nb, _ptr(bfacets),
                                         _ptr(bdirichlet), na, _ptr(axis_vertices)))
        self._keep = None
        self.dim, self.nv, self.nt = dim, nv, nt

    def space_build(self, order):
        out = [C.c_int64() for _ in range(4)]
        self._ck(self.lib.remo_space_build(self.h, int(order), *[C.byref(o) for o in out]))
        self.ndof, self._nnz, self.ne, self.nf = (o.value for o in out)
        self.order = int(order)
        return self.ndof, self._nnz

    @property
    def nnz(self):
        """Non-zeros of the CSR pattern.  The pattern is built on demand (the element-wise PCG path never needs it), so
        the first read after `space_build` may cost a kernel pass."""
        if not self._nnz and self.ndof:
            n = C.c_int64()
            self._ck(self.lib.remo_matrix_nnz(self.h, C.byref(n)))
            self._nnz = n.value
        return self._nnz

    def topology(self):
        nle = 6 if self.dim == 3 else 3
        edges = np.empty((self.ne, 2), np.int32)
        faces = np.empty((self.nf if self.dim == 3 else 0, 3), np.int32)
        ee = np.empty((self.nt, nle), np.int32)
        ef = np.empty((self.nt, 4), np.int32) if (self.dim == 3 and self.order == 3) else None
        self._ck(self.lib.remo_topology_get(self.h, _ptr(edges), _ptr(faces) if faces.size else None, _ptr(ee), _ptr(ef)))
        return edges, faces, ee, ef

    def assemble(self, sigma):
        s = _np(sigma, np.float64)
        self._ck(self.lib.remo_assemble(self.h, s.shape[0], _ptr(s)))

    def matrix(self, values=True):
        rowptr = np.empty(self.ndof + 1, np.int64)
        col = np.empty(self.nnz, np.int32)
        val = np.empty(self.nnz, np.float64) if values else None
        self._ck(self.lib.remo_matrix_get(self.h, _ptr(rowptr), _ptr(col), _ptr(val)))
        return rowptr, col, val

    def dirichlet(self):
        m = np.empty(self.ndof, np.uint8)
        self._ck(self.lib.remo_dirichlet_get(self.h, _ptr(m)))
        return m.astype(bool)

    def precond_setup(self, kind):
        self._ck(self.lib.remo_precond_setup(self.h, PRECOND[kind] if isinstance(kind, str) else int(kind)))

    def precond_get(self, vertex_block=False):
        """(dinv, vertex block CSR or None, [(rows, nnz) per level of the aggregation hierarchy]) -- parity export."""
        dinv = np.empty(self.ndof, np.float64)
        nlev = C.c_int(16)
        rows, nnzs = np.zeros(16, np.int64), np.zeros(16, np.int64)
        vv = None
        if vertex_block:
            n = self.nv + 2 * self.ne
            vv = (np.empty(self.nv + 1, np.int64), np.empty(n, np.int32), np.empty(n, np.float64))
        self._ck(self.lib.remo_precond_get(self.h, _ptr(dinv), *([_ptr(a) for a in vv] if vv else [None, None, None]), C.byref(nlev),
                                            _ptr(rows), _ptr(nnzs)))
        return dinv, vv, [(int(rows[l]), int(nnzs[l])) for l in range(nlev.value)]

    # ---- right-hand sides / solve / sampling
    def rhs_point_sources(self, src_ptr, src_z, src_fac):
        sp, sz, sf = _np(src_ptr, np.int64), _np(src_z, np.float64), _np(src_fac, np.float64)
        nrhs = sp.shape[0] - 1
        self._ck(self.lib.remo_rhs_point_sources(self.h, nrhs, _ptr(sp), _ptr(sz), _ptr(sf)))
        self.nrhs = nrhs

    def rhs(self, r):
        f = np.empty(self.ndof, np.float64)
        self._ck(self.lib.remo_rhs_get(self.h, int(r), _ptr(f)))
        return f

    def solve(self, rtol=1e-10, maxit=1000, raise_on_noconv=True):
        iters = np.zeros(self.nrhs, np.int32)
        relres = np.zeros(self.nrhs, np.float64)
        rc = self._ck(self.lib.remo_solve(self.h, float(rtol), int(maxit), _ptr(iters), _ptr(relres)), allow=(ERR_NOCONV,))
        if rc == ERR_NOCONV and raise_on_noconv:
            raise RemoError(rc, "PCG did not reach rtol=%g within %d iterations (relres %s)" % (rtol, maxit, relres))
        return iters, relres

    def sample_axis(self, z, rhs=None):
        z = _np(np.atleast_1d(z), np.float64)
        r = None if rhs is None else _np(np.broadcast_to(rhs, z.shape), np.int32)
        out = np.empty(z.shape[0], np.float64)
        self._ck(self.lib.remo_sample_axis(self.h, z.shape[0], _ptr(r), _ptr(z), _ptr(out)))
        return out

    def apparent_resistivity(self, pt_rhs, z0, z1, k, scale, out=None):
        pr, a0, a1, kk = _np(pt_rhs, np.int32), _np(z0, np.float64), _np(z1, np.float64), _np(k, np.float64)
        ra = np.empty(pr.shape[0], np.float64) if out is None else out
        self._ck(self.lib.remo_apparent_resistivity(self.h, pr.shape[0], _ptr(pr), _ptr(a0), _ptr(a1), _ptr(kk), float(scale), _ptr(ra)))
        return ra

    def solution(self, r=0):
        u = np.empty(self.ndof, np.float64)
        self._ck(self.lib.remo_solution_get(self.h, int(r), _ptr(u)))
        return u

    # ---- measurement
    def kernel_time(self, which, nrhs, reps):
        ms = C.c_float()
        self._ck(self.lib.remo_kernel_time(self.h, int(which), int(nrhs), int(reps), C.byref(ms)))
        return ms.value

    def spmm_apply(self, p):
        """One launch of the PCG SpMM on host search directions p (ndof x nrhs) -> (Q = A P, fused dots p.q)."""
        p = np.ascontiguousarray(p, dtype=np.float64)
        if p.ndim != 2 or p.shape[0] != self.ndof:
            raise ValueError("spmm_apply: p must be ndof x nrhs")
        q = np.empty_like(p)
        pq = np.empty(p.shape[1])
        self._ck(self.lib.remo_spmm_apply(self.h, p.shape[1], p.ctypes.data, q.ctypes.data, pq.ctypes.data))
        return q, pq

    def spmm_kind(self):
        """0 = CSR, 1 = SELL copy, 2 = element-wise product (for the right-hand sides currently set)."""
        return int(self.lib.remo_spmm_kind(self.h))

    def set_option(self, name, value):
        self._ck(self.lib.remo_set_option(self.h, name.encode(), float(value)))

    def profile(self, on=True):
        self._ck(self.lib.remo_profile(self.h, 1 if on else 0))

    def profile_get(self):
        ms, n = C.c_double(), C.c_int64()
        self._ck(self.lib.remo_profile_get(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def launch_count(self):
        return int(self.lib.remo_launch_count(self.h))

    def stage_times(self):
        t = np.zeros(7, np.float32)
        self._ck(self.lib.remo_stage_times(self.h, _ptr(t)))
        return dict(zip(("mesh_set", "space_build", "assemble", "precond_setup", "rhs", "solve", "sample"), t.tolist()))
