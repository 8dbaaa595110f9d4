"""ctypes binding of the CPU oracle (oracle/osqp_ref.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by the product package.  PARITY UNPINNED
(see oracle/osqp_ref.h): a restatement of OSQP 0.6.x, not genuine OSQP.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "build", "libosqp_ref.so")

I64 = np.int64
_pd = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_pi = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_pi32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class Settings(C.Structure):
    _fields_ = [
        ("rho", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double),
        ("scaling", C.c_int), ("adaptive_rho", C.c_int), ("adaptive_rho_interval", C.c_int),
        ("adaptive_rho_tolerance", C.c_double), ("max_iter", C.c_int),
        ("eps_abs", C.c_double), ("eps_rel", C.c_double),
        ("eps_prim_inf", C.c_double), ("eps_dual_inf", C.c_double),
        ("scaled_termination", C.c_int), ("check_termination", C.c_int), ("warm_start", C.c_int),
    ]


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc, no dependencies)."""
    src = os.path.join(_HERE, "osqp_ref.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(os.path.join(_HERE, "osqp_ref.h"))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    L.oref_default_settings.argtypes = [C.POINTER(Settings)]
    L.oref_kkt_ordering.argtypes = [C.c_longlong, C.c_longlong, _pi, _pi, _pi, _pi, _pi]
    L.oref_setup.restype = C.c_void_p
    L.oref_setup.argtypes = [C.c_longlong, C.c_longlong, _pi, _pi, _pd, C.c_void_p, _pi, _pi, _pd,
                             _pd, _pd, C.POINTER(Settings), C.c_void_p]
    L.oref_cleanup.argtypes = [C.c_void_p]
    L.oref_warm_start_x.argtypes = [C.c_void_p, _pd]
    L.oref_update_A.argtypes = [C.c_void_p, _pd]
    L.oref_update_bounds.argtypes = [C.c_void_p, _pd, _pd]
    L.oref_solve.argtypes = [C.c_void_p]
    for f in ("oref_x", "oref_y", "oref_D", "oref_E", "oref_delta_y", "oref_delta_x"):
        getattr(L, f).restype = C.POINTER(C.c_double)
        getattr(L, f).argtypes = [C.c_void_p]
    for f in ("oref_status", "oref_exit_code", "oref_iter", "oref_rho_updates", "oref_num_factorizations"):
        getattr(L, f).restype = C.c_int
        getattr(L, f).argtypes = [C.c_void_p]
    for f in ("oref_rho", "oref_pri_res", "oref_dua_res", "oref_obj_val", "oref_c"):
        getattr(L, f).restype = C.c_double
        getattr(L, f).argtypes = [C.c_void_p]
    L.oref_nnz_L.restype = C.c_longlong
    L.oref_nnz_L.argtypes = [C.c_void_p]
    L.oref_solve_batch.restype = C.c_int
    L.oref_solve_batch.argtypes = [C.c_longlong, C.c_longlong, C.c_longlong, _pi, _pi, _pd, _pi, _pi,
                                   _pd, _pd, _pd, C.c_void_p, C.POINTER(Settings), C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.oref_solve_batch_passes.restype = C.c_int
    L.oref_solve_batch_passes.argtypes = [C.c_longlong, C.c_longlong, C.c_longlong, _pi, _pi, _pd, _pi, _pi,
                                          _pd, _pd, _pd, C.c_void_p, C.POINTER(Settings), C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.oref_max_threads.restype = C.c_int
    _lib = L
    return L


def default_settings(**kw):
    s = Settings()
    lib().oref_default_settings(C.byref(s))
    for k, v in kw.items():
        if not hasattr(s, k):
            raise KeyError(k)
        setattr(s, k, v)
    return s


def _csc(M):
    """scipy sparse / (indptr, indices, data) -> contiguous int64/float64 triple."""
    if hasattr(M, "tocsc"):
        M = M.tocsc()
        M.sort_indices()
        return (np.ascontiguousarray(M.indptr, dtype=I64), np.ascontiguousarray(M.indices, dtype=I64),
                np.ascontiguousarray(M.data, dtype=np.float64))
    p, i, x = M
    return (np.ascontiguousarray(p, dtype=I64), np.ascontiguousarray(i, dtype=I64),
            np.ascontiguousarray(x, dtype=np.float64))


class Solver:
    """Mirror of the reference's QPSolver call surface (src/osqp-wrapper.h:12-60) on the oracle."""

    def __init__(self, P, A, l, u, q=None, settings=None, perm=None):
        L = lib()
        self.Pp, self.Pi, self.Px = _csc(P)
        self.Ap, self.Ai, self.Ax = _csc(A)
        self.n = len(self.Pp) - 1
        self.m = len(l)
        self.settings = settings if settings is not None else default_settings()
        l = np.ascontiguousarray(l, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        qp = None
        if q is not None:
            self._q = np.ascontiguousarray(q, dtype=np.float64)
            qp = self._q.ctypes.data
        pp = None
        if perm is not None:
            self._perm = np.ascontiguousarray(perm, dtype=I64)
            pp = self._perm.ctypes.data
        self.h = L.oref_setup(self.n, self.m, self.Pp, self.Pi, self.Px, qp, self.Ap, self.Ai, self.Ax,
                              l, u, C.byref(self.settings), pp)
        if not self.h:
            raise ValueError("oref_setup failed (l > u or singular KKT)")

    def __del__(self):
        if getattr(self, "h", None):
            lib().oref_cleanup(self.h)
            self.h = None

    def warm_start_x(self, x):
        lib().oref_warm_start_x(self.h, np.ascontiguousarray(x, dtype=np.float64))

    def update(self, Ax=None, l=None, u=None):
        """reference src/osqp-wrapper.h:33-43: A values first, then bounds."""
        L = lib()
        if Ax is not None:
            if L.oref_update_A(self.h, np.ascontiguousarray(Ax, dtype=np.float64)):
                raise ValueError("update_A failed")
        if l is not None:
            if L.oref_update_bounds(self.h, np.ascontiguousarray(l, dtype=np.float64),
                                    np.ascontiguousarray(u, dtype=np.float64)):
                raise ValueError("update_bounds failed")

    def solve(self):
        L = lib()
        L.oref_solve(self.h)
        return self.exit_code, self.x

    @property
    def x(self):
        return np.ctypeslib.as_array(lib().oref_x(self.h), shape=(self.n,)).copy()

    @property
    def y(self):
        return np.ctypeslib.as_array(lib().oref_y(self.h), shape=(self.m,)).copy()

    @property
    def prim_inf_cert(self):
        return np.ctypeslib.as_array(lib().oref_delta_y(self.h), shape=(self.m,)).copy()

    @property
    def dual_inf_cert(self):
        return np.ctypeslib.as_array(lib().oref_delta_x(self.h), shape=(self.n,)).copy()

    @property
    def D(self):
        return np.ctypeslib.as_array(lib().oref_D(self.h), shape=(self.n,)).copy()

    @property
    def E(self):
        return np.ctypeslib.as_array(lib().oref_E(self.h), shape=(self.m,)).copy()

    c = property(lambda self: lib().oref_c(self.h))
    status = property(lambda self: lib().oref_status(self.h))
    exit_code = property(lambda self: lib().oref_exit_code(self.h))
    iter = property(lambda self: lib().oref_iter(self.h))
    rho_updates = property(lambda self: lib().oref_rho_updates(self.h))
    rho = property(lambda self: lib().oref_rho(self.h))
    pri_res = property(lambda self: lib().oref_pri_res(self.h))
    dua_res = property(lambda self: lib().oref_dua_res(self.h))
    obj_val = property(lambda self: lib().oref_obj_val(self.h))
    nnz_L = property(lambda self: lib().oref_nnz_L(self.h))
    num_factorizations = property(lambda self: lib().oref_num_factorizations(self.h))


def kkt_ordering(n, m, Pp, Pi, Ap, Ai):
    perm = np.zeros(n + m, dtype=I64)
    lib().oref_kkt_ordering(n, m, np.ascontiguousarray(Pp, dtype=I64), np.ascontiguousarray(Pi, dtype=I64),
                            np.ascontiguousarray(Ap, dtype=I64), np.ascontiguousarray(Ai, dtype=I64), perm)
    return perm


def solve_batch(n, m, P, A_pattern, Ax, l, u, x0=None, settings=None, nthreads=0, want_y=False, passes=1):
    """Cold setup + warm_start_x + solve for every problem of a batch that shares (P, pattern(A)).

    Ax: [batch, nnzA], l/u: [batch, m], x0: [batch, n] or None.
    passes > 1: that many passes over the batch as one continuous work queue (bench.py's CPU arm).
    Returns dict(x, y, exit_code, iters, rho_updates)."""
    L = lib()
    Pp, Pi, Px = _csc(P)
    Ap, Ai = (np.ascontiguousarray(A_pattern[0], dtype=I64), np.ascontiguousarray(A_pattern[1], dtype=I64))
    Ax = np.ascontiguousarray(Ax, dtype=np.float64)
    l = np.ascontiguousarray(l, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    batch = Ax.shape[0]
    settings = settings if settings is not None else default_settings()
    x = np.zeros((batch, n))
    y = np.zeros((batch, m)) if want_y else None
    code = np.zeros(batch, dtype=np.int32)
    iters = np.zeros(batch, dtype=np.int32)
    rup = np.zeros(batch, dtype=np.int32)
    x0p = None
    if x0 is not None:
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        x0p = x0.ctypes.data
    rc = L.oref_solve_batch_passes(batch, n, m, Pp, Pi, Px, Ap, Ai, Ax, l, u, x0p, C.byref(settings), nthreads, int(passes),
                                   x.ctypes.data, y.ctypes.data if want_y else None, code.ctypes.data,
                            iters.ctypes.data, rup.ctypes.data)
    if rc:
        raise ValueError("oref_solve_batch: a setup failed")
    return dict(x=x, y=y, exit_code=code, iters=iters, rho_updates=rup)


def max_threads():
    return lib().oref_max_threads()
