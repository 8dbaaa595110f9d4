"""ctypes binding of tests/cpp/emulate.cpp, the sequential CPU walk-through of the device
algorithm (test infrastructure: validates the product's symbolic analysis and the reduced
block-tridiagonal formulation against the oracle without a GPU)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        out = os.path.join(ROOT, "tests", "cpp", "build")
        os.makedirs(out, exist_ok=True)
        so = os.path.join(out, "libgomp_emul.so")
        srcs = [os.path.join(ROOT, "tests", "cpp", "emulate.cpp"), os.path.join(ROOT, "osqp-solver_b200", "csrc", "symbolic.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-ffp-contract=off", "-I",
                                   os.path.join(ROOT, "osqp-solver_b200", "csrc"), "-o", so, srcs[0]])
        L = C.CDLL(so)
        L.emul_create.restype = C.c_void_p
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def settings_array(st):
    """oracle Settings struct -> the 15 doubles emul_create takes."""
    return np.array([st.rho, st.sigma, st.alpha, st.scaling, st.adaptive_rho, st.adaptive_rho_interval,
                     st.adaptive_rho_tolerance, st.max_iter, st.eps_abs, st.eps_rel, st.eps_prim_inf, st.eps_dual_inf,
                     st.scaled_termination, st.check_termination, st.warm_start], dtype=np.float64)


class Emulator:
    def __init__(self, batch, n, m, Pp, Pi, Px, Ap, Ai, settings, var_order=None, block_size=0):
        self.batch, self.n, self.m = batch, n, m
        self._keep = [np.ascontiguousarray(a, dtype=np.int64) for a in (Pp, Pi, Ap, Ai)]
        self._Px = np.ascontiguousarray(Px, dtype=np.float64)
        vo = None if var_order is None else np.ascontiguousarray(var_order, dtype=np.int64)
        err = C.create_string_buffer(512)
        st = settings_array(settings)
        self.h = lib().emul_create(C.c_int64(batch), C.c_int64(n), C.c_int64(m), _p(self._keep[0]), _p(self._keep[1]),
                                   _p(self._Px), _p(self._keep[2]), _p(self._keep[3]), _p(st), _p(vo), C.c_int(block_size),
                                   err, C.c_int(512))
        if not self.h:
            raise ValueError(err.value.decode())
        d = np.zeros(8, dtype=np.int32)
        lib().emul_dims(C.c_void_p(self.h), _p(d))
        self.m_used, self.bs, self.T, self.n_pad, self.roww, self.colw, self.ow, self.owT = (int(v) for v in d)

    def __del__(self):
        if getattr(self, "h", None):
            lib().emul_destroy(C.c_void_p(self.h))
            self.h = None

    def set_data(self, Ax=None, l=None, u=None, q=None, cold=True):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (Ax, l, u, q)]
        rc = lib().emul_set_data(C.c_void_p(self.h), *[_p(a) for a in arrs], C.c_int(1 if cold else 0))
        if rc:
            raise ValueError("l > u")

    def warm_start_x(self, x0):
        lib().emul_warm_start_x(C.c_void_p(self.h), _p(np.ascontiguousarray(x0, dtype=np.float64)))

    def solve(self):
        lib().emul_solve(C.c_void_p(self.h))
        x = np.zeros((self.batch, self.n))
        y = np.zeros((self.batch, self.m))
        code = np.zeros(self.batch, dtype=np.int32)
        iters = np.zeros(self.batch, dtype=np.int32)
        rup = np.zeros(self.batch, dtype=np.int32)
        rho = np.zeros(self.batch)
        nf = np.zeros(self.batch, dtype=np.int32)
        lib().emul_get(C.c_void_p(self.h), _p(x), _p(y), _p(code), _p(iters), _p(rup), _p(rho), _p(nf))
        return dict(x=x, y=y, exit_code=code, iters=iters, rho_updates=rup, rho=rho, nfactor=nf)
