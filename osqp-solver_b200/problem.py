"""ctypes view of libgomp_host (include/gomp_problem.h): synthetic GOMP trajectory QPs.

The generator runs the C++ facade's ConstraintBuilder<N> (the drop-in for reference
src/constraints/constraint-builder.h) for every problem of a batch; see DESIGN.md for the
concrete instance definition.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class ProblemConfig(C.Structure):
    _fields_ = [
        ("n_dim", C.c_int32), ("waypoints", C.c_int32), ("time_step", C.c_double),
        ("pos_limit", C.c_double), ("vel_limit", C.c_double), ("acc_limit", C.c_double),
        ("n_balls", C.c_int32), ("n_obstacles", C.c_int32), ("z_floor", C.c_double), ("seed", C.c_uint64),
    ]


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libgomp_host.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(path)
        for name in ("gomp_problem_default_config", "gomp_problem_dims", "gomp_problem_pattern",
                     "gomp_problem_generate", "gomp_problem_relinearize", "gomp_problem_check", "gomp_problem_fk"):
            getattr(L, name).restype = C.c_int
        _LIB = L
    return _LIB


def default_config(**kw):
    cfg = ProblemConfig()
    lib().gomp_problem_default_config(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise KeyError(k)
        setattr(cfg, k, v)
    return cfg


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Pattern:
    """Shared structure of one problem family: dims, CSC patterns of A and P, stage order."""

    def __init__(self, cfg):
        self.cfg = cfg
        n, m, nnzA, nnzP = (C.c_int64() for _ in range(4))
        rc = lib().gomp_problem_dims(C.byref(cfg), C.byref(n), C.byref(m), C.byref(nnzA), C.byref(nnzP))
        if rc:
            raise ValueError("invalid gomp_problem_config")
        self.n, self.m, self.nnzA, self.nnzP = n.value, m.value, nnzA.value, nnzP.value
        self.A_colptr = np.zeros(self.n + 1, dtype=np.int64)
        self.A_rowidx = np.zeros(self.nnzA, dtype=np.int64)
        self.P_colptr = np.zeros(self.n + 1, dtype=np.int64)
        self.P_rowidx = np.zeros(self.nnzP, dtype=np.int64)
        self.P_val = np.zeros(self.nnzP, dtype=np.float64)
        self.var_order = np.zeros(self.n, dtype=np.int64)
        rc = lib().gomp_problem_pattern(C.byref(cfg), _p(self.A_colptr), _p(self.A_rowidx), _p(self.P_colptr),
                                        _p(self.P_rowidx), _p(self.P_val), _p(self.var_order))
        if rc:
            raise ValueError("gomp_problem_pattern failed")
        self.block_size = 2 * cfg.n_dim

    def P_scipy(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.P_val, self.P_rowidx, self.P_colptr), shape=(self.n, self.n))

    def A_scipy(self, A_val):
        import scipy.sparse as sp
        return sp.csc_matrix((A_val, self.A_rowidx, self.A_colptr), shape=(self.m, self.n))


def generate(pattern, first_id, count, nthreads=0, out=None):
    """A_val [count, nnzA], l, u [count, m], x0 [count, n], start_goal [count, 2N].

    `out` may hold preallocated arrays (e.g. views of pinned torch tensors) under the same keys."""
    cfg = pattern.cfg
    out = dict(out or {})
    shapes = dict(A_val=(count, pattern.nnzA), l=(count, pattern.m), u=(count, pattern.m), x0=(count, pattern.n),
                  start_goal=(count, 2 * cfg.n_dim))
    for k, shp in shapes.items():
        if k not in out:
            out[k] = np.empty(shp, dtype=np.float64)
        assert out[k].shape == shp and out[k].dtype == np.float64 and out[k].flags.c_contiguous
    rc = lib().gomp_problem_generate(C.byref(cfg), C.c_int64(first_id), C.c_int64(count), _p(out["A_val"]), _p(out["l"]),
                                     _p(out["u"]), _p(out["x0"]), _p(out["start_goal"]), C.c_int32(nthreads))
    if rc:
        raise ValueError("gomp_problem_generate failed")
    return out


def relinearize(pattern, first_id, x, nthreads=0, out=None):
    """One SQP round (reference src/gomp-solver.h:83-87): new A_val, l, u around trajectories x [count, n]."""
    cfg = pattern.cfg
    x = np.ascontiguousarray(x, dtype=np.float64)
    count = x.shape[0]
    out = dict(out or {})
    shapes = dict(A_val=(count, pattern.nnzA), l=(count, pattern.m), u=(count, pattern.m))
    for k, shp in shapes.items():
        if k not in out:
            out[k] = np.empty(shp, dtype=np.float64)
    rc = lib().gomp_problem_relinearize(C.byref(cfg), C.c_int64(first_id), C.c_int64(count), _p(x), _p(out["A_val"]),
                                        _p(out["l"]), _p(out["u"]), C.c_int32(nthreads))
    if rc:
        raise ValueError("gomp_problem_relinearize failed")
    return out


def check(pattern, first_id, x, nthreads=0):
    """isSolutionOK (reference src/gomp-solver.h:141-199) per trajectory."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    ok = np.zeros(x.shape[0], dtype=np.int32)
    rc = lib().gomp_problem_check(C.byref(pattern.cfg), C.c_int64(first_id), C.c_int64(x.shape[0]), _p(x), _p(ok),
                                  C.c_int32(nthreads))
    if rc:
        raise ValueError("gomp_problem_check failed")
    return ok.astype(bool)


def fk(cfg, q, ball=0, want_jac=False):
    q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, cfg.n_dim)
    xyz = np.zeros((q.shape[0], 3))
    jac = np.zeros((q.shape[0], 3, cfg.n_dim)) if want_jac else None
    rc = lib().gomp_problem_fk(C.byref(cfg), C.c_int32(ball), C.c_int64(q.shape[0]), _p(q), _p(xyz), _p(jac))
    if rc:
        raise ValueError("gomp_problem_fk failed")
    return (xyz, jac) if want_jac else xyz


def start_goal(pattern, first_id, count, nthreads=0, out=None):
    """start / goal joint vectors of the problems only: [count, 2N] (what device-side assembly needs from the host)."""
    cfg = pattern.cfg
    sg = out if out is not None else np.empty((count, 2 * cfg.n_dim), dtype=np.float64)
    assert sg.shape == (count, 2 * cfg.n_dim) and sg.dtype == np.float64 and sg.flags.c_contiguous
    rc = lib().gomp_problem_generate(C.byref(cfg), C.c_int64(first_id), C.c_int64(count), None, None, None, None, _p(sg), C.c_int32(nthreads))
    if rc:
        raise ValueError("gomp_problem_generate failed")
    return sg


def scene(pattern, first_id, count):
    """Obstacle lines of the problems: [count, n_obstacles, 5] = unit direction (x, y), point (x, y, z)."""
    cfg = pattern.cfg
    lines = np.zeros((count, max(cfg.n_obstacles, 1), 5), dtype=np.float64)[:, :cfg.n_obstacles]
    lines = np.ascontiguousarray(lines)
    rc = lib().gomp_problem_scene(C.byref(cfg), C.c_int64(first_id), C.c_int64(count), _p(lines) if lines.size else None)
    if rc:
        raise ValueError("gomp_problem_scene failed")
    return lines


def cart_rows(pattern):
    """Cartesian row table in withObstacles order: (rows [R] int32, kidx [R, N] int32)."""
    cfg = pattern.cfg
    nr = C.c_int64()
    if lib().gomp_problem_cart_rows(C.byref(cfg), C.byref(nr), None, None):
        raise ValueError("gomp_problem_cart_rows failed")
    rows = np.zeros(nr.value, dtype=np.int32)
    kidx = np.zeros((nr.value, cfg.n_dim), dtype=np.int32)
    rc = lib().gomp_problem_cart_rows(C.byref(cfg), C.byref(nr), _p(rows), _p(kidx))
    if rc:
        raise ValueError(f"gomp_problem_cart_rows failed ({rc})")
    return rows, kidx


class DeviceScene:
    """Device-side withObstacles / isSolutionOK (include/gomp_problem.h, implemented in libgomp_cuda):
    re-linearises the Cartesian rows of A_val / l / u of `count` problems in place, on the GPU.
    Arguments of relinearize / check are CUDA torch tensors (or raw device pointers as ints)."""

    def __init__(self, pattern, first_id, count, device=-1):
        """first_id = None: the obstacle lines are not uploaded but derived on the device by `assemble`."""
        from . import solver as _solver
        self._L = _solver.lib()
        self.pattern, self.count = pattern, count
        lines = scene(pattern, first_id, count) if first_id is not None else np.zeros((0,))
        rows, kidx = cart_rows(pattern)
        self.h = C.c_void_p()
        rc = self._L.gomp_device_scene_create(C.byref(self.h), C.byref(pattern.cfg), C.c_int64(count), C.c_int64(pattern.n),
                                              C.c_int64(pattern.m), C.c_int64(pattern.nnzA), _p(lines) if lines.size else None,
                                              C.c_int64(len(rows)), _p(rows), _p(kidx), C.c_int32(device))
        if rc:
            raise RuntimeError(f"gomp_device_scene_create failed ({rc}): no CUDA device or table mismatch")

    @staticmethod
    def _dp(t):
        if t is None:
            return None
        return C.c_void_p(t.data_ptr() if hasattr(t, "data_ptr") else int(t))

    def relinearize(self, x, A_val, l, u, x_fallback=None, stream=None):
        rc = self._L.gomp_device_relinearize(self.h, self._dp(x), self._dp(x_fallback), self._dp(A_val), self._dp(l), self._dp(u),
                                             C.c_void_p(stream or 0))
        if rc:
            raise RuntimeError(f"gomp_device_relinearize failed ({rc})")

    def set_template(self, template=None):
        """(A_val, l, u) of any one problem of the configuration: the problem-independent rows of device-side assembly."""
        if template is None:
            template = generate(self.pattern, 0, 1)
        t = [np.ascontiguousarray(np.asarray(template[k]).reshape(-1), dtype=np.float64) for k in ("A_val", "l", "u")]
        assert t[0].size == self.pattern.nnzA and t[1].size == self.pattern.m and t[2].size == self.pattern.m
        rc = self._L.gomp_device_scene_set_template(self.h, _p(t[0]), _p(t[1]), _p(t[2]))
        if rc:
            raise RuntimeError(f"gomp_device_scene_set_template failed ({rc})")

    def assemble(self, start_goal, A_val, l, u, x0, stream=None):
        """gomp_problem_generate on the device: start_goal [count, 2N] -> A_val, l, u, x0 (all CUDA tensors)."""
        rc = self._L.gomp_device_assemble(self.h, self._dp(start_goal), self._dp(A_val), self._dp(l), self._dp(u), self._dp(x0),
                                          C.c_void_p(stream or 0))
        if rc:
            raise RuntimeError(f"gomp_device_assemble failed ({rc})")

    def check(self, x, ok, stream=None):
        rc = self._L.gomp_device_check(self.h, self._dp(x), self._dp(ok), C.c_void_p(stream or 0))
        if rc:
            raise RuntimeError(f"gomp_device_check failed ({rc})")

    def close(self):
        if self.h:
            self._L.gomp_device_scene_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
