"""ctypes view of libgomp_cuda (include/gomp_cuda.h) plus a thin batch-solver class.

`BatchQPSolver` mirrors the reference's QPSolver call surface (src/osqp-wrapper.h:12-60:
ctor / setWarmStart / solve / update) for a whole batch of QPs sharing one sparsity pattern.
Array arguments may be numpy arrays (host) or torch CUDA tensors (device pointers are passed
through unchanged).  There is no CPU fallback: without the CUDA library or a device, every call
raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

EXIT_NAMES = ["kOptimal", "kPrimalInfeasible", "kDualInfeasible", "kOptimalInaccurate", "kPrimalInfeasibleInaccurate",
              "kDualInfeasibleInaccurate", "kMaxIterations", "kInterrupted", "kTimeLimitReached", "kNonConvex", "kUnknown"]

EXPORTS = ["gomp_settings_default", "gomp_batch_create", "gomp_batch_setup", "gomp_batch_warm_start_x", "gomp_batch_solve",
           "gomp_batch_update", "gomp_batch_set_active", "gomp_batch_set_max_ctas", "gomp_batch_get", "gomp_batch_get_async", "gomp_batch_get_info", "gomp_batch_get_certificates", "gomp_batch_info", "gomp_batch_sync",
           "gomp_batch_destroy", "gomp_batch_device_views", "gomp_device_malloc", "gomp_device_free", "gomp_device_memcpy",
           "gomp_measure_peaks", "gomp_last_error_string", "gomp_device_count", "gomp_version",
           "gomp_mg_create", "gomp_mg_setup", "gomp_mg_warm_start_x", "gomp_mg_update", "gomp_mg_solve", "gomp_mg_get", "gomp_mg_shard",
           "gomp_mg_num_shards", "gomp_mg_destroy"]
# include/gomp_problem.h, device section (same library)
DEVICE_EXPORTS = ["gomp_device_scene_create", "gomp_device_scene_set_template", "gomp_device_assemble", "gomp_device_relinearize",
                  "gomp_device_check", "gomp_device_scene_destroy"]


class Settings(C.Structure):
    _fields_ = [
        ("rho", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double),
        ("scaling", C.c_int32), ("adaptive_rho", C.c_int32), ("adaptive_rho_interval", C.c_int32),
        ("adaptive_rho_tolerance", C.c_double), ("max_iter", C.c_int32),
        ("eps_abs", C.c_double), ("eps_rel", C.c_double), ("eps_prim_inf", C.c_double), ("eps_dual_inf", C.c_double),
        ("scaled_termination", C.c_int32), ("check_termination", C.c_int32), ("warm_start", C.c_int32),
        ("verbose", C.c_int32), ("device", C.c_int32), ("stream", C.c_void_p), ("linsys", C.c_int32),
        ("block_size", C.c_int32), ("var_order", C.c_void_p), ("threads_per_cta", C.c_int32), ("borrow_inputs", C.c_int32),
        ("prefer_throughput", C.c_int32),
    ]


class Info(C.Structure):
    _fields_ = [
        ("batch", C.c_int64), ("n", C.c_int64), ("m", C.c_int64), ("m_used", C.c_int64), ("nnzA", C.c_int64), ("nnzP", C.c_int64),
        ("block_size", C.c_int32), ("num_blocks", C.c_int32), ("offdiag_width", C.c_int32),
        ("linsys", C.c_int32), ("threads_per_cta", C.c_int32), ("ctas_per_sm", C.c_int32), ("num_sms", C.c_int32), ("grid", C.c_int32),
        ("qps_per_sm", C.c_int32), ("reserved0", C.c_int32),
        ("smem_bytes", C.c_int64), ("workspace_bytes", C.c_int64), ("kernel_launches", C.c_int64),
        ("last_solve_ms", C.c_double), ("last_h2d_ms", C.c_double),
        ("flops_per_iter", C.c_int64), ("smem_bytes_per_iter", C.c_int64),
        ("flops_per_factor", C.c_int64), ("smem_bytes_per_factor", C.c_int64),
        ("flops_per_scaling", C.c_int64), ("smem_bytes_per_scaling", C.c_int64),
        ("flops_per_check", C.c_int64), ("smem_bytes_per_check", C.c_int64),
        ("phase_cycles", C.c_double * 8), ("chain_cycles", C.c_double), ("assemble_cycles", C.c_double),
    ]


def lib_path():
    # GOMP_CUDA_LIB: an alternative build of the same library (A/B measurements of kernel variants)
    return os.environ.get("GOMP_CUDA_LIB") or os.path.join(_HERE, "libgomp_cuda.so")


def lib():
    """Load libgomp_cuda.so; fails loudly when it has not been built (no fallback path exists)."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: build it with `make -C osqp-solver_b200` (or __graft_entry__.build())")
        L = C.CDLL(path)
        for name in EXPORTS + DEVICE_EXPORTS:
            getattr(L, name)  # every symbol of include/gomp_cuda.h must be exported
        L.gomp_last_error_string.restype = C.c_char_p
        L.gomp_version.restype = C.c_char_p
        _LIB = L
    return _LIB


class GompError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libgomp_cuda error {code}: {msg}")
        self.code = code


def _check(rc):
    if rc != 0:
        raise GompError(rc, lib().gomp_last_error_string().decode())


def default_settings(**kw):
    s = Settings()
    _check(lib().gomp_settings_default(C.byref(s)))
    for k, v in kw.items():
        if not hasattr(s, k):
            raise KeyError(k)
        setattr(s, k, v)
    return s


def _ptr(a, dtype=np.float64):
    """numpy array / torch tensor / None -> (void*, keepalive)."""
    if a is None:
        return None, None
    if hasattr(a, "data_ptr"):  # torch tensor (host or CUDA)
        assert a.is_contiguous()
        return C.c_void_p(a.data_ptr()), a
    arr = np.ascontiguousarray(a, dtype=dtype)
    return C.c_void_p(arr.ctypes.data), arr


class BatchQPSolver:
    """A batch of QPs  min 1/2 x'Px  s.t. l <= Ax <= u  sharing P and the pattern of A."""

    def __init__(self, n, m, P_colptr, P_rowidx, P_val, A_colptr, A_rowidx, A_val=None, l=None, u=None, q=None, batch=None,
                 settings=None, var_order=None, block_size=0):
        L = lib()
        self.n, self.m = int(n), int(m)
        self.nnzA = int(A_colptr[-1])
        if batch is None:
            batch = A_val.shape[0]
        self.batch = int(batch)
        self.settings = settings if settings is not None else default_settings()
        self._keep = []
        if var_order is not None:
            vo = np.ascontiguousarray(var_order, dtype=np.int64)
            self._keep.append(vo)
            self.settings.var_order = vo.ctypes.data
            self.settings.block_size = int(block_size)
        elif block_size:
            self.settings.block_size = int(block_size)
        idx = [np.ascontiguousarray(a, dtype=np.int64) for a in (P_colptr, P_rowidx, A_colptr, A_rowidx)]
        pv = np.ascontiguousarray(P_val, dtype=np.float64)
        self._keep += idx + [pv]
        ptrs = [_ptr(a) for a in (A_val, l, u, q)]
        self.h = C.c_void_p()
        _check(L.gomp_batch_create(C.byref(self.h), C.byref(self.settings), C.c_int64(self.batch), C.c_int64(self.n), C.c_int64(self.m),
                                   C.c_void_p(idx[0].ctypes.data), C.c_void_p(idx[1].ctypes.data), C.c_void_p(pv.ctypes.data),
                                   C.c_void_p(idx[2].ctypes.data), C.c_void_p(idx[3].ctypes.data),
                                   ptrs[0][0], ptrs[1][0], ptrs[2][0], ptrs[3][0]))
        self._borrowed = ptrs if self.settings.borrow_inputs else None

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            lib().gomp_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setup(self, A_val, l, u, q=None):
        """Fresh QPSolver on the same pattern (reference src/gomp-solver.h:61-64): cold start."""
        ptrs = [_ptr(a) for a in (A_val, l, u, q)]
        _check(lib().gomp_batch_setup(self.h, ptrs[0][0], ptrs[1][0], ptrs[2][0], ptrs[3][0]))
        if self.settings.borrow_inputs:
            self._borrowed = ptrs

    def set_warm_start(self, x0):
        p, keep = _ptr(x0)
        _check(lib().gomp_batch_warm_start_x(self.h, p))

    def update(self, A_val=None, l=None, u=None, q=None):
        """reference src/osqp-wrapper.h:33-43: new A values (same pattern) and bounds; iterates kept."""
        ptrs = [_ptr(a) for a in (A_val, l, u, q)]
        _check(lib().gomp_batch_update(self.h, ptrs[0][0], ptrs[1][0], ptrs[2][0], ptrs[3][0]))
        if self.settings.borrow_inputs:
            self._borrowed = ptrs

    def set_active(self, active):
        """Mask for the following solves: int32 [batch] (numpy or torch CUDA tensor), None = all problems."""
        p, keep = _ptr(active, dtype=np.int32)
        _check(lib().gomp_batch_set_active(self.h, p))
        self._active_keep = keep

    def set_max_ctas(self, max_ctas):
        """Cap on the CTAs of the following solves (0 = all): how a caller with several batches in flight shares the SMs."""
        _check(lib().gomp_batch_set_max_ctas(self.h, C.c_int32(int(max_ctas))))

    def solve_async(self):
        _check(lib().gomp_batch_solve(self.h))

    def sync(self):
        _check(lib().gomp_batch_sync(self.h))

    def get(self, x=None, y=None, want_y=False, exit_code=None, iters=None, rho_updates=None):
        """Fetch results into numpy arrays (allocated here unless given; torch tensors accepted)."""
        if x is None:
            x = np.empty((self.batch, self.n))
        if y is None and want_y:
            y = np.empty((self.batch, self.m))
        if exit_code is None:
            exit_code = np.empty(self.batch, dtype=np.int32)
        if iters is None:
            iters = np.empty(self.batch, dtype=np.int32)
        if rho_updates is None:
            rho_updates = np.empty(self.batch, dtype=np.int32)
        px, _ = _ptr(x) if hasattr(x, "data_ptr") else (C.c_void_p(x.ctypes.data), x)
        py = None
        if y is not None:
            py = C.c_void_p(y.data_ptr()) if hasattr(y, "data_ptr") else C.c_void_p(y.ctypes.data)
        _check(lib().gomp_batch_get(self.h, px, py, C.c_void_p(exit_code.ctypes.data), C.c_void_p(iters.ctypes.data),
                                    C.c_void_p(rho_updates.ctypes.data)))
        return dict(x=x, y=y, exit_code=exit_code, iters=iters, rho_updates=rho_updates)

    def get_async(self, x=None, y=None, exit_code=None, iters=None, rho_updates=None):
        """Queue the result copies on the handle's stream without waiting (device tensors / pinned host arrays)."""
        ptrs = [_ptr(a, dtype=np.float64 if k < 2 else np.int32)[0] for k, a in enumerate((x, y, exit_code, iters, rho_updates))]
        _check(lib().gomp_batch_get_async(self.h, *ptrs))

    def solve(self, want_y=False):
        """QPSolver::solve for the whole batch: run, wait, fetch."""
        self.solve_async()
        return self.get(want_y=want_y)

    def diagnostics(self):
        pri = np.empty(self.batch)
        dua = np.empty(self.batch)
        rho = np.empty(self.batch)
        _check(lib().gomp_batch_get_info(self.h, C.c_void_p(pri.ctypes.data), C.c_void_p(dua.ctypes.data), C.c_void_p(rho.ctypes.data), None))
        return dict(pri_res=pri, dua_res=dua, rho=rho)

    def certificates(self, primal=True, dual=False):
        """Infeasibility certificates of the last solve: (prim_inf_cert [batch, m] | None, dual_inf_cert [batch, n] | None);
        NaN rows for problems whose status is not (inaccurately) primal / dual infeasible."""
        pc = np.empty((self.batch, self.m)) if primal else None
        dc = np.empty((self.batch, self.n)) if dual else None
        _check(lib().gomp_batch_get_certificates(self.h, None if pc is None else C.c_void_p(pc.ctypes.data),
                                                 None if dc is None else C.c_void_p(dc.ctypes.data)))
        return pc, dc

    def info(self):
        inf = Info()
        _check(lib().gomp_batch_info(self.h, C.byref(inf)))
        out = {k: getattr(inf, k) for k, _ in Info._fields_}
        out["phase_cycles"] = list(inf.phase_cycles)
        return out


class MultiGpuSolver:
    """One batch interleaved over several GPUs of this process (include/gomp_cuda.h, gomp_mg_*): same call surface as
    BatchQPSolver with whole-batch arrays in the caller's order."""

    def __init__(self, n, m, P_colptr, P_rowidx, P_val, A_colptr, A_rowidx, batch, devices=None, settings=None, var_order=None,
                 block_size=0):
        L = lib()
        self.n, self.m, self.batch = int(n), int(m), int(batch)
        self.settings = settings if settings is not None else default_settings()
        self._keep = []
        if var_order is not None:
            vo = np.ascontiguousarray(var_order, dtype=np.int64)
            self._keep.append(vo)
            self.settings.var_order = vo.ctypes.data
        if block_size:
            self.settings.block_size = int(block_size)
        idx = [np.ascontiguousarray(a, dtype=np.int64) for a in (P_colptr, P_rowidx, A_colptr, A_rowidx)]
        pv = np.ascontiguousarray(P_val, dtype=np.float64)
        self._keep += idx + [pv]
        dv = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
        self.h = C.c_void_p()
        _check(L.gomp_mg_create(C.byref(self.h), C.byref(self.settings), None if dv is None else C.c_void_p(dv.ctypes.data),
                                C.c_int32(0 if dv is None else len(dv)), C.c_int64(self.batch), C.c_int64(self.n), C.c_int64(self.m),
                                C.c_void_p(idx[0].ctypes.data), C.c_void_p(idx[1].ctypes.data), C.c_void_p(pv.ctypes.data),
                                C.c_void_p(idx[2].ctypes.data), C.c_void_p(idx[3].ctypes.data)))
        self.num_shards = L.gomp_mg_num_shards(self.h)

    def setup(self, A_val, l, u, q=None):
        ptrs = [_ptr(a) for a in (A_val, l, u, q)]
        _check(lib().gomp_mg_setup(self.h, *[p[0] for p in ptrs]))

    def update(self, A_val=None, l=None, u=None, q=None):
        ptrs = [_ptr(a) for a in (A_val, l, u, q)]
        _check(lib().gomp_mg_update(self.h, *[p[0] for p in ptrs]))

    def set_warm_start(self, x0):
        p, keep = _ptr(x0)
        _check(lib().gomp_mg_warm_start_x(self.h, p))

    def solve_async(self):
        _check(lib().gomp_mg_solve(self.h))

    def get(self, want_y=False):
        x = np.empty((self.batch, self.n))
        y = np.empty((self.batch, self.m)) if want_y else None
        code, iters, rup = (np.empty(self.batch, dtype=np.int32) for _ in range(3))
        _check(lib().gomp_mg_get(self.h, C.c_void_p(x.ctypes.data), None if y is None else C.c_void_p(y.ctypes.data),
                                 C.c_void_p(code.ctypes.data), C.c_void_p(iters.ctypes.data), C.c_void_p(rup.ctypes.data)))
        return dict(x=x, y=y, exit_code=code, iters=iters, rho_updates=rup)

    def solve(self, want_y=False):
        self.solve_async()
        return self.get(want_y=want_y)

    def shard(self, k):
        dev, cnt = C.c_int32(), C.c_int64()
        _check(lib().gomp_mg_shard(self.h, C.c_int32(k), None, C.byref(dev), C.byref(cnt)))
        return dev.value, cnt.value

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            lib().gomp_mg_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def measure_peaks():
    """(fp64 TFLOP/s, shared-memory GB/s) measured on the current device."""
    f, s = C.c_double(), C.c_double()
    _check(lib().gomp_measure_peaks(C.byref(f), C.byref(s)))
    return f.value, s.value


def device_count():
    return lib().gomp_device_count()
