"""The C ABI library loads, exports every symbol include/gomp_cuda.h declares, and refuses to
compute without a CUDA device (no CPU fallback).  No GPU needed."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from common import ROOT, package


def declared_symbols(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gomp_[a-z_0-9]+)\s*\(", text)))


def test_cuda_library_exports_header():
    pkg = package()
    L = pkg.solver.lib()
    syms = declared_symbols("gomp_cuda.h")
    assert set(syms) == set(pkg.solver.EXPORTS)
    for s in syms:
        assert hasattr(L, s), s


def test_host_library_exports_header():
    """gomp_problem.h: the generator lives in libgomp_host, the device-side re-linearisation in libgomp_cuda."""
    pkg = package()
    host, cuda = pkg.problem.lib(), pkg.solver.lib()
    syms = declared_symbols("gomp_problem.h")
    assert any(s.startswith("gomp_device_") for s in syms) and any(s.startswith("gomp_problem_") for s in syms)
    for s in syms:
        if s == "gomp_device_scene":  # the opaque handle type, not a function
            continue
        assert hasattr(cuda if s.startswith("gomp_device_") else host, s), s


def test_unsupported_settings_are_refused_not_ignored():
    """warm_start = 0, threads_per_cta and adaptive_rho_interval = 0 used to be silently ignored."""
    pkg = package()
    P, S = pkg.problem, pkg.solver
    pat = P.Pattern(P.default_config(waypoints=8))
    for kw, code in ((dict(warm_start=0), 4), (dict(threads_per_cta=128), 4), (dict(adaptive_rho_interval=0), 1),
                     (dict(adaptive_rho_interval=30), 1), (dict(adaptive_rho_tolerance=0.5), 1)):
        with pytest.raises(S.GompError) as e:
            S.BatchQPSolver(pat.n, pat.m, pat.P_colptr, pat.P_rowidx, pat.P_val, pat.A_colptr, pat.A_rowidx, batch=2,
                            settings=S.default_settings(**kw), var_order=pat.var_order, block_size=pat.block_size)
        assert e.value.code == code, kw


def test_settings_defaults_match_osqp():
    pkg = package()
    s = pkg.solver.default_settings()
    assert (s.rho, s.sigma, s.alpha, s.scaling, s.adaptive_rho, s.max_iter) == (0.1, 1e-6, 1.6, 10, 1, 4000)
    assert (s.eps_abs, s.eps_rel, s.eps_prim_inf, s.eps_dual_inf) == (1e-3, 1e-3, 1e-4, 1e-4)
    assert (s.check_termination, s.adaptive_rho_interval, s.adaptive_rho_tolerance, s.warm_start) == (25, 25, 5.0, 1)
    assert (s.linsys, s.borrow_inputs, s.prefer_throughput) == (0, 0, 0)


def test_no_cpu_fallback_without_device():
    pkg = package()
    if pkg.solver.device_count() > 0:
        pytest.skip("a CUDA device is present")
    P = pkg.problem
    pat = P.Pattern(P.default_config(waypoints=8))
    d = P.generate(pat, 0, 2)
    with pytest.raises(pkg.solver.GompError) as e:
        pkg.solver.BatchQPSolver(pat.n, pat.m, pat.P_colptr, pat.P_rowidx, pat.P_val, pat.A_colptr, pat.A_rowidx,
                                 d["A_val"], d["l"], d["u"], var_order=pat.var_order, block_size=pat.block_size)
    assert e.value.code == 3  # GOMP_ERR_NO_DEVICE


def test_invalid_arguments_are_rejected():
    pkg = package()
    L = pkg.solver.lib()
    h = C.c_void_p()
    rc = L.gomp_batch_create(C.byref(h), None, C.c_int64(0), C.c_int64(4), C.c_int64(2), None, None, None, None, None, None, None, None, None)
    assert rc == 1 and not h.value
    assert b"" != L.gomp_last_error_string()
