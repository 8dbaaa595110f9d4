"""Parity of the automatically chosen back end against the CPU oracle over a list of horizons (development check).
usage: python tools/shape_sweep.py [count] [waypoints ...]      env GOMP_SWEEP_LINSYS=<n> forces a back end"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
S, P = pkg.solver, pkg.problem
count = int(sys.argv[1]) if len(sys.argv) > 1 else 12
shapes = [int(a) for a in sys.argv[2:]] or [8, 16, 24, 32, 40, 48, 56, 64, 72, 80]
linsys = int(os.environ.get("GOMP_SWEEP_LINSYS", "0"))
O = entry.load_oracle()
for w in shapes:
    cfg = P.default_config(waypoints=w)
    pat = P.Pattern(cfg)
    d = P.generate(pat, 300, count)
    s = S.BatchQPSolver(pat.n, pat.m, pat.P_colptr, pat.P_rowidx, pat.P_val, pat.A_colptr, pat.A_rowidx, d["A_val"], d["l"], d["u"],
                        settings=S.default_settings(linsys=linsys), var_order=pat.var_order, block_size=pat.block_size)
    s.set_warm_start(d["x0"])
    got = s.solve(want_y=True)
    info = s.info()
    s.close()
    ref = O.solve_batch(pat.n, pat.m, pat.P_scipy(), (pat.A_colptr, pat.A_rowidx), d["A_val"], d["l"], d["u"], d["x0"], want_y=True)
    ec = int((got["exit_code"] != ref["exit_code"]).sum())
    it = int((got["iters"] != ref["iters"]).sum())
    ok = (ref["exit_code"] == 0) & (got["exit_code"] == 0)
    dx = -1.0
    if ok.any():
        scale = np.abs(ref["x"][ok]).max(axis=1, keepdims=True)
        dx = float((np.abs(got["x"][ok] - ref["x"][ok]) / scale).max())
    print(f"W={w}: back end {info['linsys']} n={pat.n} m={pat.m} exit mismatches {ec} iter mismatches {it} max rel dx {dx:.2e}", flush=True)
