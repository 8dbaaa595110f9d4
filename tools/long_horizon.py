"""BASELINE.json configs[3]: long-horizon GOMP QPs (default N = 7, H = 400) on the streamed back end, timed
beside the CPU oracle on the same inputs.
usage: python tools/long_horizon.py [waypoints] [batch ...]        (prints one JSON line per batch size)
Phase breakdown: build with `make -C osqp-solver_b200 PROFILE=1` and run with GOMP_PROFILE_PHASES=1."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

pkg, O = entry.load_package(), entry.load_oracle()
P, S = pkg.problem, pkg.solver
W = int(sys.argv[1]) if len(sys.argv) > 1 else 400
batches = [int(a) for a in sys.argv[2:]] or [1, 32]
pat = P.Pattern(P.default_config(waypoints=W))
for batch in batches:
    d = P.generate(pat, 0, batch)
    s = S.BatchQPSolver(pat.n, pat.m, pat.P_colptr, pat.P_rowidx, pat.P_val, pat.A_colptr, pat.A_rowidx, d["A_val"], d["l"], d["u"],
                        var_order=pat.var_order, block_size=pat.block_size)
    best, wall = 1e30, 1e30
    for r in range(4):
        t0 = time.perf_counter()
        s.setup(d["A_val"], d["l"], d["u"])
        s.set_warm_start(d["x0"])
        out = s.solve()
        wall = min(wall, time.perf_counter() - t0)
        info = s.info()
        best = min(best, info["last_solve_ms"])
    nthreads = min(batch, O.max_threads())
    t0 = time.perf_counter()
    ref = O.solve_batch(pat.n, pat.m, pat.P_scipy(), (pat.A_colptr, pat.A_rowidx), d["A_val"], d["l"], d["u"], d["x0"], nthreads=nthreads)
    cpu_s = time.perf_counter() - t0
    print(json.dumps({
        "workload": f"N=7 H={W} n={pat.n} m={pat.m} nnzA={pat.nnzA}", "batch": batch, "linsys": info["linsys"],
        "gpu_kernel_ms": round(best, 3), "gpu_e2e_ms_host_buffers": round(wall * 1e3, 3),
        "gpu_qps_kernel": round(batch / best * 1e3, 1), "cpu_oracle_s": round(cpu_s, 3), "cpu_threads": nthreads,
        "cpu_qps": round(batch / cpu_s, 2), "iters_mean": float(out["iters"].mean()),
        "status_equal": bool((out["exit_code"] == ref["exit_code"]).all()), "iters_equal": bool((out["iters"] == ref["iters"]).all()),
        "grid": info["grid"], "smem_bytes": info["smem_bytes"]}))
    ph = info["phase_cycles"]
    if sum(ph) > 0:
        names = ['load', 'scale', 'factor(incl. assemble)', 'rhs', 'solve', 'update', 'info', 'store']
        its, runs = out['iters'].sum() * 4, batch * 4
        print('  per ADMM iteration: ' + ', '.join(f'{n} {ph[k] / its:.0f}' for k, n in enumerate(names) if n in ('rhs', 'solve', 'update')),
              '| per QP: ' + ', '.join(f'{n} {ph[k] / runs:.0f}' for k, n in enumerate(names) if n not in ('rhs', 'solve', 'update')),
              '| assemble per QP %.0f' % (info['assemble_cycles'] / runs), '| chain (warp 0) per iteration %.0f' % (info['chain_cycles'] / its))
    s.close()
