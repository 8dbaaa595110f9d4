"""Throughput of the two back ends on the headline shape (N=7, H=50) for a tail-free slice of the workload:
problems whose solve reaches max_iter are located first (resident kernel) and the longest run of ids without one is used.
usage: python tools/backend_compare.py [scan] [repeats]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

pkg = entry.load_package()
P, S = pkg.problem, pkg.solver
scan = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
repeats = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pat = P.Pattern(P.default_config())


def run(d, linsys, n):
    s = S.BatchQPSolver(pat.n, pat.m, pat.P_colptr, pat.P_rowidx, pat.P_val, pat.A_colptr, pat.A_rowidx, d["A_val"], d["l"], d["u"],
                        settings=S.default_settings(linsys=linsys), var_order=pat.var_order, block_size=pat.block_size)
    best = 1e30
    for r in range(n):
        s.setup(d["A_val"], d["l"], d["u"])
        s.set_warm_start(d["x0"])
        out = s.solve()
        info = s.info()
        best = min(best, info["last_solve_ms"])
    s.close()
    return best, out, info


d = P.generate(pat, 0, scan)
_, out, _ = run(d, 1, 1)
long_ids = np.nonzero(out["iters"] >= 1000)[0]
edges = np.concatenate([[-1], long_ids, [scan]])
k = int(np.argmax(np.diff(edges)))
first, count = int(edges[k] + 1), int(edges[k + 1] - edges[k] - 1)
count = (count // 296) * 296 or count
print(f"ids with >= 1000 iterations among 0..{scan - 1}: {long_ids.tolist()}; using ids {first}..{first + count - 1} ({count} QPs)")
d = P.generate(pat, first, count)
ref = None
for linsys in (1, 2):
    ms, out, info = run(d, linsys, repeats)
    print(f"linsys {linsys}: {count} QPs in {ms:.3f} ms -> {count / ms * 1e3:.0f} QP/s | ctas/SM {info['ctas_per_sm']} x {info['threads_per_cta']} threads, "
          f"smem {info['smem_bytes']} B, mean iters {out['iters'].mean():.1f}, max {out['iters'].max()}")
    if ref is None:
        ref = out
    else:
        same = (out["iters"] == ref["iters"]).all() and (out["exit_code"] == ref["exit_code"]).all()
        ok = ref["exit_code"] == 0
        err = np.abs(out["x"][ok] - ref["x"][ok]).max() / np.abs(ref["x"][ok]).max()
        print(f"   same status / iterations as the resident kernel: {same}; max rel dx {err:.2e}")
