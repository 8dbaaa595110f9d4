"""Development check of the two-slot TMEM back end (linsys 4): parity against the CPU oracle and the resident kernel on the
headline shape, then kernel time for one 4096-QP batch per back end.  python tools/duo_check.py [count] [big]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
S = pkg.solver
count = int(sys.argv[1]) if len(sys.argv) > 1 else 320
big = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
cfg = pkg.problem.default_config()
pat = pkg.problem.Pattern(cfg)


def run(d, linsys, n=1, **kw):
    s = S.BatchQPSolver(pat.n, pat.m, pat.P_colptr, pat.P_rowidx, pat.P_val, pat.A_colptr, pat.A_rowidx, d["A_val"], d["l"], d["u"],
                        settings=S.default_settings(linsys=linsys, **kw), var_order=pat.var_order, block_size=pat.block_size)
    out = None
    ms = []
    for _ in range(n):
        s.setup(d["A_val"], d["l"], d["u"])
        s.set_warm_start(d["x0"])
        out = s.solve(want_y=True)
        ms.append(s.info()["last_solve_ms"])
    info = s.info()
    s.close()
    return out, info, ms


d = pkg.problem.generate(pat, 0, count)
t0 = time.time()
got, info, ms = run(d, 4)
print(f"duo: linsys {info['linsys']} threads {info['threads_per_cta']} smem {info['smem_bytes']} grid {info['grid']} "
      f"kernel {ms[-1]:.3f} ms wall {time.time() - t0:.1f}s", flush=True)
O = entry.load_oracle()
ref = O.solve_batch(pat.n, pat.m, pat.P_scipy(), (pat.A_colptr, pat.A_rowidx), d["A_val"], d["l"], d["u"], d["x0"], want_y=True)
res, _, _ = run(d, 1)
for name, other in (("oracle", ref), ("resident", res)):
    ec = (got["exit_code"] != other["exit_code"]).sum()
    it = (got["iters"] != other["iters"]).sum()
    ru = (got["rho_updates"] != other["rho_updates"]).sum()
    ok = (other["exit_code"] == 0) & (got["exit_code"] == 0)
    scale = np.abs(other["x"][ok]).max(axis=1, keepdims=True)
    dx = (np.abs(got["x"][ok] - other["x"][ok]) / scale).max() if ok.any() else -1
    ys = np.maximum(np.abs(other["y"][ok]).max(axis=1, keepdims=True), 1e-3)
    dy = (np.abs(got["y"][ok] - other["y"][ok]) / ys).max() if ok.any() else -1
    print(f"vs {name}: exit mismatches {ec}, iter mismatches {it}, rho_update mismatches {ru}, max rel dx {dx:.3e}, dy {dy:.3e}", flush=True)
    if ec or it:
        bad = np.nonzero((got["exit_code"] != other["exit_code"]) | (got["iters"] != other["iters"]))[0][:10]
        print("  first bad:", [(int(b), int(got["exit_code"][b]), int(other["exit_code"][b]), int(got["iters"][b]), int(other["iters"][b])) for b in bad])
print("iters histogram (duo):", np.unique(got["iters"], return_counts=True))

if big:
    d = pkg.problem.generate(pat, 0, big)
    for linsys, kw in ((4, {}), (1, {}), (0, {"prefer_throughput": 1})):
        if linsys == 0:
            os.environ["GOMP_NO_DUO"] = "1"
        out, info, ms = run(d, linsys, n=3, **kw)
        print(f"linsys {linsys} ({info['linsys']}): {big} QPs kernel ms {['%.2f' % m for m in ms]} -> {big / min(ms) * 1e3:.0f} QP/s solo; "
              f"mean iters {out['iters'].mean():.1f}", flush=True)
