"""Host-side timeline of the pipelined end-to-end loop of bench.py (PIPE handles, host buffers): per step the time its submit
returned, the time its collect returned and the device time of its solve (first CTA of the main launch -> last CTA of the batch).
usage: python tools/pipe_trace.py [steps] [host|dev]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

pkg = entry.load_package()
P, S = pkg.problem, pkg.solver
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 36
mode = sys.argv[2] if len(sys.argv) > 2 else "host"
PIPE, B = 12, 4096
LANES = int(os.environ.get("GOMP_BENCH_LANES", "4"))
pat = P.Pattern(P.default_config())
host = {k: torch.empty(shape, dtype=torch.float64).pin_memory() for k, shape in
        dict(A_val=(B, pat.nnzA), l=(B, pat.m), u=(B, pat.m), x0=(B, pat.n), start_goal=(B, 14)).items()}
P.generate(pat, 0, B, out={k: v.numpy() for k, v in host.items()})
src = host if mode == "host" else {k: host[k].cuda() for k in ("A_val", "l", "u", "x0")}
streams = [torch.cuda.Stream() for _ in range(PIPE)]
sol = [S.BatchQPSolver(pat.n, pat.m, pat.P_colptr, pat.P_rowidx, pat.P_val, pat.A_colptr, pat.A_rowidx, batch=B,
                       settings=S.default_settings(borrow_inputs=0 if mode == "host" else 1, stream=streams[k].cuda_stream, prefer_throughput=1),
                       var_order=pat.var_order, block_size=pat.block_size) for k in range(PIPE)]
xh = [torch.empty((B, pat.n), dtype=torch.float64).pin_memory() for _ in range(PIPE)]
code, iters, rup = (np.empty(B, dtype=np.int32) for _ in range(3))


def run(n, log):
    t0 = time.perf_counter()
    sub, col, dev = {}, {}, {}

    def collect(k):
        sol[k % PIPE].get(x=xh[k % PIPE].numpy(), exit_code=code, iters=iters, rho_updates=rup)
        col[k] = (time.perf_counter() - t0) * 1e3
        dev[k] = sol[k % PIPE].info()["last_solve_ms"]

    for k in range(n):
        if k >= PIPE:
            collect(k - PIPE)
        s = sol[k % PIPE]
        lanes = max(1, min(LANES, k + 1, n - k))
        s.set_max_ctas(-(-148 // lanes) if lanes > 1 else 0)
        s.setup(src["A_val"], src["l"], src["u"])
        s.set_warm_start(src["x0"])
        s.solve_async()
        sub[k] = (time.perf_counter() - t0) * 1e3
    for k in range(max(0, n - PIPE), n):
        collect(k)
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t0) * 1e3
    if log:
        for k in range(n):
            print(f"step {k:3d} submitted {sub[k]:8.2f} collected {col[k]:8.2f} (+{col[k] - col.get(k - 1, 0):6.2f}) main launch {dev[k]:7.2f} ms")
        print(f"total {tot:.1f} ms, {n * B / tot * 1e3:.0f} QP/s ({mode} inputs)")


run(PIPE, False)
torch.cuda.synchronize()
run(steps, True)
