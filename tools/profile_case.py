"""Small fixed case for ncu: a few waves of the 7-DoF H=50 workload through the CUDA hot path.
usage: python tools/profile_case.py [batch] [repeats]
Phase breakdown: build with `make -C osqp-solver_b200 PROFILE=1` (compiles the on-device timers in) and run with
GOMP_PROFILE_PHASES=1; the production build prints the kernel time only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

pkg = entry.load_package()
P, S = pkg.problem, pkg.solver
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 296
repeats = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pat = P.Pattern(P.default_config())
d = P.generate(pat, 0, batch)
# GOMP_PROFILE_THROUGHPUT=1: the two-QPs-per-SM mode of the streamed back end (what bench.py measures), else the resident kernel
st = S.default_settings(prefer_throughput=int(os.environ.get("GOMP_PROFILE_THROUGHPUT", "0")))
s = S.BatchQPSolver(pat.n, pat.m, pat.P_colptr, pat.P_rowidx, pat.P_val, pat.A_colptr, pat.A_rowidx, d["A_val"], d["l"], d["u"],
                    settings=st, var_order=pat.var_order, block_size=pat.block_size)
for r in range(repeats):
    s.setup(d["A_val"], d["l"], d["u"])
    s.set_warm_start(d["x0"])
    out = s.solve()
    info = s.info()
    print(f"run {r}: batch {batch} kernel {info['last_solve_ms']:.3f} ms -> {batch / info['last_solve_ms'] * 1e3:.0f} QP/s; "
          f"mean iters {out['iters'].mean():.1f}; codes {out['exit_code'][:8].tolist()}")
ph = info['phase_cycles']
if sum(ph) > 0:
    names = ['load', 'scale', 'factor', 'rhs', 'solve', 'update', 'info', 'store']
    tot = sum(ph)
    its = out['iters'].sum() * repeats
    print('chain cycles per iteration (warp 0): %.0f' % (info['chain_cycles'] / its))
    print('assemble share of factor: %.1f%%' % (100 * info['assemble_cycles'] / max(ph[2], 1)))
    print('phase cycles (sum over QPs and repeats):', {n: f'{v / tot * 100:.1f}%' for n, v in zip(names, ph)})
    print('per ADMM iteration: ' + ', '.join(f'{n} {ph[k] / its:.0f}' for k, n in enumerate(names) if n in ('rhs', 'solve', 'update')),
          '| per QP: ' + ', '.join(f'{n} {ph[k] / (batch * repeats):.0f}' for k, n in enumerate(names) if n in ('load', 'scale', 'factor', 'info', 'store')))
s.close()
