"""Throughput of the streamed back end (GOMP_LINSYS_BLOCKTRI_GLOBAL) on the resident kernel's shape (N=7, H=50).
usage: python tools/stream_on_resident_shape.py [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

pkg = entry.load_package()
P, S = pkg.problem, pkg.solver
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
pat = P.Pattern(P.default_config())
d = P.generate(pat, 0, batch)
for linsys in (1, 2):
    s = S.BatchQPSolver(pat.n, pat.m, pat.P_colptr, pat.P_rowidx, pat.P_val, pat.A_colptr, pat.A_rowidx, d["A_val"], d["l"], d["u"],
                        settings=S.default_settings(linsys=linsys), var_order=pat.var_order, block_size=pat.block_size)
    for r in range(2):
        s.setup(d["A_val"], d["l"], d["u"])
        s.set_warm_start(d["x0"])
        out = s.solve()
        info = s.info()
    print(f"linsys {linsys}: batch {batch} kernel {info['last_solve_ms']:.3f} ms -> {batch / info['last_solve_ms'] * 1e3:.0f} QP/s; "
          f"ctas/SM {info['ctas_per_sm']} threads {info['threads_per_cta']} smem {info['smem_bytes']} mean iters {out['iters'].mean():.1f}")
    ph = info['phase_cycles']
    if sum(ph) > 0:
        its = out['iters'].sum() * 2
        names = ['load', 'scale', 'factor', 'rhs', 'solve', 'update', 'info', 'store']
        print('   per iteration: ' + ', '.join(f'{n} {ph[k] / its:.0f}' for k, n in enumerate(names) if n in ('rhs', 'solve', 'update')),
              '| per QP: ' + ', '.join(f'{n} {ph[k] / (batch * 2):.0f}' for k, n in enumerate(names) if n not in ('rhs', 'solve', 'update')),
              '| chain (warp 0) per iteration %.0f' % (info['chain_cycles'] / its))
    s.close()
