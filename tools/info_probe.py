import os, sys
sys.path.insert(0, '/root/repo')
import __graft_entry__ as entry
pkg = entry.load_package()
P, S = pkg.problem, pkg.solver
pat = P.Pattern(P.default_config())
d = P.generate(pat, 0, 148)
for ct in (25, 5, 1):
    st = S.default_settings(check_termination=ct, adaptive_rho=0)
    s = S.BatchQPSolver(pat.n, pat.m, pat.P_colptr, pat.P_rowidx, pat.P_val, pat.A_colptr, pat.A_rowidx, d["A_val"], d["l"], d["u"],
                        settings=st, var_order=pat.var_order, block_size=pat.block_size)
    s.set_warm_start(d["x0"])
    out = s.solve()
    info = s.info()
    ph = info['phase_cycles']
    its = out['iters'].sum()
    checks = (out['iters'] // ct).sum()
    print(f"check_termination={ct}: iters {its}, checks {checks}, info cycles per check {ph[6]/checks:.0f}, per-iter rhs {ph[3]/its:.0f} solve {ph[4]/its:.0f} update {ph[5]/its:.0f}")
    s.close()
