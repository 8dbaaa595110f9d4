"""Handle creation (symbolic analysis, uploads, allocations) against solve time for the reference example's shape (N = 6, W = 802), with\nthe cluster mode (default) and without (GOMP_RUIZ_CLUSTER=0).  usage: python tools/create_time.py"""
import sys,time,os; sys.path.insert(0,"."); import __graft_entry__ as e, numpy as np
pkg=e.load_package(); P,S=pkg.problem,pkg.solver
pat=P.Pattern(P.default_config(n_dim=6, waypoints=802, n_balls=2, n_obstacles=0)); d=P.generate(pat,0,1)
def mk():
    return S.BatchQPSolver(pat.n,pat.m,pat.P_colptr,pat.P_rowidx,pat.P_val,pat.A_colptr,pat.A_rowidx,d["A_val"],d["l"],d["u"],var_order=pat.var_order,block_size=pat.block_size)
s=mk(); s.set_warm_start(d["x0"]); s.solve(); s.close()
for rep in range(4):
    t0=time.perf_counter(); s=mk(); t1=time.perf_counter(); s.set_warm_start(d["x0"]); s.solve(); t2=time.perf_counter(); s.solve(); t3=time.perf_counter(); s.close(); t4=time.perf_counter()
    print(os.environ.get("GOMP_RUIZ_CLUSTER","auto"), "create %.1f ms, first solve %.1f ms, second solve %.1f ms, destroy %.1f ms" % ((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3,(t4-t3)*1e3))
