"""Golden vectors from GENUINE OSQP for the solver path -- the harness that pins oracle/osqp_ref.c the moment the `osqp`
Python package (OSQP + QDLDL + AMD, the stack behind the reference's osqp-cpp) is importable.

It is not importable in the build container (no network, no wheel), which is why the oracle's header says "parity unpinned".
Run where it is:
    python tests/golden/make_osqp_fixtures.py            # writes tests/golden/osqp_fixtures.npz
and commit the file; tests/test_osqp_fixtures.py then asserts the oracle against it on every run (and generates the
vectors on the fly, uncommitted, whenever `osqp` can be imported but the file is absent).

Cases (the sequences the reference drives through QPSolver, src/osqp-wrapper.h:16-54 and src/gomp-solver.h:61-88):
  * GOMP QPs from the repo's generator at small horizons: setup -> warm_start_x -> solve,
    then update_A + update_bounds (one SQP re-linearisation) -> solve, twice;
  * an infeasible GOMP QP (status, NaN solution);
  * random block-structured QPs with q != 0 at tight tolerances.
Settings = osqp-cpp defaults with adaptive_rho_interval pinned to 25 (OSQP's 0 derives it from wall-clock time) and
verbose off.  Recorded per solve: status_val, iterations, rho updates, x, y, pri_res, dua_res."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SETTINGS = dict(rho=0.1, sigma=1e-6, alpha=1.6, scaling=10, adaptive_rho=True, adaptive_rho_interval=25, adaptive_rho_tolerance=5.0,
                max_iter=4000, eps_abs=1e-3, eps_rel=1e-3, eps_prim_inf=1e-4, eps_dual_inf=1e-4, scaled_termination=False,
                check_termination=25, warm_start=True, verbose=False, polish=False)


def cases():
    """Deterministic inputs, independent of osqp: list of dicts(name, P, A pattern, steps=[(Ax, l, u, x0 or None)], q, eps)."""
    import scipy.sparse as sp
    from common import package, random_block_qp
    pkg = package()
    P = pkg.problem
    out = []
    for name, kw, first, count in (("gomp_n7_w12", dict(waypoints=12), 3, 6), ("gomp_n6_w20", dict(n_dim=6, waypoints=20), 40, 4),
                                   ("gomp_n7_w50", dict(), 7, 3)):
        pat = P.Pattern(P.default_config(**kw))
        d = P.generate(pat, first, count)
        for b in range(count):
            steps = [(d["A_val"][b], d["l"][b], d["u"][b], d["x0"][b])]
            rng = np.random.default_rng(100 + b)
            x = d["x0"][b]
            for r in range(2):  # two re-linearisations around perturbed trajectories (inputs must not depend on a solver's output)
                x = x + 0.03 * rng.standard_normal(x.shape)
                upd = P.relinearize(pat, first + b, x[None, :], nthreads=1)
                steps.append((upd["A_val"][0], upd["l"][0], upd["u"][0], None))
            out.append(dict(name=f"{name}_{b}", n=pat.n, m=pat.m, P=pat.P_scipy(), Ap=pat.A_colptr, Ai=pat.A_rowidx, steps=steps, q=None, eps=1e-3))
    # infeasible by construction: a position row with crossing equality pins
    pat = P.Pattern(P.default_config(waypoints=12))
    d = P.generate(pat, 11, 1)
    l, u = d["l"][0].copy(), d["u"][0].copy()
    rows = np.flatnonzero((l == u) & (np.abs(l) > 1e-9))[:2]
    A = pat.A_scipy(d["A_val"][0]).tocsr()
    if len(rows):
        l[rows[0]] += 50.0
        u[rows[0]] += 50.0   # start pinned far outside the position box of its neighbours' dynamics
    out.append(dict(name="gomp_infeasible", n=pat.n, m=pat.m, P=pat.P_scipy(), Ap=pat.A_colptr, Ai=pat.A_rowidx,
                    steps=[(d["A_val"][0], l, u, d["x0"][0])], q=None, eps=1e-3))
    rng = np.random.default_rng(5)
    Pm, A, Ax, l, u, q = random_block_qp(rng, 6, 5, 4)
    for b in range(4):
        out.append(dict(name=f"random_block_{b}", n=Pm.shape[0], m=A.shape[0], P=sp.csc_matrix(Pm), Ap=A.indptr.astype(np.int64),
                        Ai=A.indices.astype(np.int64), steps=[(Ax[b], l[b], u[b], None)], q=q[b], eps=1e-7))
    return out


def run_osqp(case):
    """The case through genuine OSQP (0.6.x or 1.x Python API): list of per-solve records."""
    import osqp
    import scipy.sparse as sp
    Ax, l, u, x0 = case["steps"][0]
    A = sp.csc_matrix((Ax, case["Ai"], case["Ap"]), shape=(case["m"], case["n"]))
    q = case["q"] if case["q"] is not None else np.zeros(case["n"])
    st = dict(SETTINGS, eps_abs=case["eps"], eps_rel=case["eps"])
    prob = osqp.OSQP()
    try:
        prob.setup(sp.triu(case["P"]).tocsc(), q, A, l, u, **st)
    except TypeError:  # osqp >= 1.0 renamed polish -> polishing
        st["polishing"] = st.pop("polish")
        prob.setup(sp.triu(case["P"]).tocsc(), q, A, l, u, **st)
    recs = []
    for k, (Ax, l, u, x0) in enumerate(case["steps"]):
        if k > 0:
            prob.update(Ax=Ax)           # reference src/osqp-wrapper.h:36: UpdateConstraintMatrix first ...
            prob.update(l=l, u=u)        # ... then SetBounds (:40)
        if x0 is not None:
            prob.warm_start(x=x0)
        r = prob.solve()
        recs.append(dict(status_val=int(r.info.status_val), iters=int(r.info.iter), rho_updates=int(r.info.rho_updates),
                         x=np.asarray(r.x, dtype=np.float64), y=np.asarray(r.y, dtype=np.float64),
                         pri_res=float(getattr(r.info, "pri_res", getattr(r.info, "prim_res", np.nan))), dua_res=float(getattr(r.info, "dua_res", getattr(r.info, "dual_res", np.nan)))))
    return recs


def generate(path=None):
    flat = {}
    for c in cases():
        for k, rec in enumerate(run_osqp(c)):
            for key, v in rec.items():
                flat[f"{c['name']}/{k}/{key}"] = np.asarray(v)
    import osqp
    flat["__osqp_version__"] = np.asarray(getattr(osqp, "__version__", "unknown"))
    if path:
        np.savez_compressed(path, **flat)
    return flat


if __name__ == "__main__":
    try:
        import osqp  # noqa: F401
    except ImportError:
        sys.exit("the `osqp` package is not importable here: the oracle stays 'parity unpinned' (see oracle/osqp_ref.h)")
    out = os.path.join(HERE, "osqp_fixtures.npz")
    generate(out)
    print("wrote", out)
