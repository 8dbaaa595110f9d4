"""Long-horizon back end (GOMP_LINSYS_BLOCKTRI_GLOBAL, osqp-solver_b200/csrc/gomp_stream.cu) through the C ABI
against the CPU oracle: BASELINE.json configs[3] (H = 400, batch 1 and small batches) and, forced onto
problems the resident kernel also solves, against the same oracle answers.  Same bar as
test_gpu_parity.py: equal status / iteration count / rho updates, x and y within 1e-6 relative.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from common import oracle, package
from test_gpu_parity import REL_TOL, compare, make_solver, oracle_batch, random_qp

pytestmark = pytest.mark.gpu

GLOBAL = 2  # GOMP_LINSYS_BLOCKTRI_GLOBAL


def test_streamed_backend_on_resident_shape():
    """configs[1] shape forced onto the streamed back end: same answers as the oracle, incl. infeasible
    and multi-check problems."""
    pkg, O = package(), oracle()
    pat = pkg.problem.Pattern(pkg.problem.default_config())
    d = pkg.problem.generate(pat, 0, 320)
    s = make_solver(pkg, pat, d, settings=pkg.solver.default_settings(linsys=GLOBAL))
    assert s.info()["linsys"] == GLOBAL
    s.set_warm_start(d["x0"])
    got = s.solve(want_y=True)
    ref = oracle_batch(O, pat, d)
    compare(got, ref)
    assert (ref["exit_code"] != 0).any() and (ref["iters"] > 25).any() and (ref["rho_updates"] > 0).any()
    s.close()


@pytest.mark.parametrize("batch", [1, 6])
def test_long_horizon_400(batch):
    """configs[3]: N = 7, H = 400 (n = 5600, m = 22372, nnz(A) = 30730); AUTO picks the streamed back end."""
    pkg, O = package(), oracle()
    pat = pkg.problem.Pattern(pkg.problem.default_config(waypoints=400))
    assert (pat.n, pat.m, pat.nnzA) == (5600, 22372, 30730)
    d = pkg.problem.generate(pat, 40, batch)
    s = make_solver(pkg, pat, d)
    info = s.info()
    assert info["linsys"] == GLOBAL and info["num_blocks"] == 400 and info["block_size"] == 14
    s.set_warm_start(d["x0"])
    got = s.solve(want_y=True)
    ref = oracle_batch(O, pat, d)
    compare(got, ref)
    s.close()


def test_long_horizon_tight_tolerance():
    """H = 200, eps 1e-8: many iterations, several rho updates (refactorisations of the streamed factor)."""
    pkg, O = package(), oracle()
    pat = pkg.problem.Pattern(pkg.problem.default_config(waypoints=200))
    d = pkg.problem.generate(pat, 7, 4)
    s = make_solver(pkg, pat, d, settings=pkg.solver.default_settings(eps_abs=1e-8, eps_rel=1e-8))
    assert s.info()["linsys"] == GLOBAL
    s.set_warm_start(d["x0"])
    got = s.solve(want_y=True)
    ref = oracle_batch(O, pat, d, settings=O.default_settings(eps_abs=1e-8, eps_rel=1e-8))
    compare(got, ref)
    assert (ref["iters"] > 50).any()
    s.close()


def test_long_horizon_sqp_updates():
    """configs[2] sequence on a long horizon: solve -> relinearise -> update(A, l, u) -> solve, 3 rounds."""
    pkg, O = package(), oracle()
    pat = pkg.problem.Pattern(pkg.problem.default_config(waypoints=160))
    B = 3
    d = pkg.problem.generate(pat, 500, B)
    s = make_solver(pkg, pat, d)
    assert s.info()["linsys"] == GLOBAL
    s.set_warm_start(d["x0"])
    Ps = pat.P_scipy()
    refs = [O.Solver(Ps, pat.A_scipy(d["A_val"][b]), d["l"][b], d["u"][b]) for b in range(B)]
    for b in range(B):
        refs[b].warm_start_x(d["x0"][b])
    for rnd in range(3):
        got = s.solve(want_y=True)
        for b in range(B):
            refs[b].solve()
        ref = dict(x=np.array([r.x for r in refs]), y=np.array([r.y for r in refs]),
                   exit_code=np.array([r.exit_code for r in refs]), iters=np.array([r.iter for r in refs]),
                   rho_updates=np.array([r.rho_updates for r in refs]))
        compare(got, ref)
        xs = np.where(np.isnan(ref["x"]), d["x0"], ref["x"])
        upd = pkg.problem.relinearize(pat, 500, xs)
        s.update(upd["A_val"], upd["l"], upd["u"])
        for b in range(B):
            refs[b].update(upd["A_val"][b], upd["l"][b], upd["u"][b])
    s.close()


@pytest.mark.parametrize("waypoints", [4, 5, 9])
def test_streamed_short_horizons(waypoints):
    """Both chains one or two stages long."""
    pkg, O = package(), oracle()
    pat = pkg.problem.Pattern(pkg.problem.default_config(waypoints=waypoints))
    d = pkg.problem.generate(pat, 0, 8)
    s = make_solver(pkg, pat, d, settings=pkg.solver.default_settings(linsys=GLOBAL))
    s.set_warm_start(d["x0"])
    got = s.solve(want_y=True)
    compare(got, oracle_batch(O, pat, d))
    s.close()


@pytest.mark.parametrize("blocks", [2, 3])
def test_streamed_two_and_three_blocks(blocks):
    """T = 2 (bottom chain empty) and T = 3 (both chains a single stage) on a hand-made chain QP:
    P = tridiag(-1, 2, -1), A = [first differences; identity], block size 4."""
    pkg, O = package(), oracle()
    rng = np.random.default_rng(5 + blocks)
    bs, B = 4, 5
    n = bs * blocks
    P = sp.csc_matrix(np.triu(2 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1)))
    Dm = (np.eye(n, k=1) - np.eye(n))[: n - 1]
    A = sp.csc_matrix(np.vstack([Dm, np.eye(n)]))
    A.sort_indices()
    m = A.shape[0]
    Ax = np.stack([A.data * (1 + 0.2 * rng.random(A.nnz)) for _ in range(B)])
    l = np.concatenate([-0.3 * np.ones((B, n - 1)), -1 + 0.5 * rng.random((B, n))], axis=1)
    u = np.concatenate([0.3 * np.ones((B, n - 1)), 1 + 0.5 * rng.random((B, n))], axis=1)
    l[:, n - 1] = u[:, n - 1] = 0.4   # x_0 pinned
    q = rng.standard_normal((B, n))
    st = pkg.solver.default_settings(eps_abs=1e-7, eps_rel=1e-7, linsys=GLOBAL)
    s = pkg.solver.BatchQPSolver(n, m, P.indptr, P.indices, P.data, A.indptr, A.indices, Ax, l, u, q=q, settings=st,
                                 var_order=np.arange(n), block_size=bs)
    assert s.info()["num_blocks"] == blocks
    got = s.solve(want_y=True)
    s.close()
    ost = O.default_settings(eps_abs=1e-7, eps_rel=1e-7)
    for b in range(B):
        Ab = sp.csc_matrix((Ax[b], A.indices, A.indptr), shape=(m, n))
        r = O.Solver(P, Ab, l[b], u[b], q=q[b], settings=ost)
        r.solve()
        assert got["exit_code"][b] == r.exit_code and got["iters"][b] == r.iter
        assert np.abs(got["x"][b] - r.x).max() < REL_TOL * max(1.0, np.abs(r.x).max())
        assert np.abs(got["y"][b] - r.y).max() < REL_TOL * max(1.0, np.abs(r.y).max())


def test_streamed_general_qp_with_linear_term():
    """One dense block (T = 1), q != 0, loose / equality / one-sided rows, on the streamed back end."""
    pkg, O = package(), oracle()
    rng = np.random.default_rng(11)
    n, m, B = 12, 18, 6
    P, A = random_qp(rng, n, m)
    A.sort_indices()
    Ax = np.stack([A.data * (1 + 0.1 * rng.standard_normal(A.nnz)) for _ in range(B)])
    l = -rng.random((B, m)) - 0.1
    u = rng.random((B, m)) + 0.1
    l[:, 3] = u[:, 3] = 0.05
    l[:, 5], u[:, 5] = -1e30, 1e30
    u[:, 7] = 1e30
    q = rng.standard_normal((B, n))
    st = pkg.solver.default_settings(eps_abs=1e-7, eps_rel=1e-7, linsys=GLOBAL)
    s = pkg.solver.BatchQPSolver(n, m, P.indptr, P.indices, P.data, A.indptr, A.indices, Ax, l, u, q=q, settings=st,
                                 var_order=np.arange(n), block_size=n)
    got = s.solve(want_y=True)
    s.close()
    ost = O.default_settings(eps_abs=1e-7, eps_rel=1e-7)
    for b in range(B):
        Ab = sp.csc_matrix((Ax[b], A.indices, A.indptr), shape=(m, n))
        r = O.Solver(P, Ab, l[b], u[b], q=q[b], settings=ost)
        r.solve()
        assert got["exit_code"][b] == r.exit_code and got["iters"][b] == r.iter
        assert np.abs(got["x"][b] - r.x).max() < REL_TOL * max(1.0, np.abs(r.x).max())
        assert np.abs(got["y"][b] - r.y).max() < REL_TOL * max(1.0, np.abs(r.y).max())


def test_prefer_throughput_selects_two_ctas_per_sm(monkeypatch):
    """Back-end selection on the headline shape.  AUTO takes the two-slot TMEM kernel (two QPs per SM, gomp_duo.cu) whatever
    the batch size; with that back end disabled (GOMP_NO_DUO), AUTO + prefer_throughput and a batch of >= 2 QPs per SM take
    the streamed kernel's small-QP mode (256 threads, two CTAs per SM) and small batches stay on the resident kernel.
    Each of them is compared with the CPU oracle."""
    pkg, O = package(), oracle()
    pat = pkg.problem.Pattern(pkg.problem.default_config())
    B = 640
    d = pkg.problem.generate(pat, 2000, B)
    ref = oracle_batch(O, pat, d)
    duo = make_solver(pkg, pat, d)
    info = duo.info()
    assert info["linsys"] == 4 and info["threads_per_cta"] == 512 and info["qps_per_sm"] == 2
    duo.set_warm_start(d["x0"])
    compare(duo.solve(want_y=True), ref)
    duo.close()
    monkeypatch.setenv("GOMP_NO_DUO", "1")
    fast = make_solver(pkg, pat, d, settings=pkg.solver.default_settings(prefer_throughput=1))
    info = fast.info()
    assert info["linsys"] == GLOBAL and info["threads_per_cta"] == 256 and info["ctas_per_sm"] == 2
    fast.set_warm_start(d["x0"])
    compare(fast.solve(want_y=True), ref)
    fast.close()
    few = {k: v[:16] for k, v in d.items()}
    s = make_solver(pkg, pat, few, settings=pkg.solver.default_settings(prefer_throughput=1))
    assert s.info()["linsys"] == 1          # a handful of QPs: latency matters, resident kernel
    s.close()
    monkeypatch.delenv("GOMP_NO_DUO")
    s = make_solver(pkg, pat, few)
    assert s.info()["linsys"] == 4
    s.close()


@pytest.mark.parametrize("kw", [dict(waypoints=4), dict(waypoints=8), dict(waypoints=16), dict(waypoints=24), dict(waypoints=40), dict(waypoints=48),
                                dict(n_dim=6, waypoints=30), dict(n_balls=2, n_obstacles=2, waypoints=36)])
def test_two_slot_backend_shapes(kw):
    """The two-slot TMEM kernel forced (linsys = 4) on shapes whose rows / columns per owner thread are below the compiled
    variant's (fewer column slots than the template holds, a single row slot, block size 12): oracle parity."""
    pkg, O = package(), oracle()
    pat = pkg.problem.Pattern(pkg.problem.default_config(**kw))
    d = pkg.problem.generate(pat, 300, 20)
    s = make_solver(pkg, pat, d, settings=pkg.solver.default_settings(linsys=4))
    assert s.info()["linsys"] == 4
    s.set_warm_start(d["x0"])
    got = s.solve(want_y=True)
    s.close()
    compare(got, oracle_batch(O, pat, d))


def test_reference_example_shape_n6_w802():
    """BASELINE.json configs[0]: the shape of the reference's only end-to-end configuration (examples/solver-example.cpp:12-16:
    N = 6 joints, 802 waypoints, two collision balls, no obstacle lines -> n = 9 624, m = 33 660, nnz(A) = 48 060) on the streamed
    back end against the oracle.  The UR5e kinematics of the example are not vendored; the synthetic chain truncated to six
    joints stands in."""
    pkg, O = package(), oracle()
    P = pkg.problem
    pat = P.Pattern(P.default_config(n_dim=6, waypoints=802, n_balls=2, n_obstacles=0))
    assert (pat.n, pat.m, pat.nnzA) == (9624, 33660, 48060)
    d = P.generate(pat, 0, 3)
    s = make_solver(pkg, pat, d)
    assert s.info()["linsys"] == GLOBAL
    s.set_warm_start(d["x0"])
    got = s.solve(want_y=True)
    compare(got, oracle_batch(O, pat, d))
    # one SQP round on top: withObstacles(x) -> update -> solve (reference src/gomp-solver.h:83-87)
    ok = got["exit_code"] == 0
    x = np.where(ok[:, None], got["x"], d["x0"])
    upd = P.relinearize(pat, 0, x)
    s.update(upd["A_val"], upd["l"], upd["u"])
    got2 = s.solve()
    s.close()
    Ps = pat.P_scipy()
    for b in range(3):
        ref = O.Solver(Ps, pat.A_scipy(d["A_val"][b]), d["l"][b], d["u"][b])
        ref.warm_start_x(d["x0"][b])
        ref.solve()
        ref.update(upd["A_val"][b], upd["l"][b], upd["u"][b])
        ref.solve()
        assert got2["exit_code"][b] == ref.exit_code and got2["iters"][b] == ref.iter
        if ref.exit_code == 0:
            assert np.abs(got2["x"][b] - ref.x).max() < REL_TOL * np.abs(ref.x).max()


def test_cluster_mode_matches_single_cta_and_is_repeatable(monkeypatch):
    """Long horizons, small batch: the QP is served by a thread-block cluster (Ruiz cluster kernel + leader / helper CTAs over
    distributed shared memory).  Against the single-CTA path on the same inputs: same status, iterations and rho
    updates, x / y to 1e-9 (only the cost normalisation's sum is ordered differently); repeated solves are bit-identical (no
    race between the cluster's CTAs); an update round works through the cluster path; all of it equals the oracle."""
    pkg, O = package(), oracle()
    P = pkg.problem
    pat = P.Pattern(P.default_config(waypoints=200))
    d = P.generate(pat, 40, 6)

    def run(cluster):
        if cluster is None:
            monkeypatch.delenv("GOMP_RUIZ_CLUSTER", raising=False)
        else:
            monkeypatch.setenv("GOMP_RUIZ_CLUSTER", str(cluster))
        s = make_solver(pkg, pat, d)
        info = s.info()
        assert info["linsys"] == GLOBAL
        outs = []
        for _ in range(3):
            s.setup(d["A_val"], d["l"], d["u"])
            s.set_warm_start(d["x0"])
            outs.append(s.solve(want_y=True))
        ok = outs[0]["exit_code"] == 0
        x = np.where(ok[:, None], outs[0]["x"], d["x0"])
        upd = P.relinearize(pat, 40, x)
        s.update(upd["A_val"], upd["l"], upd["u"])
        second = s.solve(want_y=True)
        s.close()
        return info, outs, second

    info_c, outs_c, second_c = run(None)        # AUTO: 6 QPs leave SMs idle -> clusters of 8
    assert info_c["grid"] == 6 * 8
    info_1, outs_1, second_1 = run(0)           # one CTA per QP, Ruiz inside the ADMM kernel
    assert info_1["grid"] == 6
    for k in ("x", "y", "exit_code", "iters", "rho_updates"):
        np.testing.assert_array_equal(outs_c[0][k], outs_c[1][k])
        np.testing.assert_array_equal(outs_c[0][k], outs_c[2][k])
    for a, b in ((outs_c[0], outs_1[0]), (second_c, second_1)):
        for k in ("exit_code", "iters", "rho_updates"):
            np.testing.assert_array_equal(a[k], b[k])
        ok = a["exit_code"] == 0
        np.testing.assert_allclose(a["x"][ok], b["x"][ok], rtol=0, atol=1e-9 * np.abs(b["x"][ok]).max())
    compare(outs_c[0], oracle_batch(O, pat, d))
    for cs in (2, 4):                           # forced smaller clusters
        _, outs, _ = run(cs)
        for k in ("exit_code", "iters", "rho_updates"):
            np.testing.assert_array_equal(outs[0][k], outs_1[0][k])


def test_cluster_mode_more_problems_than_clusters(monkeypatch):
    """Forced clusters of 8 with more QPs (40) than clusters of that size are co-resident (14-16 on a B200: a cluster lives inside
    one GPC): leaders pull several QPs from the work queue one after the other, the helpers follow through the control words;
    oracle parity."""
    pkg, O = package(), oracle()
    P = pkg.problem
    pat = P.Pattern(P.default_config(waypoints=100))
    d = P.generate(pat, 700, 40)
    monkeypatch.setenv("GOMP_RUIZ_CLUSTER", "8")
    s = make_solver(pkg, pat, d, settings=pkg.solver.default_settings(linsys=GLOBAL))
    grid = s.info()["grid"]
    assert grid % 8 == 0 and 8 <= grid // 8 < 40, grid
    s.set_warm_start(d["x0"])
    got = s.solve(want_y=True)
    s.close()
    compare(got, oracle_batch(O, pat, d))
