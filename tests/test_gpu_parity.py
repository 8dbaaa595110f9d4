"""CUDA hot path (through the C ABI, include/gomp_cuda.h) against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): same solver status, iteration counts within +-10% (here: equal),
primal/dual solutions within 1e-6 relative -- demonstrated where both solvers converge to the
optimum (tight eps) and, at the default eps = 1e-3, as agreement of the iterates themselves because
the two implementations perform the same ADMM steps (different linear-system factorisations).
The oracle is a restatement of OSQP, not genuine OSQP: parity unpinned (oracle/osqp_ref.h).
"""
import numpy as np
import pytest
import scipy.sparse as sp

from common import oracle, package

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6  # north_star: primal/dual within 1e-6 relative


_BACKEND = [0]  # linear-system back end the `env` tests run on (set by the fixture): 0 = AUTO, 1 = resident kernel


def make_solver(pkg, pat, d, settings=None, **kw):
    if settings is None:
        settings = pkg.solver.default_settings()
    if settings.linsys == 0:
        settings.linsys = _BACKEND[0]
    return pkg.solver.BatchQPSolver(pat.n, pat.m, pat.P_colptr, pat.P_rowidx, pat.P_val, pat.A_colptr, pat.A_rowidx,
                                    d["A_val"], d["l"], d["u"], settings=settings, var_order=pat.var_order,
                                    block_size=pat.block_size, **kw)


def compare(got, ref, want_y=True):
    assert (got["exit_code"] == ref["exit_code"]).all(), (got["exit_code"], ref["exit_code"])
    assert (got["iters"] == ref["iters"]).all(), (got["iters"], ref["iters"])
    assert (got["rho_updates"] == ref["rho_updates"]).all()
    ok = ref["exit_code"] == 0
    bad = ~ok & np.isin(ref["exit_code"], [1, 2, 4, 5])  # infeasible: NaN solution on both sides
    assert np.isnan(got["x"][bad]).all() and np.isnan(ref["x"][bad]).all()
    if ok.any():
        scale = np.abs(ref["x"][ok]).max(axis=1, keepdims=True)
        assert (np.abs(got["x"][ok] - ref["x"][ok]) / scale).max() < REL_TOL
        if want_y:
            ys = np.maximum(np.abs(ref["y"][ok]).max(axis=1, keepdims=True), 1e-3)
            assert (np.abs(got["y"][ok] - ref["y"][ok]) / ys).max() < REL_TOL


@pytest.fixture(scope="module", params=[0, 1], ids=["auto-two-slot", "resident"])
def env(request):
    """Every test on the headline shape runs twice: on the back end AUTO picks for it (the two-slot TMEM kernel,
    gomp_duo.cu -- what bench.py measures) and on the register-resident kernel (gomp_kernel.cuh)."""
    pkg, O = package(), oracle()
    cfg = pkg.problem.default_config()
    pat = pkg.problem.Pattern(cfg)
    _BACKEND[0] = request.param
    yield pkg, O, pat
    _BACKEND[0] = 0


def oracle_batch(O, pat, d, settings=None):
    return O.solve_batch(pat.n, pat.m, pat.P_scipy(), (pat.A_colptr, pat.A_rowidx), d["A_val"], d["l"], d["u"], d["x0"],
                         settings=settings, want_y=True)


def test_gomp_batch_default_settings(env):
    """BASELINE configs[1] shape (7-DoF, H=50), 320 problems incl. multi-check and infeasible ones."""
    pkg, O, pat = env
    d = pkg.problem.generate(pat, 0, 320)
    s = make_solver(pkg, pat, d)
    s.set_warm_start(d["x0"])
    got = s.solve(want_y=True)
    ref = oracle_batch(O, pat, d)
    compare(got, ref)
    assert (ref["exit_code"] != 0).any() and (ref["iters"] > 25).any() and (ref["rho_updates"] > 0).any()  # paths exercised
    info = s.info()
    assert info["m_used"] == 1558 and info["block_size"] == 14 and info["num_blocks"] == 50
    assert info["linsys"] == (4 if _BACKEND[0] == 0 else 1)   # AUTO = two QPs per SM with the factor in tensor memory
    s.close()


def test_gomp_tight_tolerance_solution_parity(env):
    """eps = 1e-8: both converge to the optimum, x and y agree to 1e-6 relative."""
    pkg, O, pat = env
    d = pkg.problem.generate(pat, 100, 24)
    st = pkg.solver.default_settings(eps_abs=1e-8, eps_rel=1e-8)
    s = make_solver(pkg, pat, d, settings=st)
    s.set_warm_start(d["x0"])
    got = s.solve(want_y=True)
    ref = oracle_batch(O, pat, d, settings=O.default_settings(eps_abs=1e-8, eps_rel=1e-8))
    compare(got, ref)
    s.close()


def test_sqp_update_sequence(env):
    """BASELINE configs[2]: solve -> withObstacles(x) -> update(A, l, u) -> solve, 5 rounds
    (reference src/gomp-solver.h:70-88); iterates and rho persist on the device."""
    pkg, O, pat = env
    B = 12
    d = pkg.problem.generate(pat, 200, B)
    s = make_solver(pkg, pat, d)
    s.set_warm_start(d["x0"])
    Ps = pat.P_scipy()
    refs = [O.Solver(Ps, pat.A_scipy(d["A_val"][b]), d["l"][b], d["u"][b]) for b in range(B)]
    for b in range(B):
        refs[b].warm_start_x(d["x0"][b])
    for rnd in range(5):
        got = s.solve(want_y=True)
        for b in range(B):
            refs[b].solve()
        ref = dict(x=np.array([r.x for r in refs]), y=np.array([r.y for r in refs]),
                   exit_code=np.array([r.exit_code for r in refs]), iters=np.array([r.iter for r in refs]),
                   rho_updates=np.array([r.rho_updates for r in refs]))
        compare(got, ref)
        xs = np.where(np.isnan(ref["x"]), d["x0"], ref["x"])
        upd = pkg.problem.relinearize(pat, 200, xs)
        s.update(upd["A_val"], upd["l"], upd["u"])
        for b in range(B):
            refs[b].update(upd["A_val"][b], upd["l"][b], upd["u"][b])
    s.close()


def test_setup_resets_state(env):
    """gomp_batch_setup == constructing a fresh QPSolver: same answer as the first cold solve."""
    pkg, O, pat = env
    d = pkg.problem.generate(pat, 300, 8)
    s = make_solver(pkg, pat, d)
    s.set_warm_start(d["x0"])
    first = s.solve()
    s.solve()  # moves iterates / rho
    s.setup(d["A_val"], d["l"], d["u"])
    s.set_warm_start(d["x0"])
    again = s.solve()
    np.testing.assert_array_equal(first["x"], again["x"])
    np.testing.assert_array_equal(first["iters"], again["iters"])
    s.close()


def test_batch_equals_single(env):
    """Determinism: a QP solved inside a batch gives bit-identical results to the same QP alone."""
    pkg, O, pat = env
    d = pkg.problem.generate(pat, 400, 40)
    s = make_solver(pkg, pat, d)
    s.set_warm_start(d["x0"])
    full = s.solve(want_y=True)
    s.close()
    for b in (0, 17, 39):
        one = {k: v[b:b + 1] for k, v in d.items()}
        s1 = make_solver(pkg, pat, one)
        s1.set_warm_start(one["x0"])
        r = s1.solve(want_y=True)
        np.testing.assert_array_equal(r["x"][0], full["x"][b])
        np.testing.assert_array_equal(r["y"][0], full["y"][b])
        assert r["iters"][0] == full["iters"][b]
        s1.close()


def test_device_resident_inputs(env):
    """Device pointers (torch CUDA tensors), borrowed in place, give the same answer as host inputs."""
    import torch
    pkg, O, pat = env
    d = pkg.problem.generate(pat, 500, 16)
    s = make_solver(pkg, pat, d)
    s.set_warm_start(d["x0"])
    host = s.solve()
    s.close()
    dd = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    st = pkg.solver.default_settings(borrow_inputs=1)
    s2 = make_solver(pkg, pat, dd, settings=st, batch=16)
    s2.set_warm_start(dd["x0"])
    s2.solve_async()
    xdev = torch.empty((16, pat.n), dtype=torch.float64, device="cuda")
    r = s2.get(x=xdev)
    np.testing.assert_array_equal(host["x"], xdev.cpu().numpy())
    np.testing.assert_array_equal(host["iters"], r["iters"])
    s2.close()


@pytest.mark.parametrize("kw", [dict(n_dim=6), dict(n_balls=2, n_obstacles=2), dict(waypoints=12), dict(n_dim=3, waypoints=20)])
def test_other_shapes(kw):
    """6-DoF (block 12, generic kernel), the heavy 2-ball/2-line variant, short horizons."""
    pkg, O = package(), oracle()
    cfg = pkg.problem.default_config(**kw)
    pat = pkg.problem.Pattern(cfg)
    d = pkg.problem.generate(pat, 0, 24)
    s = make_solver(pkg, pat, d)
    s.set_warm_start(d["x0"])
    got = s.solve(want_y=True)
    ref = oracle_batch(O, pat, d)
    compare(got, ref)
    s.close()


def random_qp(rng, n, m, density=0.4):
    M = rng.standard_normal((n, n))
    P = sp.csc_matrix(np.triu(M @ M.T + 0.1 * np.eye(n)))
    A = rng.standard_normal((m, n)) * (rng.random((m, n)) < density)
    A[0, :] = rng.standard_normal(n)  # at least one dense row keeps every column coupled
    return P, sp.csc_matrix(A)


def test_general_qp_with_linear_term():
    """Generic path: no ordering hint (RCM + detected block size), q != 0, batch of random QPs with
    a shared pattern; includes loose, equality and inequality rows."""
    pkg, O = package(), oracle()
    rng = np.random.default_rng(7)
    n, m, B = 12, 18, 10
    P, A = random_qp(rng, n, m)
    A.sort_indices()
    nnz = A.nnz
    Ax = np.stack([A.data * (1 + 0.1 * rng.standard_normal(nnz)) for _ in range(B)])
    l = -rng.random((B, m)) - 0.1
    u = rng.random((B, m)) + 0.1
    l[:, 3] = u[:, 3] = 0.05      # equality row
    l[:, 5], u[:, 5] = -1e30, 1e30  # loose row
    u[:, 7] = 1e30                  # one-sided
    q = rng.standard_normal((B, n))
    st = pkg.solver.default_settings(eps_abs=1e-7, eps_rel=1e-7)
    s = pkg.solver.BatchQPSolver(n, m, P.indptr, P.indices, P.data, A.indptr, A.indices, Ax, l, u, q=q, settings=st)
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


def test_infeasible_and_unbounded():
    """Known-answer status tests: primal infeasible (x in [1,2] and x in [-5,0]), dual infeasible
    (min -x s.t. x >= 0), empty row whose bounds exclude 0."""
    pkg, O = package(), oracle()
    S = pkg.solver
    P = sp.csc_matrix(np.eye(2))
    A = sp.csc_matrix(np.array([[1.0, 0.0], [1.0, 0.0], [0.0, 1.0]]))
    s = S.BatchQPSolver(2, 3, P.indptr, P.indices, P.data, A.indptr, A.indices, A.data[None, :],
                        np.array([[1.0, -5.0, -1.0]]), np.array([[2.0, 0.0, 1.0]]))
    r = s.solve()
    ref = O.Solver(P, A, np.array([1.0, -5.0, -1.0]), np.array([2.0, 0.0, 1.0]))
    ref.solve()
    assert r["exit_code"][0] == ref.exit_code == 1 and r["iters"][0] == ref.iter
    assert np.isnan(r["x"]).all()
    # the certificate osqp-cpp would hand out (primal_infeasibility_certificate): unit infinity norm, A'v = 0, support < 0
    pc, dc = s.certificates(primal=True, dual=True)
    np.testing.assert_allclose(pc[0], ref.prim_inf_cert, rtol=1e-9, atol=1e-12)
    assert np.isnan(dc).all() and abs(np.abs(pc[0]).max() - 1.0) < 1e-12
    s.close()

    P0 = sp.csc_matrix((1, 1))
    A1 = sp.csc_matrix(np.array([[1.0]]))
    s = S.BatchQPSolver(1, 1, P0.indptr, P0.indices, P0.data, A1.indptr, A1.indices, A1.data[None, :],
                        np.array([[0.0]]), np.array([[1e30]]), q=np.array([[-1.0]]))
    r = s.solve()
    ref = O.Solver(P0, A1, np.array([0.0]), np.array([1e30]), q=np.array([-1.0]))
    ref.solve()
    assert r["exit_code"][0] == ref.exit_code == 2 and r["iters"][0] == ref.iter
    pc, dc = s.certificates(primal=True, dual=True)
    np.testing.assert_allclose(dc[0], ref.dual_inf_cert, rtol=1e-9, atol=1e-12)
    assert np.isnan(pc).all()
    s.close()

    A2 = sp.csc_matrix(np.array([[1.0, 0.0], [0.0, 0.0]]))  # second row empty, bounds [1, 2] exclude 0
    s = S.BatchQPSolver(2, 2, P.indptr, P.indices, P.data, A2.indptr, A2.indices, A2.data[None, :],
                        np.array([[-1.0, 1.0]]), np.array([[1.0, 2.0]]))
    r = s.solve()
    assert r["exit_code"][0] == 1
    pc, _ = s.certificates()
    np.testing.assert_array_equal(pc[0], [0.0, -1.0])   # the empty row's unit vector, signed towards the violated bound
    s.close()


def test_primal_infeasibility_certificates_on_gomp_batch(env):
    """the certificates through the API on the headline shape.  The batch holds primal infeasible QPs
    (~2 % of the workload): their certificate equals the oracle's delta_y and certifies infeasibility of the QP the caller
    passed (A'v ~ 0, u'v+ + l'v- < 0); every other problem reports NaN."""
    pkg, O, pat = env
    B = 160
    d = pkg.problem.generate(pat, 0, B)
    s = make_solver(pkg, pat, d)
    s.set_warm_start(d["x0"])
    got = s.solve()
    pc, _ = s.certificates()
    s.close()
    inf = np.isin(got["exit_code"], [1, 4])
    assert inf.any() and np.isnan(pc[~inf]).all() and not np.isnan(pc[inf]).any()
    Ps = pat.P_scipy()
    for b in np.flatnonzero(inf):
        A = pat.A_scipy(d["A_val"][b])
        ref = O.Solver(Ps, A, d["l"][b], d["u"][b])
        ref.warm_start_x(d["x0"][b])
        ref.solve()
        assert ref.exit_code == got["exit_code"][b]
        v = pc[b]
        np.testing.assert_allclose(v, ref.prim_inf_cert, rtol=1e-6, atol=1e-9)
        assert abs(np.abs(v).max() - 1.0) < 1e-12
        lo, hi = np.maximum(d["l"][b], -1e30), np.minimum(d["u"][b], 1e30)
        support = hi @ np.maximum(v, 0.0) + lo @ np.minimum(v, 0.0)
        assert support < 0 and np.abs(A.T @ v).max() < 1e-3 * abs(support) + 1e-3


def test_update_rejects_bad_bounds(env):
    """QPSolver::update throws std::invalid_argument on a failed SetBounds (reference
    src/osqp-wrapper.h:40-42); the C ABI returns GOMP_ERR_INVALID_ARGUMENT."""
    pkg, O, pat = env
    d = pkg.problem.generate(pat, 0, 2)
    s = make_solver(pkg, pat, d)
    bad_l = d["u"] + 1.0
    with pytest.raises(pkg.solver.GompError) as e:
        s.update(None, bad_l, d["u"])
    assert e.value.code == 1
    s.close()
    # device-resident bounds (they used to go unchecked): validated on the device, reported by the next get / sync
    import torch
    dev = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    s = make_solver(pkg, pat, dev, settings=pkg.solver.default_settings(borrow_inputs=1), batch=2)
    s.set_warm_start(dev["x0"])
    good = s.solve()
    assert (good["exit_code"] >= 0).all()
    bad_dev = dev["l"].clone()
    bad_dev[1, 7] = dev["u"][1, 7] + 1.0
    s.update(None, bad_dev, dev["u"])       # asynchronous: nothing to report yet
    s.solve_async()
    with pytest.raises(pkg.solver.GompError) as e:
        s.get()
    assert e.value.code == 1
    s.update(None, dev["l"], dev["u"])      # valid data again: the flag clears
    s.solve()
    s.close()


def test_full_size_batch_properties(env):
    """BASELINE configs[1] at full size (4096 QPs): size-independent properties instead of the
    oracle -- every optimal solution satisfies its own stopping criterion when re-evaluated on the
    host in fp64, iteration counts are multiples of check_termination, and a repeated solve of the
    same batch is bit-identical."""
    pkg, O, pat = env
    B = 4096
    d = pkg.problem.generate(pat, 0, B)
    s = make_solver(pkg, pat, d)
    s.set_warm_start(d["x0"])
    r = s.solve(want_y=True)
    assert (r["iters"] % 25 == 0).all() and (r["iters"] >= 25).all()
    ok = r["exit_code"] == 0
    assert ok.mean() > 0.85
    P = pat.P_scipy()
    idx = np.flatnonzero(ok)[:: max(1, ok.sum() // 200)]  # ~200 spread samples
    for b in idx:
        A = pat.A_scipy(d["A_val"][b])
        x, y = r["x"][b], r["y"][b]
        Ax = A @ x
        zproj = np.clip(Ax, d["l"][b], d["u"][b])
        pri = np.abs(Ax - zproj).max()
        dua = np.abs(P @ x + A.T @ y).max()
        # OSQP stops on ||Ax - z||, z the ADMM iterate; the projection distance is bounded by it
        assert pri <= 1e-3 + 1e-3 * np.abs(Ax).max() + 1e-9
        assert dua <= 1e-3 + 1e-3 * max(np.abs(P @ x).max(), np.abs(A.T @ y).max()) + 1e-9
    s.setup(d["A_val"], d["l"], d["u"])
    s.set_warm_start(d["x0"])
    r2 = s.solve(want_y=True)
    np.testing.assert_array_equal(r["x"], r2["x"])
    np.testing.assert_array_equal(r["iters"], r2["iters"])
    s.close()


@pytest.mark.parametrize("linsys", [0, 2])
@pytest.mark.parametrize("bs,T", [(3, 5), (6, 9), (10, 12), (16, 7), (32, 3), (5, 1), (4, 2), (14, 20), (2, 40)])
def test_random_block_tridiagonal_qps(bs, T, linsys):
    """Generic patterns on both back ends (resident kernel variants 2_1_0 / 4_2_0 / 8_4_0 with run-time block size;
    streamed kernel for even block sizes): random block sizes and chain lengths, q != 0, equality / loose /
    one-sided / empty rows.  Same bar as the GOMP shapes: equal status, iterations, rho updates; x, y to 1e-6."""
    from common import random_block_qp
    if linsys == 2 and bs % 2:
        pytest.skip("streamed back end needs an even block size (16-byte bulk copies)")
    pkg, O = package(), oracle()
    rng = np.random.default_rng(100 * bs + T)
    B = 6
    P, A, Ax, l, u, q = random_block_qp(rng, bs, T, B)
    n, m = P.shape[0], A.shape[0]
    st = pkg.solver.default_settings(eps_abs=1e-6, eps_rel=1e-6, linsys=linsys)
    s = pkg.solver.BatchQPSolver(n, m, P.indptr, P.indices, P.data, A.indptr, A.indices, Ax, l, u, q=q, settings=st,
                                 var_order=np.arange(n), block_size=bs)
    info = s.info()
    assert (info["block_size"], info["num_blocks"]) == (bs, T) and info["linsys"] == (2 if linsys == 2 else 1)
    got = s.solve(want_y=True)
    s.close()
    ost = O.default_settings(eps_abs=1e-6, eps_rel=1e-6)
    for b in range(B):
        Ab = sp.csc_matrix((Ax[b], A.indices, A.indptr), shape=(m, n))
        r = O.Solver(P, Ab, l[b], u[b], q=q[b], settings=ost)
        r.solve()
        assert got["exit_code"][b] == r.exit_code and got["iters"][b] == r.iter and got["rho_updates"][b] == r.rho_updates
        if r.exit_code == 0:
            assert np.abs(got["x"][b] - r.x).max() < REL_TOL * max(1.0, np.abs(r.x).max())
            assert np.abs(got["y"][b] - r.y).max() < REL_TOL * max(1.0, np.abs(r.y).max())


def test_active_mask_freezes_problems(env):
    """gomp_batch_set_active: masked problems keep iterates, rho and results; the others solve as usual
    (also exercises the longest-first queue order of the second solve)."""
    pkg, O, pat = env
    B = 24
    d = pkg.problem.generate(pat, 700, B)
    s = make_solver(pkg, pat, d)
    s.set_warm_start(d["x0"])
    first = s.solve(want_y=True)
    ref = make_solver(pkg, pat, d)          # unmasked twin
    ref.set_warm_start(d["x0"])
    ref.solve()
    upd = pkg.problem.relinearize(pat, 700, np.where(np.isnan(first["x"]), d["x0"], first["x"]))
    mask = (np.arange(B) % 3 != 0).astype(np.int32)
    s.set_active(mask)
    s.update(upd["A_val"], upd["l"], upd["u"])
    second = s.solve(want_y=True)
    ref.update(upd["A_val"], upd["l"], upd["u"])
    want = ref.solve(want_y=True)
    off, on = mask == 0, mask == 1
    for k in ("x", "y"):
        np.testing.assert_array_equal(second[k][off], first[k][off])       # frozen (NaN-aware)
        np.testing.assert_array_equal(second[k][on], want[k][on])           # same as without a mask
    np.testing.assert_array_equal(second["iters"][off], first["iters"][off])
    np.testing.assert_array_equal(second["iters"][on], want["iters"][on])
    np.testing.assert_array_equal(second["exit_code"][on], want["exit_code"][on])
    s.set_active(None)
    third = s.solve()
    assert (third["iters"][off] > 0).all()
    s.close()
    ref.close()


def test_max_ctas_and_exit_code_paths(env):
    """gomp_batch_set_max_ctas: a batch served by fewer CTAs (how several batches in flight share the SMs, bench.py's
    lanes) gives bit-identical results, cold and in the warm re-solve that orders its queue by the previous iteration
    counts; the exit codes mapped on the host (gomp_batch_get with a host array) equal the ones mapped on the device
    (gomp_batch_get_async into a device array); a negative cap is refused."""
    import torch
    pkg, O, pat = env
    B = 640
    d = pkg.problem.generate(pat, 2100, B)
    st = pkg.solver.default_settings(prefer_throughput=1)
    full, capped = make_solver(pkg, pat, d, settings=st), make_solver(pkg, pat, d, settings=st)
    capped.set_max_ctas(5)
    outs = []
    for s in (full, capped):
        s.set_warm_start(d["x0"])
        outs.append(s.solve(want_y=True))
    ref = oracle_batch(O, pat, d)
    compare(outs[0], ref)
    for k in ("x", "y", "iters", "exit_code", "rho_updates"):
        np.testing.assert_array_equal(outs[0][k], outs[1][k])
    code_dev = torch.empty(B, dtype=torch.int32, device="cuda")
    capped.get_async(exit_code=code_dev)
    capped.sync()
    np.testing.assert_array_equal(code_dev.cpu().numpy(), outs[1]["exit_code"])
    upd = pkg.problem.relinearize(pat, 2100, np.where(np.isnan(outs[0]["x"]), d["x0"], outs[0]["x"]))
    capped.set_max_ctas(0)
    full.set_max_ctas(37)
    again = []
    for s in (full, capped):
        s.update(upd["A_val"], upd["l"], upd["u"])
        again.append(s.solve(want_y=True))
    for k in ("x", "y", "iters", "exit_code", "rho_updates"):
        np.testing.assert_array_equal(again[0][k], again[1][k])
    with pytest.raises(Exception):
        full.set_max_ctas(-1)
    full.close()
    capped.close()


def test_full_batch_kkt_properties():
    """BASELINE configs[1] at its full size (4096 QPs, more than the oracle is asked to solve in a test) through properties that
    do not need the oracle: every QP reported optimal satisfies OSQP's own termination criteria, evaluated here from the
    returned x and y in plain numpy (primal residual dist(Ax, [l, u]) and dual residual |Px + A'y| against eps_abs + eps_rel x
    the norms OSQP uses, dual signs on inactive sides); every primal infeasible one comes with a certificate that proves it;
    a random sample of 96 is compared with the oracle on top."""
    pkg, O = package(), oracle()
    P = pkg.problem
    pat = P.Pattern(P.default_config())
    B = 4096
    d = P.generate(pat, 0, B)
    s = make_solver(pkg, pat, d, settings=pkg.solver.default_settings(linsys=4))   # the kernel bench.py measures
    assert s.info()["linsys"] == 4
    s.set_warm_start(d["x0"])
    got = s.solve(want_y=True)
    cert, _ = s.certificates()
    s.close()
    Pm = pat.P_scipy()
    eps = 1e-3
    n_opt = n_inf = 0
    for b in range(B):
        A = pat.A_scipy(d["A_val"][b])
        l, u = np.maximum(d["l"][b], -1e30), np.minimum(d["u"][b], 1e30)
        code = got["exit_code"][b]
        if code in (0, 3):
            x, y = got["x"][b], got["y"][b]
            ax, px, aty = A @ x, Pm @ x, A.T @ y
            z = np.clip(ax, l, u)
            slack = 10.0 if code == 3 else 1.0   # kOptimalInaccurate: the 10x tolerances of the max_iter check
            # (the unscaled criteria of OSQP; its z is not returned, so dist(Ax, [l, u]) <= |Ax - z_osqp| is the testable form)
            assert np.abs(ax - z).max() <= slack * (eps + eps * max(np.abs(ax).max(), np.abs(z).max())) * 1.0001, b
            assert np.abs(px + aty).max() <= slack * (eps + eps * max(np.abs(px).max(), np.abs(aty).max())) * 1.0001, b
            # dual signs: y <= 0 only where the lower bound can be active, y >= 0 only where the upper one can
            tol = slack * 1e-2 * max(1.0, np.abs(y).max())
            assert (y[u > 1e29] <= tol).all() and (y[l < -1e29] >= -tol).all(), b
            n_opt += 1
        elif code in (1, 4):
            v = cert[b]
            assert np.isnan(got["x"][b]).all() and abs(np.abs(v).max() - 1.0) < 1e-12
            support = u @ np.maximum(v, 0.0) + l @ np.minimum(v, 0.0)
            assert support < 0 and np.abs(A.T @ v).max() <= 1e-3 * max(1.0, abs(support)), b
            n_inf += 1
        else:
            raise AssertionError((b, code))
    assert n_opt > 0.95 * B and n_inf > 0
    pick = np.random.default_rng(0).choice(B, 96, replace=False)
    sub = {k: v[pick] for k, v in d.items()}
    ref = oracle_batch(O, pat, sub)
    compare({k: got[k][pick] for k in ("x", "y", "exit_code", "iters", "rho_updates")}, ref)
