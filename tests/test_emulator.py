"""The sequential CPU walk-through of the device algorithm (tests/cpp/emulate.cpp: product symbolic analysis
+ reduced block-tridiagonal KKT + two-way block factorisation) against the oracle (KKT LDL^T): validates,
without a GPU, the index maps the kernels stream and that the reduced formulation reproduces OSQP's
iterates -- same status, same iteration counts, same rho updates, x and y to 1e-6."""
import numpy as np
import pytest
import scipy.sparse as sp

import emul
from common import oracle, package


def run_pair(pkg, O, cfg_kw, count, first=0, **settings):
    P = pkg.problem
    pat = P.Pattern(P.default_config(**cfg_kw))
    d = P.generate(pat, first, count)
    st = O.default_settings(**settings)
    E = emul.Emulator(count, pat.n, pat.m, pat.P_colptr, pat.P_rowidx, pat.P_val, pat.A_colptr, pat.A_rowidx, st,
                      var_order=pat.var_order, block_size=pat.block_size)
    E.set_data(d["A_val"], d["l"], d["u"])
    E.warm_start_x(d["x0"])
    got = E.solve()
    ref = O.solve_batch(pat.n, pat.m, pat.P_scipy(), (pat.A_colptr, pat.A_rowidx), d["A_val"], d["l"], d["u"], d["x0"],
                        settings=st, want_y=True)
    return pat, d, E, got, ref


def check(got, ref):
    np.testing.assert_array_equal(got["exit_code"], ref["exit_code"])
    np.testing.assert_array_equal(got["iters"], ref["iters"])
    np.testing.assert_array_equal(got["rho_updates"], ref["rho_updates"])
    ok = ref["exit_code"] == 0
    if not ok.any():
        return
    scale = np.abs(ref["x"][ok]).max(axis=1, keepdims=True)
    assert (np.abs(got["x"][ok] - ref["x"][ok]) / scale).max() < 1e-6
    assert np.abs(got["y"][ok] - ref["y"][ok]).max() < 1e-6 * max(1.0, np.abs(ref["y"][ok]).max())


def test_symbolic_structure_of_the_headline_shape():
    pkg, O = package(), oracle()
    pat, d, E, got, ref = run_pair(pkg, O, {}, 2)
    assert (E.m_used, E.bs, E.T, E.n_pad, E.ow, E.owT) == (1558, 14, 50, 700, 2, 2)  # n=700, 1558 used rows
    check(got, ref)


@pytest.mark.parametrize("kw", [dict(), dict(n_dim=6), dict(n_balls=2, n_obstacles=2), dict(waypoints=12), dict(n_dim=3, waypoints=9)])
def test_emulator_matches_oracle(kw):
    pkg, O = package(), oracle()
    _, _, _, got, ref = run_pair(pkg, O, kw, 12)
    check(got, ref)


def test_tight_tolerance():
    pkg, O = package(), oracle()
    _, _, _, got, ref = run_pair(pkg, O, dict(waypoints=20), 6, eps_abs=1e-8, eps_rel=1e-8)
    check(got, ref)


def test_sqp_update_rounds():
    pkg, O = package(), oracle()
    P = pkg.problem
    pat, d, E, got, _ = run_pair(pkg, O, dict(waypoints=16), 4)
    refs = [O.Solver(pat.P_scipy(), pat.A_scipy(d["A_val"][b]), d["l"][b], d["u"][b]) for b in range(4)]
    for b in range(4):
        refs[b].warm_start_x(d["x0"][b])
        refs[b].solve()
    for rnd in range(3):
        xs = np.array([np.where(np.isnan(r.x), d["x0"][b], r.x) for b, r in enumerate(refs)])
        upd = P.relinearize(pat, 0, xs)
        E.set_data(upd["A_val"], upd["l"], upd["u"], cold=False)
        got = E.solve()
        for b in range(4):
            refs[b].update(upd["A_val"][b], upd["l"][b], upd["u"][b])
            refs[b].solve()
            assert got["exit_code"][b] == refs[b].exit_code and got["iters"][b] == refs[b].iter
            if refs[b].exit_code == 0:
                assert np.abs(got["x"][b] - refs[b].x).max() < 1e-6 * max(1.0, np.abs(refs[b].x).max())


def test_generic_pattern_without_ordering_hint():
    """No var_order: reverse Cuthill-McKee + detected block size must give a valid block-tridiagonal form."""
    O = oracle()
    rng = np.random.default_rng(11)
    n, m = 12, 16
    M = rng.standard_normal((n, n))
    P = sp.csc_matrix(np.triu(M @ M.T + 0.1 * np.eye(n)))
    A = sp.csc_matrix(rng.standard_normal((m, n)) * (rng.random((m, n)) < 0.4))
    A.sort_indices()
    l, u = -rng.random(m) - 0.1, rng.random(m) + 0.1
    q = rng.standard_normal(n)
    st = O.default_settings(eps_abs=1e-7, eps_rel=1e-7)
    E = emul.Emulator(1, n, m, P.indptr, P.indices, P.data, A.indptr, A.indices, st)
    E.set_data(A.data[None, :], l[None, :], u[None, :], q=q[None, :])
    got = E.solve()
    ref = O.Solver(P, A, l, u, q=q, settings=st)
    ref.solve()
    assert got["exit_code"][0] == ref.exit_code and got["iters"][0] == ref.iter
    assert np.abs(got["x"][0] - ref.x).max() < 1e-6 * max(1.0, np.abs(ref.x).max())


def test_non_block_tridiagonal_order_is_rejected():
    pkg = package()
    P = pkg.problem
    pat = P.Pattern(P.default_config(waypoints=8))
    O = oracle()
    with pytest.raises(ValueError):
        emul.Emulator(1, pat.n, pat.m, pat.P_colptr, pat.P_rowidx, pat.P_val, pat.A_colptr, pat.A_rowidx, O.default_settings(),
                      var_order=np.arange(pat.n), block_size=pat.block_size)  # [all q | all v] order is not block tridiagonal


@pytest.mark.parametrize("bs,T", [(3, 5), (6, 9), (10, 12), (16, 7), (32, 3), (5, 1), (4, 2)])
def test_random_block_tridiagonal_qps(bs, T):
    """Generic patterns (no GOMP structure): random block sizes / chain lengths, q != 0, all row types."""
    from common import random_block_qp
    pkg, O = package(), oracle()
    rng = np.random.default_rng(100 * bs + T)
    B = 4
    P, A, Ax, l, u, q = random_block_qp(rng, bs, T, B)
    n, m = P.shape[0], A.shape[0]
    st = O.default_settings(eps_abs=1e-6, eps_rel=1e-6)
    E = emul.Emulator(B, n, m, P.indptr, P.indices, P.data, A.indptr, A.indices, st, var_order=np.arange(n), block_size=bs)
    assert (E.bs, E.T) == (bs, T)
    E.set_data(Ax, l, u, q)
    got = E.solve()
    for b in range(B):
        Ab = sp.csc_matrix((Ax[b], A.indices, A.indptr), shape=(m, n))
        r = O.Solver(P, Ab, l[b], u[b], q=q[b], settings=st)
        r.solve()
        assert got["exit_code"][b] == r.exit_code and got["iters"][b] == r.iter and got["rho_updates"][b] == r.rho_updates
        if r.exit_code == 0:
            assert np.abs(got["x"][b] - r.x).max() < 1e-6 * max(1.0, np.abs(r.x).max())
            assert np.abs(got["y"][b] - r.y).max() < 1e-6 * max(1.0, np.abs(r.y).max())


def test_long_horizon_shape_on_the_emulator():
    """BASELINE configs[3] shape (N = 7, H = 400) through the product's symbolic analysis and the sequential
    walk-through of the device algorithm: same status, iterations and solution as the oracle (no GPU needed)."""
    pkg, O = package(), oracle()
    pat, d, E, got, ref = run_pair(pkg, O, dict(waypoints=400), 2, first=40)
    assert (pat.n, pat.m, pat.nnzA) == (5600, 22372, 30730)
    assert (E.bs, E.T, E.n_pad, E.ow, E.owT) == (14, 400, 5600, 2, 2)
    check(got, ref)
