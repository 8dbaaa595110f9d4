"""The oracle (oracle/osqp_ref.c, a CPU restatement of OSQP 0.6) against answers that do not depend on
it: closed-form QPs, KKT optimality conditions, scipy's trust-constr, infeasibility certificates.
Genuine OSQP is not available offline, so this is the strongest pin there is: parity unpinned."""
import numpy as np
import pytest
import scipy.sparse as sp
from scipy.optimize import LinearConstraint, minimize

from common import oracle

O = oracle()


def test_default_settings_are_osqp_defaults():
    s = O.default_settings()
    assert (s.rho, s.sigma, s.alpha, s.scaling, s.max_iter, s.check_termination) == (0.1, 1e-6, 1.6, 10, 4000, 25)
    assert s.adaptive_rho_interval == 25  # pinned; OSQP's own default (0) is wall-clock driven


def test_closed_form_box_qp():
    # min 1/2 sum (x_i - c_i)^2  s.t. 0 <= x <= 1  ->  x = clip(c, 0, 1)
    c = np.array([-0.5, 0.3, 1.7, 0.9])
    P = sp.identity(4, format="csc")
    A = sp.identity(4, format="csc")
    s = O.Solver(P, A, np.zeros(4), np.ones(4), q=-c, settings=O.default_settings(eps_abs=1e-9, eps_rel=1e-9))
    code, x = s.solve()
    assert code == 0
    np.testing.assert_allclose(x, np.clip(c, 0, 1), atol=1e-7)
    # duals: y_i = -(x_i - c_i) on active bounds
    np.testing.assert_allclose(s.y, -(x - c), atol=1e-6)


def test_equality_constrained_closed_form():
    # min 1/2 x'x  s.t. a'x = 1  ->  x = a / (a'a)
    a = np.array([1.0, 2.0, -1.0])
    s = O.Solver(sp.identity(3, format="csc"), sp.csc_matrix(a[None, :]), np.array([1.0]), np.array([1.0]),
                 settings=O.default_settings(eps_abs=1e-10, eps_rel=1e-10))
    code, x = s.solve()
    assert code == 0
    np.testing.assert_allclose(x, a / (a @ a), atol=1e-8)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_qp_kkt_and_scipy(seed):
    rng = np.random.default_rng(seed)
    n, m = 10, 15
    M = rng.standard_normal((n, n))
    P = sp.csc_matrix(M @ M.T + 0.1 * np.eye(n))
    A = sp.csc_matrix(rng.standard_normal((m, n)) * (rng.random((m, n)) < 0.5))
    q = rng.standard_normal(n)
    l, u = -rng.random(m), rng.random(m)
    s = O.Solver(P, A, l, u, q=q, settings=O.default_settings(eps_abs=1e-9, eps_rel=1e-9, max_iter=20000))
    code, x = s.solve()
    y = s.y
    assert code == 0
    Ax = A @ x
    assert max(0.0, (l - Ax).max(), (Ax - u).max()) < 1e-7          # primal feasibility
    assert np.abs(P @ x + q + A.T @ y).max() < 1e-7                  # stationarity
    assert np.abs(np.minimum(y, 0) * (Ax - l)).max() < 1e-6          # complementary slackness
    assert np.abs(np.maximum(y, 0) * (u - Ax)).max() < 1e-6
    r = minimize(lambda v: 0.5 * v @ (P @ v) + q @ v, np.zeros(n), jac=lambda v: P @ v + q, hess=lambda v: P.toarray(),
                 constraints=[LinearConstraint(A.toarray(), l, u)], method="trust-constr",
                 options=dict(gtol=1e-12, xtol=1e-14, maxiter=5000))
    np.testing.assert_allclose(x, r.x, atol=1e-6)


def test_primal_and_dual_infeasibility_certificates():
    s = O.Solver(sp.identity(2, format="csc"), sp.csc_matrix(np.array([[1.0, 0], [1.0, 0]])), np.array([1.0, -5.0]), np.array([2.0, 0.0]))
    code, x = s.solve()
    assert code == 1 and np.isnan(x).all()      # kPrimalInfeasible, NaN solution (OSQP store_solution)
    s = O.Solver(sp.csc_matrix((1, 1)), sp.csc_matrix(np.array([[1.0]])), np.array([0.0]), np.array([1e30]), q=np.array([-1.0]))
    code, x = s.solve()
    assert code == 2                              # kDualInfeasible


def test_max_iterations_and_inaccurate_status():
    rng = np.random.default_rng(3)
    n, m = 8, 12
    M = rng.standard_normal((n, n))
    P = sp.csc_matrix(M @ M.T + 0.01 * np.eye(n))
    A = sp.csc_matrix(rng.standard_normal((m, n)))
    s = O.Solver(P, A, -np.ones(m), np.ones(m), q=rng.standard_normal(n),
                 settings=O.default_settings(eps_abs=1e-14, eps_rel=1e-14, max_iter=50))
    code, _ = s.solve()
    assert code in (3, 6) and s.iter == 50        # kOptimalInaccurate or kMaxIterations after exactly max_iter


def test_warm_start_and_update_sequence_keeps_state():
    """update_A + update_bounds keep x, z, y and rho (warm start); a re-solve of unchanged data converges
    immediately; an exact warm start needs no more iterations than a cold start."""
    rng = np.random.default_rng(4)
    n, m = 6, 9
    M = rng.standard_normal((n, n))
    P = sp.csc_matrix(M @ M.T + np.eye(n))
    A = sp.csc_matrix(rng.standard_normal((m, n)))
    A.sort_indices()
    l, u = -np.ones(m), np.ones(m)
    s = O.Solver(P, A, l, u, q=rng.standard_normal(n))
    s.solve()
    it_cold = s.iter
    s.update(A.data, l, u)
    s.solve()
    assert s.iter <= it_cold and s.exit_code == 0
    s.update(A.data * 1.01, 0.9 * l, 0.9 * u)
    code, x = s.solve()
    assert code == 0
    Ab = sp.csc_matrix((A.data * 1.01, A.indices, A.indptr), shape=(m, n))
    assert ((Ab @ x) <= 0.9 * u + 2e-3).all() and ((Ab @ x) >= 0.9 * l - 2e-3).all()


def test_kkt_ordering_is_a_permutation():
    rng = np.random.default_rng(5)
    A = sp.random(20, 12, 0.3, random_state=1, format="csc")
    P = sp.identity(12, format="csc")
    perm = O.kkt_ordering(12, 20, P.indptr, P.indices, A.indptr, A.indices)
    assert sorted(perm.tolist()) == list(range(32))


def test_infeasibility_certificates_are_normalised():
    """OSQP store_solution: delta_y / delta_x of an infeasible problem are handed out with unit infinity norm (what osqp-cpp's
    primal_infeasibility_certificate / dual_infeasibility_certificate expose) and certify infeasibility."""
    import scipy.sparse as sp
    O = oracle()
    P = sp.csc_matrix(np.eye(2))
    A = sp.csc_matrix(np.array([[1.0, 0.0], [1.0, 0.0], [0.0, 1.0]]))
    l, u = np.array([1.0, -5.0, -1.0]), np.array([2.0, 0.0, 1.0])
    s = O.Solver(P, A, l, u)
    s.solve()
    assert s.exit_code == 1
    v = s.prim_inf_cert
    assert abs(np.abs(v).max() - 1.0) < 1e-12 and np.abs(A.T @ v).max() < 1e-6
    assert u @ np.maximum(v, 0) + l @ np.minimum(v, 0) < 0
    s = O.Solver(sp.csc_matrix((1, 1)), sp.csc_matrix(np.array([[1.0]])), np.array([0.0]), np.array([1e30]), q=np.array([-1.0]))
    s.solve()
    assert s.exit_code == 2
    w = s.dual_inf_cert
    assert abs(np.abs(w).max() - 1.0) < 1e-12 and -1.0 * w[0] < 0
