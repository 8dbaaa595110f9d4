"""The synthetic GOMP workload: sizes from the reference's row formula, FK / Jacobian consistency, the
reference's warm start and SQP re-linearisation semantics, determinism of the generator."""
import numpy as np
import pytest

from common import package


def test_headline_dimensions_match_reference_formula():
    pkg = package()
    P = pkg.problem
    pat = P.Pattern(P.default_config())
    N, W, O, B = 7, 50, 1, 1
    m = N * (W - 1) + N * (3 * W - 3 + W * (3 + O * B))   # reference constraint-builder.h:43-44
    assert (pat.n, pat.m, pat.nnzA, pat.nnzP) == (2 * N * W, m, 3780, 1036)
    assert pat.m == 2772 and (np.bincount(pat.A_rowidx, minlength=pat.m) > 0).sum() == 1558
    heavy = P.Pattern(P.default_config(n_balls=2, n_obstacles=2))
    assert (heavy.m, heavy.nnzA) == (3822, 4830)  # B=2 (one gripper), O=2


def test_fk_jacobian_finite_differences():
    pkg = package()
    P = pkg.problem
    cfg = P.default_config()
    rng = np.random.default_rng(0)
    q = rng.uniform(-2.5, 2.5, (20, 7))
    for ball in (0, 1):
        xyz, J = P.fk(cfg, q, ball=ball, want_jac=True)
        for j in range(7):
            dq = q.copy()
            dq[:, j] += 1e-6
            fd = (P.fk(cfg, dq, ball=ball) - xyz) / 1e-6
            np.testing.assert_allclose(fd, J[:, :, j], atol=2e-5)
    np.testing.assert_allclose(P.fk(cfg, np.zeros(7))[0], [0.088, 0.0, 0.926], atol=1e-12)  # home pose of the Panda-like chain


def test_generator_is_deterministic_and_warm_start_is_linspace():
    pkg = package()
    P = pkg.problem
    pat = P.Pattern(P.default_config(waypoints=10))
    a = P.generate(pat, 5, 3, nthreads=1)
    b = P.generate(pat, 0, 8, nthreads=4)
    for k in ("A_val", "l", "u", "x0"):
        np.testing.assert_array_equal(a[k], b[k][5:8])
    N, W = 7, 10
    s, g = a["start_goal"][0, :N], a["start_goal"][0, N:]
    want = np.concatenate([np.linspace(s, g, W).ravel(), np.zeros(N * W)])   # reference gomp-solver.h:105-116
    np.testing.assert_allclose(a["x0"][0], want, atol=1e-15)
    # start pinned at waypoint 0, goal pinned at waypoint W-3 (reference gomp-solver.h:131-133)
    off = N * (W - 1)
    np.testing.assert_array_equal(a["l"][0, off:off + N], s)
    np.testing.assert_array_equal(a["u"][0, off + N * (W - 3):off + N * (W - 2)], g)


def test_relinearize_keeps_pattern_and_changes_only_cartesian_rows():
    pkg = package()
    P = pkg.problem
    pat = P.Pattern(P.default_config(waypoints=10))
    d = P.generate(pat, 0, 2)
    x = d["x0"] + 0.01
    upd = P.relinearize(pat, 0, x)
    N, W = 7, 10
    first_cart = N * (W - 1) + N * (3 * W - 3)   # reference constraint-builder.h:92
    rows = pat.A_rowidx
    dyn = rows < first_cart
    np.testing.assert_array_equal(upd["A_val"][:, dyn], d["A_val"][:, dyn])
    assert (upd["A_val"][:, ~dyn] != d["A_val"][:, ~dyn]).any()
    np.testing.assert_array_equal(upd["l"][:, :first_cart], d["l"][:, :first_cart])
    assert P.check(pat, 0, d["x0"]).shape == (2,)
