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


def numpy_relinearize(P, pat, first_id, x, A_val, l, u):
    """What gomp_relin.cu does per (problem, ball, waypoint), in numpy, from the same scene / row tables."""
    cfg = pat.cfg
    N, W, O = cfg.n_dim, cfg.waypoints, cfg.n_obstacles
    lines = P.scene(pat, first_id, x.shape[0])
    rows, kidx = P.cart_rows(pat)
    balls = ([dict(ball=1, radius=0.15, gripper=False)] if cfg.n_balls >= 2 else []) + [dict(ball=0, radius=0.05, gripper=True)]
    r = 0
    for b in balls:
        for w in range(W):
            q = x[:, w * N:(w + 1) * N]
            p, J = P.fk(cfg, q, ball=b["ball"], want_jac=True)
            pm = P.fk(cfg, x[:, (w - 1) * N:w * N], ball=b["ball"]) if w > 0 else None
            pp = P.fk(cfg, x[:, (w + 1) * N:(w + 2) * N], ball=b["ball"]) if w + 1 < W else None

            def write(axis, lo, hi):
                nonlocal r
                for j in range(N):
                    A_val[:, kidx[r, j]] = J[:, axis, j]
                l[:, rows[r]] = lo + b["radius"]
                u[:, rows[r]] = hi - b["radius"]
                r += 1

            def lin(limit, axis):
                return limit - p[:, axis] + np.einsum("bj,bj->b", J[:, axis, :], q)

            if b["gripper"]:
                box_lo, box_hi = [-1e30, -1e30, cfg.z_floor], [1e30, 1e30, 1e30]
                for axis in range(3):
                    write(axis, lin(box_lo[axis], axis), lin(box_hi[axis], axis))
            for k in range(O):
                D, Apt = lines[:, k, :2], lines[:, k, 2:]

                def dist(pt):
                    along = (pt[:, 0] - Apt[:, 0]) * D[:, 0] + (pt[:, 1] - Apt[:, 1]) * D[:, 1]
                    return np.stack([Apt[:, 0] + along * D[:, 0] - pt[:, 0], Apt[:, 1] + along * D[:, 1] - pt[:, 1]], axis=1)
                dv = dist(p)
                hit = np.hypot(dv[:, 0], dv[:, 1]) < b["radius"]
                if pm is not None:
                    hit |= (dist(pm) * dv).sum(axis=1) < 0
                if pp is not None:
                    hit |= (dv * dist(pp)).sum(axis=1) < 0
                lo = np.where(hit, lin(p[:, 2] + (Apt[:, 2] - p[:, 2]), 2), -1e30)
                write(2, lo, np.full_like(lo, 1e30))
    assert r == len(rows)


@pytest.mark.parametrize("kw", [dict(), dict(n_balls=2, n_obstacles=2, n_dim=6, waypoints=12)])
def test_device_relinearize_tables_reproduce_host_withObstacles(kw):
    """The inputs of the device kernel (gomp_problem_scene, gomp_problem_cart_rows) and its per-waypoint recipe give
    the same A_val / l / u as ConstraintBuilder::withObstacles on the host (which only rewrites the Cartesian rows)."""
    pkg = package()
    P = pkg.problem
    pat = P.Pattern(P.default_config(**kw))
    first, B = 3, 6
    d = P.generate(pat, first, B)
    rng = np.random.default_rng(1)
    x = d["x0"] + 0.05 * rng.standard_normal(d["x0"].shape)
    want = P.relinearize(pat, first, x)
    A_val, l, u = d["A_val"].copy(), d["l"].copy(), d["u"].copy()
    numpy_relinearize(P, pat, first, x, A_val, l, u)
    np.testing.assert_allclose(A_val, want["A_val"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(l, want["l"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(u, want["u"], rtol=1e-13, atol=1e-13)
    assert (want["l"] > -1e29).sum() > (d["l"] > -1e29).sum() or (want["l"] != d["l"]).any()


def test_horizon_selection_rule():
    """reference src/gomp-solver.h:42-51: last kOptimal horizon before the first failed QP; kUnknown does not stop the walk."""
    pkg = package()
    codes = [[0, 0, 10, 1, 10],    # longest horizon
             [0, 10, 0, 0, 10],
             [0, 1, 0, 0, 10],
             [6, 0, 10, 0, 10]]    # shortest
    out, pick = pkg.sqp.select_horizon(codes)
    assert out.tolist() == [0, 0, 0, 10, 10]
    assert pick.tolist() == [2, 0, 2, -1, -1]
