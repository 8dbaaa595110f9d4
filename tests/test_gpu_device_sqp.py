"""Device-side withObstacles / isSolutionOK (osqp-solver_b200/csrc/gomp_relin.cu) against the
host facade (ConstraintBuilder::withObstacles through libgomp_host), and BASELINE.json configs[2] -- the
warm-started SQP re-solve sequence -- run without leaving the GPU, checked round by round against the oracle
fed with the very arrays the device produced."""
import numpy as np
import pytest

from common import oracle, package
from test_gpu_parity import compare, make_solver

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kw", [dict(), dict(n_balls=2, n_obstacles=2), dict(n_dim=6, waypoints=12), dict(n_dim=3, waypoints=20, n_obstacles=0)])
def test_device_relinearize_matches_host(kw):
    import torch
    pkg = package()
    P = pkg.problem
    pat = P.Pattern(P.default_config(**kw))
    first, B = 11, 64
    d = P.generate(pat, first, B)
    rng = np.random.default_rng(2)
    x = d["x0"] + 0.05 * rng.standard_normal(d["x0"].shape)
    x[5] = np.nan                      # an infeasible QP hands back NaN: that problem uses the fallback trajectory
    xf = np.where(np.isnan(x), d["x0"], x)
    want = P.relinearize(pat, first, xf)
    dev = {k: torch.from_numpy(d[k]).cuda() for k in ("A_val", "l", "u", "x0")}
    sc = P.DeviceScene(pat, first, B)
    sc.relinearize(torch.from_numpy(x).cuda(), dev["A_val"], dev["l"], dev["u"], x_fallback=dev["x0"])
    torch.cuda.synchronize()
    A_val, l, u = (dev[k].cpu().numpy() for k in ("A_val", "l", "u"))
    # same classification (which rows are active) and the same numbers up to sin/cos rounding
    assert ((l > -1e29) == (want["l"] > -1e29)).all() and ((u < 1e29) == (want["u"] < 1e29)).all()
    np.testing.assert_allclose(A_val, want["A_val"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(l, want["l"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(u, want["u"], rtol=1e-13, atol=1e-13)
    assert (want["l"] != d["l"]).any()
    # isSolutionOK
    ok = torch.empty(B, dtype=torch.int32, device="cuda")
    sc.check(torch.from_numpy(x).cuda(), ok)
    torch.cuda.synchronize()
    host_ok = P.check(pat, first, xf)
    host_ok[5] = False                 # NaN trajectory: not a solution
    np.testing.assert_array_equal(ok.cpu().numpy().astype(bool), host_ok)
    sc.close()


def test_sqp_sequence_stays_on_device():
    """configs[2]: solve -> withObstacles(x) -> update(A, l, u) -> solve ... with x, A_val, l, u resident in HBM
    (borrowed by the solver, rewritten in place by the re-linearisation kernel on the solver's stream)."""
    import torch
    pkg, O = package(), oracle()
    P = pkg.problem
    pat = P.Pattern(P.default_config())
    first, B, rounds = 200, 24, 4
    d = P.generate(pat, first, B)
    stream = torch.cuda.Stream()
    dev = {k: torch.from_numpy(d[k]).cuda() for k in ("A_val", "l", "u", "x0")}
    st = pkg.solver.default_settings(borrow_inputs=1)
    st.stream = stream.cuda_stream
    s = make_solver(pkg, pat, dev, settings=st, batch=B)
    s.set_warm_start(dev["x0"])
    sc = P.DeviceScene(pat, first, B)
    xdev = torch.empty((B, pat.n), dtype=torch.float64, device="cuda")
    ydev = torch.empty((B, pat.m), dtype=torch.float64, device="cuda")
    ok = torch.empty(B, dtype=torch.int32, device="cuda")
    Ps = pat.P_scipy()
    refs = [O.Solver(Ps, pat.A_scipy(d["A_val"][b]), d["l"][b], d["u"][b]) for b in range(B)]
    for b in range(B):
        refs[b].warm_start_x(d["x0"][b])
    seen_active = False
    for rnd in range(rounds):
        s.solve_async()
        got = s.get(x=xdev, y=ydev)
        got["x"], got["y"] = xdev.cpu().numpy(), ydev.cpu().numpy()
        for b in range(B):
            refs[b].solve()
        ref = dict(x=np.array([r.x for r in refs]), y=np.array([r.y for r in refs]),
                   exit_code=np.array([r.exit_code for r in refs]), iters=np.array([r.iter for r in refs]),
                   rho_updates=np.array([r.rho_updates for r in refs]))
        compare(got, ref)
        # next round: everything on the device, same stream as the solver
        sc.check(xdev, ok, stream=stream.cuda_stream)
        sc.relinearize(xdev, dev["A_val"], dev["l"], dev["u"], x_fallback=dev["x0"], stream=stream.cuda_stream)
        s.update(dev["A_val"], dev["l"], dev["u"])
        stream.synchronize()
        A_val, l, u = (dev[k].cpu().numpy() for k in ("A_val", "l", "u"))
        seen_active |= bool((l != d["l"]).any())
        for b in range(B):                      # the oracle gets the arrays the device produced
            refs[b].update(A_val[b], l[b], u[b])
        xs = np.where(np.isnan(got["x"]), d["x0"], got["x"])
        np.testing.assert_array_equal(ok.cpu().numpy().astype(bool) | np.isnan(got["x"][:, 0]), P.check(pat, first, xs) | np.isnan(got["x"][:, 0]))
    assert seen_active
    s.close()
    sc.close()


def test_batched_sqp_driver_matches_per_problem_reference_loop():
    """GOMPSolver::run for one horizon (reference src/gomp-solver.h:57-91) over a batch: per-problem exit code,
    number of QP solves and trajectory equal those of the reference's loop run problem by problem on the CPU
    (oracle QP solver + host withObstacles / isSolutionOK)."""
    pkg, O = package(), oracle()
    P = pkg.problem
    pat = P.Pattern(P.default_config())
    first, B, max_rounds = 0, 40, 12
    drv = pkg.sqp.BatchSQP(pat, first, B)
    out = drv.run(max_rounds=max_rounds)
    x_dev = out["x"].cpu().numpy()
    d = drv.host
    Ps = pat.P_scipy()
    for b in range(B):
        ref = O.Solver(Ps, pat.A_scipy(d["A_val"][b]), d["l"][b], d["u"][b])
        ref.warm_start_x(d["x0"][b])
        code, rounds, x = 10, 0, None
        for rnd in range(max_rounds):
            ref.solve()
            rounds += 1
            x = ref.x
            if ref.exit_code != 0:
                break
            if P.check(pat, first + b, x[None, :])[0]:
                code = 0
                break
            upd = P.relinearize(pat, first + b, x[None, :])
            ref.update(upd["A_val"][0], upd["l"][0], upd["u"][0])
        assert out["exit_code"][b] == code, (b, out["exit_code"][b], code)
        assert out["rounds"][b] == rounds, (b, out["rounds"][b], rounds)
        if code == 0:
            assert np.abs(x_dev[b] - x).max() < 1e-6 * np.abs(x).max()
    assert (out["exit_code"] == 0).sum() >= B // 2 and (out["rounds"] > 1).any() and len(set(out["rounds"].tolist())) > 1
    drv.close()


def test_cpp_batch_solver_matches_python_driver():
    """GOMPBatchSolver (osqp-solver_b200/include/gomp/gomp-batch-solver.h: host C++ over the two C ABIs, no CUDA
    headers) gives the same per-problem exit codes and round counts as sqp.BatchSQP."""
    import os
    import subprocess
    from common import ROOT
    pkg = package()
    P = pkg.problem
    pkg_dir = os.path.join(ROOT, "osqp-solver_b200")
    exe = os.path.join(pkg_dir, "build", "batch-sqp-example")
    subprocess.check_call(["make", "-C", pkg_dir, "example"], stdout=subprocess.DEVNULL)
    B, first, max_rounds = 48, 100, 10
    out = subprocess.run([exe, str(B), str(first), str(max_rounds), "rerun"], capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[2].strip() == "rerun_equal 1"   # GOMPBatchSolver::run is re-entrant
    pairs = [tuple(int(v) for v in tok.split(":")) for tok in out[1].split()[1:]]
    drv = pkg.sqp.BatchSQP(P.Pattern(P.default_config()), first, B)
    ref = drv.run(max_rounds=max_rounds)
    drv.close()
    assert pairs == list(zip(ref["exit_code"].tolist(), ref["rounds"].tolist()))
    assert f"qp_solves {ref['total_qp_solves']}" in out[0]


def test_shortest_horizon_batch():
    """Time-shrink outer loop (reference src/gomp-solver.h:38-55) as one batch per horizon: the returned horizon is
    the shortest kOptimal one, its trajectory passes the host isSolutionOK at that horizon, and horizons longer than
    the resident kernel's limit run on the streamed back end."""
    pkg = package()
    P = pkg.problem
    cfg = P.default_config(waypoints=80)
    B = 12
    out = pkg.sqp.shortest_horizon(cfg, 300, B, max_rounds=8)
    assert out["horizons"] == [80, 72, 64, 56, 48, 40, 32, 24, 16, 8]
    assert (out["exit_code"] == 0).sum() >= B // 2
    import copy
    for b in range(B):
        if out["exit_code"][b] != 0:
            assert out["waypoints"][b] == 0 and out["x"][b] is None
            continue
        w = int(out["waypoints"][b])
        c = copy.copy(cfg)
        c.waypoints = w
        pat = P.Pattern(c)
        x = out["x"][b].copy()
        assert x.shape == (pat.n,)
        x[pat.n // 2:] *= cfg.time_step   # back to per-step velocities for the check
        assert P.check(pat, 300 + b, x[None, :])[0]


def test_chained_time_shrink_matches_reference_loop():
    """GOMPSolver::run(start, goal) with the reference's chained warm starts (src/gomp-solver.h:38-55, the slicing of :43-44
    included): the batched driver against the same loop run problem by problem on the CPU (oracle QP solver + host
    withObstacles / isSolutionOK) -- per-horizon exit codes, the horizon kept, and the returned trajectory."""
    import copy
    pkg, O = package(), oracle()
    P = pkg.problem
    cfg = P.default_config(waypoints=40)
    first, B, max_rounds = 500, 10, 8
    N, W = cfg.n_dim, cfg.waypoints
    out = pkg.sqp.shortest_horizon_chained(cfg, first, B, max_rounds=max_rounds)
    assert out["horizons"] == [40, 36, 32, 28, 24, 20, 16, 12, 8, 4]
    pats = {}
    for w in out["horizons"]:
        c = copy.copy(cfg)
        c.waypoints = w
        pats[w] = P.Pattern(c)
    full = P.generate(pats[W], first, B)
    for b in range(B):
        last, code, kept = full["x0"][b].copy(), 10, 0                       # :39-40
        for k, w in enumerate(out["horizons"]):                              # :41-42
            pat = pats[w]
            warm = np.concatenate([last[:w * N], last[w * N:2 * w * N]])     # :43-44
            d = P.relinearize(pat, first + b, warm[None, :])                 # :58 initConstraints(..., warm_start, waypoints)
            ref = O.Solver(pat.P_scipy(), pat.A_scipy(d["A_val"][0]), d["l"][0], d["u"][0])
            ref.warm_start_x(warm)                                           # :65
            hcode, sol = 10, warm
            for rnd in range(max_rounds):                                    # :70
                ref.solve()
                sol = ref.x
                if ref.exit_code != 0:                                       # :72-76
                    break
                if P.check(pat, first + b, sol[None, :])[0]:                 # :77-81
                    hcode = 0
                    break
                upd = P.relinearize(pat, first + b, sol[None, :])            # :83-87
                ref.update(upd["A_val"][0], upd["l"][0], upd["u"][0])
            assert out["codes_per_horizon"][k][b] == hcode, (b, w, out["codes_per_horizon"][k][b], hcode)
            if hcode == 0:                                                   # :48-51
                code, kept, last = 0, w, sol.copy()
        assert out["exit_code"][b] == code and out["waypoints"][b] == kept
        want = last.copy()
        want[want.size // 2:] /= cfg.time_step                               # :52
        assert out["x"][b].shape == want.shape
        assert np.abs(out["x"][b] - want).max() <= 1e-6 * max(1.0, np.abs(want).max())
    assert (out["exit_code"] == 0).any() and len(set(out["waypoints"].tolist())) > 1


def test_batch_sqp_run_is_reentrant():
    """a second run() used to start from the last round's linearisation with stale duals and rho; it now
    starts over (fresh QPSolver per run in the reference, src/gomp-solver.h:61-65) and reproduces the first run."""
    pkg = package()
    P = pkg.problem
    drv = pkg.sqp.BatchSQP(P.Pattern(P.default_config()), 60, 24)
    a = drv.run(max_rounds=6)
    xa = a["x"].cpu().numpy().copy()
    b = drv.run(max_rounds=6)
    assert (a["rounds"] > 1).any()
    np.testing.assert_array_equal(a["exit_code"], b["exit_code"])
    np.testing.assert_array_equal(a["rounds"], b["rounds"])
    np.testing.assert_array_equal(np.nan_to_num(xa), np.nan_to_num(b["x"].cpu().numpy()))
    drv.close()
