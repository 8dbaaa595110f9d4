"""Batched GOMPSolver::run for one horizon (reference src/gomp-solver.h:57-91) on the GPU.

Every problem of the batch runs the reference's sequential-convex-programming loop

    solve -> not kOptimal: give up | isSolutionOK: done | withObstacles(x) -> update -> again   (<= 100 rounds)

with the data plane on the device: the QPs are solved by libgomp_cuda, the feasibility re-check and the
re-linearisation are the kernels of gomp_relin.cu writing A_val / l / u in place, and a per-problem mask
(gomp_batch_set_active) freezes problems that have finished while the others iterate.  The host sees three
int32 per problem and round (exit code, ok flag, mask).  Needs torch for device buffers.
"""
import numpy as np

from . import problem as _problem
from . import solver as _solver

MAX_ITERATIONS = 100  # reference src/gomp-solver.h:10
K_OPTIMAL, K_UNKNOWN = 0, 10


class BatchSQP:
    def __init__(self, pattern, first_id, count, settings=None, device=None, warm_start=None):
        """warm_start [count, n] (optional): the trajectory each problem is warm-started from AND linearised around, as
        GOMPSolver::run(start, goal, waypoints, warm_start) does (reference src/gomp-solver.h:57-65, initConstraints);
        default: the straight line of gomp-solver.h:105-116."""
        import torch
        self.torch = torch
        self.pattern, self.first_id, self.count = pattern, first_id, count
        self.stream = torch.cuda.Stream()
        st = settings if settings is not None else _solver.default_settings()
        st.borrow_inputs = 1
        st.stream = self.stream.cuda_stream
        if warm_start is None:
            host = _problem.generate(pattern, first_id, count)
        else:
            warm_start = np.ascontiguousarray(warm_start, dtype=np.float64)
            assert warm_start.shape == (count, pattern.n)
            host = _problem.relinearize(pattern, first_id, warm_start)
            host["x0"] = warm_start
        self.host = host
        self.dev = {k: torch.from_numpy(host[k]).cuda() for k in ("A_val", "l", "u", "x0")}
        self.solver = _solver.BatchQPSolver(pattern.n, pattern.m, pattern.P_colptr, pattern.P_rowidx, pattern.P_val,
                                            pattern.A_colptr, pattern.A_rowidx, self.dev["A_val"], self.dev["l"], self.dev["u"],
                                            batch=count, settings=st, var_order=pattern.var_order, block_size=pattern.block_size)
        self.scene = _problem.DeviceScene(pattern, first_id, count)
        # every buffer the solver borrows is created and written on the solver's stream (the stream is non-blocking: nothing
        # orders it against torch's default stream)
        with torch.cuda.stream(self.stream):
            self.x = torch.empty((count, pattern.n), dtype=torch.float64, device="cuda")
            self.ok = torch.empty(count, dtype=torch.int32, device="cuda")
            self.active = torch.ones(count, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()  # the uploads above ran on the default stream
        self._ran = False

    def run(self, max_rounds=MAX_ITERATIONS):
        """Returns dict(exit_code [count] (kOptimal / kUnknown as GOMPSolver::run reports them), qp_exit_code (the last
        QP status of each problem), rounds [count] (QP solves used), x [count, n] torch CUDA tensor, total_qp_solves)."""
        torch = self.torch
        B = self.count
        s = self.solver
        if self._ran:
            # a second run starts over like the first one: the rounds rewrote A_val / l / u in place and left iterates, duals
            # and rho of the last round on the device -- the reference builds a fresh QPSolver per run (gomp-solver.h:61-65)
            with torch.cuda.stream(self.stream):
                for k in ("A_val", "l", "u"):
                    self.dev[k].copy_(torch.from_numpy(self.host[k]), non_blocking=False)
            s.setup(self.dev["A_val"], self.dev["l"], self.dev["u"])
        self._ran = True
        s.set_warm_start(self.dev["x0"])
        running = np.ones(B, dtype=bool)
        code = np.full(B, K_UNKNOWN, dtype=np.int32)
        qp_code = np.full(B, K_UNKNOWN, dtype=np.int32)
        rounds = np.zeros(B, dtype=np.int32)
        total = 0
        with torch.cuda.stream(self.stream):
            self.active.fill_(1)
        s.set_active(self.active)
        sc = self.stream.cuda_stream
        for rnd in range(max_rounds):
            s.solve_async()
            got = s.get(x=self.x)                          # x stays on the device; codes come to the host
            self.scene.check(self.x, self.ok, stream=sc)
            with torch.cuda.stream(self.stream):
                ok = self.ok.cpu().numpy().astype(bool)     # (synchronises the stream)
            total += int(running.sum())
            rounds[running] += 1
            qp_code[running] = got["exit_code"][running]
            failed = running & (got["exit_code"] != K_OPTIMAL)          # gomp-solver.h:74-77: no solution for this horizon
            done = running & ~failed & ok                               # :78-82
            code[done] = K_OPTIMAL
            running &= ~(failed | done)
            if not running.any():
                break
            with torch.cuda.stream(self.stream):
                self.active.copy_(torch.from_numpy(running.astype(np.int32)), non_blocking=False)
            # :83-87 withObstacles + update for everybody still running (finished problems are masked out of the solve;
            # re-linearising their rows is harmless and keeps the kernel branch-free)
            self.scene.relinearize(self.x, self.dev["A_val"], self.dev["l"], self.dev["u"], x_fallback=self.dev["x0"], stream=sc)
            s.update(self.dev["A_val"], self.dev["l"], self.dev["u"])
        self.stream.synchronize()
        return dict(exit_code=code, qp_exit_code=qp_code, rounds=rounds, x=self.x, total_qp_solves=total)

    def close(self):
        self.solver.close()
        self.scene.close()


SEGMENTS = 10  # reference src/gomp-solver.h:11


def shortest_horizon_chained(cfg, first_id, count, max_rounds=MAX_ITERATIONS, settings_factory=None):
    """GOMPSolver::run(start, goal) with the reference's OWN semantics (src/gomp-solver.h:38-55) for a batch: the horizons
    W*i/SEGMENTS, i = SEGMENTS..1, are solved one after the other (one batch over all problems per horizon), and every horizon
    is warm-started from -- and linearised around -- the slices of the last kOptimal trajectory exactly as :43-44 takes them:
    the first w*N entries, then the w*N entries that follow them (for a shrinking horizon those are NOT the velocities of the
    previous solution but the tail of its positions and the head of its velocities; kept verbatim, as the facade's
    GOMPSolver<N>::run does).  A horizon's run reports kOptimal or kUnknown only (:66-90), so no horizon is skipped (:46-48).

    Returns the dict of `shortest_horizon` plus codes_per_horizon [H, count]."""
    import copy
    W, N = cfg.waypoints, cfg.n_dim
    pat_full = _problem.Pattern(cfg)
    last = list(_problem.generate(pat_full, first_id, count)["x0"])   # calcWarmStart: [linspace | 0] at the full horizon
    code = np.full(count, K_UNKNOWN, dtype=np.int32)
    chosen = np.zeros(count, dtype=np.int32)
    horizons, codes, total = [], [], 0
    for i in range(SEGMENTS, 0, -1):
        w = W * i // SEGMENTS
        if w < 4:
            break                                                      # (the reference asserts waypoints >= 4)
        c = copy.copy(cfg)
        c.waypoints = w
        pat = _problem.Pattern(c)
        warm = np.stack([np.concatenate([v[:w * N], v[w * N:2 * w * N]]) for v in last])
        drv = BatchSQP(pat, first_id, count, settings=settings_factory() if settings_factory else None, warm_start=warm)
        out = drv.run(max_rounds=max_rounds)
        x = out["x"].cpu().numpy()
        drv.close()
        total += out["total_qp_solves"]
        horizons.append(w)
        codes.append(out["exit_code"].copy())
        for b in np.flatnonzero(out["exit_code"] == K_OPTIMAL):
            code[b], chosen[b], last[b] = K_OPTIMAL, w, x[b].copy()
    xs = []
    for b in range(count):
        v = last[b].copy()
        v[v.size // 2:] /= cfg.time_step                               # :52 velocities per second
        xs.append(v)
    return dict(exit_code=code, waypoints=chosen, x=xs, horizons=horizons, qp_solves=total, codes_per_horizon=np.array(codes))


def shortest_horizon(cfg, first_id, count, max_rounds=MAX_ITERATIONS, settings_factory=None):
    """GOMPSolver::run(start, goal) -- the time-shrink outer loop of reference src/gomp-solver.h:38-55 -- for a batch, with every
    horizon independent of the others (see `shortest_horizon_chained` for the reference's chained warm starts).

    The reference solves the horizons W*i/SEGMENTS, i = SEGMENTS..1, one after the other, keeps the last kOptimal
    trajectory and stops at the first horizon whose QP fails.  Given a warm start the horizons are independent, so
    here every horizon is one batch over all problems (BatchSQP; long ones run on the streamed back end) and the
    reference's selection rule is applied per problem afterwards.  Difference to the reference, on purpose: each
    horizon starts from its own straight-line warm start instead of the previous horizon's sliced solution
    (gomp-solver.h:43-44), which is what makes them independent.

    Returns dict(exit_code [count], waypoints [count] (horizon of the returned trajectory, 0 = none),
    x (list of per-problem numpy trajectories, velocities per second as the reference returns them), horizons,
    qp_solves)."""
    import copy
    W = cfg.waypoints
    horizons = [W * i // SEGMENTS for i in range(SEGMENTS, 0, -1) if W * i // SEGMENTS >= 4]
    per_h = []
    total = 0
    for w in horizons:
        c = copy.copy(cfg)
        c.waypoints = w
        pat = _problem.Pattern(c)
        drv = BatchSQP(pat, first_id, count, settings=settings_factory() if settings_factory else None)
        out = drv.run(max_rounds=max_rounds)
        per_h.append(dict(code=out["exit_code"].copy(), x=out["x"].cpu().numpy(), n=pat.n))
        total += out["total_qp_solves"]
        drv.close()
    code, pick = select_horizon([res["code"] for res in per_h])
    chosen = np.zeros(count, dtype=np.int32)
    xs = [None] * count
    for b in range(count):
        if pick[b] >= 0:
            res = per_h[pick[b]]
            chosen[b] = horizons[pick[b]]
            x = res["x"][b].copy()
            x[res["n"] // 2:] /= cfg.time_step     # gomp-solver.h:52-54 velocities per second
            xs[b] = x
    return dict(exit_code=code, waypoints=chosen, x=xs, horizons=horizons, qp_solves=total)


def select_horizon(codes_per_horizon):
    """The selection rule of reference src/gomp-solver.h:42-51 applied to per-horizon exit codes (longest horizon first):
    walk the horizons, stop at the first code that is neither kOptimal nor kUnknown, remember the last kOptimal one.
    Returns (exit_code [count], index of the chosen horizon or -1 [count])."""
    codes = np.asarray(codes_per_horizon)
    count = codes.shape[1]
    out = np.full(count, K_UNKNOWN, dtype=np.int32)
    pick = np.full(count, -1, dtype=np.int32)
    for b in range(count):
        for k in range(codes.shape[0]):
            ec = codes[k, b]
            if ec != K_OPTIMAL and ec != K_UNKNOWN:
                break
            if ec == K_OPTIMAL:
                out[b], pick[b] = K_OPTIMAL, k
    return out, pick
