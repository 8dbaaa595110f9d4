#!/usr/bin/env python
"""bench.py -- GOMP QPs solved per second on B200 (BASELINE.json metric), one JSON line.

A "step" is one pass of the hot path over one batch: for every QP of the batch a fresh OSQP
problem (cold start: Ruiz scaling, rho classification, KKT factorisation), the straight-line
primal warm start, and the ADMM solve to OSQP's default tolerances -- exactly what the
reference does per QP at src/gomp-solver.h:61-71 (QPSolver ctor + setWarmStart + solve).

  value    whole-job QPs/s with inputs already resident in HBM (device pointers borrowed
           in place), device-timed with CUDA events on the launching stream, max over ranks
  e2e      the same work through the C ABI with HOST (pinned) buffers: H2D of A_val, l, u, x0 and
           D2H of x, status, iterations inside the timed region
  roofline achieved vs measured shared-memory bandwidth (the bound SURVEY.md 8d names for this
           kernel: min(FP64 pipe, shared memory)); FP64 and HBM fractions alongside
  cpu_baseline  the oracle (CPU restatement of OSQP, kind "port") timed on the box's host cores on
           a bounded sample of the same workload, fed as ONE continuous work queue (no thread idles
           at a pass boundary), the same problem ids as the GPU arm
  extra    configs[2] (warm SQP x10 on the device), configs[3] (H = 400, batch 1 and 148) and
           configs[4] (1 048 576 problems sharded over the ranks, assembled on the device from start /
           goal: strong scaling), each bounded to a few seconds and checked against the oracle

--impl reference times that CPU oracle on all host threads instead (the real reference cannot be
built offline: it needs osqp-cpp/OSQP/Eigen via CMake FetchContent).

Workload: BASELINE.json configs[1] -- 4096 synthetic 7-DoF GOMP QPs, horizon H=50, per GPU
(weak scaling across ranks; problems are independent, NCCL gathers every step's results on rank 0
inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

BATCH_PER_GPU = 4096


def rank_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# batches in flight per GPU: consecutive steps rotate over PIPE solver handles / streams (measured on B200:
# 2 -> 107 k, 4 -> 130 k, 6 -> 142 k, 8 -> 142 k QP/s with the resident kernel; final build, two QPs per SM, 256 steps:
# 12 -> 241.5 k, 16 -> 241.4 k, 24 -> 241.6 k -- more only adds memory); every step still is a full cold setup + warm
# start + solve.  The timed region ends with the drain of the last batches (their max_iter QPs take 70 ms each), so
# few steps under-report the steady state: 32 steps 197 k, 128 steps 213 k on one build; 64 -> 233 k, 256 -> 241 k now.
PIPE = int(os.environ.get("GOMP_BENCH_PIPE", "12"))
# SMs shared between the batches in flight (gomp_batch_set_max_ctas, include/gomp_cuda.h): a batch's persistent CTAs hold their
# SMs until its work queue is empty, so with one CTA per SM the batches run one after the other and each ends with its two
# max_iter QPs (~45 ms) alone on their SMs.  LANES batches co-resident on num_sms / LANES CTAs each last longer than their
# longest solve, and the stragglers of one overlap with the bulk of the others.  Measured on B200 (20 / 96 steps, `value`):
# LANES 1: 286 k / 315 k, 2: 292 k / 327 k, 3: 297 k / 333 k, 4: 307 k / 337 k, 5: 295 k / 338 k, 6: 303 k / 341 k, 8: 297 k /
# 321 k, 12: 291 k / 260 k QP/s (a handle's next batch waits for its previous one: lanes need PIPE > LANES handles).
# Host-buffer leg (`e2e`): its batches become ready one H2D copy (6 ms) apart and the host submits a batch only after it has
# collected the one PIPE steps before, so the pipeline knows its depth: the first batches (fewer than LANES in flight) and the
# last ones (fewer than LANES to come) get the wider share -- 20 steps: 268 k without, 281 k with that ramp.  The device-
# resident leg enqueues every step before the first one has finished; there the ramp only delays the lanes (307 k -> 301 k).
LANES = int(os.environ.get("GOMP_BENCH_LANES", "4"))
GATHER = os.environ.get("GOMP_BENCH_GATHER", "1") != "0"  # A/B switch: the per-step NCCL gather of the multi-GPU runs


def lane_ctas(k, nsteps, num_sms, ramp):
    """CTA cap of step k of a run of nsteps pipelined steps (0 = no cap); ramp bit 0: first batches, bit 1: last batches."""
    lanes = max(1, min(LANES, k + 1 if ramp & 1 else LANES, nsteps - k if ramp & 2 else LANES))
    return -(-num_sms // lanes) if lanes > 1 else 0


def workload(pkg):
    P = pkg.problem
    cfg = P.default_config()
    pat = P.Pattern(cfg)
    desc = {"workload": "BASELINE.json configs[1]: batch of 4096 synthetic 7-DoF GOMP QPs per GPU, H=50, pos/vel/acc bounds + "
                        "1 gripper ball + 1 horizontal-line obstacle; cold setup + warm start + solve, OSQP default settings "
                        "(eps 1e-3, adaptive_rho_interval pinned to 25)",
            "n_dim": 7, "waypoints": 50, "n": pat.n, "m": pat.m, "nnzA": pat.nnzA, "batch_per_gpu": BATCH_PER_GPU,
            "l2_policy": "inputs larger than L2 (305 MB of A_val/l/u per step vs 126 MB L2)",
            "batches_in_flight": PIPE, "sm_lanes": LANES}
    return cfg, pat, desc


def host_cores():
    """(logical CPUs usable by this process, physical cores behind them)."""
    logical = len(os.sched_getaffinity(0))
    phys = None
    try:
        import psutil
        phys = psutil.cpu_count(logical=False)
    except Exception:
        pass
    return logical, phys


def time_oracle(pkg, O, pat, count, first_id, nthreads, passes=1, data=None):
    """`passes` passes over problems first_id .. first_id + count - 1 as one continuous work queue (oracle/osqp_ref.c
    oref_solve_batch_passes): every work item is a full cold setup + warm start + solve of one QP."""
    d = data if data is not None else pkg.problem.generate(pat, first_id, count)
    t0 = time.perf_counter()
    r = O.solve_batch(pat.n, pat.m, pat.P_scipy(), (pat.A_colptr, pat.A_rowidx), d["A_val"], d["l"], d["u"], d["x0"], nthreads=nthreads,
                      passes=passes)
    dt = time.perf_counter() - t0
    return count * passes / dt, dt, r


CPU_NOTE = ("one QP per thread from one continuous work queue over all passes (dynamic, no pass boundary), KKT ordering computed once per "
            "call (favours the CPU); oracle = CPU restatement of OSQP 0.6, not genuine OSQP/qdldl")


def run_reference(args):
    """CPU arm: the oracle on all host threads.  A step is one pass over the GPU arm's 4096 problems (ids 0..4095); all timed
    steps are ONE call, i.e. one continuous queue of steps x 4096 work items, so no thread idles while a long QP of the
    previous step finishes -- the CPU counterpart of the GPU arm's batches in flight."""
    rank, local_rank, world = rank_env()
    if rank != 0:
        return
    pkg = entry.load_package()
    O = entry.load_oracle()
    cfg, pat, desc = workload(pkg)
    logical, phys = host_cores()
    cores = min(O.max_threads(), logical)
    sample = BATCH_PER_GPU
    d = pkg.problem.generate(pat, 0, sample)
    # bounded: at ~0.3 ms of CPU time per QP and core a pass is 4096 / (cores * 3k) s; cap the passes so the run ends in minutes
    warm = max(1, min(args.warmup, 2))
    steps = max(1, min(args.steps, 24))
    time_oracle(pkg, O, pat, sample, 0, cores, passes=warm, data=d)
    value, total, _ = time_oracle(pkg, O, pat, sample, 0, cores, passes=steps, data=d)
    line = {"impl": "reference", "metric": "GOMP QPs solved/sec", "value": value, "unit": "QP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": desc,
            "cpu_baseline": {"value": value, "unit": "QP/s", "cores": cores, "logical_cpus": logical, "physical_cores": phys, "kind": "port",
                             "sample": f"{steps} timed passes (after {warm} untimed) over the GPU arm's {sample} QPs (problem ids 0..{sample - 1}), "
                                       f"{total:.2f} s wall; " + CPU_NOTE},
            "e2e": {"value": value, "unit": "QP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def bind_near_gpu(index):
    """Multi-rank runs: keep this rank's threads (and therefore the pages of its pinned host buffers) on the NUMA node of its
    GPU -- eight ranks streaming 328 MB per step each out of one node's memory was the e2e limiter of round 1."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True,
                             text=True, timeout=20).stdout.strip()
        bus = out.lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        cpus = open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
            return cpus
    except Exception:
        pass
    return None


def strong_scaling(pkg, pat, rank, world, local_rank, torch, dist):
    """BASELINE.json configs[4]: TOTAL problems of configs[1] sharded over the ranks (contiguous id ranges), strong scaling.
    Per chunk the host hands over start / goal only (16 N bytes per QP); A_val, l, u and the warm start are assembled on
    the device (gomp_device_assemble), solved, and x / exit code / iterations return to pinned host memory.  One pass."""
    P, S = pkg.problem, pkg.solver
    total = int(os.environ.get("GOMP_BENCH_STRONG_PROBLEMS", str(1 << 20)))
    if total <= 0:
        return None
    chunk = int(os.environ.get("GOMP_BENCH_STRONG_CHUNK", "16384"))
    nh = int(os.environ.get("GOMP_BENCH_STRONG_HANDLES", "4"))
    lanes = int(os.environ.get("GOMP_BENCH_STRONG_LANES", "2"))   # chunks co-resident on the SMs (gomp_batch_set_max_ctas)
    lo, hi = pkg.sharding.shard_range(total, rank, world)   # contiguous id ranges, sizes differ by at most one
    count = hi - lo
    chunk = max(1, min(chunk, count))
    nchunks = (count + chunk - 1) // chunk
    N = pat.cfg.n_dim
    t_gen = time.perf_counter()
    sg_host = torch.empty((max(count, 1), 2 * N), dtype=torch.float64).pin_memory()
    if count:
        P.start_goal(pat, lo, count, out=sg_host.numpy()[:count])
    t_gen = time.perf_counter() - t_gen
    streams = [torch.cuda.Stream() for _ in range(nh)]
    solvers, scenes, bufs, outs = [], [], [], []
    tmpl = P.generate(pat, 0, 1)
    for k in range(nh):
        st = S.default_settings(borrow_inputs=1, stream=streams[k].cuda_stream, device=local_rank)
        solvers.append(S.BatchQPSolver(pat.n, pat.m, pat.P_colptr, pat.P_rowidx, pat.P_val, pat.A_colptr, pat.A_rowidx, batch=chunk,
                                       settings=st, var_order=pat.var_order, block_size=pat.block_size))
        if lanes > 1:
            solvers[-1].set_max_ctas(-(-solvers[-1].info()["num_sms"] // lanes))
        sc = P.DeviceScene(pat, None, chunk, device=local_rank)
        sc.set_template(tmpl)
        scenes.append(sc)
        bufs.append({k_: torch.empty(shape, dtype=torch.float64, device="cuda") for k_, shape in
                     dict(A_val=(chunk, pat.nnzA), l=(chunk, pat.m), u=(chunk, pat.m), x0=(chunk, pat.n), sg=(chunk, 2 * N)).items()})
        outs.append(dict(x=torch.empty((chunk, pat.n), dtype=torch.float64).pin_memory(),
                         ints=torch.empty((2, chunk), dtype=torch.int32).pin_memory()))
    hist = np.zeros(11, dtype=np.int64)
    iters_sum = 0

    def submit(c):
        k = c % nh
        a, b = c * chunk, min(count, (c + 1) * chunk)
        if b - a < chunk:   # ragged last chunk: pad with the chunk's first problem (results of the padding are dropped)
            sg_host_chunk = torch.cat([sg_host[a:b], sg_host[a:a + 1].expand(chunk - (b - a), 2 * N)]).pin_memory()
        else:
            sg_host_chunk = sg_host[a:b]
        with torch.cuda.stream(streams[k]):
            bufs[k]["sg"].copy_(sg_host_chunk, non_blocking=True)
        bufs[k]["keep"] = sg_host_chunk
        scenes[k].assemble(bufs[k]["sg"], bufs[k]["A_val"], bufs[k]["l"], bufs[k]["u"], bufs[k]["x0"], stream=streams[k].cuda_stream)
        solvers[k].setup(bufs[k]["A_val"], bufs[k]["l"], bufs[k]["u"])
        solvers[k].set_warm_start(bufs[k]["x0"])
        solvers[k].solve_async()
        solvers[k].get_async(x=outs[k]["x"], exit_code=outs[k]["ints"][0], iters=outs[k]["ints"][1])

    def collect(c):
        nonlocal iters_sum
        k = c % nh
        solvers[k].sync()
        valid = min(count, (c + 1) * chunk) - c * chunk
        ints = outs[k]["ints"].numpy()
        hist[:] += np.bincount(ints[0][:valid], minlength=11)[:11]
        iters_sum += int(ints[1][:valid].sum())

    # untimed: one small chunk through every handle (kernel load, workspace sizing)
    if nchunks:
        for c in range(min(nh, nchunks)):
            submit(c)
        for c in range(min(nh, nchunks)):
            collect(c)
    hist[:] = 0
    iters_sum = 0
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    ev0.record(main)
    for s_ in streams:
        s_.wait_event(ev0)
    t0 = time.perf_counter()
    for c in range(nchunks):
        if c >= nh:
            collect(c - nh)
        submit(c)
    for c in range(max(0, nchunks - nh), nchunks):
        collect(c)
    for s_ in streams:
        e = torch.cuda.Event()
        e.record(s_)
        main.wait_event(e)
    ev1.record(main)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    for sv in solvers:
        sv.close()
    for sc in scenes:
        sc.close()
    stats = torch.tensor([dev_ms, wall * 1e3, t_gen * 1e3, float(iters_sum)] + hist.astype(np.float64).tolist(), dtype=torch.float64, device="cuda")
    allr = [stats]
    if world > 1:
        dist.barrier()
        allr = [torch.empty_like(stats) for _ in range(world)]
        dist.all_gather(allr, stats)
    if rank != 0:
        return None
    rows = np.stack([t.cpu().numpy() for t in allr])
    dev_max, wall_max = float(rows[:, 0].max()), float(rows[:, 1].max())
    return {"workload": f"BASELINE.json configs[4]: {total} problems of configs[1] sharded over {world} GPU(s), contiguous id ranges, "
                        f"chunks of {chunk} through {nh} handles per GPU; A_val / l / u / x0 assembled on the device from start / goal",
            "problems": total, "n_gpus": world, "scaling": "strong",
            "device_ms_max_over_ranks": dev_max, "wall_ms_max_over_ranks": wall_max,
            "qps_device_timed": total / (dev_max * 1e-3), "qps_wall": total / (wall_max * 1e-3),
            "rank_device_ms": [round(float(v), 1) for v in rows[:, 0]],
            "h2d_bytes_per_qp": 16 * N, "d2h_bytes_per_qp": 8 * pat.n + 8,
            "host_start_goal_generation_ms_untimed": float(rows[:, 2].max()),
            "exit_code_hist": rows[:, 4:].sum(axis=0).astype(np.int64).tolist(), "iters_mean": float(rows[:, 3].sum() / total)}


def extra_sqp_sequence(pkg, O, torch, batch=4096, rounds=10, cpu_sample=64):
    """BASELINE.json configs[2]: `batch` QPs x `rounds` warm-started SQP rounds (withObstacles(x) -> update -> solve, reference
    src/gomp-solver.h:70-88 without the early exit) with the data plane on the device; the oracle runs a sample of it."""
    import concurrent.futures as cf
    P, S = pkg.problem, pkg.solver
    pat = P.Pattern(P.default_config())
    d = P.generate(pat, 0, batch)

    def device_path():
        stream = torch.cuda.Stream()
        dev = {k: torch.from_numpy(d[k]).cuda() for k in ("A_val", "l", "u", "x0")}
        st = S.default_settings(borrow_inputs=1, stream=stream.cuda_stream)
        s = S.BatchQPSolver(pat.n, pat.m, pat.P_colptr, pat.P_rowidx, pat.P_val, pat.A_colptr, pat.A_rowidx, dev["A_val"], dev["l"], dev["u"],
                            batch=batch, settings=st, var_order=pat.var_order, block_size=pat.block_size)
        sc = P.DeviceScene(pat, 0, batch)
        x = torch.empty((batch, pat.n), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s.set_warm_start(dev["x0"])
        hist = []
        for r in range(rounds):
            s.solve_async()
            got = s.get(x=x)
            hist.append((got["exit_code"].copy(), got["iters"].copy()))
            if r + 1 < rounds:
                sc.relinearize(x, dev["A_val"], dev["l"], dev["u"], x_fallback=dev["x0"], stream=stream.cuda_stream)
                s.update(dev["A_val"], dev["l"], dev["u"])
        stream.synchronize()
        dt = time.perf_counter() - t0
        linsys = s.info()["linsys"]
        s.close()
        sc.close()
        return dt, hist, linsys

    device_path()
    t_dev, hist, linsys = device_path()
    Ps = pat.P_scipy()
    nthreads = min(O.max_threads(), len(os.sched_getaffinity(0)))

    def one(b):
        ref = O.Solver(Ps, pat.A_scipy(d["A_val"][b]), d["l"][b], d["u"][b])
        ref.warm_start_x(d["x0"][b])
        out = []
        for r in range(rounds):
            ref.solve()
            out.append((ref.exit_code, ref.iter))
            if r + 1 < rounds:
                xs = np.where(np.isnan(ref.x), d["x0"][b], ref.x)
                upd = P.relinearize(pat, b, xs[None, :], nthreads=1)
                ref.update(upd["A_val"][0], upd["l"][0], upd["u"][0])
        return out
    n_cpu = min(cpu_sample, batch)
    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(nthreads) as ex:
        cpu = list(ex.map(one, range(n_cpu)))
    t_cpu = time.perf_counter() - t0
    # the host facade re-linearises for the oracle, the device kernel for the GPU (they differ through sin / cos rounding), so the
    # first round is the bit-comparable one; later rounds are reported as the share of (problem, round) pairs that agree
    code_eq = np.mean([[hist[r][0][b] == cpu[b][r][0] for r in range(rounds)] for b in range(n_cpu)])
    iter_eq = np.mean([[hist[r][1][b] == cpu[b][r][1] for r in range(rounds)] for b in range(n_cpu)])
    gpu_rate, cpu_rate = batch * rounds / t_dev, n_cpu * rounds / t_cpu
    return {"workload": f"BASELINE.json configs[2]: {batch} QPs x {rounds} warm-started SQP rounds (N=7, H=50), data plane on the device",
            "qp_solves_per_s": gpu_rate, "seconds": t_dev, "linsys": linsys,
            "cpu_oracle": {"sample": n_cpu, "threads": nthreads, "qp_solves_per_s": cpu_rate, "seconds": t_cpu}, "vs_cpu": gpu_rate / cpu_rate,
            "round0_status_equal": bool(all(hist[0][0][b] == cpu[b][0][0] for b in range(n_cpu))),
            "round0_iterations_equal": bool(all(hist[0][1][b] == cpu[b][0][1] for b in range(n_cpu))),
            "all_rounds_status_agreement": float(code_eq), "all_rounds_iterations_agreement": float(iter_eq)}


def extra_long_horizon(pkg, O, W=400, batches=(1, 32, 148), cpu_sample=16, name="BASELINE.json configs[3]", **cfg_kw):
    """BASELINE.json configs[3]: N=7, H=400 on the streamed back end, batch 1 and one QP per SM, beside the oracle (also used for
    configs[0], the reference example's shape: N=6, W=802, two balls, no obstacle lines)."""
    P, S = pkg.problem, pkg.solver
    pat = P.Pattern(P.default_config(waypoints=W, **cfg_kw))
    nmax = max(batches)
    d = P.generate(pat, 0, nmax)
    out = {"workload": f"{name}: N={pat.cfg.n_dim}, H={W} (n={pat.n}, m={pat.m}, nnzA={pat.nnzA})"}
    results = None
    for batch in batches:
        sub = {k: d[k][:batch] for k in ("A_val", "l", "u", "x0")}
        s = S.BatchQPSolver(pat.n, pat.m, pat.P_colptr, pat.P_rowidx, pat.P_val, pat.A_colptr, pat.A_rowidx, sub["A_val"], sub["l"], sub["u"],
                            var_order=pat.var_order, block_size=pat.block_size)
        best = 1e30
        for _ in range(3):
            s.setup(sub["A_val"], sub["l"], sub["u"])
            s.set_warm_start(sub["x0"])
            results = s.solve()
            best = min(best, s.info()["last_solve_ms"])
        out[f"ms_b{batch}"] = best
        out[f"qps_b{batch}"] = batch / best * 1e3
        out[f"ctas_per_qp_b{batch}"] = s.info()["grid"] // min(batch, s.info()["grid"])  # > 1: a thread-block cluster per QP
        out[f"iters_max_b{batch}"] = int(results["iters"].max())
        out["linsys"] = s.info()["linsys"]
        s.close()
    n_cpu = min(cpu_sample, nmax)
    nthreads = min(n_cpu, O.max_threads(), len(os.sched_getaffinity(0)))
    t0 = time.perf_counter()
    ref = O.solve_batch(pat.n, pat.m, pat.P_scipy(), (pat.A_colptr, pat.A_rowidx), d["A_val"][:n_cpu], d["l"][:n_cpu], d["u"][:n_cpu], d["x0"][:n_cpu],
                        nthreads=nthreads)
    t_cpu = time.perf_counter() - t0
    t1 = time.perf_counter()
    O.solve_batch(pat.n, pat.m, pat.P_scipy(), (pat.A_colptr, pat.A_rowidx), d["A_val"][:1], d["l"][:1], d["u"][:1], d["x0"][:1], nthreads=1)
    t_cpu1 = time.perf_counter() - t1
    out["cpu_oracle"] = {"sample": n_cpu, "threads": nthreads, "qps": n_cpu / t_cpu, "ms_b1": t_cpu1 * 1e3}
    out["vs_cpu_b1"] = t_cpu1 * 1e3 / out[f"ms_b{batches[0]}"]
    out[f"vs_cpu_b{nmax}"] = out[f"qps_b{nmax}"] / (n_cpu / t_cpu)
    out["status_equal"] = bool((results["exit_code"][:n_cpu] == ref["exit_code"]).all())
    out["iterations_equal"] = bool((results["iters"][:n_cpu] == ref["iters"]).all())
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = rank_env()
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner with printf when the communicator is created: keep it off stdout, which carries
        # the one JSON line (fd 1 points at stderr until the communicator exists)
        # The solver's persistent CTAs own every SM's register file, so a collective's kernel can only start on an SM a solver CTA
        # has just left, and holds it while it runs: keep NCCL's point-to-point traffic (the gather is send / recv) on few
        # channels = few CTAs.  161 MB per step into rank 0 at 8 GPUs needs ~10 GB/s, a fraction of one NVLink channel pair.
        os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "2")
        os.environ.setdefault("NCCL_MIN_P2P_NCHANNELS", "1")
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    pkg = entry.load_package()
    S = pkg.solver
    if S.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: libgomp_cuda has no CPU fallback")
    cfg, pat, desc = workload(pkg)
    B = BATCH_PER_GPU
    first_id = rank * B
    full_affinity = os.sched_getaffinity(0)
    numa_cpus = bind_near_gpu(local_rank) if world > 1 else None

    # ---- synthetic inputs: generated on the host into pinned memory, then made resident in HBM
    host = {k: torch.empty(shape, dtype=torch.float64).pin_memory() for k, shape in
            dict(A_val=(B, pat.nnzA), l=(B, pat.m), u=(B, pat.m), x0=(B, pat.n), start_goal=(B, 2 * cfg.n_dim)).items()}
    pkg.problem.generate(pat, first_id, B, out={k: v.numpy() for k, v in host.items()})
    dev = {k: host[k].cuda() for k in ("A_val", "l", "u", "x0")}
    # PIPE batches in flight: consecutive steps rotate over PIPE handles on PIPE streams, so the few
    # QPs of a batch that need thousands of ADMM iterations (max_iter = 4000) overlap with the next batch
    # instead of leaving 146 SMs idle.  Every step still is a full cold setup + warm start + solve.
    streams = [torch.cuda.Stream() for _ in range(PIPE)]
    main = torch.cuda.current_stream()

    def make(borrow, stream):
        st = S.default_settings(borrow_inputs=1 if borrow else 0, stream=stream.cuda_stream, device=local_rank,
                                linsys=int(os.environ.get("GOMP_BENCH_LINSYS", "0")),  # 0 = AUTO (the measured configuration)
                                # several independent batches are kept in flight: let AUTO take the two-QPs-per-SM mode
                                prefer_throughput=int(os.environ.get("GOMP_BENCH_THROUGHPUT", "1")))
        return S.BatchQPSolver(pat.n, pat.m, pat.P_colptr, pat.P_rowidx, pat.P_val, pat.A_colptr, pat.A_rowidx,
                               batch=B, settings=st, var_order=pat.var_order, block_size=pat.block_size)

    solvers = [make(True, streams[k]) for k in range(PIPE)]
    solver = solvers[0]
    x_dev = torch.empty((B, pat.n), dtype=torch.float64, device="cuda")
    code = np.empty(B, dtype=np.int32)
    iters = np.empty(B, dtype=np.int32)
    rup = np.empty(B, dtype=np.int32)

    num_sms = solver.info()["num_sms"]

    def step_resident(k=0, nsteps=1):
        sv = solvers[k % PIPE]
        sv.set_max_ctas(lane_ctas(k, nsteps, num_sms, int(os.environ.get("GOMP_BENCH_LANE_RAMP", "0"))) if nsteps > 1 else 0)
        sv.setup(dev["A_val"], dev["l"], dev["u"])
        sv.set_warm_start(dev["x0"])
        sv.solve_async()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N > 1: the problems are independent, the only collective is the gather of every step's results {x, exit code,
    # iterations} on rank 0 (NCCL over NVLink), queued behind the step's solve on the step's stream
    if world > 1:
        xg = [torch.empty((B, pat.n), dtype=torch.float64, device="cuda") for _ in range(PIPE)]
        ig = [torch.empty((2, B), dtype=torch.int32, device="cuda") for _ in range(PIPE)]
        xg_all = [[torch.empty_like(xg[0]) for _ in range(world)] for _ in range(PIPE)] if rank == 0 else None
        ig_all = [[torch.empty_like(ig[0]) for _ in range(world)] for _ in range(PIPE)] if rank == 0 else None

    def gather_step(k):
        j = k % PIPE
        solvers[j].get_async(x=xg[j], exit_code=ig[j][0], iters=ig[j][1])
        with torch.cuda.stream(streams[j]):
            dist.gather(xg[j], xg_all[j] if rank == 0 else None, dst=0)
            dist.gather(ig[j], ig_all[j] if rank == 0 else None, dst=0)

    def run_steps(nsteps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(main)
        for st_ in streams:
            st_.wait_event(ev0)
        for k in range(nsteps):
            step_resident(k, nsteps)
            if world > 1 and GATHER:
                gather_step(k)
        for st_ in streams:
            e = torch.cuda.Event()
            e.record(st_)
            main.wait_event(e)
        ev1.record(main)
        return ev0, ev1

    run_steps(max(args.warmup, PIPE))  # every handle has run once before the timed region (first use sizes its workspace)
    barrier()
    launches0 = sum(sv.info()["kernel_launches"] for sv in solvers)
    sampler = ClockSampler(local_rank)
    sampler.start()
    kernel_ms = []
    barrier()
    ev0, ev1 = run_steps(args.steps)  # N > 1: includes the NCCL gather of every step's results
    barrier()
    clocks = sampler.stop()
    total_ms = ev0.elapsed_time(ev1)
    gathered_ok = None
    if world > 1 and rank == 0 and GATHER:
        # the last step's gathered block of every rank against this rank's own results (all ranks solve their own ids; rank 0's
        # block must be its own x bit for bit, the others must carry valid exit codes)
        j = (args.steps - 1) % PIPE
        own = torch.empty_like(xg[j])
        solvers[j].get(x=own)
        codes = torch.stack([t[0] for t in ig_all[j]]).cpu().numpy()
        gathered_ok = bool(torch.equal(torch.nan_to_num(own), torch.nan_to_num(xg_all[j][0]))) and bool(((codes >= 0) & (codes <= 10)).all())
    r = solver.get(x=x_dev, exit_code=code, iters=iters, rho_updates=rup)
    info = solver.info()
    launches = sum(sv.info()["kernel_launches"] for sv in solvers) - launches0
    # per-launch duration of the dominant (only) kernel, CUDA events on its stream, live, one batch alone
    for _ in range(3):
        step_resident(0)
        solver.sync()
        kernel_ms.append(solver.info()["last_solve_ms"])
    kernel_avg_ms = float(np.mean(kernel_ms))

    rank_ms, rank_mhz = [total_ms], [clocks.get("sm_mhz")]
    if world > 1:
        # every rank's device time and median SM clock (diagnostics), then the maximum over ranks (the contract's time)
        mine = torch.tensor([total_ms, float(clocks.get("sm_mhz") or 0.0)], dtype=torch.float64, device="cuda")
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        rank_ms = [float(t[0].item()) for t in allr]
        rank_mhz = [float(t[1].item()) for t in allr]
        total_ms = max(rank_ms)
    value = world * B * args.steps / (total_ms * 1e-3)

    # ---- end to end through the C ABI with host buffers (copies inside the timed region)
    # The same rotation over PIPE handles as above, but every call takes HOST pointers: each step copies its
    # A_val, l, u, x0 from pinned host memory (H2D on the handle's stream), solves, and its x / status /
    # iteration counts are read back to the host before the handle is reused (D2H).  Wall clock, host side.
    for sv in solvers:
        sv.close()
    e2e_solvers = [make(False, streams[k]) for k in range(PIPE)]
    x_hosts = [torch.empty((B, pat.n), dtype=torch.float64).pin_memory() for _ in range(PIPE)]

    def e2e_submit(k, nsteps):
        sv = e2e_solvers[k % PIPE]
        sv.set_max_ctas(lane_ctas(k, nsteps, num_sms, int(os.environ.get("GOMP_BENCH_E2E_LANE_RAMP", "3"))))
        sv.setup(host["A_val"], host["l"], host["u"])
        sv.set_warm_start(host["x0"])
        sv.solve_async()

    def e2e_collect(k):
        e2e_solvers[k % PIPE].get(x=x_hosts[k % PIPE].numpy(), exit_code=code, iters=iters, rho_updates=rup)

    def e2e_run(nsteps):
        for k in range(nsteps):
            if k >= PIPE:
                e2e_collect(k - PIPE)
            e2e_submit(k, nsteps)
        for k in range(max(0, nsteps - PIPE), nsteps):
            e2e_collect(k)

    e2e_run(max(PIPE, args.warmup))
    barrier()
    e2e_steps = max(PIPE, args.steps)
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * B * e2e_steps / (e2e_ms * 1e-3)
    h2d = 8 * B * (pat.nnzA + 2 * pat.m + pat.n)
    d2h = B * (8 * pat.n + 12)

    for sv in e2e_solvers:
        sv.close()
    del e2e_solvers, x_hosts, dev
    torch.cuda.empty_cache()
    strong = None
    if os.environ.get("GOMP_BENCH_EXTRA", "1") != "0":
        strong = strong_scaling(pkg, pat, rank, world, local_rank, torch, dist)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    os.sched_setaffinity(0, full_affinity)

    # ---- roofline of the ADMM kernel (work model: DESIGN.md section 5 / gomp_batch_info)
    it = r["iters"].astype(np.float64)
    nfact = 1.0 + r["rho_updates"].astype(np.float64)
    checks = np.ceil(it / 25.0)
    smem_bytes = (info["smem_bytes_per_scaling"] + nfact * info["smem_bytes_per_factor"] + it * info["smem_bytes_per_iter"]
                  + checks * info["smem_bytes_per_check"]).sum()
    flops = (info["flops_per_scaling"] + nfact * info["flops_per_factor"] + it * info["flops_per_iter"] + checks * info["flops_per_check"]).sum()
    hbm_bytes = B * (8 * (pat.nnzA + 2 * pat.m + 2 * pat.n) + 8 * (pat.n + pat.m))  # inputs + x0 + outputs
    fp64_peak, smem_peak = S.measure_peaks()
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback"
    if os.path.exists(peaks_path):
        hbm_peak, hbm_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    # One launch = one step's 4096 QPs.  Launches of consecutive steps overlap on the GPU (PIPE handles), so the time a launch
    # costs in steady state is the step time of the timed region -- `frac` is formed with it and is reproducible from this
    # line's own fields (achieved = algorithmic_smem_bytes / ms_per_step).  `frac_solo` uses the duration of one launch
    # running alone (CUDA events on its stream), which is dominated by the tail of its max_iter QPs.
    step_s = total_ms / args.steps * 1e-3
    ksec = kernel_avg_ms * 1e-3
    linsys = int(info["linsys"])
    kernel_name = {4: "gomp_duo_kernel<9,4,14>", 2: "gomp_admm_stream_kernel<14,256,true>"}.get(linsys, "gomp_admm_kernel<4,2,14>")
    # DRAM bytes and pipe utilisation of one launch from the committed ncu --set full capture (same kernel, same workload)
    traffic, traffic_src, ncu = None, None, {}
    tpath = os.path.join(ROOT, "profiles", {4: "r02_ncu_duo.json", 2: "r01_ncu_traffic_stream.json"}.get(linsys, "r01_ncu_traffic.json"))
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) * B / tj["qps_per_launch"]
        traffic_src = "dram__bytes_read.sum + dram__bytes_write.sum per launch, " + tj["capture"]
        ncu = {k: tj[k] for k in ("warp_instructions_per_qp", "smem_wavefronts_per_qp") if k in tj}
        if "smem_wavefronts" in tj and "warp_instructions" in tj:
            # the capture's per-launch counts spread over this run's measured step time at the measured SM clock: share of the
            # LSU's shared-memory wavefront slots (1 per SM and cycle) and of the warp-issue slots (1 per SM sub-partition and cycle)
            cycles = step_s * 1e6 * float(clocks.get("sm_mhz") or 1965.0)
            scale = B / tj["qps_per_launch"]
            ncu["lsu_wavefront_pct"] = 100.0 * tj["smem_wavefronts"] * scale / (info["num_sms"] * cycles)
            ncu["issue_pct"] = 100.0 * tj["warp_instructions"] * scale / (4 * info["num_sms"] * cycles)
            ncu["utilisation_source"] = "counts of " + tj["report"] + " / (slots in this run's ms_per_step at clocks.sm_mhz)"
    roofline = {"bound": "smem", "achieved": smem_bytes / step_s / 1e9, "peak": smem_peak, "unit": "GB/s",
                "frac": smem_bytes / step_s / 1e9 / smem_peak, "frac_solo": smem_bytes / ksec / 1e9 / smem_peak,
                "algorithmic_smem_bytes": float(smem_bytes), "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_hbm_bytes": hbm_bytes,
                "peak_source": "measured live (gomp_measure_peaks: conflict-free 16-byte shared-memory loads on all SMs)",
                "kernel": kernel_name, "kernel_ms": total_ms / args.steps, "kernel_ms_solo": kernel_avg_ms,
                "fp64": {"achieved": flops / step_s / 1e12, "peak": fp64_peak, "unit": "TFLOP/s", "frac": flops / step_s / 1e12 / fp64_peak,
                         "peak_source": "measured live (gomp_measure_peaks: 8 independent DFMA chains per thread)"},
                "hbm": {"achieved": hbm_bytes / step_s / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_bytes / step_s / 1e9 / hbm_peak,
                        "peak_source": hbm_src},
                "note": "SURVEY.md 8d: not an HBM- or tensor-bound kernel; bound = min(FP64 pipe, shared-memory bandwidth), "
                        "real limiter = dependency chain of the block substitutions",
                "mean_iters": float(it.mean()), "mean_factorizations": float(nfact.mean())}
    roofline.update(ncu)

    line = {"metric": "GOMP QPs solved/sec", "value": value, "unit": "QP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": desc,
            "e2e": {"value": e2e_value, "unit": "QP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "ms_per_step": e2e_ms / e2e_steps},
            "gpu_launches": int(launches),
            "ranks": {"device_ms": [round(v, 3) for v in rank_ms], "sm_mhz": rank_mhz, "numa_cpus_rank0": numa_cpus,
                      "per_step_gather": None if world == 1 else {"collective": "NCCL gather of {x, exit code, iterations} of every step on rank 0, "
                                                                                "inside the timed region",
                                                                  "bytes_per_step_into_rank0": (world - 1) * B * (8 * pat.n + 8),
                                                                  "last_step_block_ok": gathered_ok}},
            "roofline": roofline,
            "clocks": clocks,
            "solver": {"linsys": linsys, "qps_per_sm": info["qps_per_sm"], "ctas_per_sm": info["ctas_per_sm"], "smem_bytes_per_cta": info["smem_bytes"],
                       "threads_per_cta": info["threads_per_cta"],
                       "grid": info["grid"], "exit_code_hist": np.bincount(r["exit_code"], minlength=11).tolist(),
                       "iters_mean": float(it.mean()), "iters_max": int(it.max())}}

    O = entry.load_oracle()
    extra = {}
    if strong is not None:
        extra["strong_scaling"] = strong
    if world == 1:
        # ---- CPU baseline on the host cores (rank 0, N = 1 only): the GPU arm's own 4096 problems, continuous queue
        logical, phys = host_cores()
        cores = min(O.max_threads(), logical)
        d_cpu = {k: host[k].numpy() for k in ("A_val", "l", "u", "x0")}
        passes = 3
        time_oracle(pkg, O, pat, 256, first_id, cores, data={k: v[:256] for k, v in d_cpu.items()})  # untimed: thread start-up, first touch
        cpu_all, cpu_dt, ref = time_oracle(pkg, O, pat, B, first_id, cores, passes=passes, data=d_cpu)
        cpu_1, _, _ = time_oracle(pkg, O, pat, 96, first_id, 1, data={k: v[:96] for k, v in d_cpu.items()})
        line["cpu_baseline"] = {"value": cpu_all, "unit": "QP/s", "cores": cores, "logical_cpus": logical, "physical_cores": phys, "kind": "port",
                                "single_core_value": cpu_1, "all_core_over_single_core": cpu_all / cpu_1,
                                "sample": f"{passes} passes over the GPU arm's {B} QPs (problem ids {first_id}..{first_id + B - 1}) on all {cores} host threads, "
                                          f"{cpu_dt:.2f} s wall after one untimed 256-QP call; " + CPU_NOTE,
                                "status_equal_to_gpu": bool((ref["exit_code"] == r["exit_code"]).all()),
                                "iterations_equal_to_gpu": bool((ref["iters"] == r["iters"]).all())}
        if os.environ.get("GOMP_BENCH_EXTRA", "1") != "0":
            try:
                extra["sqp_sequence"] = extra_sqp_sequence(pkg, O, torch)
            except Exception as e:  # an extra must never take the headline line down
                extra["sqp_sequence"] = {"error": repr(e)}
            try:
                extra["long_horizon_h400"] = extra_long_horizon(pkg, O)
            except Exception as e:
                extra["long_horizon_h400"] = {"error": repr(e)}
            try:
                extra["example_shape_n6_w802"] = extra_long_horizon(pkg, O, W=802, batches=(1, 16), cpu_sample=16, n_dim=6, n_balls=2, n_obstacles=0,
                                                                    name="BASELINE.json configs[0] (examples/solver-example.cpp shape, synthetic chain)")
            except Exception as e:
                extra["example_shape_n6_w802"] = {"error": repr(e)}
    line["extra"] = extra
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=12)
    ap.add_argument("--impl", default="gomp_cuda", choices=["gomp_cuda", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
