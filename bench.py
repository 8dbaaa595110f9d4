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
           a bounded sample of the same workload

--impl reference times that CPU oracle on all host threads instead (the real reference cannot be
built offline: it needs osqp-cpp/OSQP/Eigen via CMake FetchContent).

Workload: BASELINE.json configs[1] -- 4096 synthetic 7-DoF GOMP QPs, horizon H=50, per GPU
(weak scaling across ranks; problems are independent, NCCL only gathers the results).
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


def workload(pkg):
    P = pkg.problem
    cfg = P.default_config()
    pat = P.Pattern(cfg)
    desc = {"workload": "BASELINE.json configs[1]: batch of 4096 synthetic 7-DoF GOMP QPs per GPU, H=50, pos/vel/acc bounds + "
                        "1 gripper ball + 1 horizontal-line obstacle; cold setup + warm start + solve, OSQP default settings "
                        "(eps 1e-3, adaptive_rho_interval pinned to 25)",
            "n_dim": 7, "waypoints": 50, "n": pat.n, "m": pat.m, "nnzA": pat.nnzA, "batch_per_gpu": BATCH_PER_GPU,
            "l2_policy": "inputs larger than L2 (305 MB of A_val/l/u per step vs 126 MB L2)",
            "batches_in_flight": PIPE}
    return cfg, pat, desc


def time_oracle(pkg, O, pat, count, first_id, nthreads):
    d = pkg.problem.generate(pat, first_id, count)
    t0 = time.perf_counter()
    r = O.solve_batch(pat.n, pat.m, pat.P_scipy(), (pat.A_colptr, pat.A_rowidx), d["A_val"], d["l"], d["u"], d["x0"], nthreads=nthreads)
    dt = time.perf_counter() - t0
    return count / dt, dt, r


def run_reference(args):
    """CPU arm: the oracle on all host threads, bounded sample of the same workload per step."""
    rank, local_rank, world = rank_env()
    if rank != 0:
        return
    pkg = entry.load_package()
    O = entry.load_oracle()
    cfg, pat, desc = workload(pkg)
    cores = O.max_threads()
    sample = 1024
    times = []
    for s in range(args.warmup + args.steps):
        qps, dt, _ = time_oracle(pkg, O, pat, sample, 0, cores)
        if s >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = sample * len(times) / total
    line = {"impl": "reference", "metric": "GOMP QPs solved/sec", "value": value, "unit": "QP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": desc,
            "cpu_baseline": {"value": value, "unit": "QP/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} QPs of the workload per step (problem ids 0..{sample - 1}), one QP per thread, KKT ordering computed once "
                                       "per batch (favours the CPU); oracle = restatement of OSQP 0.6, not genuine OSQP/qdldl"},
            "e2e": {"value": value, "unit": "QP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = rank_env()
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner with printf when the communicator is created: keep it off stdout, which carries
        # the one JSON line (fd 1 points at stderr until the communicator exists)
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

    def step_resident(k=0):
        sv = solvers[k % PIPE]
        sv.setup(dev["A_val"], dev["l"], dev["u"])
        sv.set_warm_start(dev["x0"])
        sv.solve_async()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gathered = [torch.empty_like(x_dev) for _ in range(world)] if (world > 1 and rank == 0) else None

    def gather_results():
        # problems are independent: the only collective is the final gather of the solutions (NCCL)
        if world > 1:
            solvers[-1].get(x=x_dev, exit_code=code, iters=iters, rho_updates=rup)
            dist.gather(x_dev, gathered, dst=0)

    def run_steps(nsteps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(main)
        for st_ in streams:
            st_.wait_event(ev0)
        for k in range(nsteps):
            step_resident(k)
        for st_ in streams:
            e = torch.cuda.Event()
            e.record(st_)
            main.wait_event(e)
        ev1.record(main)
        return ev0, ev1

    run_steps(args.warmup)
    barrier()
    gather_results()
    barrier()
    launches0 = sum(sv.info()["kernel_launches"] for sv in solvers)
    sampler = ClockSampler(local_rank)
    sampler.start()
    kernel_ms = []
    barrier()
    ev0, ev1 = run_steps(args.steps)
    gather_results()  # N > 1: the NCCL gather of the last batch's solutions, inside the timed region
    if world > 1:
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record(main)
    barrier()
    clocks = sampler.stop()
    total_ms = ev0.elapsed_time(ev1)
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

    def e2e_submit(k):
        sv = e2e_solvers[k % PIPE]
        sv.setup(host["A_val"], host["l"], host["u"])
        sv.set_warm_start(host["x0"])
        sv.solve_async()

    def e2e_collect(k):
        e2e_solvers[k % PIPE].get(x=x_hosts[k % PIPE].numpy(), exit_code=code, iters=iters, rho_updates=rup)

    def e2e_run(nsteps):
        for k in range(nsteps):
            if k >= PIPE:
                e2e_collect(k - PIPE)
            e2e_submit(k)
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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the ADMM kernel (work model: DESIGN.md / gomp_batch_info)
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
    ksec = kernel_avg_ms * 1e-3
    # DRAM bytes of one launch from the committed ncu --set full capture (same kernel, same workload)
    traffic, traffic_src = None, None
    streamed = int(info["linsys"]) == 2  # AUTO + prefer_throughput: the two-QPs-per-SM mode of the streamed kernel
    tpath = os.path.join(ROOT, "profiles", "r01_ncu_traffic_stream.json" if streamed else "r01_ncu_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) * B / tj["qps_per_launch"]
        traffic_src = "dram__bytes_read.sum + dram__bytes_write.sum per launch, " + tj["capture"]
    roofline = {"bound": "smem", "achieved": smem_bytes / ksec / 1e9, "peak": smem_peak, "unit": "GB/s",
                "frac": smem_bytes / ksec / 1e9 / smem_peak, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_hbm_bytes": hbm_bytes,
                "peak_source": "measured live (gomp_measure_peaks: conflict-free 16-byte shared-memory loads on all SMs)",
                "kernel": "gomp_admm_stream_kernel<14,256,true>" if streamed else "gomp_admm_kernel<4,2,14>", "kernel_ms": kernel_avg_ms,
                "fp64": {"achieved": flops / ksec / 1e12, "peak": fp64_peak, "unit": "TFLOP/s", "frac": flops / ksec / 1e12 / fp64_peak,
                         "peak_source": "measured live (gomp_measure_peaks: 8 independent DFMA chains per thread)"},
                "hbm": {"achieved": hbm_bytes / ksec / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_bytes / ksec / 1e9 / hbm_peak,
                        "peak_source": hbm_src},
                "note": "SURVEY.md 8d: not an HBM- or tensor-bound kernel; bound = min(FP64 pipe, shared-memory bandwidth), "
                        "real limiter = dependency chain of the block substitutions",
                "mean_iters": float(it.mean()), "mean_factorizations": float(nfact.mean())}

    # ---- CPU baseline on the host cores (bounded sample)
    O = entry.load_oracle()
    cores = O.max_threads()
    sample, passes = 1024, 3
    time_oracle(pkg, O, pat, 256, first_id, cores)          # untimed: thread start-up, first-touch of the work arrays
    cpu_dt = 0.0
    for _ in range(passes):
        _, dt, ref = time_oracle(pkg, O, pat, sample, first_id, cores)
        cpu_dt += dt
    cpu_all = passes * sample / cpu_dt
    cpu_1, _, _ = time_oracle(pkg, O, pat, 64, first_id, 1)
    same_code = bool((ref["exit_code"] == r["exit_code"][:sample]).all())
    same_iters = bool((ref["iters"] == r["iters"][:sample]).all())

    line = {"metric": "GOMP QPs solved/sec", "value": value, "unit": "QP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": desc,
            "e2e": {"value": e2e_value, "unit": "QP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "ms_per_step": e2e_ms / e2e_steps},
            "gpu_launches": int(launches),
            "ranks": {"device_ms": [round(v, 3) for v in rank_ms], "sm_mhz": rank_mhz},
            "roofline": roofline,
            "cpu_baseline": {"value": cpu_all, "unit": "QP/s", "cores": cores, "kind": "port", "single_core_value": cpu_1,
                             "sample": f"{sample} QPs of the same workload (problem ids {first_id}..{first_id + sample - 1}) on all {cores} host threads, "
                                       f"one QP per thread, {passes} passes after one untimed pass, {cpu_dt:.2f} s wall; oracle = CPU restatement of OSQP 0.6 (not genuine OSQP/qdldl), "
                                       "KKT ordering shared across the batch",
                             "status_equal_to_gpu": same_code, "iterations_equal_to_gpu": same_iters},
            "clocks": clocks,
            "solver": {"ctas_per_sm": info["ctas_per_sm"], "smem_bytes_per_cta": info["smem_bytes"], "threads_per_cta": info["threads_per_cta"],
                       "grid": info["grid"], "exit_code_hist": np.bincount(r["exit_code"], minlength=11).tolist(),
                       "iters_mean": float(it.mean()), "iters_max": int(it.max())}}
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
