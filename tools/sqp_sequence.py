"""BASELINE.json configs[2]: warm-started SQP re-solve sequence -- `batch` QPs x `rounds` rounds of
withObstacles(x) -> update(A, l, u) -> solve (reference src/gomp-solver.h:70-88 without the early exit), timed
  (a) with the data plane on the device (gomp_relin.cu re-linearises in place, nothing crosses PCIe),
  (b) through host buffers (D2H x, libgomp_host withObstacles on all cores, H2D A_val / l / u), and
  (c) on the CPU oracle for a sample of the batch (all host threads).
usage: python tools/sqp_sequence.py [batch] [rounds] [cpu_sample]       (prints one JSON line)"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

pkg, O = entry.load_package(), entry.load_oracle()
P, S = pkg.problem, pkg.solver
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cpu_sample = int(sys.argv[3]) if len(sys.argv) > 3 else 128
pat = P.Pattern(P.default_config())
d = P.generate(pat, 0, batch)


def make(dev_arrays, stream):
    # GOMP_SQP_THROUGHPUT=1: two QPs per SM (streamed kernel, small-QP mode) instead of the resident kernel
    st = S.default_settings(borrow_inputs=1, prefer_throughput=int(os.environ.get("GOMP_SQP_THROUGHPUT", "0")))
    st.stream = stream.cuda_stream
    s = S.BatchQPSolver(pat.n, pat.m, pat.P_colptr, pat.P_rowidx, pat.P_val, pat.A_colptr, pat.A_rowidx, dev_arrays["A_val"],
                        dev_arrays["l"], dev_arrays["u"], batch=batch, settings=st, var_order=pat.var_order, block_size=pat.block_size)
    return s


def device_path():
    stream = torch.cuda.Stream()
    dev = {k: torch.from_numpy(d[k]).cuda() for k in ("A_val", "l", "u", "x0")}
    s = make(dev, stream)
    sc = P.DeviceScene(pat, 0, batch)
    x = torch.empty((batch, pat.n), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.set_warm_start(dev["x0"])
    iters = 0
    for r in range(rounds):
        s.solve_async()
        got = s.get(x=x)
        iters += int(got["iters"].sum())
        if r + 1 < rounds:
            sc.relinearize(x, dev["A_val"], dev["l"], dev["u"], x_fallback=dev["x0"], stream=stream.cuda_stream)
            s.update(dev["A_val"], dev["l"], dev["u"])
    stream.synchronize()
    dt = time.perf_counter() - t0
    codes = got["exit_code"].copy()
    s.close(); sc.close()
    return dt, iters, codes, x.cpu().numpy()


def host_path():
    stream = torch.cuda.Stream()
    pin = {k: torch.from_numpy(d[k].copy()).pin_memory() for k in ("A_val", "l", "u")}
    dev = {k: pin[k].cuda() for k in pin}
    dev["x0"] = torch.from_numpy(d["x0"]).cuda()
    s = make(dev, stream)
    xh = torch.empty((batch, pat.n), dtype=torch.float64).pin_memory()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.set_warm_start(dev["x0"])
    for r in range(rounds):
        s.solve_async()
        got = s.get(x=xh.numpy())
        if r + 1 < rounds:
            xs = np.where(np.isnan(xh.numpy()), d["x0"], xh.numpy())
            P.relinearize(pat, 0, xs, out={k: pin[k].numpy() for k in pin})
            with torch.cuda.stream(stream):
                for k in pin:
                    dev[k].copy_(pin[k], non_blocking=True)
            s.update(dev["A_val"], dev["l"], dev["u"])
    stream.synchronize()
    dt = time.perf_counter() - t0
    codes = got["exit_code"].copy()
    s.close()
    return dt, codes, xh.numpy().copy()


def cpu_path(nsample):
    import concurrent.futures as cf
    Ps = pat.P_scipy()
    nthreads = O.max_threads()

    def one(b):
        ref = O.Solver(Ps, pat.A_scipy(d["A_val"][b]), d["l"][b], d["u"][b])
        ref.warm_start_x(d["x0"][b])
        for r in range(rounds):
            ref.solve()
            if r + 1 < rounds:
                xs = np.where(np.isnan(ref.x), d["x0"][b], ref.x)
                upd = P.relinearize(pat, b, xs[None, :], nthreads=1)
                ref.update(upd["A_val"][0], upd["l"][0], upd["u"][0])
        return ref.exit_code
    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(nthreads) as ex:
        codes = list(ex.map(one, range(nsample)))
    return time.perf_counter() - t0, nthreads, np.array(codes)


device_path()  # warm-up (context, kernels, allocator)
t_dev, iters, codes_dev, x_dev = device_path()
t_host, codes_host, x_host = host_path()
t_cpu, nthreads, codes_cpu = cpu_path(min(cpu_sample, batch))
both = (codes_dev == 0) & (codes_host == 0)
print(json.dumps({
    "workload": f"configs[2]: {batch} QPs x {rounds} SQP rounds (N=7, H=50)", "qp_solves": batch * rounds,
    "device_path_s": round(t_dev, 4), "device_qp_solves_per_s": round(batch * rounds / t_dev, 1),
    "host_roundtrip_path_s": round(t_host, 4), "host_roundtrip_qp_solves_per_s": round(batch * rounds / t_host, 1),
    "cpu_oracle": {"sample": int(min(cpu_sample, batch)), "threads": nthreads, "s": round(t_cpu, 3),
                   "qp_solves_per_s": round(min(cpu_sample, batch) * rounds / t_cpu, 1), "note": "ctypes oracle, one problem per Python thread (GIL released inside the C calls)"},
    "mean_admm_iters_per_solve": round(iters / (batch * rounds), 2),
    "final_codes_equal_device_vs_host_path": bool((codes_dev == codes_host).all()),
    "max_rel_dx_device_vs_host_path": float(np.nanmax(np.abs(x_dev[both] - x_host[both])) / np.nanmax(np.abs(x_host[both]))) if both.any() else None,
    "final_codes_equal_device_vs_cpu_sample": bool((codes_dev[:len(codes_cpu)] == codes_cpu).all())}))
