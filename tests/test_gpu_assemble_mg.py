"""Device-side problem assembly (BASELINE.json configs[4]: only start / goal cross PCIe; gomp_device_assemble in
osqp-solver_b200/csrc/gomp_relin.cu) against the host generator that drives the facade's ConstraintBuilder, and the
multi-GPU handle gomp_mg_* (include/gomp_cuda.h) against the single-GPU handle."""
import numpy as np
import pytest

from common import oracle, package
from test_gpu_parity import compare, make_solver

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kw", [dict(), dict(n_balls=2, n_obstacles=2), dict(n_dim=6, waypoints=12), dict(n_dim=3, waypoints=20, n_obstacles=0)])
def test_device_assembly_matches_host_generator(kw):
    import torch
    pkg = package()
    P = pkg.problem
    pat = P.Pattern(P.default_config(**kw))
    first, B = 37, 96
    want = P.generate(pat, first, B)
    sc = P.DeviceScene(pat, None, B)                  # no host-side scene: the obstacle lines are derived on the device
    sc.set_template(P.generate(pat, 5, 1))            # any problem of the configuration serves as template
    sg = torch.from_numpy(P.start_goal(pat, first, B)).cuda()
    dev = {k: torch.full(want[k].shape, float("nan"), dtype=torch.float64, device="cuda") for k in ("A_val", "l", "u", "x0")}
    sc.assemble(sg, dev["A_val"], dev["l"], dev["u"], dev["x0"])
    torch.cuda.synchronize()
    got = {k: v.cpu().numpy() for k, v in dev.items()}
    np.testing.assert_array_equal(got["x0"], want["x0"])              # linspace: same arithmetic, bit-exact
    rows, kidx = P.cart_rows(pat)
    cart = np.zeros(pat.m, dtype=bool)
    cart[rows] = True
    cart_val = np.zeros(pat.nnzA, dtype=bool)
    cart_val[kidx.reshape(-1)] = True
    # everything outside the Cartesian rows is exact (template + pinned start / goal rows)
    np.testing.assert_array_equal(got["l"][:, ~cart], want["l"][:, ~cart])
    np.testing.assert_array_equal(got["u"][:, ~cart], want["u"][:, ~cart])
    np.testing.assert_array_equal(got["A_val"][:, ~cart_val], want["A_val"][:, ~cart_val])
    # Cartesian rows: same active set, values up to sin / cos rounding
    assert ((got["l"] > -1e29) == (want["l"] > -1e29)).all() and ((got["u"] < 1e29) == (want["u"] < 1e29)).all()
    np.testing.assert_allclose(got["A_val"], want["A_val"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(got["l"], want["l"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(got["u"], want["u"], rtol=1e-12, atol=1e-12)
    sc.close()


def test_device_assembled_batch_solves_like_the_oracle():
    """start / goal -> device assembly -> solve, no QP data from the host; oracle fed with the host generator's arrays."""
    import torch
    pkg, O = package(), oracle()
    P, S = pkg.problem, pkg.solver
    pat = P.Pattern(P.default_config())
    first, B = 4096, 64
    want = P.generate(pat, first, B)
    sc = P.DeviceScene(pat, None, B)
    sc.set_template()
    stream = torch.cuda.Stream()
    dev = {k: torch.empty(want[k].shape, dtype=torch.float64, device="cuda") for k in ("A_val", "l", "u", "x0")}
    sg = torch.from_numpy(P.start_goal(pat, first, B)).cuda()
    torch.cuda.synchronize()
    st = S.default_settings(borrow_inputs=1, stream=stream.cuda_stream)
    solver = S.BatchQPSolver(pat.n, pat.m, pat.P_colptr, pat.P_rowidx, pat.P_val, pat.A_colptr, pat.A_rowidx, batch=B, settings=st,
                             var_order=pat.var_order, block_size=pat.block_size)
    sc.assemble(sg, dev["A_val"], dev["l"], dev["u"], dev["x0"], stream=stream.cuda_stream)
    solver.setup(dev["A_val"], dev["l"], dev["u"])
    solver.set_warm_start(dev["x0"])
    got = solver.solve(want_y=True)
    ref = O.solve_batch(pat.n, pat.m, pat.P_scipy(), (pat.A_colptr, pat.A_rowidx), want["A_val"], want["l"], want["u"], want["x0"], want_y=True)
    compare(got, ref)
    solver.close()
    sc.close()


def _mg_vs_single(devices, B=75):
    pkg, O = package(), oracle()
    P, S = pkg.problem, pkg.solver
    pat = P.Pattern(P.default_config())
    d = P.generate(pat, 900, B)
    single = make_solver(pkg, pat, d)
    single.set_warm_start(d["x0"])
    want = single.solve(want_y=True)
    mg = S.MultiGpuSolver(pat.n, pat.m, pat.P_colptr, pat.P_rowidx, pat.P_val, pat.A_colptr, pat.A_rowidx, batch=B, devices=devices,
                          var_order=pat.var_order, block_size=pat.block_size)
    assert mg.num_shards == len(devices)
    assert sum(mg.shard(k)[1] for k in range(mg.num_shards)) == B
    mg.setup(d["A_val"], d["l"], d["u"])
    mg.set_warm_start(d["x0"])
    got = mg.solve(want_y=True)
    # the same kernel on the same inputs: bit-identical, whatever device a problem landed on
    for k in ("exit_code", "iters", "rho_updates"):
        np.testing.assert_array_equal(got[k], want[k])
    np.testing.assert_array_equal(got["x"], want["x"])
    np.testing.assert_array_equal(got["y"], want["y"])
    # an SQP-style update round through the sharded handle
    ok = want["exit_code"] == 0
    x = np.where(ok[:, None], want["x"], d["x0"])
    upd = P.relinearize(pat, 900, x)
    single.update(upd["A_val"], upd["l"], upd["u"])
    want2 = single.solve()
    mg.update(upd["A_val"], upd["l"], upd["u"])
    got2 = mg.solve()
    for k in ("exit_code", "iters", "rho_updates"):
        np.testing.assert_array_equal(got2[k], want2[k])
    np.testing.assert_array_equal(got2["x"], want2["x"])
    # and against the oracle
    ref = O.solve_batch(pat.n, pat.m, pat.P_scipy(), (pat.A_colptr, pat.A_rowidx), d["A_val"], d["l"], d["u"], d["x0"], want_y=True)
    compare(got, ref)
    # l > u is refused before anything is launched (the reference throws in update, osqp-wrapper.h:36-42)
    bad_l = upd["l"].copy()
    bad_l[3, 10] = upd["u"][3, 10] + 1.0
    with pytest.raises(S.GompError) as e:
        mg.update(None, bad_l, upd["u"])
    assert e.value.code == 1
    mg.close()
    single.close()


def test_multi_gpu_handle_interleaving_on_one_device():
    """Three shards on device 0: exercises the strided scatter / gather without needing a second GPU."""
    _mg_vs_single([0, 0, 0])


def test_multi_gpu_handle_two_devices_equals_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    _mg_vs_single([0, 1], B=301)
