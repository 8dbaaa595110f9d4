"""N > 1 host logic on CPU: two gloo ranks own contiguous shards of a batch, solve them independently
(here with the oracle standing in for the device), and one gather reassembles the results in problem
order -- identical to the single-process answer.  No collective touches the solve itself."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import oracle, package

TOTAL = 7  # odd on purpose: shards of 4 and 3


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg, O = package(), oracle()
    P = pkg.problem
    pat = P.Pattern(P.default_config(waypoints=12))
    first, last = pkg.sharding.shard_range(TOTAL, rank, world)
    d = P.generate(pat, first, last - first)
    r = O.solve_batch(pat.n, pat.m, pat.P_scipy(), (pat.A_colptr, pat.A_rowidx), d["A_val"], d["l"], d["u"], d["x0"], nthreads=1)
    got = pkg.sharding.gather_results(torch.from_numpy(r["x"]), torch.from_numpy(r["exit_code"]), torch.from_numpy(r["iters"]), TOTAL)
    if rank == 0:
        np.savez(out_path, x=got[0].numpy(), code=got[1].numpy(), iters=got[2].numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition_the_batch():
    pkg = package()
    for total in (1, 7, 4096, 1048576):
        for world in (1, 2, 4, 8):
            spans = [pkg.sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gather_matches_single_process(tmp_path):
    pkg, O = package(), oracle()
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    P = pkg.problem
    pat = P.Pattern(P.default_config(waypoints=12))
    d = P.generate(pat, 0, TOTAL)
    ref = O.solve_batch(pat.n, pat.m, pat.P_scipy(), (pat.A_colptr, pat.A_rowidx), d["A_val"], d["l"], d["u"], d["x0"], nthreads=1)
    np.testing.assert_array_equal(got["code"], ref["exit_code"])
    np.testing.assert_array_equal(got["iters"], ref["iters"])
    np.testing.assert_array_equal(got["x"], ref["x"])
