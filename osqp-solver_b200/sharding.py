"""Multi-GPU sharding of a QP batch: contiguous slices of the problem index per rank, no exchange
during the solve (the QPs are independent), one gather of the results at the end
(SURVEY.md 8e).  Works with any torch.distributed backend: NCCL on the GPUs, gloo in the CPU tests."""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """[first, last) of the problems rank `rank` owns: contiguous, sizes differ by at most one."""
    base, extra = divmod(total, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def gather_results(x, exit_code, iters, total, dst=0):
    """Collect per-rank results (x [count, n], exit_code [count], iters [count]; torch tensors on the
    backend's device) on rank `dst` in problem order.  Returns (x, exit_code, iters) on dst, None elsewhere.
    Shards may differ in size by one, so every rank pads to the largest shard."""
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
    cap = max(counts)

    def padded(t):
        if t.shape[0] == cap:
            return t.contiguous()
        pad = torch.zeros((cap - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        return torch.cat([t, pad]).contiguous()

    outs = []
    for t in (x, exit_code, iters):
        t = padded(t)
        bufs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
        dist.gather(t, bufs, dst=dst)
        if rank == dst:
            outs.append(torch.cat([bufs[r][:counts[r]] for r in range(world)]))
    return tuple(outs) if rank == dst else None
