"""Sharding helpers for the multi-GPU form of the hot path: every unit of work is independent
(row blocks of the IoU matrix, images for NMS / detection), so ranks take contiguous shards and NO data-path
collective is needed; results are gathered only when a caller wants the whole thing on every rank."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """contiguous [lo, hi) of n units for `rank`; sizes differ by at most one, earlier ranks take the remainder"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_rows(local, n_total):
    """gather row shards [n_r, ...] (shard_range layout) into [n_total, ...] on every rank (optional epilogue)"""
    world = dist.get_world_size()
    rank = dist.get_rank()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], 0)


def max_over_ranks(value, device):
    """timing convention of bench.py: the job takes as long as its slowest rank"""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_gradients(params):
    """Data-parallel training (train.py:168-176 wraps the model in DistributedDataParallel): average the gradients of
    `params` over the ranks with ONE all-reduce of the flattened buffer (NCCL on GPUs, gloo in the CPU tests).  Every rank
    must call it the same number of times.  No-op without an initialised process group / with world size 1."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat)
    flat.div_(dist.get_world_size())
    for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
        g.copy_(f)
