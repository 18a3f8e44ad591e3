"""Multi-GPU plumbing: environments are independent, so the path shards by global env id with no
data-path collective.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is
used only to all-gather the int64[8] episode statistics, to max-reduce timings and to broadcast policy
weights (the packed DecomposedDQN(198) operand stages, ~1.1 MB) from the learner's rank -- never inside the step."""
from __future__ import annotations


def shard_range(total_envs: int, rank: int, world: int):
    """Contiguous global env ids [base, base + n) owned by `rank`; the shards tile [0, total)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    per, rem = divmod(int(total_envs), int(world))
    n = per + (1 if rank < rem else 0)
    base = rank * per + min(rank, rem)
    return base, n


def gather_stats(stats, group=None):
    """All-gather the per-rank stats vector -> [world, 8] on every rank (config 4)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return stats[None].clone()
    world = dist.get_world_size(group)
    out = [torch.zeros_like(stats) for _ in range(world)]
    dist.all_gather(out, stats, group=group)
    return torch.stack(out)


def max_over_ranks(values, group=None):
    """Element-wise MAX over ranks of a 1-D float64 tensor (device timings)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(values, op=dist.ReduceOp.MAX, group=group)
    return values


def broadcast_policy(tensors, src=0, group=None):
    """Broadcast the policy weights (e.g. AfterstateMLP.wpack and .bias) from rank `src` in place; returns the
    same tensors.  NCCL over NVLink on GPUs (one ncclBroadcast per tensor), a no-op without a process group."""
    import torch.distributed as dist

    import torch

    if dist.is_available() and dist.is_initialized():
        for t in tensors:
            # bytes on the wire: the packed weights are int16 bit patterns of bf16, which neither NCCL nor gloo
            # has a reduction type for (a broadcast needs none)
            dist.broadcast(t.view(torch.uint8) if t.is_contiguous() else t, src=src, group=group)
    return tensors


def merge_stats(all_stats):
    """[world, NUM_STATS] -> dict of job-wide totals (slot 6, max_actions, is a max not a sum)."""
    from . import _cabi

    tot = all_stats.sum(0).tolist()
    tot[6] = int(all_stats[:, 6].max().item())
    return {k: int(v) for k, v in zip(_cabi.STAT_NAMES, tot)}
