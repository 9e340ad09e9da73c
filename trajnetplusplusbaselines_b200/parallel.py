"""Multi-GPU plumbing: scenes are independent units.

Inference and the classical simulators shard contiguous scene ranges across ranks with NO
data-path collective; training adds exactly one all-reduce(SUM) per step over a single flat fp32
gradient bucket (the reference has no multi-device path at all, so this is new surface:
one process per GPU, torch.distributed over NCCL; gloo on CPU for the host-logic tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_scenes(batch_split, world_size, rank):
    """Contiguous scene range of `rank`, balanced by sum N_b^2 (pair work of the grid pooling).

    Returns (scene_lo, scene_hi, track_lo, track_hi, local_batch_split)."""
    bs = np.asarray([int(v) for v in batch_split], dtype=np.int64)
    sizes = np.diff(bs)
    cost = np.cumsum(sizes.astype(np.float64) ** 2)
    total = cost[-1] if len(cost) else 0.0
    B = len(sizes)
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        b = int(np.searchsorted(cost, target, side="left")) + 1 if B else 0
        b = min(max(b, bounds[-1]), B)
        bounds.append(b)
    bounds.append(B)
    lo, hi = bounds[rank], bounds[rank + 1]
    t_lo, t_hi = int(bs[lo]), int(bs[hi])
    return lo, hi, t_lo, t_hi, (bs[lo:hi + 1] - bs[lo]).tolist()


def allreduce_gradients(parameters, group=None, local_scenes=None, global_scenes=None):
    """One all-reduce(SUM) over a flat fp32 bucket of every existing gradient (parameters whose
    grad is None, e.g. goal_embedding.*, are skipped identically on all ranks).

    The reference multiplies the mean loss by batch_size (trainer.py:263); with each rank scaling
    by its LOCAL scene count the summed gradient equals the single-process one for equal shards."""
    params = [p for p in parameters if p.grad is not None]
    if not params:
        return 0
    flat = torch.cat([p.grad.reshape(-1).to(torch.float32) for p in params])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return flat.numel()
