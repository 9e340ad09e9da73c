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


def allreduce_gradients(parameters, group=None):
    """One all-reduce(SUM) per step over a single flat fp32 bucket.

    The bucket is laid out over the FIXED list of parameters that require grad, so every rank enters the
    collective with the same size whatever happened locally: a rank whose shard was empty (more ranks than
    scenes) or that ran no backward contributes zeros instead of skipping the call (which would leave the other
    ranks blocked).  One extra float per parameter says "some rank has a gradient for it": parameters without a
    gradient on ANY rank (e.g. goal_embedding.*) keep `grad is None`, exactly as in a single process, so the
    optimizer (weight decay) still skips them.

    The reference multiplies the mean loss by batch_size (trainer.py:263); with each rank scaling by its LOCAL
    scene count the summed gradient equals the single-process one."""
    params = [p for p in parameters if p.requires_grad]
    if not params:
        return 0
    device = params[0].device
    pieces = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(torch.float32) for p in params]
    flags = torch.tensor([0.0 if p.grad is None else 1.0 for p in params], dtype=torch.float32, device=device)
    flat = torch.cat(pieces + [flags])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    # a rank that ran a backward keeps `None` where it had none (every rank runs the same graph); only a rank
    # without any local gradient reads the flags (one small device-to-host copy) to learn which parameters exist
    local_any = any(p.grad is not None for p in params)
    have = None if local_any else flat[-len(params):].tolist()
    off = 0
    dst, src = [], []
    for i, p in enumerate(params):
        n = p.numel()
        if p.grad is not None:
            dst.append(p.grad)
            src.append(flat[off:off + n].view_as(p))
        elif have is not None and have[i] > 0:
            p.grad = flat[off:off + n].view_as(p).to(p.dtype).clone()
        off += n
    if dst:
        torch._foreach_copy_(dst, src)           # one fused launch instead of one copy kernel per parameter
    return flat.numel()
