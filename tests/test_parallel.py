"""Host logic of the multi-GPU path on CPU: scene sharding and the flat-bucket gradient
all-reduce with world_size = 2 over gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from trajnetplusplusbaselines_b200.parallel import allreduce_gradients, shard_scenes


def test_shard_scenes_partitions_every_scene_once():
    rng = np.random.RandomState(0)
    sizes = rng.randint(1, 40, size=57)
    bs = np.concatenate([[0], np.cumsum(sizes)])
    for world in (1, 2, 3, 8):
        seen = []
        costs = []
        for rank in range(world):
            lo, hi, t_lo, t_hi, local = shard_scenes(bs, world, rank)
            seen += list(range(lo, hi))
            assert t_lo == bs[lo] and t_hi == bs[hi]
            assert local[0] == 0 and local[-1] == t_hi - t_lo and len(local) == hi - lo + 1
            costs.append(float((sizes[lo:hi].astype(float) ** 2).sum()))
        assert seen == list(range(len(sizes)))
        if world > 1:
            assert max(costs) <= 2.0 * (sum(costs) / world) + sizes.max() ** 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
    frozen = torch.nn.Linear(2, 2)            # never used -> grad None, must be skipped on all ranks
    x = torch.arange(24, dtype=torch.float32).reshape(4, 6) * (rank + 1)
    model(x).sum().backward()
    n = allreduce_gradients(list(model.parameters()) + list(frozen.parameters()))
    assert all(p.grad is None for p in frozen.parameters())        # no gradient on any rank -> stays None
    first = [p.grad.clone() for p in model.parameters()]
    # second step: rank 1 has an empty shard (runs no backward at all) and must still enter the collective
    model.zero_grad(set_to_none=True)
    if rank == 0:
        model(x).sum().backward()
    allreduce_gradients(list(model.parameters()) + list(frozen.parameters()))
    second = [p.grad.clone() for p in model.parameters()]
    assert all(p.grad is None for p in frozen.parameters())
    out[rank] = (n, first, second)
    dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
    expect = None
    for rank in range(world):
        model.zero_grad()
        x = torch.arange(24, dtype=torch.float32).reshape(4, 6) * (rank + 1)
        model(x).sum().backward()
        g = [p.grad.clone() for p in model.parameters()]
        expect = g if expect is None else [a + b for a, b in zip(expect, g)]
    model.zero_grad()
    model(torch.arange(24, dtype=torch.float32).reshape(4, 6)).sum().backward()
    only_rank0 = [p.grad.clone() for p in model.parameters()]
    n_params = sum(p.numel() for p in model.parameters()) + 2 * 2 + 2          # + the unused Linear(2, 2)
    for rank in range(world):
        n, grads, second = out[rank]
        assert n == n_params + 6                    # fixed layout: every parameter + one "has a gradient" flag each
        for a, b in zip(grads, expect):
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-6)
        for a, b in zip(second, only_rank0):        # the idle rank received rank 0's gradients
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-6)


class _ConstantVelocity:
    """Stand-in with the reference's predictor call signature (CPU)."""

    def __call__(self, paths, scene_goal, n_predict=12, modes=1, predict_all=True, obs_length=9, start_length=0, args=None):
        from trajnetplusplusbaselines_b200.data import paths_to_xy
        xy = paths_to_xy(paths)
        v = xy[obs_length - 1] - xy[obs_length - 2]
        pred = xy[obs_length - 1][None] + np.arange(1, n_predict + 1)[:, None, None] * v[None]
        return {0: [pred[:, 0], pred[:, 1:]]}


def _write_test_file(filename, sizes, seed=0):
    from trajnetplusplusbaselines_b200.data import SceneRow, TrackRow, trajnet_line
    rng = np.random.RandomState(seed)
    with open(filename, "w") as f:
        for sid, n in enumerate(sizes):
            start = rng.randn(n, 2) * 3.0
            vel = rng.randn(n, 2) * 0.2
            frames = [1000 * sid + 10 * t for t in range(21)]
            f.write(trajnet_line(SceneRow(sid, 100 * sid, frames[0], frames[-1], 2.5, 0)) + "\n")
            for p in range(n):
                for t, fr in enumerate(frames):
                    f.write(trajnet_line(TrackRow(fr, 100 * sid + p, start[p, 0] + vel[p, 0] * t, start[p, 1] + vel[p, 1] * t)) + "\n")


def _eval_worker(rank, world, port, infile, outfile):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from trajnetplusplusbaselines_b200.evaluator import evaluate_file
    n = evaluate_file(_ConstantVelocity(), infile, outfile, chunk=2)
    assert n == 7
    assert os.path.exists(outfile) and not os.path.exists("%s.part%d" % (outfile, rank))     # after the final barrier
    dist.destroy_process_group()


def test_sharded_evaluate_file_world2_gloo_is_byte_identical(tmp_path):
    """Two ranks, contiguous scene ranges, parts concatenated by rank 0 == the single-process file, byte for byte; three
    ranks with explicit (rank, world_size) and one empty shard likewise."""
    from trajnetplusplusbaselines_b200.evaluator import evaluate_file
    infile = str(tmp_path / "in.ndjson")
    _write_test_file(infile, [3, 1, 6, 2, 2, 9, 4])
    single = str(tmp_path / "single.ndjson")
    assert evaluate_file(_ConstantVelocity(), infile, single, rank=0, world_size=1) == 7
    sharded = str(tmp_path / "sharded.ndjson")
    mp.spawn(_eval_worker, args=(2, _free_port(), infile, sharded), nprocs=2, join=True)
    assert open(sharded, "rb").read() == open(single, "rb").read()
    # explicit ranks without a process group (no barrier): run the non-zero ranks first, rank 0 assembles
    small = str(tmp_path / "small.ndjson")
    _write_test_file(small, [5, 1])
    one = str(tmp_path / "one.ndjson")
    evaluate_file(_ConstantVelocity(), small, one, rank=0, world_size=1)
    three = str(tmp_path / "three.ndjson")
    for rank in (2, 1, 0):                           # more ranks than scenes: one shard is empty
        evaluate_file(_ConstantVelocity(), small, three, rank=rank, world_size=3)
    assert open(three, "rb").read() == open(one, "rb").read()
