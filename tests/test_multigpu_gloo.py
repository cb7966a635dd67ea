"""CPU test of the N>1 host logic (world_size 2, gloo): sequences shard one per rank with no data-path collective;
the only exchange is an all-gather of the per-frame poses.  The CPU oracle pipeline stands in for the
device path here (no GPU in this container); bench.py uses the same gather with NCCL."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    from oracle import frontend as ofe
    from oracle import synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cam = dict(width=320, height=200, fx=260.0, fy=260.0, cx=160.0, cy=100.0, b=0.3)
    world_ = synth.World(seed=20 + rank, length=30.0, n_quads=80, n_segs=40, half_width=5.0, half_height=2.5)
    frames = [(L, R) for L, R, _ in synth.stream(cam, 3, world=world_, seed=100 + rank, step=0.1)]   # rank r <-> sequence r
    res = ofe.run_sequence(cam, frames, dict(orb_nfeatures=300, lsd_nfeatures=60, orb_nlevels=2))
    poses = torch.from_numpy(np.stack([r["DT"].reshape(16) for r in res]))
    gathered = torch.zeros((world * len(res), 16), dtype=torch.float64)
    dist.all_gather_into_tensor(gathered, poses)
    dist.barrier()
    q.put((rank, poses.numpy(), gathered.numpy()))
    dist.destroy_process_group()


def test_pose_allgather_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out.sort(key=lambda t: t[0])
    own = np.concatenate([o[1] for o in out])
    for _, _, g in out:
        assert np.array_equal(g, own)             # every rank holds every sequence's poses, rank-major
    assert not np.array_equal(out[0][1], out[1][1])   # the two ranks really tracked different sequences
