"""CPU: frame sharding + the pose all-gather, world_size 2 and 3 over gloo (the N > 1 host logic of bench.py and
FramePipeline). The GPU path uses the same functions over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import rgbd_slam_b200 as rs


def test_frame_shard_partition():
    for n in (0, 1, 7, 32, 255, 256, 257):
        for world in (1, 2, 3, 4, 8):
            spans = [rs.sharding.frame_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
            assert sizes == rs.sharding.shard_sizes(n, world)
    with pytest.raises(ValueError):
        rs.sharding.frame_shard(8, 2, 2)


def test_gather_without_process_group_is_identity():
    p = torch.arange(21, dtype=torch.float64).view(3, 7)
    assert torch.equal(rs.sharding.gather_poses(p, 3), p)
    with pytest.raises(ValueError):
        rs.sharding.gather_poses(p, 5)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = rs.sharding.frame_shard(n_frames, rank, world)
        # each rank "solves" its shard: pose of global frame f = [f, 2f, 3f, 1, 0, 0, 0] + rank-independent noise
        f = torch.arange(a, b, dtype=torch.float64)
        local = torch.zeros((b - a, 7), dtype=torch.float64)
        local[:, 0], local[:, 1], local[:, 2], local[:, 3] = f, 2 * f, 3 * f, 1.0
        allp = rs.sharding.gather_poses(local, n_frames)
        q.put((rank, allp.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames", [(2, 32), (2, 33), (3, 8)])
def test_pose_all_gather_gloo(world, n_frames):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    f = np.arange(n_frames, dtype=np.float64)
    expect = np.zeros((n_frames, 7))
    expect[:, 0], expect[:, 1], expect[:, 2], expect[:, 3] = f, 2 * f, 3 * f, 1.0
    for r in range(world):
        np.testing.assert_array_equal(results[r], expect)
