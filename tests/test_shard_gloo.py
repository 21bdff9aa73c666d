"""World-size-2 (gloo, CPU) test of the multi-GPU host logic: rollouts are partitioned by id with no data-path
collective, noise depends on the global rollout id only, and the one optional gather restores global order."""
import os
import socket

import torch
import torch.multiprocessing as mp

from gtav_b200.shard import gather_rollouts, rollout_noise, shard_rollouts


def test_partition_covers_every_rollout_once():
    for n in (1, 2, 7, 8, 64):
        for world in (1, 2, 4, 8):
            owned = [shard_rollouts(n, r, world) for r in range(world)]
            flat = sorted(i for o in owned for i in o)
            assert flat == list(range(n))
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1


def test_noise_is_independent_of_the_world_size():
    full = rollout_noise(list(range(4)), 3, (2, 5), "cpu")
    for world in (2, 4):
        for rank in range(world):
            ids = shard_rollouts(4, rank, world)
            assert torch.equal(rollout_noise(ids, 3, (2, 5), "cpu"), full[ids])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rollouts, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids = shard_rollouts(n_rollouts, rank, world)
        # a stand-in for "frames of rollout r": values that encode r, computed with NO communication
        noise = rollout_noise(ids, 2, (3,), "cpu")
        local = noise.sum(dim=(1, 2), keepdim=False).reshape(len(ids), 1) + torch.tensor(ids).reshape(-1, 1) * 1000.0
        out = gather_rollouts(local, n_rollouts, rank, world)
        if rank == 0:
            ret.put(out.clone())
        else:
            assert out is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_gather_world_size_2_matches_single_process():
    n = 5
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, ret)) for r in range(2)]
    for p in procs:
        p.start()
    out = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ids = list(range(n))
    noise = rollout_noise(ids, 2, (3,), "cpu")
    expect = noise.sum(dim=(1, 2)).reshape(n, 1) + torch.tensor(ids).reshape(-1, 1) * 1000.0
    assert torch.equal(out, expect)
