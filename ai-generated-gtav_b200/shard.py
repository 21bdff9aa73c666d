"""Multi-GPU story of the hot path: independent rollouts, batch-sharded over the ranks of one box, one process per
GPU, NO collective on the data path.

The reference has no multi-GPU inference at all (generate.py:133 hard-codes B = 1; its only parallelism is DDP
in training, train_dit.py:182-188).  Rollouts never interact - attention, conditioning and the DDIM update are
per sample - so the natural partition is by rollout:

    rollout r  ->  rank r mod world            (shard_rollouts)
    its noise  ->  generator seeded base + r   (rollout_generator: results do not depend on `world`)

and the only exchange that can ever be wanted is ONE gather of the finished uint8 frames (22.1 MB per 32-frame
rollout) or latents to rank 0 after sampling (gather_rollouts: torch.distributed.all_gather over NCCL/NVLink on
GPUs, gloo in the CPU tests).  It is outside the timed region of bench.py.
"""
from __future__ import annotations

import torch


def shard_rollouts(n_rollouts: int, rank: int, world: int) -> list[int]:
    """Rollout ids owned by `rank`: r with r % world == rank (round-robin keeps per-rank counts within one)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_rollouts, world))


def rollout_generators(ids, device, base_seed: int = 1000):
    """One torch generator per owned rollout, seeded by the GLOBAL rollout id."""
    return [torch.Generator(device=device).manual_seed(base_seed + r) for r in ids]


def rollout_noise(ids, n_frames: int, frame_shape, device, base_seed: int = 1000) -> torch.Tensor:
    """N(0,1) draws for the owned rollouts, [len(ids), n_frames, *frame_shape]: rollout r's noise comes from its own
    generator (seed base + r, drawn frame by frame like generate.py:201), so it is identical whatever the world size
    and whichever rank owns r."""
    gens = rollout_generators(ids, device, base_seed)
    out = torch.empty((len(ids), n_frames, *frame_shape), dtype=torch.float32, device=device)
    for j, g in enumerate(gens):
        for f in range(n_frames):
            out[j, f] = torch.randn(frame_shape, device=device, generator=g)
    return out


def gather_rollouts(local: torch.Tensor, n_rollouts: int, rank: int, world: int, group=None) -> torch.Tensor | None:
    """local: [len(shard_rollouts(n, rank, world)), ...] results of this rank's rollouts -> on rank 0 the tensor
    [n_rollouts, ...] in global rollout order; None elsewhere.  One all_gather of equally padded shards."""
    if world == 1:
        return local
    import torch.distributed as dist
    per = (n_rollouts + world - 1) // world
    pad = torch.zeros((per, *local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    if rank != 0:
        return None
    out = torch.empty((n_rollouts, *local.shape[1:]), dtype=local.dtype, device=local.device)
    for rk in range(world):
        ids = shard_rollouts(n_rollouts, rk, world)
        if ids:
            out[ids] = parts[rk][: len(ids)]
    return out
