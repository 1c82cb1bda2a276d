"""Env sharding across GPUs: contiguous env blocks per rank and the single observation all-gather.

The reference makes no collective call at all (Isaac Lab's ``--distributed`` launches one process per GPU and the RL library
all-reduces gradients); envs are fully independent, so the only cross-GPU traffic this engine adds is ONE all-gather of the
final observation tensor per step, over NCCL / NVLink (gloo in the CPU tests)."""

from __future__ import annotations

import torch
import torch.distributed as dist


def env_shard(num_envs: int, rank: int, world: int) -> tuple[int, int]:
    """[start, stop) of the contiguous env block of ``rank``; blocks differ by at most one env."""
    base, rem = divmod(num_envs, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_gather_obs(local: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """Gathers equally sized env shards (dim 0) from every rank into one tensor, rank-major = env order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if out is None:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out
