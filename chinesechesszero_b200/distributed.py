"""Multi-GPU plumbing for the self-play path: games shard per GPU, nothing else is shared.

One process per GPU (torchrun / torch.distributed), each with its own arena, weight replica and RNG
stream; there is NO collective on the hot path.  The helpers here cover the little
that is global: the rank's seed, disjoint ``game_{k}`` indices so per-rank replay shards merge into
the reference's numbering (collect.py:146-167), and the barrier / max-over-ranks timing that
``bench.py`` needs.  Backend-agnostic (``nccl`` on GPUs, ``gloo`` in the CPU tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_info():
    """(rank, local_rank, world_size) from the torchrun environment (1 process => (0, 0, 1))."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init(backend: str | None = None, device=None) -> tuple[int, int, int]:
    rank, local_rank, world = shard_info()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, **kw)
    return rank, local_rank, world


def rank_seed(base_seed: int, rank: int) -> int:
    """Independent, reproducible noise stream per rank."""
    return int(base_seed) + 1_000_003 * int(rank)


def global_game_index(local_index: int, rank: int, world: int, start: int = 0) -> int:
    """Disjoint game numbers: rank r owns start + r, start + r + world, ..."""
    return int(start) + int(rank) + int(world) * int(local_index)


def barrier() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device="cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device="cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(units_this_rank: float, elapsed_ms_this_rank: float, device="cpu") -> float:
    """Whole-job units/s: all ranks' units over the slowest rank's time."""
    total = sum_over_ranks(units_this_rank, device)
    worst = max_over_ranks(elapsed_ms_this_rank, device)
    return total / worst * 1e3


def shutdown() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
