"""Synthetic reachable positions generated ON THE DEVICE with the product kernels (no oracle):
breadth-first expansion from the start position -- K1 for the legal moves of a whole level, K2
(``ccz_board_push``) for the children.  Used by bench.py for the configs[1] workload (~1M perft-3/4
positions) and as a size-independent check: the level sizes are the published Xiangqi perft numbers
44 / 1,920 / 79,666 / 3,290,240 / 133,312,995."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def expand_level(boards: torch.Tensor):
    """All children of ``boards`` [n,96] in generation order -> ([m,96] uint8, counts [n] int16)."""
    ids, counts, _, _ = _lib.movegen_encode(boards, planes=False)
    c = counts.to(torch.int64)
    parents = torch.repeat_interleave(torch.arange(boards.shape[0], device=boards.device), c)
    first = torch.cumsum(c, 0) - c
    slot = torch.arange(parents.shape[0], device=boards.device) - first[parents]
    moves = ids[parents, slot].contiguous()
    children = boards[parents].contiguous()
    _lib.board_push(children, moves, None)
    return children, counts


def perft_levels(depth: int, device="cuda"):
    """[level_1, ..., level_depth] board tensors from the start position."""
    level = _lib.boards_start(1, device)
    out = []
    for _ in range(depth):
        level, _ = expand_level(level)
        out.append(level)
    return out


def perft_count(depth: int, device="cuda", root=None) -> int:
    """perft(depth) = number of leaf nodes: expand depth-1 levels, sum the move counts of the last.  ``root``: a
    96-byte board record (default: the start position)."""
    if root is None:
        level = _lib.boards_start(1, device)
    else:
        level = torch.as_tensor(np.ascontiguousarray(root, dtype=np.uint8).reshape(1, _lib.BOARD_BYTES)).to(device)
    for _ in range(depth - 1):
        level, _ = expand_level(level)
    _, counts, _, _ = _lib.movegen_encode(level, planes=False)
    return int(counts.to(torch.int64).sum())


def bench_positions(n_target: int = 1 << 20, seed: int = 0, device="cuda") -> torch.Tensor:
    """configs[1] workload: all 79,666 perft-3 leaves + a seeded uniform sample of perft-4 leaves."""
    levels = perft_levels(4, device)
    l3, l4 = levels[2], levels[3]
    if n_target <= l3.shape[0]:
        return l3[:n_target].contiguous()
    pick = np.sort(np.random.default_rng(seed).choice(l4.shape[0], size=n_target - l3.shape[0], replace=False))
    return torch.cat([l3, l4[torch.from_numpy(pick).to(l4.device)]]).contiguous()
