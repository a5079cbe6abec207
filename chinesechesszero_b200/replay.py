"""Replay output of the self-play path.

Turns finished games (``selfplay.GameRecord``) into the arrays the reference stores per game
(collect.py:64-169): ``states (2T,17,7,10,9) float16``, ``mcts_probs (2T,2086) float64``,
``winners (2T,) float64`` -- first T rows the game's samples, next T their file-mirrored twins
(collect.py:131) -- with the densification done on the device by K8 (``ccz_replay_pack``).

``states_mode``
  "reference": byte-for-byte what the reference saves: every sample of a game
               carries the FINAL 8-deep history (game.py:234-237 aliasing) and the turn plane is
               all ones (collect.py:28,78-81);
  "per_move":  sample i carries the history as of move i (slot 0 = the position searched, most
               recent first, padded with the initial position) and the true side-to-move plane.

``NpyReplayWriter`` appends to the ``states.npy / mcts.npy / winners.npy / meta.json`` layout that
the reference's trainer actually reads (convert.py:85-99, train.py:95-100).
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import _lib


def history_boards(rec, states_mode: str = "per_move"):
    """(hist [T,8,96] uint8, turn_plane [T] uint8) for a GameRecord."""
    boards = rec.boards
    t = boards.shape[0]
    idx = np.arange(t)[:, None] - np.arange(8)[None, :]          # slot k = position searched at move i-k
    hist = boards[np.clip(idx, 0, None)]                          # before move 0: the initial position
    if states_mode == "reference":
        # history after the last update_states_history(): slot 0 = position before the last move
        hist = np.broadcast_to(hist[t - 1], (t, 8, _lib.BOARD_BYTES)).copy()
        turn = np.ones(t, dtype=np.uint8)
    elif states_mode == "per_move":
        turn = rec.turns.astype(np.uint8)
    else:
        raise ValueError(f"unknown states_mode {states_mode!r}")
    return np.ascontiguousarray(hist), turn


def pack_game(rec, states_mode: str = "per_move", device="cuda"):
    """GameRecord -> (states f16 (2T,17,7,10,9), mcts_probs f64 (2T,2086), winners f64 (2T,)) NumPy,
    computed by ccz_replay_pack on the device."""
    t = len(rec)
    hist, turn = history_boards(rec, states_mode)
    acts = np.full((t, _lib.MAX_MOVES), -1, dtype=np.int16)
    probs = np.zeros((t, _lib.MAX_MOVES), dtype=np.float64)
    counts = np.zeros(t, dtype=np.int16)
    for i, (a, p) in enumerate(zip(rec.acts, rec.probs)):
        counts[i] = len(a)
        acts[i, : len(a)] = a
        probs[i, : len(a)] = p
    dev = torch.device(device)
    states, pi = _lib.replay_pack(torch.from_numpy(hist).to(dev), torch.from_numpy(turn).to(dev),
                                  torch.from_numpy(acts).to(dev), torch.from_numpy(probs).to(dev),
                                  torch.from_numpy(counts).to(dev))
    winners = np.concatenate([rec.z, rec.z]).astype(np.float64)
    return states.cpu().numpy(), pi.cpu().numpy(), winners


class NpyReplayWriter:
    """Accumulates packed games and writes the npy triple + meta.json (convert.py:85-99 dtypes:
    states float16, mcts float64, winners float32)."""

    def __init__(self, out_dir: str):
        self.out_dir = out_dir
        self._states, self._mcts, self._winners = [], [], []
        self.games = 0

    def add(self, states, mcts_probs, winners):
        self._states.append(np.asarray(states, dtype=np.float16))
        self._mcts.append(np.asarray(mcts_probs, dtype=np.float64))
        self._winners.append(np.asarray(winners, dtype=np.float32))
        self.games += 1

    def flush(self):
        os.makedirs(self.out_dir, exist_ok=True)
        states = np.concatenate(self._states) if self._states else np.zeros((0, 17, 7, 10, 9), np.float16)
        mcts = np.concatenate(self._mcts) if self._mcts else np.zeros((0, _lib.N_ACTIONS), np.float64)
        winners = np.concatenate(self._winners) if self._winners else np.zeros((0,), np.float32)
        np.save(os.path.join(self.out_dir, "states.npy"), states)
        np.save(os.path.join(self.out_dir, "mcts.npy"), mcts)
        np.save(os.path.join(self.out_dir, "winners.npy"), winners)
        meta = {  # same keys as convert.py:89-97
            "total_count": int(states.shape[0]),
            "states_shape": list(states.shape), "states_dtype": str(states.dtype),
            "mcts_shape": list(mcts.shape), "mcts_dtype": str(mcts.dtype),
            "winners_shape": list(winners.shape), "winners_dtype": str(winners.dtype),
        }
        with open(os.path.join(self.out_dir, "meta.json"), "w", encoding="utf-8") as f:
            json.dump(meta, f, ensure_ascii=False, indent=2)
        return states.shape[0]
