"""Golden numbers for the training step from the UNMODIFIED reference train.py
(TrainPipeline.policy_update, CPU fp32 branch, train.py:186-207) on a deterministic 64-sample batch:
   python scripts/make_golden.py train
The batch is rebuilt identically in the tests by ``train_batch()`` (oracle-only ingredients)."""
from __future__ import annotations

import json
import os
import re
import tempfile

import numpy as np
import torch

from oracle import load_reference, mcts_oracle, replay_oracle
from oracle import cchess_shim as cs
from tests import positions


def train_batch(n_moves: int = 32):
    """(states f16 (2n,17,7,10,9), probs f64 (2n,2086), winners f32 (2n,)) from a random playout."""
    boards = positions.random_playout_positions(1, n_moves, seed=9)[:n_moves]
    dense, turns = [], []
    for rec in boards:
        ids = mcts_oracle.legal_ids(cs.Board.from_record(rec))
        p = np.zeros(2086)
        w, _ = mcts_oracle.fake_policy_arrays(rec, "hash")
        p[ids] = w[ids].astype(np.float64)
        dense.append(p / p.sum())
        turns.append(bool(rec[90]))
    z = np.array([1.0 if t else -1.0 for t in turns])
    states, probs, winners = replay_oracle.pack_reference(boards, np.stack(dense), turns, z, "per_move")
    return states, probs, winners.astype(np.float32)


def main(golden_dir):
    ref_train = load_reference.load("train")
    states, probs, winners = train_batch()
    captured = []
    ref_train.log = lambda msg, *a, **k: captured.append(str(msg))
    with tempfile.TemporaryDirectory() as td:
        np.save(os.path.join(td, "states.npy"), states)
        np.save(os.path.join(td, "mcts.npy"), probs)
        np.save(os.path.join(td, "winners.npy"), winners)
        torch.manual_seed(0)
        torch.set_num_threads(8)
        pipe = ref_train.TrainPipeline(init_model=None)
        pipe.data_dir = td
        pipe.batch_size = len(states)
        pipe.num_workers = 0
        avg_loss, avg_entropy = pipe.policy_update()
    line = [m for m in captured if m.startswith("kl:")][-1]
    vals = {k: float(v) for k, v in re.findall(r"(\w+):(-?[\d.]+(?:e-?\d+)?)", line)}
    out = dict(source="unmodified /root/reference/train.py TrainPipeline.policy_update, CPU fp32, seed 0, one batch of 64",
               n=len(states), avg_loss=float(avg_loss), avg_entropy=float(avg_entropy), log=vals,
               lr_multiplier_after=float(pipe.lr_multiplier))
    with open(os.path.join(golden_dir, "train_reference.json"), "w") as f:
        json.dump(out, f)
    print(out)
