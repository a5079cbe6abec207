"""Reference-shaped façade of ``collect.py`` (collect.py:26-198): the self-play data pipeline.

``CollectPipeline(init_model).run()`` keeps the reference's shape -- load the model (random init when
the file is missing, collect.py:49-56), play, preprocess + mirror (collect.py:64-131), append to the
replay store, bump the game counter -- but plays ``n_games`` games concurrently in lockstep on the
GPU (``selfplay.SelfPlayEngine``) and lets K8 do the densification.  Games are appended in the order
they finish; each finished game becomes one ``game_{k}`` unit exactly like one reference iteration.

Output: the ``states.npy / mcts.npy / winners.npy / meta.json`` layout the reference's trainer reads
(train.py:95-100) via ``replay.NpyReplayWriter``; the ``data.h5`` container itself (collect.py:146-167)
is the next row of SURVEY.md §8(f) -- no HDF5 library exists in this image.
"""
from __future__ import annotations

import argparse
import os

from . import replay
from .net import PolicyValueNet
from .parameters import C_PUCT, DATA_DIR, MODEL_DIR, PLAYOUT
from .selfplay import SelfPlayEngine


class CollectPipeline:
    def __init__(self, init_model=None, n_games=4096, n_playout=PLAYOUT, c_puct=C_PUCT, data_dir=DATA_DIR,
                 states_mode="reference", seed=0, node_cap=None, max_game_moves=None, net_kwargs=None):
        self.temp = 1.0
        self.n_playout = n_playout
        self.c_puct = c_puct
        self.init_model = init_model
        self.n_games = n_games
        self.iters = 0
        self.episode_len = 0
        self.policy_value_net = None
        self.engine = None
        self.data_dir = data_dir
        self.states_mode = states_mode
        self.seed = seed
        self.node_cap = node_cap
        self.max_game_moves = max_game_moves
        self.net_kwargs = net_kwargs or {}
        self.writer = replay.NpyReplayWriter(data_dir)

    def load_model(self):
        """collect.py:47-62: load once; fall back to random init when the model cannot be loaded."""
        if self.policy_value_net is None:
            model_path = self.init_model if self.init_model else os.path.join(MODEL_DIR, "current_policy.pkl")
            try:
                self.policy_value_net = PolicyValueNet(model=model_path, **self.net_kwargs)
            except Exception:
                self.policy_value_net = PolicyValueNet(**self.net_kwargs)
            self.engine = SelfPlayEngine(self.policy_value_net.evaluator(), n_games=self.n_games,
                                         n_playout=self.n_playout, c_puct=self.c_puct, temp=self.temp,
                                         seed=self.seed, node_cap=self.node_cap,
                                         max_game_moves=self.max_game_moves)

    def collect_data(self, is_shown=False):
        """One lockstep move in every game slot; every game that finishes is packed (preprocess +
        flip, collect.py:141-142) and appended.  Returns the running game count (collect.py:176)."""
        self.load_model()
        for rec in self.engine.play_move():
            states, probs, winners = replay.pack_game(rec, self.states_mode, device=self.engine.device)
            self.writer.add(states, probs, winners)
            self.episode_len = len(rec)
            self.iters += 1
        return self.iters

    def run(self, is_shown=False, max_games=None):
        try:
            while max_games is None or self.iters < max_games:
                self.collect_data(is_shown=is_shown)
        except KeyboardInterrupt:
            pass
        finally:
            self.writer.flush()
        return self.iters


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="collect Xiangqi self-play data on the GPU")
    parser.add_argument("--show", action="store_true", default=False)
    parser.add_argument("--model", type=str, default="current_policy.pkl")
    parser.add_argument("--games", type=int, default=4096)
    parser.add_argument("--max-games", type=int, default=None)
    args = parser.parse_args()
    CollectPipeline(init_model=args.model, n_games=args.games).run(is_shown=args.show, max_games=args.max_games)
