"""Self-play -> train loop on one GPU (BASELINE config 5): GPU-generated games feed the
PolicyValueNet training step, and the updated weights go straight back into the lockstep evaluator.

The reference runs ``collect.py`` and ``train.py`` as two processes that meet on disk
(``data/data.h5`` -> ``convert.py`` -> npy triple; ``models/current_policy.pkl``; README.md:23-48).
Here both halves share one ``PolicyValueNet`` in one process: ``CollectPipeline`` (lockstep games,
K8 packing, replay files in the reference's layouts) -> ``TrainPipeline.policy_update`` (reference
semantics, bf16) -> ``BatchedEvaluator.refresh`` (BN re-folded) -> next batch of games.  The on-disk
hand-off files are still written so either half can be swapped for the reference's own script.
"""
from __future__ import annotations

import os

from .collect import CollectPipeline
from .parameters import MODEL_DIR
from .train import TrainPipeline


class SelfPlayTrainLoop:
    def __init__(self, n_games=4096, n_playout=400, data_dir="data", model_dir=MODEL_DIR, batch_size=512,
                 games_per_iteration=None, seed=0, net_kwargs=None, max_game_moves=None, node_cap=None,
                 states_mode="per_move"):
        self.collect = CollectPipeline(n_games=n_games, n_playout=n_playout, data_dir=data_dir, seed=seed,
                                       net_kwargs=net_kwargs, max_game_moves=max_game_moves, node_cap=node_cap,
                                       states_mode=states_mode)
        self.collect.load_model()
        self.train = TrainPipeline(data_dir=data_dir, batch_size=batch_size, net_kwargs=net_kwargs)
        # one network for both halves
        self.train.policy_value_net = self.collect.policy_value_net
        self.games_per_iteration = games_per_iteration or n_games
        self.model_dir = model_dir
        self.iterations = 0

    def iterate(self):
        """Collect >= games_per_iteration finished games, run one training pass, refresh the evaluator."""
        target = self.collect.iters + self.games_per_iteration
        while self.collect.iters < target:
            self.collect.collect_data()
        self.train.dataset = None                      # drop the old memory maps, then re-open the grown npy triple
        self.collect.flush()
        samples = self.collect.npy.rows
        loss, entropy = self.train.policy_update()
        pv = self.collect.policy_value_net
        pv.policy_value_net.eval()
        pv.evaluator().refresh(pv.policy_value_net)    # weights back into the lockstep forward
        os.makedirs(self.model_dir, exist_ok=True)
        pv.save_model(os.path.join(self.model_dir, "current_policy.pkl"))
        self.iterations += 1
        return {"games": self.collect.iters, "samples": samples, "loss": loss, "entropy": entropy,
                "lr_multiplier": self.train.lr_multiplier}

    def close(self):
        self.collect.close()
