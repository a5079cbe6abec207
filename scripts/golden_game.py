"""Golden vectors for the game loop and the replay post-processing, produced by the UNMODIFIED
reference game.py (Game.start_self_play), mcts.py (MCTS_AI.get_action) and collect.py
(CollectPipeline.preprocess / flip_data) driven by the cchess shim and the deterministic "hash"
stand-in policy, with the global NumPy RNG seeded:   python scripts/make_golden.py game
"""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np

from oracle import cchess_shim as cs
from oracle import load_reference, mcts_oracle

N_PLAYOUT, SEED = 24, 7


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main(golden_dir):
    ref_mcts, ref_game, ref_collect = load_reference.load("mcts", "game", "collect")
    base = mcts_oracle.make_policy("hash")

    def policy_value_fn(board, red_states=None, black_states=None):
        ids, probs, value = base(board)
        return zip(ids, probs), np.array([[value]], dtype=np.float32)

    np.random.seed(SEED)
    ai = ref_mcts.MCTS_AI(policy_value_fn, c_puct=5, n_playout=N_PLAYOUT, is_selfplay=True)
    game = ref_game.Game(cs.Board())
    play_data = game.start_self_play(ai, is_shown=False)
    moves = [m.uci() for m in game.board.move_stack]
    z = [float(d[3]) for d in play_data]
    probs = np.stack([d[2] for d in play_data])
    pipe = ref_collect.CollectPipeline.__new__(ref_collect.CollectPipeline)
    pipe.board = cs.Board()  # collect.py:28: the pipeline's own never-moved board
    data = pipe.flip_data(pipe.preprocess(list(play_data)))
    states = np.array([s for s, _, _ in data])
    mcts_probs = np.array([p for _, p, _ in data])
    winners = np.array([w for _, _, w in data])
    out = dict(source="unmodified reference game.py / mcts.py / collect.py with oracle.cchess_shim",
               n_playout=N_PLAYOUT, seed=SEED, moves=moves, z=z, probs_sha=sha(probs),
               states_shape=list(states.shape), states_dtype=str(states.dtype), states_sha=sha(states),
               mcts_probs_shape=list(mcts_probs.shape), mcts_probs_dtype=str(mcts_probs.dtype),
               mcts_probs_sha=sha(mcts_probs), winners=winners.tolist(),
               final_fen=game.board.fen(), episode_len=pipe.episode_len)
    path = os.path.join(golden_dir, "game_reference.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("game length", len(moves), "z[0]", z[0], "states", states.shape, states.dtype, "probs", mcts_probs.shape,
          mcts_probs.dtype, "final", game.board.fen())
