"""Pin the oracle's game loop and replay post-processing (oracle/collect_oracle-style loop,
oracle/replay_oracle.py) to the UNMODIFIED reference game.py / mcts.py / collect.py through
tests/golden/game_reference.json (a seeded 406-move self-play game with the "hash" stand-in policy)."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import cchess_shim as cs
from oracle import mcts_oracle, replay_oracle

EPS, ALPHA = 0.25, 0.2


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load(golden_dir):
    with open(os.path.join(golden_dir, "game_reference.json")) as f:
        return json.load(f)


def oracle_seeded_game(n_playout, seed):
    """game.py:148-237 + mcts.py:203-224 with the global NumPy RNG, on the oracle search."""
    np.random.seed(seed)
    board = cs.Board()
    search = mcts_oracle.FlatMCTS(mcts_oracle.make_policy("hash"), c_puct=5, n_playout=n_playout)
    boards, probs_dense, turns, moves = [], [], [], []
    move_count = 0
    while True:
        move_count += 1
        temp = 1.0 if move_count <= 30 else max(0.1, 1.0 * 0.5)
        acts, probs = search.get_move_probs(board, temp)
        move_probs = np.zeros(2086)
        move_probs[list(acts)] = probs
        move = int(np.random.choice(acts, p=(1 - EPS) * probs + EPS * np.random.dirichlet(ALPHA * np.ones(len(probs)))))
        search.update_with_move(move)
        move_probs = move_probs / np.sum(move_probs)
        boards.append(board.record())
        probs_dense.append(move_probs)
        turns.append(board.turn)
        moves.append(move)
        board.push(mcts_oracle.move_from_id(move))
        tie = board.is_insufficient_material() or board.is_fourfold_repetition() or board.is_sixty_moves()
        if board.is_game_over() or tie:
            outcome = board.outcome() if board.is_game_over() else None
            z = np.zeros(len(turns))
            if outcome and outcome.winner is not None:
                z = np.array([1.0 if t == outcome.winner else -1.0 for t in turns])
            return np.stack(boards), np.stack(probs_dense), turns, moves, z, board


@pytest.fixture(scope="module")
def game(golden_dir):
    gold = load(golden_dir)
    return gold, oracle_seeded_game(gold["n_playout"], gold["seed"])


def test_oracle_game_equals_reference(game):
    gold, (boards, probs, turns, moves, z, board) = game
    ucis = [mcts_oracle.move_from_id(m).uci() for m in moves]
    assert ucis == gold["moves"]
    assert z.tolist() == gold["z"]
    assert sha(probs) == gold["probs_sha"]
    assert board.fen() == gold["final_fen"]


def test_replay_oracle_equals_reference_preprocess_and_flip(game):
    gold, (boards, probs, turns, moves, z, board) = game
    states, mcts_probs, winners = replay_oracle.pack_reference(boards, probs, turns, z, "reference")
    assert list(states.shape) == gold["states_shape"] and str(states.dtype) == gold["states_dtype"]
    assert sha(states) == gold["states_sha"]
    assert sha(mcts_probs) == gold["mcts_probs_sha"]
    assert winners.tolist() == gold["winners"]
