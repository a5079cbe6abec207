"""Game-loop parity (game.py:133-237): the lockstep SelfPlayEngine in deterministic mode (no noise,
first most-visited move) vs the same procedure run on the CPU with the oracle search and the shim
board: identical moves, visit distributions, sides to move, winner and z for whole games."""
import numpy as np
import pytest
import torch

from oracle import cchess_shim as cs
from oracle import mcts_oracle
from tests import positions
from tests.test_mcts_gpu import fake_evaluator

pytestmark = pytest.mark.gpu


def oracle_selfplay(root_record, kind, n_playout, max_moves=400):
    """game.py:148-237 + mcts.py:203-233 in deterministic mode, on the CPU oracle."""
    board = cs.Board.from_record(root_record)
    search = mcts_oracle.FlatMCTS(mcts_oracle.make_policy(kind), c_puct=5, n_playout=n_playout)
    pis, turns, moves = [], [], []
    move_count = 0
    while True:
        move_count += 1
        temp = 1.0 if move_count <= 30 else max(0.1, 1.0 * 0.5)
        acts, probs = search.get_move_probs(board, temp)
        move_probs = np.zeros(2086)
        move_probs[list(acts)] = probs
        move = int(acts[int(np.argmax(probs))])
        search.update_with_move(move)
        move_probs = move_probs / np.sum(move_probs)
        pis.append(move_probs)
        turns.append(board.turn)
        moves.append(move)
        board.push(mcts_oracle.move_from_id(move))
        tie = board.is_insufficient_material() or board.is_fourfold_repetition() or board.is_sixty_moves()
        if board.is_game_over() or tie or move_count >= max_moves:
            outcome = board.outcome() if board.is_game_over() else None
            z = np.zeros(len(turns))
            winner = outcome.winner if outcome else None
            if winner is not None:
                z = np.array([1.0 if t == winner else -1.0 for t in turns])
            return moves, pis, turns, winner, z


ROOTS = [
    ("mate_in_one", "3k5/9/9/9/9/9/9/9/4R4/R4K3 w", 0),
    ("sixty", None, 116),
    ("endgame", "4k4/4a4/9/9/4p4/9/9/4C4/4A4/3K1R3 w", 100),
    ("rook_ending", "3k5/9/9/9/9/9/9/9/9/R3K4 b", 90),
]


def test_deterministic_games_match_oracle():
    from chinesechesszero_b200.selfplay import SelfPlayEngine

    recs = []
    for _, fen, clock in ROOTS:
        r = cs.start_record() if fen is None else positions.record_from_fen(fen)
        r[91] = clock
        recs.append(r)
    recs = np.stack(recs)
    n_playout = 60
    eng = SelfPlayEngine(fake_evaluator("hash"), n_games=len(recs), n_playout=n_playout, deterministic=True,
                         node_cap=32768, max_game_moves=40)
    eng.search.set_roots(recs)
    done = {}
    for _ in range(80):
        for rec in eng.play_move():
            done.setdefault(rec.slot, rec)
        if len(done) == len(recs):
            break
    assert len(done) == len(recs), f"only {sorted(done)} finished"
    for g, rec in done.items():
        moves, pis, turns, winner, z = oracle_selfplay(recs[g], "hash", n_playout, max_moves=40)
        assert rec.moves.tolist() == moves, ROOTS[g][0]
        assert rec.turns.tolist() == turns
        assert rec.winner == winner
        assert np.array_equal(rec.z, z)
        for i in range(len(moves)):
            dense = np.zeros(2086)
            dense[rec.acts[i]] = rec.probs[i]
            assert np.array_equal(dense, pis[i]), (ROOTS[g][0], i)


def test_noise_mode_runs_and_is_seed_reproducible():
    from chinesechesszero_b200.selfplay import SelfPlayEngine

    outs = []
    for _ in range(2):
        eng = SelfPlayEngine(fake_evaluator("hash"), n_games=6, n_playout=24, seed=11, node_cap=8192)
        eng.play(3)
        outs.append([eng.current_moves(g) for g in range(6)])
    assert outs[0] == outs[1]
    assert len({tuple(m) for m in outs[0]}) > 1  # noise makes the slots diverge
