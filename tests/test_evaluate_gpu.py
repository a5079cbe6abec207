"""Evaluation play (game.py:77-130 with two non-self-play MCTS_AI players, mcts.py:203-233): the lockstep
``EvaluationMatch`` in deterministic mode against the same two-player loop on the CPU oracle (flat search +
shim board): identical move sequences, lengths and winners; A plays RED in the first half of the games and
BLACK in the second."""
import numpy as np
import pytest
import torch

from oracle import cchess_shim as cs
from oracle import mcts_oracle
from tests.test_mcts_gpu import fake_evaluator

pytestmark = pytest.mark.gpu


def oracle_match_game(kind_red, kind_black, n_playout, max_plies):
    board = cs.Board()
    players = {True: mcts_oracle.FlatMCTS(mcts_oracle.make_policy(kind_red), c_puct=5, n_playout=n_playout),
               False: mcts_oracle.FlatMCTS(mcts_oracle.make_policy(kind_black), c_puct=5, n_playout=n_playout)}
    moves = []
    for _ in range(max_plies):
        search = players[board.turn]
        acts, probs = search.get_move_probs(board, 1e-3)     # mcts.py:208
        move = int(acts[int(np.argmax(probs))])               # Choice(p) at temp 1e-3 = the most visited move
        search.update_with_move(-1)                           # mcts.py:226-227: fresh tree every move
        moves.append(move)
        board.push(mcts_oracle.move_from_id(move))
        if board.is_game_over():                              # game.py:118
            return moves, board.outcome().winner
    return moves, "unfinished"


def test_match_equals_the_reference_two_player_loop():
    from chinesechesszero_b200.evaluate import EvaluationMatch

    n_playout, max_plies = 24, 10
    m = EvaluationMatch(fake_evaluator("hash"), fake_evaluator("uniform"), n_games=4, n_playout=n_playout,
                        deterministic=True, max_plies=max_plies, node_cap=8192)
    res = m.play()
    a_red, _ = oracle_match_game("hash", "uniform", n_playout, max_plies)
    b_red, _ = oracle_match_game("uniform", "hash", n_playout, max_plies)
    assert res.moves[0] == a_red and res.moves[1] == a_red      # A = RED in games [0, G/2)
    assert res.moves[2] == b_red and res.moves[3] == b_red      # B = RED in games [G/2, G)
    assert res.games == 4 and res.unfinished == 4 and res.win_ratio == 0.5


def test_match_runs_to_the_end_and_scores_like_the_oracle():
    """Whole games with the bf16 evaluators of two small random nets: every move is legal, the recorded result
    is the oracle's outcome of the replayed game."""
    from chinesechesszero_b200.evaluate import EvaluationMatch
    from chinesechesszero_b200.net import BatchedEvaluator, Net

    torch.manual_seed(0)
    ea = BatchedEvaluator(Net(num_channels=32, resblocks_num=1).cuda().eval())
    eb = BatchedEvaluator(Net(num_channels=32, resblocks_num=1).cuda().eval())
    m = EvaluationMatch(ea, eb, n_games=8, n_playout=8, seed=1, max_plies=60)
    res = m.play()
    assert res.games == 8 and res.wins_a + res.wins_b + res.draws + res.unfinished == 8
    for g in range(8):
        board = cs.Board()
        for mv in res.moves[g]:
            assert mv in mcts_oracle.legal_ids(board)
            assert not board.is_game_over()
            board.push(mcts_oracle.move_from_id(mv))
        assert len(res.moves[g]) == res.plies[g]
        if board.is_game_over():
            w = board.outcome().winner
            a_is_red = g < 4
            expect = None if w is None else ("a" if bool(w) == a_is_red else "b")
            assert res.winners[g] == expect
        else:
            assert res.plies[g] == 60 and res.winners[g] is None


def test_start_play_facade_returns_a_winner_or_draw():
    from chinesechesszero_b200.game import Game
    from chinesechesszero_b200.mcts import MCTS_AI
    from chinesechesszero_b200.net import PolicyValueNet

    torch.manual_seed(0)
    np.random.seed(0)
    net = PolicyValueNet(num_channels=32, resblocks_num=1)
    p1 = MCTS_AI(net.policy_value_fn, c_puct=5, n_playout=6)
    p0 = MCTS_AI(net.policy_value_fn, c_puct=5, n_playout=6)
    w = Game().start_play(p1, p0, is_shown=False, max_moves=12)
    assert w in (True, False, -1) and p1.player == 1 and p0.player == 0
