"""reference_arm -- the UNMODIFIED reference timed on host cores (bench.py ``--impl reference`` and
``cpu_baseline``; BASELINE INFRASTRUCTURE ONLY).

Drives the reference's own classes exactly as ``collect.py`` / ``game.py`` do: ``PolicyValueNet``
(random init, what collect.py:51-56 falls back to; ``use_gpu=False`` = the CPU branch net.py:190-200),
``MCTS_AI(policy_value_fn, c_puct=5, n_playout, is_selfplay=True)`` (collect.py:57-62) and, per move,
the body of ``Game.start_self_play``'s loop (game.py:155-201): temperature schedule, ``get_action``,
normalise, ``update_states_history``, ``board.push``.  ``start_self_play`` itself only returns after a
whole game (hundreds of moves x seconds), so the loop is stepped from here.  The only stand-in is the
board: ``cchess`` is not installable, so ``oracle.cchess_shim`` (C-backed, cheaper than the pure-Python
package: the timing errs in the reference's favour) is installed under that name.
"""
from __future__ import annotations

import os
import sys

import numpy as np


def available() -> bool:
    from . import load_reference

    return load_reference.available()


class ReferenceSelfPlay:
    def __init__(self, n_playout: int = 400, threads: int | None = None, seed: int = 0):
        import torch

        from . import load_reference

        if threads:
            torch.set_num_threads(threads)
        tools, mcts, net, game = load_reference.load("tools", "mcts", "net", "game")
        import cchess  # the shim, installed by load_reference

        torch.manual_seed(seed)
        np.random.seed(seed)
        self._tools, self._cchess = tools, cchess
        self.policy_value_net = net.PolicyValueNet(use_gpu=False)
        self.player = mcts.MCTS_AI(self.policy_value_net.policy_value_fn, c_puct=5, n_playout=n_playout, is_selfplay=True)
        self.game = game.Game(cchess.Board())
        self.temp = 1.0
        self.new_game()

    def new_game(self):
        self.game.board = self._cchess.Board()          # game.py:148-149
        self.game.reset_states_history()
        self.move_count = 0
        self.mcts_probs, self.current_players = [], []

    def play_move(self):
        g = self.game
        self.move_count += 1
        current_temp = self.temp if self.move_count <= 30 else max(0.1, self.temp * 0.5)   # game.py:159
        move, move_probs = self.player.get_action(g.board, temp=current_temp, return_prob=True)  # game.py:178
        move_probs = move_probs / np.sum(move_probs)                                        # game.py:188-190
        self.mcts_probs.append(move_probs)
        self.current_players.append(g.board.turn)
        g.update_states_history()                                                           # game.py:198
        g.board.push(self._cchess.Move.from_uci(self._tools.move_id2move_action[move]))     # game.py:201
        if g.board.is_game_over() or self._tools.is_tie(g.board):                           # game.py:208
            self.player.reset_player()
            self.new_game()
            return move, True
        return move, False
