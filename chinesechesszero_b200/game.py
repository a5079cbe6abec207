"""Reference-shaped façade of ``game.py`` (game.py:10-237) for ONE game on the device board.

``Game.start_self_play(player, is_shown=False, temp=1.0, game_index=None)`` follows the reference
loop line by line -- temperature schedule (game.py:159), ``get_action`` (game.py:178),
renormalisation (game.py:187-190), sample recording and history update BEFORE the push
(game.py:196-201), terminal test (game.py:208), z (game.py:213-219) -- and returns the same list of
``(red_states, black_states, mcts_prob, winner_z)`` tuples, including the reference's aliasing of the
two history lists across all samples (game.py:234-237).
``Game.start_play(player1, player0, is_shown=False)`` is the two-player loop of game.py:77-130
(the batched form is ``evaluate.EvaluationMatch``).  Board drawing (``graphic``) is out of scope.
"""
from __future__ import annotations

import numpy as np

from . import tools
from .board import RED, Board, is_tie


class Game:
    def __init__(self, board=None):
        self.board = board if board is not None else Board()
        self.red_states = None
        self.black_states = None
        self.reset_states_history()

    def reset_states_history(self):
        """game.py:23-35: eight copies of the initial position's planes per side."""
        init_red, init_black = tools.decode_board(self.board)
        self.red_states = [init_red.copy() for _ in range(8)]
        self.black_states = [init_black.copy() for _ in range(8)]

    def update_states_history(self):
        """game.py:37-44: most recent first."""
        red, black = tools.decode_board(self.board)
        self.red_states.pop()
        self.red_states.insert(0, red)
        self.black_states.pop()
        self.black_states.insert(0, black)

    def start_play(self, player1, player0, is_shown=False, max_moves=None):
        """game.py:77-130: player1 = RED moves first; returns the winner (True = RED, False = BLACK) or -1
        for a draw.  Like the reference, only ``board.is_game_over()`` ends the game."""
        self.board = Board(device=self.board.device)
        player1.set_player_idx(1)
        player0.set_player_idx(0)
        players = {RED: player1, (not RED): player0}
        n = 0
        while True:
            move = players[self.board.turn].get_action(self.board)
            self.update_states_history()
            self.board.push(int(move))
            n += 1
            if self.board.is_game_over():
                outcome = self.board.outcome()
                return outcome.winner if outcome.winner is not None else -1
            if max_moves is not None and n >= max_moves:
                return -1

    def start_self_play(self, player, is_shown=False, temp=1.0, game_index=None, max_moves=None):
        self.board = Board(device=self.board.device)
        self.reset_states_history()
        mcts_probs, current_players = [], []
        move_count = 0
        while True:
            move_count += 1
            current_temp = temp if move_count <= 30 else max(0.1, temp * 0.5)
            move, move_probs = player.get_action(self.board, temp=current_temp, return_prob=True)
            prob_sum = np.sum(move_probs)
            if prob_sum > 0:
                move_probs = move_probs / prob_sum
            else:
                continue
            mcts_probs.append(move_probs)
            current_players.append(self.board.turn)
            self.update_states_history()
            self.board.push(int(move))
            over = self.board.is_game_over()
            if over or is_tie(self.board) or (max_moves is not None and move_count >= max_moves):
                outcome = self.board.outcome() if over else None
                winner_z = np.zeros(len(current_players))
                if outcome and outcome.winner is not None:
                    for i, player_id in enumerate(current_players):
                        winner_z[i] = 1 if player_id == outcome.winner else -1
                player.reset_player()
                return [(self.red_states, self.black_states, mcts_probs[i], winner_z[i])
                        for i in range(len(mcts_probs))]
