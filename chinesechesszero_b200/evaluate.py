"""Batched model-vs-model evaluation play: the lockstep counterpart of ``Game.start_play``
(game.py:77-130) with two ``MCTS_AI`` players (mcts.py:181-233, ``is_selfplay=False``).

Per game the reference does: RED = player1, BLACK = player0; the side to move runs ``n_playout``
playouts on a FRESH tree (non-self-play ``get_action`` ends with ``update_with_move(-1)``,
mcts.py:226-227), takes ``move ~ Choice(acts, p = softmax(log(N+1e-10)/temp))`` with temp = 1e-3
(the most-visited move up to ties), pushes it, and the game ends on ``board.is_game_over()``
(game.py:118-130; ``is_tie`` is NOT consulted in start_play); winner = outcome().winner, -1 = draw.

Here G games run in lockstep on one arena.  Games [0, G/2) have evaluator A as RED, games [G/2, G)
have evaluator B as RED, so every ply evaluates one half of the leaf batch with A and the other with
B (``SplitEvaluator``); all games are at the same ply, nothing is refilled, finished slots idle on the
start position and are ignored.  ``train.py:314-319``'s ``policy_evaluate`` is a stub in the reference;
``win_ratio`` below is what it was meant to return (wins + draws/2 over games).
"""
from __future__ import annotations

import dataclasses

import numpy as np
import torch

from . import _lib
from .search import LockstepSearch, visit_softmax
from .tools import outcome_winner_flags


class SplitEvaluator:
    """Evaluator protocol of ``LockstepSearch`` over two nets: ``first`` evaluates leaves [0, split),
    ``second`` leaves [split, G).  Both must return logits (``net.BatchedEvaluator``)."""

    def __init__(self, first, second, split: int):
        self.first, self.second, self.split = first, second, int(split)
        self.needs_planes = getattr(first, "needs_planes", True) or getattr(second, "needs_planes", True)

    def swapped(self) -> "SplitEvaluator":
        return SplitEvaluator(self.second, self.first, self.split)

    def __call__(self, planes, leaf_boards):
        k = self.split
        p0 = None if planes is None else planes[:k]
        p1 = None if planes is None else planes[k:]
        l0, k0, v0 = self.first(p0, leaf_boards[:k])
        l1, k1, v1 = self.second(p1, leaf_boards[k:])
        if k0 != k1:
            raise _lib.CczError("SplitEvaluator: both evaluators must return the same policy kind")
        return torch.cat([l0, l1]), k0, torch.cat([v0, v1])


@dataclasses.dataclass
class MatchResult:
    wins_a: int
    wins_b: int
    draws: int
    unfinished: int          # games cut by max_plies (counted as draws in win_ratio)
    plies: np.ndarray        # (G,) plies played per game
    winners: list            # per game: "a", "b", None (draw / unfinished)
    moves: list              # per game: action ids played

    @property
    def games(self) -> int:
        return self.wins_a + self.wins_b + self.draws + self.unfinished

    @property
    def win_ratio(self) -> float:
        """Score of A: (wins + draws / 2) / games -- the quantity train.py's ``policy_evaluate`` stub stands for."""
        return (self.wins_a + 0.5 * (self.draws + self.unfinished)) / max(1, self.games)


class EvaluationMatch:
    def __init__(self, evaluator_a, evaluator_b, n_games: int, n_playout: int = 400, c_puct: float = 5.0,
                 temp: float = 1e-3, node_cap: int | None = None, device="cuda", seed: int = 0,
                 deterministic: bool = False, max_plies: int = 400):
        if n_games < 2 or n_games % 2:
            raise ValueError("n_games must be even: each evaluator plays RED in half of the games")
        self.n_games, self.n_playout = int(n_games), int(n_playout)
        self.temp, self.deterministic, self.max_plies = float(temp), bool(deterministic), int(max_plies)
        if node_cap is None:  # fresh tree every move: n_playout expansions of <= 119 children
            node_cap = max(4096, self.n_playout * 64 + 256)
        self.search = LockstepSearch(n_games, node_cap=node_cap, device=device, c_puct=c_puct)
        self.device = self.search.device
        half = self.n_games // 2
        self._red_to_move = SplitEvaluator(evaluator_a, evaluator_b, half)   # RED's ply: A in [0,half), B in [half,G)
        self._black_to_move = self._red_to_move.swapped()
        self.rng = np.random.default_rng(seed)
        g = self.n_games
        self._flag_out = (torch.empty((g, _lib.MAX_MOVES), dtype=torch.int16, device=self.device),
                          torch.empty((g,), dtype=torch.int16, device=self.device),
                          torch.empty((g,), dtype=torch.uint8, device=self.device), None)

    def _choose(self, acts, visits, counts) -> np.ndarray:
        """mcts.py:165 + 225: move ~ Choice(acts, p = softmax(log(N + 1e-10) / temp)); deterministic = first max."""
        chosen = np.empty(self.n_games, dtype=np.int16)
        for g in range(self.n_games):
            n = int(counts[g])
            if n <= 0:
                raise _lib.CczError(f"game {g}: root has no children after search")
            p = visit_softmax(visits[g, :n], self.temp)
            k = int(np.argmax(p)) if self.deterministic else int(self.rng.choice(n, p=p))
            chosen[g] = acts[g, k]
        return chosen

    def play(self) -> MatchResult:
        s, g, half = self.search, self.n_games, self.n_games // 2
        s.reset()
        alive = np.ones(g, dtype=bool)
        winners: list = [None] * g
        finished = np.zeros(g, dtype=bool)
        plies = np.zeros(g, dtype=np.int64)
        moves: list[list[int]] = [[] for _ in range(g)]
        a_is_red = np.arange(g) < half
        for ply in range(self.max_plies):
            if not alive.any():
                break
            red_to_move = ply % 2 == 0  # every live game is at the same ply from the start position
            s.run(self._red_to_move if red_to_move else self._black_to_move, self.n_playout)
            s.check_status()
            acts, visits, counts = (t.cpu().numpy() for t in s.root_visits())
            chosen = self._choose(acts, visits, counts)
            chosen[~alive] = -1  # idle slots restart from the start position and are ignored
            s.advance(torch.from_numpy(chosen).to(self.device))
            # non-self-play players drop their tree after every move (mcts.py:226-227)
            s.advance(torch.full((g,), -2, dtype=torch.int16, device=self.device))
            _lib.movegen_encode(s.root_boards, planes=False, out=self._flag_out)
            flags = self._flag_out[2].cpu().numpy()
            turn_red = s.root_boards[:, 90].cpu().numpy() != 0
            for i in np.nonzero(alive)[0]:
                moves[i].append(int(chosen[i]))
                plies[i] += 1
                fl = int(flags[i])
                if fl & (_lib.FLAG_TIE_MASK | _lib.FLAG_NOMOVES):  # board.is_game_over(), game.py:118
                    w = outcome_winner_flags(fl, bool(turn_red[i]))
                    winners[i] = None if w is None else ("a" if bool(w) == bool(a_is_red[i]) else "b")
                    finished[i] = True
                    alive[i] = False
            if (~alive).any():  # park finished slots on the start position
                s.reset(torch.from_numpy((~alive).astype(np.uint8)).to(self.device))
        wa = sum(1 for i in range(g) if finished[i] and winners[i] == "a")
        wb = sum(1 for i in range(g) if finished[i] and winners[i] == "b")
        dr = sum(1 for i in range(g) if finished[i] and winners[i] is None)
        return MatchResult(wins_a=wa, wins_b=wb, draws=dr, unfinished=int((~finished).sum()), plies=plies,
                           winners=winners, moves=moves)


def policy_evaluate(net_a, net_b, n_games: int = 64, n_playout: int = 400, **kw) -> float:
    """What ``TrainPipeline.policy_evaluate`` (train.py:314-319, a stub returning 0.6) stands for: the score
    of ``net_a`` against ``net_b`` (``net.PolicyValueNet`` objects) over ``n_games`` lockstep games."""
    match = EvaluationMatch(net_a.evaluator(), net_b.evaluator(), n_games=n_games, n_playout=n_playout, **kw)
    return match.play().win_ratio
