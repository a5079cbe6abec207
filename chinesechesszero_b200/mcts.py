"""Reference-shaped façade of ``mcts.py`` (mcts.py:81-233) for ONE game over the CUDA arena.

``MCTS`` / ``MCTS_AI`` keep the reference's constructor arguments, attributes and method names
(``get_move_probs``, ``update_with_move``, ``get_action``, ``reset_player``, ``set_player_idx``,
``.mcts.n_playout``, ``.agent``).  The tree lives in a one-game ``search.LockstepSearch``; move
selection uses the global NumPy RNG exactly like the reference (mcts.py:218-227) so a seeded run
draws the same moves.  For throughput use ``selfplay.SelfPlayEngine`` (thousands of games).

``policy_value_fn`` may be
  * the bound ``PolicyValueNet.policy_value_fn`` of this package (its bf16 batched evaluator is used),
  * a device evaluator ``f(planes, leaf_boards) -> (policy, kind, values)`` marked with
    ``f.device_evaluator = True``,
  * or a host callable ``f(board) -> (iterable[(id, prob)], value)`` taking a ``board.Board``-like
    view (slow: one host round trip per playout, as in the reference).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .board import Board
from .search import LockstepSearch, visit_softmax

C_PUCT, EPS, ALPHA = 5, 0.25, 0.2  # parameters.py:8-12


class _LeafView:
    """What a host policy callable sees of a leaf: record(), turn, legal ids."""

    def __init__(self, rec, ids):
        self._rec, self._ids = rec, ids

    def record(self):
        return self._rec

    @property
    def turn(self):
        return bool(self._rec[90])

    def legal_ids(self):
        return self._ids


def _as_evaluator(policy_value_fn):
    owner = getattr(policy_value_fn, "__self__", None)
    if owner is not None and hasattr(owner, "evaluator") and callable(owner.evaluator):
        return owner.evaluator()
    if getattr(policy_value_fn, "device_evaluator", False):
        return policy_value_fn

    def host_evaluator(planes, leaf_boards, move_ids=None, counts=None):
        recs = leaf_boards.cpu().numpy()
        pol = np.zeros((recs.shape[0], _lib.N_ACTIONS), dtype=np.float32)
        val = np.zeros((recs.shape[0],), dtype=np.float32)
        for i, rec in enumerate(recs):
            act_probs, v = policy_value_fn(_LeafView(rec, None))
            for a, p in act_probs:
                pol[i, a] = p
            val[i] = np.asarray(v, dtype=np.float32).reshape(-1)[0]
        dev = leaf_boards.device
        return torch.from_numpy(pol).to(dev), _lib.POLICY_PROBS, torch.from_numpy(val).to(dev)

    return host_evaluator


class MCTS:
    def __init__(self, policy_value_fn, c_puct=5, n_playout=10000, node_cap=None, device="cuda"):
        self.policy = policy_value_fn
        self._evaluator = _as_evaluator(policy_value_fn)
        self.c_puct = c_puct
        self.n_playout = n_playout
        if node_cap is None:
            # one game owns the whole page pool; it starts at twice the worst case of one search and doubles
            # whenever the kept sub-tree (mcts.py:168-178 keeps it without bound) leaves less than that free
            node_cap = max(65536, 2 * _lib.search_pages(int(n_playout), 11) << 11)
        self._search = LockstepSearch(1, nodes_per_game=node_cap, device=device, c_puct=float(c_puct))
        from .net import BatchedEvaluator

        if isinstance(self._evaluator, BatchedEvaluator):
            # batch-1 playouts are launch-bound: replay the captured step (K3, K1, forward, K4/K5)
            self._search.enable_graphs(self._evaluator)

    def _sync_root(self, board: Board) -> None:
        """Point the arena's root at ``board``; the tree is kept when the position is the one the
        previous ``update_with_move`` led to (tree reuse, mcts.py:168-175)."""
        a = self._search.arena
        same = bool(torch.equal(a.root_boards[0], board._board[0])) and bool(torch.equal(a.root_keys[0], board._keys[0]))
        if not same:
            self._search.set_roots(board.record()[None])
            self._search.arena.root_keys.copy_(board._keys)

    def get_move_probs(self, board, temp=1e-3, red_states=None, black_states=None, on_playout=None):
        """mcts.py:131-166: n_playout playouts, then softmax(1/temp*log(visits+1e-10)) over the root
        children in generation order.  Returns (acts tuple, probs float64 array)."""
        self._sync_root(board)
        interval = max(1, self.n_playout // 100)  # progress throttle, mcts.py:148-160
        remaining = self.n_playout
        while remaining > 0:
            chunk = min(interval, remaining) if on_playout is not None else remaining
            self._search.run(self._evaluator, chunk)
            remaining -= chunk
            if on_playout is not None:
                try:
                    on_playout(chunk)
                except Exception:
                    pass
        self._search.check_status()
        acts, visits, counts = self._search.root_visits()
        n = int(counts[0])
        acts = tuple(int(x) for x in acts[0, :n].cpu().numpy())
        visits = visits[0, :n].cpu().numpy()
        return acts, visit_softmax(visits, temp)

    def update_with_move(self, last_move):
        """mcts.py:168-178: keep the chosen child's sub-tree, or start over (-1 / unknown move)."""
        mv = int(last_move)
        if mv < 0:
            self._search.advance(np.array([-2], dtype=np.int16))
        else:
            self._search.advance(np.array([mv], dtype=np.int16))


class MCTS_AI:
    def __init__(self, policy_value_fn, c_puct=5, n_playout=2000, is_selfplay=False, **kw):
        self.mcts = MCTS(policy_value_fn, c_puct, n_playout, **kw)
        self.is_selfplay = is_selfplay
        self.agent = "AI"

    def set_player_idx(self, p):
        self.player = p

    def reset_player(self):
        self.mcts.update_with_move(-1)

    def get_action(self, board, temp=1e-3, return_prob=False, on_playout=None):
        """mcts.py:203-233.  Returns the action id (and the 2086-vector of un-noised visit
        probabilities when ``return_prob``)."""
        move_probs = np.zeros(_lib.N_ACTIONS)
        acts, probs = self.mcts.get_move_probs(board, temp, on_playout=on_playout)
        move_probs[list(acts)] = probs
        if self.is_selfplay:
            move = np.random.choice(acts, p=(1 - EPS) * probs + EPS * np.random.dirichlet(ALPHA * np.ones(len(probs))))
            self.mcts.update_with_move(move)
        else:
            move = np.random.choice(acts, p=probs)
            self.mcts.update_with_move(-1)
        if return_prob:
            return move, move_probs
        return move
