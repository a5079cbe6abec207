"""Lockstep batched MCTS: G independent trees, one playout per tree per step.

Device counterpart of the reference's ``MCTS`` class (mcts.py:81-178) for thousands of concurrent
games.  One lockstep step = for every game: select (K3) -> movegen+encode of the leaf (K1) ->
evaluate the whole leaf batch once -> expand + backup (K4/K5).  Exactly one leaf per game per step,
no virtual loss, so each tree goes through the same sequence of states as the reference's
sequential playouts.

An *evaluator* is any callable ``evaluator(planes, leaf_boards) -> (policy, policy_kind, values)``
with ``planes`` the bf16 (G,17,7,10,9) net input (``None`` when the evaluator has an attribute
``needs_planes == False``), ``leaf_boards`` the (G,96) uint8 records,
``policy`` float32 (G,2086) (probabilities or logits, see ``_lib.POLICY_*``) and ``values``
float32 (G,), all device tensors.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


class LockstepSearch:
    def __init__(self, n_games: int, node_cap: int = 32768, device="cuda", c_puct: float = 5.0):
        self.n_games = int(n_games)
        self.node_cap = int(node_cap)
        self.device = torch.device(device)
        self.c_puct = float(c_puct)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().ccz_init(), "ccz_init")
        # two arenas: advance() compacts the kept sub-tree from one into the other
        self._arenas = [_lib.Arena(n_games, node_cap, self.device), _lib.Arena(n_games, node_cap, self.device)]
        self._cur = 0
        self._graphs = None
        self._graph_evaluator = None
        g, dev = self.n_games, self.device
        self.leaf_boards = torch.zeros((g, _lib.BOARD_BYTES), dtype=torch.uint8, device=dev)
        self.leaf_nodes = torch.zeros((g,), dtype=torch.int32, device=dev)
        self.move_ids = torch.zeros((g, _lib.MAX_MOVES), dtype=torch.int16, device=dev)
        self.counts = torch.zeros((g,), dtype=torch.int16, device=dev)
        self.flags = torch.zeros((g,), dtype=torch.uint8, device=dev)
        self.planes = torch.zeros((g, 17, 7, 10, 9), dtype=torch.bfloat16, device=dev)
        self.root_acts = torch.zeros((g, _lib.MAX_MOVES), dtype=torch.int16, device=dev)
        self.root_visit_counts = torch.zeros((g, _lib.MAX_MOVES), dtype=torch.int32, device=dev)
        self.root_counts = torch.zeros((g,), dtype=torch.int16, device=dev)
        self.reset()

    # ------------------------------------------------------------------------------------
    @property
    def arena(self) -> _lib.Arena:
        return self._arenas[self._cur]

    @property
    def root_boards(self) -> torch.Tensor:
        return self.arena.root_boards

    def reset(self, mask=None) -> None:
        """Games back to the start position with a fresh root (mcts.py:94, game.py:148); ``mask``
        (uint8 per game) restricts the reset to finished slots."""
        if mask is not None:
            mask = torch.as_tensor(mask, dtype=torch.uint8).to(self.device).contiguous()
        _lib.mcts_reset(self.arena, mask)

    def set_roots(self, records) -> None:
        """History-less root positions from (G,96) board records; trees are dropped."""
        rec = torch.as_tensor(np.ascontiguousarray(records), dtype=torch.uint8).to(self.device)
        if rec.shape != (self.n_games, _lib.BOARD_BYTES):
            raise ValueError(f"records must be ({self.n_games},96)")
        a = self.arena
        _lib.mcts_reset(a)
        a.root_boards.copy_(rec)
        a.root_keys.copy_(_lib.board_keys_init(a.root_boards))

    # ------------------------------------------------------------------------------------
    def select_and_encode(self, planes: bool = True) -> None:
        a = self.arena
        _lib.mcts_select(a, self.c_puct, self.leaf_boards, self.leaf_nodes)
        _lib.movegen_encode(self.leaf_boards, planes=planes,
                            out=(self.move_ids, self.counts, self.flags, self.planes if planes else None))

    def expand_backup(self, policy: torch.Tensor, policy_kind: int, values: torch.Tensor) -> None:
        _lib.mcts_expand_backup(self.arena, self.leaf_nodes, policy, policy_kind, values, self.move_ids,
                                self.counts, self.flags)

    def step(self, evaluator) -> None:
        """One playout in every game (MCTS.playout, mcts.py:101-129)."""
        # an evaluator that works from the board records (net.BatchedEvaluator with K10) spares K1 the planes
        need = getattr(evaluator, "needs_planes", True)
        self.select_and_encode(planes=need)
        policy, kind, values = evaluator(self.planes if need else None, self.leaf_boards)
        self.expand_backup(policy, kind, values)

    def run(self, evaluator, n_playout: int) -> None:
        if self._graphs is not None and self._graph_evaluator is evaluator and n_playout > 0:
            done = 0
            graph = self._graphs.get(self._cur)
            if graph is None:
                # the first playout of this run doubles as the eager warm-up (lazy inits, cuDNN plans);
                # capture itself records the launches without executing them
                self.step(evaluator)
                done = 1
                torch.cuda.synchronize(self.device)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self.step(evaluator)
                self._graphs[self._cur] = graph
            for _ in range(n_playout - done):
                graph.replay()
            return
        for _ in range(n_playout):
            self.step(evaluator)

    # ---- CUDA graphs: one captured lockstep step per arena side --------------------------------
    def enable_graphs(self, evaluator) -> None:
        """Capture K3 -> K1 -> evaluator -> K4/K5 into a CUDA graph (one per ping-pong arena, captured
        on first use) and replay it in ``run``.  The evaluator must be capture-safe and static
        (device-only work, no host synchronisation), e.g. ``net.BatchedEvaluator``.  It removes the
        per-step launch overhead (~170 launches), which matters when the leaf batch is small."""
        self._graphs = {}
        self._graph_evaluator = evaluator

    def disable_graphs(self) -> None:
        self._graphs = None
        self._graph_evaluator = None

    # ------------------------------------------------------------------------------------
    def root_visits(self):
        """(acts int16 (G,128), visits int32 (G,128), counts int16 (G,)) device tensors; children in
        generation order (mcts.py:163-164)."""
        _lib.mcts_root_visits(self.arena, self.root_acts, self.root_visit_counts, self.root_counts)
        return self.root_acts, self.root_visit_counts, self.root_counts

    def advance(self, chosen) -> None:
        """update_with_move per game (mcts.py:168-178) + board.push of the move; -1 = new game,
        -2 = keep the position but drop the tree."""
        ch = torch.as_tensor(chosen, dtype=torch.int16).to(self.device).contiguous()
        if ch.shape != (self.n_games,):
            raise ValueError("chosen must have one entry per game")
        src, dst = self._arenas[self._cur], self._arenas[1 - self._cur]
        _lib.mcts_advance(src, dst, ch)
        self._cur = 1 - self._cur

    def check_status(self) -> None:
        st = self.arena.status
        if bool((st != 0).any()):
            bad = torch.nonzero(st).flatten()[:8].tolist()
            raise _lib.CczError(
                f"MCTS arena overflow (node_cap={self.node_cap}) in games {bad}: results invalid, raise node_cap")

    def memory_bytes(self) -> int:
        return sum(a.bytes() for a in self._arenas)


def visit_softmax(visits: np.ndarray, temp: float) -> np.ndarray:
    """softmax(1/temp * log(visits + 1e-10)) in float64 (mcts.py:165, tools.py:126-129)."""
    x = 1.0 / temp * np.log(np.asarray(visits) + 1e-10)
    p = np.exp(x - np.max(x))
    p /= np.sum(p)
    return p
