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
    """``nodes_per_game`` is the AVERAGE node budget of a game: all trees draw pages of ``2**page_shift``
    nodes from one pool of ``n_games * nodes_per_game`` nodes (24 B each), so a game that keeps a large
    sub-tree (forced replies keep almost everything, mcts.py:168-178) borrows what the others do not
    need.  The pool grows (``ensure_capacity``: a bigger pool, trees migrated) up to ``max_pool_nodes``
    when a caller that synchronises anyway asks for it; the device-side guard ``ccz_mcts_reserve`` runs
    before every search and, as the last resort, drops the sub-trees of the largest games -- counted in
    ``pool_stats()['trees_dropped']`` -- so that no expansion can fail.  ``node_cap`` is the round-1
    name of ``nodes_per_game``."""

    def __init__(self, n_games: int, nodes_per_game: int | None = None, device="cuda", c_puct: float = 5.0,
                 page_shift: int = 11, max_pool_nodes: int | None = None, node_cap: int | None = None,
                 max_pages_per_game: int | None = None):
        self.n_games = int(n_games)
        if nodes_per_game is None:
            nodes_per_game = node_cap if node_cap is not None else 65536
        self.nodes_per_game = int(nodes_per_game)
        self.page_shift = int(page_shift)
        self.device = torch.device(device)
        self.c_puct = float(c_puct)
        page = 1 << self.page_shift
        hard = ((1 << 31) - 1) >> self.page_shift  # pool indices are int32
        self.max_pool_pages = hard if max_pool_nodes is None else max(self.n_games, min(hard, int(max_pool_nodes) // page))
        n_pages = min(self.max_pool_pages, max(self.n_games, -(-self.n_games * self.nodes_per_game // page)))
        self._max_pages_per_game = max_pages_per_game
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().ccz_init(), "ccz_init")
        self._arena = _lib.Arena(self.n_games, n_pages, self.page_shift, self._game_pages(n_pages), self.device)
        self.pool_grown = 0
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

    def _game_pages(self, n_pages: int) -> int:
        """Page-list capacity of one game: the whole pool for small pools, 8192 pages (16.8 M nodes) otherwise."""
        if self._max_pages_per_game is not None:
            return int(self._max_pages_per_game)
        return max(2, min(n_pages, 8192))

    # ------------------------------------------------------------------------------------
    @property
    def arena(self) -> _lib.Arena:
        return self._arena

    @property
    def node_cap(self) -> int:
        return self.nodes_per_game

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

    # ---- pool capacity ---------------------------------------------------------------------
    def search_pages(self, n_playout: int) -> int:
        return _lib.search_pages(n_playout, self.page_shift)

    def _grow(self, n_pages: int) -> bool:
        """Replace the pool by one of ``n_pages`` pages and move every tree over (ccz_mcts_migrate)."""
        n_pages = min(int(n_pages), self.max_pool_pages)
        old = self._arena
        if n_pages <= old.n_pages:
            return False
        need = (n_pages << self.page_shift) * _lib.NODE_BYTES + (64 << 20)
        if torch.cuda.mem_get_info(self.device)[0] < need:
            if not getattr(self, "_warned_no_memory", False):
                self._warned_no_memory = True
                import warnings

                warnings.warn(f"MCTS page pool cannot grow to {n_pages} pages ({need >> 20} MiB needed): from here on the "
                              "device-side guard drops the largest kept sub-trees when the pool runs short "
                              "(pool_stats()['trees_dropped'] counts them)")
            return False
        new = _lib.Arena(self.n_games, n_pages, self.page_shift, self._game_pages(n_pages), self.device)
        _lib.mcts_migrate(old, new)
        self._arena = new
        self.pool_grown += 1
        if self._graphs is not None:  # captured graphs hold the old pool's pointers
            torch.cuda.current_stream(self.device).synchronize()
            self._graphs = {}
        return True

    def ensure_capacity(self, n_playout: int, ctl: np.ndarray | None = None, ctl_lag: int = 0) -> None:
        """Make room for a search of ``n_playout`` playouts in every game.  Geometry first (the pool must
        hold the worst case of a fresh tree per game); then, if ``ctl`` (a host copy of ``arena.pool_ctl``
        taken after the last advance) shows fewer free pages than the worst case, the pool doubles; a
        snapshot that is ``ctl_lag`` searches old is charged the worst case of those searches as well."""
        w = self.search_pages(n_playout)
        floor = self.n_games * (w + 1)
        pages_at_snapshot = self.arena.n_pages
        if self.arena.n_pages < floor or self.arena.max_pages < w + 1:
            ok = floor <= self.max_pool_pages and self._grow(min(max(floor, 2 * self.arena.n_pages), self.max_pool_pages))
            if not ok or self.arena.max_pages < w + 1:
                raise _lib.CczError(
                    f"MCTS pool too small: {self.n_games} games x {n_playout} playouts need at least {floor} pages of "
                    f"{1 << self.page_shift} nodes ({w + 1} per game), the pool has {self.arena.n_pages} "
                    f"(raise nodes_per_game / max_pool_nodes)")
        if ctl is not None:
            # pages added by the geometric growth above are all free
            free = int(ctl[_lib.CTL_TAIL] - ctl[_lib.CTL_HEAD]) + self.arena.n_pages - pages_at_snapshot
            need = self.n_games * w * (1 + int(ctl_lag))
            if free < need:
                used = self.arena.n_pages - free
                self._grow(max(2 * self.arena.n_pages, used + need))

    def pool_stats(self) -> dict:
        st = self.arena.pool_stats()
        st["pool_grown"] = self.pool_grown
        st["pool_bytes"] = (self.arena.n_pages << self.page_shift) * _lib.NODE_BYTES
        return st

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

    def run(self, evaluator, n_playout: int, ctl: np.ndarray | None = None, may_sync: bool = True,
            ctl_lag: int = 0) -> None:
        """``n_playout`` playouts in every game.  Capacity first: ``ctl`` is a host copy of
        ``arena.pool_ctl`` taken after the last advance (SelfPlayEngine reads it back with the move's other
        results); without it the counters are read here (one 64-byte device-to-host copy) unless
        ``may_sync=False`` (the device-resident path: no host synchronisation; it passes the newest snapshot
        that has already arrived, ``ctl_lag`` searches old, and the pool grows that much earlier)."""
        if n_playout <= 0:
            return
        if ctl is None and may_sync:
            ctl = self.arena.pool_ctl.cpu().numpy()
        self.ensure_capacity(n_playout, ctl, ctl_lag)
        # device-side guard: after it no expansion of this search can fail (see ccz_mcts_reserve)
        _lib.mcts_reserve(self.arena, self.search_pages(n_playout))
        if self._graphs is not None and self._graph_evaluator is evaluator and n_playout > 0:
            done = 0
            graph = self._graphs.get(0)
            if graph is None:
                # the first playout of this run doubles as the eager warm-up (lazy inits, cuDNN plans);
                # capture itself records the launches without executing them
                self.step(evaluator)
                done = 1
                torch.cuda.synchronize(self.device)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self.step(evaluator)
                self._graphs[0] = graph
            for _ in range(n_playout - done):
                graph.replay()
            return
        for _ in range(n_playout):
            self.step(evaluator)

    # ---- CUDA graphs: one captured lockstep step per arena side --------------------------------
    def enable_graphs(self, evaluator) -> None:
        """Capture K3 -> K1 -> evaluator -> K4/K5 into a CUDA graph (captured on first use, again after
        the pool has grown) and replay it in ``run``.  The evaluator must be capture-safe and static
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
        _lib.mcts_advance(self.arena, ch)

    def check_status(self) -> None:
        """Raises when a leaf was left unexpanded (pool exhausted in spite of the guard) -- the only event
        that takes a search off the reference's sequence without restarting it.  Dropped sub-trees are
        reported by ``pool_stats()`` and the per-game status bits, not raised."""
        st = self.arena.status
        if bool(((st & _lib.STATUS_EXPAND_FAILED) != 0).any()):
            bad = torch.nonzero(st & _lib.STATUS_EXPAND_FAILED).flatten()[:8].tolist()
            raise _lib.CczError(
                f"MCTS page pool exhausted ({self.arena.n_pages} pages of {1 << self.page_shift} nodes) in games {bad}: "
                "leaves were left unexpanded, raise nodes_per_game / max_pool_nodes")

    def memory_bytes(self) -> int:
        return self.arena.bytes()


def visit_softmax(visits: np.ndarray, temp: float) -> np.ndarray:
    """softmax(1/temp * log(visits + 1e-10)) in float64 (mcts.py:165, tools.py:126-129)."""
    x = 1.0 / temp * np.log(np.asarray(visits) + 1e-10)
    p = np.exp(x - np.max(x))
    p /= np.sum(p)
    return p
