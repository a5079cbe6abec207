"""Lockstep self-play over thousands of concurrent games: the device counterpart of
``Game.start_self_play`` + ``MCTS_AI.get_action`` (game.py:133-237, mcts.py:203-233).

Every game slot follows the reference's per-game procedure:
  * temperature 1.0 for the first 30 moves, then max(0.1, temp*0.5)            (game.py:159)
  * n_playout playouts on the (reused) tree, pi = softmax(log(N+1e-10)/temp)    (mcts.py:150-166)
  * move ~ Choice(acts, p = 0.75*pi + 0.25*Dirichlet(0.2)); the noise touches the move choice
    only, the stored target is the un-noised pi                                (mcts.py:216-224)
  * the sample (pi, side to move) is recorded and the history updated BEFORE the move is pushed
    (game.py:196-201); the game ends on is_game_over() or is_tie()             (game.py:208)
  * z = +1/-1 per recorded side when outcome() has a winner, else 0            (game.py:213-219)
Noise is host-injected from a seeded NumPy Generator (one stream per engine, ``seed + rank`` in
multi-GPU runs); ``deterministic=True`` switches noise off and takes the first most-visited move,
which is the mode the parity tests compare against the reference's search.
Finished slots are refilled from the start position immediately, so all slots stay in lockstep.
"""
from __future__ import annotations

import dataclasses
from typing import Callable

import numpy as np
import torch

from . import _lib
from .search import LockstepSearch, visit_softmax
from .tools import outcome_winner_flags

EPS, ALPHA = 0.25, 0.2  # parameters.py:10-12


@dataclasses.dataclass
class GameRecord:
    """One finished game in sparse form.  ``boards[i]`` is the position searched at move i (the
    record BEFORE the i-th move was pushed); ``acts[i]``/``probs[i]`` the un-noised visit
    distribution over the root children; ``turns[i]`` the side to move (True = RED);
    ``moves[i]`` the action id played; ``winner`` True/False/None; ``z[i]`` as in game.py:213-219."""

    boards: np.ndarray          # (T, 96) uint8
    acts: list                  # T arrays int16 (n_i,)
    probs: list                 # T arrays float64 (n_i,)
    turns: np.ndarray           # (T,) bool
    moves: np.ndarray           # (T,) int16
    winner: object
    z: np.ndarray               # (T,) float64
    final_flags: int
    slot: int = -1

    def __len__(self):
        return int(self.moves.shape[0])


class SelfPlayEngine:
    def __init__(self, evaluator: Callable, n_games: int, n_playout: int = 400, c_puct: float = 5.0,
                 temp: float = 1.0, node_cap: int | None = None, device="cuda", seed: int = 0,
                 deterministic: bool = False, max_game_moves: int | None = None, use_graphs: bool = False,
                 resident_ring: int = 16, nodes_per_game: int | None = None, page_shift: int = 11,
                 max_pool_nodes: int | None = None):
        self.evaluator = evaluator
        self.n_games, self.n_playout = int(n_games), int(n_playout)
        self.temp = float(temp)
        self.deterministic = bool(deterministic)
        self.max_game_moves = max_game_moves
        self.resident_ring = int(resident_ring)
        if nodes_per_game is None:
            nodes_per_game = node_cap  # round-1 name
        if nodes_per_game is None:
            # AVERAGE budget per game of the shared page pool (search.LockstepSearch): per move a tree gains
            # n_playout expansions of ~41 (at most 119) children on top of the sub-tree kept by the last advance().
            # Measured with a near-uniform policy at 400 playouts: median live tree 23k nodes, p99 35k, and a
            # tail of forced-reply sequences beyond 65k -- which a pool absorbs because only the
            # SUM over the games has to fit (soak test: 60 moves, median peak 50k, p99 95k, max 127k).  8x the
            # mean growth per move = the worst-case reservation of every game (119 children per playout, ~51k
            # nodes at 400 playouts) on top of an 80k mean live tree; the pool doubles on demand.
            nodes_per_game = max(4096, int(self.n_playout * 41 * 8))
        self.search = LockstepSearch(n_games, nodes_per_game=nodes_per_game, device=device, c_puct=c_puct,
                                     page_shift=page_shift, max_pool_nodes=max_pool_nodes)
        if use_graphs:  # replay the lockstep step from CUDA graphs (capture-safe evaluators only)
            self.search.enable_graphs(evaluator)
        self.rng = np.random.default_rng(seed)
        self.device = self.search.device
        g = self.n_games
        self.move_count = np.zeros(g, dtype=np.int64)          # moves already played in each slot
        # per-move frames of the whole batch (boards, sparse pi, turns, moves); a slot's current game is
        # the frame range [_start[g], _n_frames)
        self._frames: list[dict] = []
        self._frame0 = 0
        self._n_frames = 0
        self._start = np.zeros(g, dtype=np.int64)
        # pinned staging for the per-move read-back / upload
        pin = dict(pin_memory=True)
        self._h_acts = torch.empty((g, _lib.MAX_MOVES), dtype=torch.int16, **pin)
        self._h_visits = torch.empty((g, _lib.MAX_MOVES), dtype=torch.int32, **pin)
        self._h_counts = torch.empty((g,), dtype=torch.int16, **pin)
        self._h_boards = torch.empty((g, _lib.BOARD_BYTES), dtype=torch.uint8, **pin)
        self._h_flags = torch.empty((g,), dtype=torch.uint8, **pin)
        self._h_chosen = torch.empty((g,), dtype=torch.int16, **pin)
        self._h_ctl = torch.zeros((_lib.CTL_WORDS,), dtype=torch.int64, **pin)
        self._ctl_valid = False
        self._d_chosen = torch.empty((g,), dtype=torch.int16, device=self.device)
        self._d_mask = torch.empty((g,), dtype=torch.uint8, device=self.device)
        self._h_mask = torch.empty((g,), dtype=torch.uint8, **pin)
        self._flag_out = (torch.empty((g, _lib.MAX_MOVES), dtype=torch.int16, device=self.device),
                          torch.empty((g,), dtype=torch.int16, device=self.device),
                          torch.empty((g,), dtype=torch.uint8, device=self.device), None)
        self.total_moves = 0
        self.total_games = 0
        self.resident_backlog: list[dict] = []   # drained rings of the resident path, oldest first
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    # ------------------------------------------------------------------------------------
    def current_temps(self) -> np.ndarray:
        """game.py:159 with move_count already incremented for the move being chosen."""
        mc = self.move_count + 1
        return np.where(mc <= 30, self.temp, max(0.1, self.temp * 0.5))

    def play_move(self) -> list[GameRecord]:
        """One lockstep move in every slot; returns the games that finished with it."""
        s = self.search
        # the pool counters read back with the last move tell whether this search fits: grow the pool (trees
        # migrate, nothing is lost) instead of letting the device-side guard drop sub-trees
        s.run(self.evaluator, self.n_playout, ctl=self._h_ctl.numpy() if self._ctl_valid else None)
        acts_d, visits_d, counts_d = s.root_visits()
        self._h_acts.copy_(acts_d, non_blocking=True)
        self._h_visits.copy_(visits_d, non_blocking=True)
        self._h_counts.copy_(counts_d, non_blocking=True)
        self._h_boards.copy_(s.root_boards, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        self.d2h_bytes += sum(t.numel() * t.element_size() for t in
                              (self._h_acts, self._h_visits, self._h_counts, self._h_boards))
        acts_np, visits_np = self._h_acts.numpy(), self._h_visits.numpy()
        counts_np, boards_np = self._h_counts.numpy().astype(np.int64), self._h_boards.numpy()
        if (counts_np <= 0).any():
            raise _lib.CczError(f"game {int(np.argmin(counts_np))}: root has no children after search")
        temps = self.current_temps()
        chosen = self._h_chosen.numpy()
        probs_pad = np.zeros((self.n_games, _lib.MAX_MOVES), dtype=np.float64)
        # games with the same child count and temperature are processed together; every row goes
        # through the reference's arithmetic unchanged (element-wise ops and contiguous row sums)
        key = counts_np * 2 + (temps != self.temp)
        for k in np.unique(key):
            rows = np.nonzero(key == k)[0]
            n, t = int(counts_np[rows[0]]), float(temps[rows[0]])
            a = acts_np[rows, :n].astype(np.int64)
            x = 1.0 / t * np.log(visits_np[rows, :n] + 1e-10)                  # mcts.py:165
            p = np.exp(x - np.max(x, axis=1, keepdims=True))                   # tools.py:126-129
            p /= np.sum(p, axis=1, keepdims=True)
            if self.deterministic:
                pick = np.argmax(p, axis=1)
            else:                                                              # mcts.py:216-222
                mix = (1 - EPS) * p + EPS * self.rng.dirichlet(ALPHA * np.ones(n), size=len(rows))
                cdf = np.cumsum(mix / np.sum(mix, axis=1, keepdims=True), axis=1)
                pick = np.minimum((cdf < self.rng.random(len(rows))[:, None]).sum(axis=1), n - 1)
            chosen[rows] = a[np.arange(len(rows)), pick]
            dense = np.zeros((len(rows), _lib.N_ACTIONS))                      # mcts.py:212-215
            dense[np.arange(len(rows))[:, None], a] = p
            probs_pad[rows, :n] = p / np.sum(dense, axis=1, keepdims=True)     # game.py:187-190
        self._frames.append(dict(boards=boards_np.copy(), acts=acts_np.copy(), probs=probs_pad,
                                 counts=counts_np, turns=boards_np[:, 90] != 0, moves=chosen.copy()))
        self._n_frames += 1
        self.move_count += 1
        self.total_moves += self.n_games
        self._d_chosen.copy_(self._h_chosen, non_blocking=True)
        self.h2d_bytes += self._h_chosen.numel() * 2
        s.advance(self._d_chosen)
        # terminal test on the new root positions (game.py:208): flags straight from K1
        _lib.movegen_encode(s.root_boards, planes=False, out=self._flag_out)
        self._h_flags.copy_(self._flag_out[2], non_blocking=True)
        self._h_boards.copy_(s.root_boards, non_blocking=True)
        self._h_ctl.copy_(s.arena.pool_ctl, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        self._ctl_valid = True
        self.d2h_bytes += self._h_flags.numel() + self._h_boards.numel() + self._h_ctl.numel() * 8
        flags = self._h_flags.numpy()
        over = (flags & (_lib.FLAG_TIE_MASK | _lib.FLAG_NOMOVES)) != 0
        if self.max_game_moves is not None:
            over |= self.move_count >= self.max_game_moves
        finished = []
        if over.any():
            mask = self._h_mask.numpy()
            mask[:] = over
            for g in np.nonzero(over)[0]:
                finished.append(self._finish(int(g), int(flags[g]), bool(self._h_boards.numpy()[g, 90])))
            self._d_mask.copy_(self._h_mask, non_blocking=True)
            self.h2d_bytes += self.n_games
            s.reset(self._d_mask)
            self.total_games += len(finished)
            self._prune_frames()
        return finished

    def _game_frames(self, g: int):
        return self._frames[int(self._start[g]) - self._frame0: self._n_frames - self._frame0]

    def current_moves(self, g: int) -> list[int]:
        """Moves played so far in slot g's current game."""
        return [int(f["moves"][g]) for f in self._game_frames(g)]

    def _finish(self, g: int, flags: int, turn_red: bool) -> GameRecord:
        winner = outcome_winner_flags(flags, turn_red) if (flags & (_lib.FLAG_TIE_MASK | _lib.FLAG_NOMOVES)) else None
        frames = self._game_frames(g)
        turns = np.array([bool(f["turns"][g]) for f in frames], dtype=bool)
        z = np.zeros(len(turns), dtype=np.float64)
        if winner is not None:
            z[:] = np.where(turns == winner, 1.0, -1.0)
        counts = [int(f["counts"][g]) for f in frames]
        rec = GameRecord(boards=np.stack([f["boards"][g] for f in frames]),
                         acts=[f["acts"][g, :n].copy() for f, n in zip(frames, counts)],
                         probs=[f["probs"][g, :n].copy() for f, n in zip(frames, counts)], turns=turns,
                         moves=np.array([f["moves"][g] for f in frames], dtype=np.int16), winner=winner, z=z,
                         final_flags=flags, slot=g)
        self._start[g] = self._n_frames
        self.move_count[g] = 0
        return rec

    def _prune_frames(self) -> None:
        keep_from = int(self._start.min())
        drop = keep_from - self._frame0
        if drop > 0:
            del self._frames[:drop]
            self._frame0 = keep_from

    # ------------------------------------------------------------------------------------
    def _resident_state(self):
        if getattr(self, "_res", None) is None:
            g, dev = self.n_games, self.device
            gen = torch.Generator(device=dev)
            gen.manual_seed(int(self.rng.integers(1 << 62)))
            self._res = dict(
                gen=gen,
                idx=torch.arange(_lib.MAX_MOVES, device=dev).view(1, -1),
                move_count=torch.zeros(g, dtype=torch.int64, device=dev),
                finished=torch.zeros((), dtype=torch.int64, device=dev),
                chosen=torch.zeros(g, dtype=torch.int16, device=dev),
                # device ring of per-move samples (what play_move() hands to the host every move):
                # positions searched, root actions, un-noised pi, counts, moves played, finished mask
                ring=self.resident_ring, head=0, filled=0, ctl_snaps=[],
                ring_boards=torch.zeros((self.resident_ring, g, _lib.BOARD_BYTES), dtype=torch.uint8, device=dev),
                ring_acts=torch.zeros((self.resident_ring, g, _lib.MAX_MOVES), dtype=torch.int16, device=dev),
                ring_pi=torch.zeros((self.resident_ring, g, _lib.MAX_MOVES), dtype=torch.float64, device=dev),
                ring_counts=torch.zeros((self.resident_ring, g), dtype=torch.int16, device=dev),
                ring_moves=torch.zeros((self.resident_ring, g), dtype=torch.int16, device=dev),
                ring_over=torch.zeros((self.resident_ring, g), dtype=torch.uint8, device=dev),
                ring_flags=torch.zeros((self.resident_ring, g), dtype=torch.uint8, device=dev),
            )
        return self._res

    def play_move_resident(self) -> None:
        """The same lockstep move with every input and output resident in HBM: visit softmax,
        seeded Dirichlet mix, sampling, advance, terminal test and slot refill all run on the device
        with no host synchronisation (the per-move sample stays in device buffers).  This is the
        path bench.py times for ``value``; ``play_move`` is the host-facing one timed for ``e2e``."""
        s, r = self.search, self._resident_state()
        g = self.n_games
        # pool counters: the snapshot enqueued one move ago has arrived by now (the host runs at most a launch
        # queue ahead of the device), the one of the last move usually has not -- no waiting either way
        ctl, lag = None, 0
        for age, (ev_done, buf) in enumerate(reversed(r["ctl_snaps"])):
            if ev_done.query():
                ctl, lag = buf.numpy(), age
                break
        grown = s.pool_grown
        s.run(self.evaluator, self.n_playout, ctl=ctl, may_sync=False, ctl_lag=lag)
        if s.pool_grown != grown:
            r["ctl_snaps"].clear()  # they describe the pool that was just replaced
        acts, visits, counts = s.root_visits()
        valid = r["idx"] < counts.view(g, 1)
        lo_temp = max(0.1, self.temp * 0.5)
        temps = torch.where(r["move_count"] < 30, self.temp, lo_temp).to(torch.float64).view(g, 1)
        x = torch.log(visits.to(torch.float64) + 1e-10) / temps
        pi = torch.softmax(x.masked_fill(~valid, float("-inf")), dim=1)
        if self.deterministic:
            choice = torch.argmax(pi, dim=1, keepdim=True)
        else:
            gam = torch._standard_gamma(torch.full((g, _lib.MAX_MOVES), ALPHA, dtype=torch.float64, device=self.device),
                                        generator=r["gen"]) * valid
            mix = (1 - EPS) * pi + EPS * gam / gam.sum(dim=1, keepdim=True).clamp_min(1e-300)
            choice = torch.multinomial(mix, 1, generator=r["gen"])
        r["chosen"].copy_(acts.gather(1, choice).view(g))
        slot = r["head"]
        r["ring_boards"][slot].copy_(s.root_boards)
        r["ring_acts"][slot].copy_(acts)
        r["ring_pi"][slot].copy_(pi)
        r["ring_counts"][slot].copy_(counts)
        r["ring_moves"][slot].copy_(r["chosen"])
        s.advance(r["chosen"])
        _lib.movegen_encode(s.root_boards, planes=False, out=self._flag_out)
        over = (self._flag_out[2] & (_lib.FLAG_TIE_MASK | _lib.FLAG_NOMOVES)) != 0
        mc = r["move_count"] + 1
        if self.max_game_moves is not None:
            over = over | (mc >= self.max_game_moves)
        self._d_mask.copy_(over.to(torch.uint8))
        r["ring_over"][slot].copy_(self._d_mask)
        r["ring_flags"][slot].copy_(self._flag_out[2])
        r["head"] = (slot + 1) % r["ring"]
        r["filled"] += 1
        if r["filled"] == r["ring"]:  # never overwrite samples: a full ring goes to the host backlog
            self.resident_backlog.append(self.drain_resident())
        s.reset(self._d_mask)
        ev_done, buf = r["ctl_snaps"].pop(0) if len(r["ctl_snaps"]) >= 3 else (torch.cuda.Event(), torch.zeros(
            (_lib.CTL_WORDS,), dtype=torch.int64, pin_memory=True))
        buf.copy_(s.arena.pool_ctl, non_blocking=True)
        ev_done.record()
        r["ctl_snaps"].append((ev_done, buf))
        r["move_count"] = torch.where(over, torch.zeros_like(mc), mc)
        r["finished"] += over.sum()
        self.total_moves += g

    def drain_resident(self) -> dict:
        """Host copies of the samples of the last ``filled`` resident moves, oldest first:
        boards (M,G,96), acts (M,G,128), pi (M,G,128) float64, counts, moves, over (slot finished with
        that move), flags (of the position after the move).  Empties the ring."""
        r = self._resident_state()
        m, ring, head = r["filled"], r["ring"], r["head"]
        order = [(head - m + i) % ring for i in range(m)]
        idx = torch.tensor(order, dtype=torch.int64, device=self.device)
        out = {k: r["ring_" + k].index_select(0, idx).cpu().numpy()
               for k in ("boards", "acts", "pi", "counts", "moves", "over", "flags")}
        r["filled"] = 0
        return out

    def pool_events(self) -> dict:
        """Cumulative pool counters (synchronises): ``trees_dropped`` and ``expand_failed`` are 0 unless
        the pool could not grow any further; see ``search.LockstepSearch``."""
        return self.search.pool_stats()

    def play(self, n_moves: int) -> list[GameRecord]:
        out = []
        for _ in range(n_moves):
            out.extend(self.play_move())
        return out
