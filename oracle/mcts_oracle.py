"""mcts_oracle -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; never imported by the product path).

A restatement of the reference's MCTS (/root/reference/mcts.py) over a flat node table, with the
reference's exact NumPy-2 numerics, so that it can run on the GPU box (where /root/reference does
not exist) as the checker for the CUDA arena kernels.  Pinned against the UNMODIFIED reference
``mcts.py`` by tests/golden/mcts_*.json (written by scripts/golden_mcts.py in the authoring
container) -- see tests/test_mcts_oracle.py.

Reference lines restated:
  Node.puct_value / select   mcts.py:41-61   -> FlatMCTS._select_child
  Node.expand                mcts.py:31-39   -> FlatMCTS._expand
  Node.update(_recursive)    mcts.py:63-78   -> FlatMCTS._backup
  MCTS.playout               mcts.py:101-129 -> FlatMCTS.playout
  MCTS.get_move_probs        mcts.py:131-166 -> FlatMCTS.get_move_probs
  MCTS.update_with_move      mcts.py:168-178 -> FlatMCTS.update_with_move
The board is any cchess-like object (oracle.cchess_shim.Board); the policy is
``policy(board) -> (ids list[int], probs np.float32[len(ids)], value np.float32)``.
"""
from __future__ import annotations

import numpy as np

from . import cchess_shim as cs

_id_of = None


def move_from_id(action: int) -> cs.Move:
    global _id_of
    if _id_of is None:
        _id_of = cs.action_table()
    _, fr, to = _id_of
    return cs.Move(int(fr[action]), int(to[action]))


def legal_ids(board) -> list[int]:
    id_of = cs.action_table()[0]
    return [int(id_of[m.from_square, m.to_square]) for m in board.legal_moves]


class FlatMCTS:
    def __init__(self, policy, c_puct=5, n_playout=400):
        self.policy = policy
        self.c_puct = c_puct
        self.n_playout = n_playout
        self.reset()

    def reset(self):
        # node table (struct of lists): the root is Node(None, 1.0) (mcts.py:94)
        self.N = [0]
        self.Q = [np.float32(0.0)]
        self.P = [np.float32(1.0)]
        self.move = [-1]
        self.parent = [-1]
        self.children = [None]  # list of child node indices in insertion order, None = leaf
        self.root = 0

    # mcts.py:41-61 -- +inf for unvisited, fp32 c_puct*P, fp64 sqrt/mul/div/add, first max wins
    def _select_child(self, node: int) -> int:
        sq = np.sqrt(self.N[node])  # np.sqrt(int) -> float64
        best, best_c = None, -1
        for c in self.children[node]:
            if self.N[c] == 0:
                score = float("inf")
            else:
                cp = np.float32(self.c_puct * self.P[c])
                score = np.float64(self.Q[c]) + cp * sq / (1 + self.N[c])
            if best is None or score > best:
                best, best_c = score, c
        return best_c

    def _expand(self, node: int, ids, probs):
        kids = []
        for a, p in zip(ids, probs):
            self.N.append(0)
            self.Q.append(np.float32(0.0))
            self.P.append(np.float32(p))
            self.move.append(int(a))
            self.parent.append(node)
            self.children.append(None)
            kids.append(len(self.N) - 1)
        self.children[node] = kids

    # mcts.py:63-78,129 -- the leaf gets -v, its parent +v, ...; Q += 1.0*(x-Q)/N in float32
    def _backup(self, node: int, leaf_value):
        x = np.float32(-leaf_value)
        cur = node
        while cur >= 0:
            self.N[cur] += 1
            q = self.Q[cur]
            self.Q[cur] = np.float32(q + np.float32(np.float32(x - q) / np.float32(self.N[cur])))
            x = np.float32(-x)
            cur = self.parent[cur]

    def playout(self, board):
        node = self.root
        while self.children[node] is not None:
            node = self._select_child(node)
            board.push(move_from_id(self.move[node]))
        ids, probs, value = self.policy(board)
        end = board.is_game_over()
        tie = (board.is_insufficient_material() or board.is_fourfold_repetition() or board.is_sixty_moves())
        if not end and not tie:
            self._expand(node, ids, probs)
            leaf_value = np.float32(value)
        elif end and tie:
            leaf_value = np.float32(0.0)
        else:
            winner = cs.RED if board.outcome().winner else cs.BLACK
            leaf_value = np.float32(1.0 if winner == board.turn else -1.0)
        self._backup(node, leaf_value)

    def root_children(self):
        kids = self.children[self.root] or []
        return [self.move[c] for c in kids], [self.N[c] for c in kids], [self.Q[c] for c in kids]

    def get_move_probs(self, board, temp=1e-3):
        for _ in range(self.n_playout):
            self.playout(board.copy())
        acts, visits, _ = self.root_children()
        x = 1.0 / temp * np.log(np.array(visits) + 1e-10)
        probs = np.exp(x - np.max(x))
        probs /= np.sum(probs)
        return tuple(acts), probs

    def update_with_move(self, last_move: int):
        kids = self.children[self.root] or []
        for c in kids:
            if self.move[c] == last_move:
                self.root = c
                self.parent[c] = -1
                return
        self.reset()


# ---- deterministic stand-in policies shared by the golden generator and all parity tests -----

def _mix(h: int) -> int:
    h &= 0xFFFFFFFFFFFFFFFF
    h ^= h >> 33
    h = (h * 0xFF51AFD7ED558CCD) & 0xFFFFFFFFFFFFFFFF
    h ^= h >> 33
    h = (h * 0xC4CEB9FE1A85EC53) & 0xFFFFFFFFFFFFFFFF
    h ^= h >> 33
    return h


def record_hash(rec: np.ndarray) -> int:
    """64-bit FNV-1a over squares + turn of a 96-byte board record (integer-only, portable)."""
    h = 0xCBF29CE484222325
    for b in np.asarray(rec, dtype=np.uint8)[:91].tolist():
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


_ACT = np.arange(1, 2087, dtype=np.uint64)


def fake_policy_arrays(rec: np.ndarray, kind: str = "hash"):
    """(probs float32[2086] summing to ~1, value float32) as a pure function of the position.

    kind "hash": pseudo-random priors and values;  kind "uniform": equal priors, value 0 (all
    PUCT scores tie, exercising first-child tie-breaking, mcts.py:59-61)."""
    if kind == "uniform":
        return np.full(2086, np.float32(1.0 / 2086), dtype=np.float32), np.float32(0.0)
    h = _mix(record_hash(rec))
    with np.errstate(over="ignore"):
        x = (_ACT * np.uint64(h | 1)) ^ np.uint64(h >> 17)
        x ^= x >> np.uint64(29)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(32)
    w = ((x >> np.uint64(40)) & np.uint64(0xFFFF)).astype(np.float64) + 1.0
    w = w ** 4  # peaky enough that priors matter
    probs = (w / w.sum()).astype(np.float32)
    value = np.float32(((h >> 11) % 2001 - 1000) / 1000.0)
    return probs, value


def make_policy(kind: str = "hash"):
    def policy(board):
        ids = legal_ids(board)
        probs, value = fake_policy_arrays(board.record(), kind)
        return ids, probs[ids], value

    return policy
