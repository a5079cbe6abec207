"""cchess_shim -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; never imported by the product path).

A clean-room stand-in for the third-party ``cchess`` package (windshadow233/python-chinese-chess,
not vendored / not pinned by the reference, README.md:21) exposing exactly the duck-typed surface the
reference touches (SURVEY.md §8c lists all call sites):

    Board(), .copy(), .push(Move), .pop(), .legal_moves, .turn, .piece_at(sq), .is_game_over(),
    .outcome().winner, .is_insufficient_material(), .is_fourfold_repetition(), .is_sixty_moves(),
    .is_check(), .is_checkmate(), .is_stalemate(), .move_stack, .peek(), .checkers(), .halfmove_clock
    Move(from_square, to_square), Move.from_uci, Move.uci, RED, BLACK

All rule arithmetic is delegated to ``oracle/xq_oracle.c`` (plain C, compiled by ``oracle/build.py``).
PARITY STATUS: move set pinned by perft KATs; generation order / outcome ordering / clock
convention are recollections of cchess ("parity unpinned", see xq_oracle.c header).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build as _build

RED = True
BLACK = False
PAWN, CANNON, ROOK, KNIGHT, BISHOP, ADVISOR, KING = range(1, 8)
PIECE_SYMBOLS = [None, "p", "c", "r", "n", "b", "a", "k"]

FLAG_CHECK, FLAG_NOMOVES, FLAG_INSUFFICIENT, FLAG_FOURFOLD, FLAG_SIXTY = 1, 2, 4, 8, 16

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.ensure_built())
        _lib.xq_perft.restype = ctypes.c_uint64
        _lib.xq_collect_leaves.restype = ctypes.c_int64
        _lib.xq_collect_leaves.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]
        _lib.xq_game_board.restype = ctypes.c_void_p
        for name in ("xq_game_init", "xq_game_copy"):
            getattr(_lib, name).argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        _lib.xq_game_push.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        _lib.xq_game_pop.argtypes = [ctypes.c_void_p]
        _lib.xq_legal_moves.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        _lib.xq_flags.argtypes = [ctypes.c_void_p, ctypes.c_int]
        _lib.xq_in_check.argtypes = [ctypes.c_void_p]
        _lib.xq_batch_movegen_encode.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 4
        _lib.xq_build_action_table.argtypes = [ctypes.c_void_p] * 3
        _lib.xq_decode_board.argtypes = [ctypes.c_void_p] * 3
        _lib.xq_encode_search_planes.argtypes = [ctypes.c_void_p] * 2
        _lib.xq_start.argtypes = [ctypes.c_void_p]
        _lib.xq_perft.argtypes = [ctypes.c_void_p, ctypes.c_int]
        _lib.xq_push.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        _lib.xq_set_order_policy.argtypes = [ctypes.c_void_p]
        _lib.xq_get_order_policy.argtypes = [ctypes.c_void_p]
    return _lib


DEFAULT_ORDER_POLICY = {"class_rank": {"p": 1, "c": 0, "r": 0, "n": 0, "b": 0, "a": 0, "k": 0},
                        "from_descending": 1, "to_descending": 1, "capture_mode": 0, "check_king_first": 0}


def order_policy_bytes(policy) -> bytes:
    """dict (see DEFAULT_ORDER_POLICY) -> the 12-byte xq_order_policy / ccz_order_policy struct."""
    ranks = [0] * 8
    for sym, r in policy["class_rank"].items():
        ranks[PIECE_SYMBOLS.index(sym)] = int(r)
    return bytes(ranks + [int(policy["from_descending"]), int(policy["to_descending"]), int(policy["capture_mode"]),
                        int(policy.get("check_king_first", 0))])


def set_order_policy(policy=None) -> None:
    """Generation order of ``legal_moves`` (None = the default, recalled cchess order)."""
    raw = None if policy is None else ctypes.create_string_buffer(order_policy_bytes(policy), 12)
    if lib().xq_set_order_policy(raw) != 0:
        raise ValueError(f"bad order policy {policy!r}")


def get_order_policy() -> dict:
    buf = ctypes.create_string_buffer(12)
    lib().xq_get_order_policy(buf)
    b = buf.raw
    return {"class_rank": {PIECE_SYMBOLS[t]: b[t] for t in range(1, 8)}, "from_descending": b[8],
            "to_descending": b[9], "capture_mode": b[10], "check_king_first": b[11]}


_FILES = "abcdefghi"


class Move:
    __slots__ = ("from_square", "to_square")

    def __init__(self, from_square: int, to_square: int):
        self.from_square = from_square
        self.to_square = to_square

    @classmethod
    def from_uci(cls, uci: str) -> "Move":
        return cls(_FILES.index(uci[0]) + 9 * int(uci[1]), _FILES.index(uci[2]) + 9 * int(uci[3]))

    def uci(self) -> str:
        f, t = self.from_square, self.to_square
        return f"{_FILES[f % 9]}{f // 9}{_FILES[t % 9]}{t // 9}"

    def __eq__(self, other):
        return isinstance(other, Move) and (self.from_square, self.to_square) == (other.from_square, other.to_square)

    def __hash__(self):
        return hash((self.from_square, self.to_square))

    def __bool__(self):
        return True

    def __repr__(self):
        return f"Move.from_uci({self.uci()!r})"


class Piece:
    __slots__ = ("piece_type", "color")

    def __init__(self, piece_type: int, color: bool):
        self.piece_type = piece_type
        self.color = color

    def symbol(self) -> str:
        s = PIECE_SYMBOLS[self.piece_type]
        return s.upper() if self.color == RED else s


class Outcome:
    __slots__ = ("termination", "winner")

    def __init__(self, termination: str, winner):
        self.termination = termination
        self.winner = winner


class Board:
    """Position + move stack.  State lives in a C ``xq_game`` struct."""

    def __init__(self, _from: "Board | None" = None):
        L = lib()
        self._g = ctypes.create_string_buffer(L.xq_game_sizeof())
        if _from is None:
            b = ctypes.create_string_buffer(96)
            L.xq_start(b)
            L.xq_game_init(self._g, b)
            self.move_stack: list[Move] = []
        else:
            L.xq_game_copy(self._g, _from._g)
            self.move_stack = list(_from.move_stack)
        self._legal = None

    # -- raw access used by our own tests -------------------------------------------------
    @classmethod
    def from_record(cls, rec) -> "Board":
        """Build a history-less board from a 96-byte device board record."""
        self = cls.__new__(cls)
        L = lib()
        self._g = ctypes.create_string_buffer(L.xq_game_sizeof())
        raw = np.ascontiguousarray(rec, dtype=np.uint8).tobytes()
        L.xq_game_init(self._g, ctypes.c_char_p(raw))
        self.move_stack = []
        self._legal = None
        return self

    def record(self) -> np.ndarray:
        """The 96-byte board record (squares, turn, clock, rep)."""
        ptr = lib().xq_game_board(self._g)
        return np.frombuffer(ctypes.string_at(ptr, 96), dtype=np.uint8).copy()

    # -- cchess surface ---------------------------------------------------------------------
    @property
    def turn(self) -> bool:
        return bool(self.record()[90])

    @property
    def halfmove_clock(self) -> int:
        return int(self.record()[91])

    def copy(self) -> "Board":
        return Board(_from=self)

    def piece_at(self, square: int):
        c = int(self.record()[square])
        return Piece(c & 7, not (c & 8)) if c else None

    def _moves(self):
        if self._legal is None:
            buf = (ctypes.c_uint16 * 128)()
            n = lib().xq_legal_moves(lib().xq_game_board(self._g), buf)
            self._legal = [Move(buf[i] >> 8, buf[i] & 255) for i in range(n)]
        return self._legal

    @property
    def legal_moves(self):
        return list(self._moves())

    def push(self, move: Move) -> None:
        if lib().xq_game_push(self._g, move.from_square, move.to_square) != 0:
            raise OverflowError("move stack full")
        self.move_stack.append(move)
        self._legal = None

    def pop(self) -> Move:
        lib().xq_game_pop(self._g)
        self._legal = None
        return self.move_stack.pop()

    def peek(self):
        return self.move_stack[-1] if self.move_stack else None

    def flags(self) -> int:
        return lib().xq_flags(lib().xq_game_board(self._g), len(self._moves()))

    def is_check(self) -> bool:
        return bool(self.flags() & FLAG_CHECK)

    def checkers(self):
        return []

    def is_checkmate(self) -> bool:
        fl = self.flags()
        return bool(fl & FLAG_CHECK) and bool(fl & FLAG_NOMOVES)

    def is_stalemate(self) -> bool:
        fl = self.flags()
        return not (fl & FLAG_CHECK) and bool(fl & FLAG_NOMOVES)

    def is_insufficient_material(self) -> bool:
        return bool(self.flags() & FLAG_INSUFFICIENT)

    def is_fourfold_repetition(self) -> bool:
        return bool(self.flags() & FLAG_FOURFOLD)

    def is_sixty_moves(self) -> bool:
        return bool(self.flags() & FLAG_SIXTY)

    def outcome(self):
        """SURVEY.md App. A.4 order: checkmate, insufficient material, stalemate (a loss for the
        side to move in Xiangqi), fourfold repetition, sixty moves."""
        fl = self.flags()
        if (fl & FLAG_CHECK) and (fl & FLAG_NOMOVES):
            return Outcome("checkmate", not self.turn)
        if fl & FLAG_INSUFFICIENT:
            return Outcome("insufficient_material", None)
        if fl & FLAG_NOMOVES:
            return Outcome("stalemate", not self.turn)
        if fl & FLAG_FOURFOLD:
            return Outcome("fourfold_repetition", None)
        if fl & FLAG_SIXTY:
            return Outcome("sixty_moves", None)
        return None

    def is_game_over(self) -> bool:
        return self.outcome() is not None

    def fen(self) -> str:
        rec = self.record()
        rows = []
        for r in range(9, -1, -1):
            row, gap = "", 0
            for f in range(9):
                c = int(rec[r * 9 + f])
                if not c:
                    gap += 1
                    continue
                if gap:
                    row += str(gap)
                    gap = 0
                s = PIECE_SYMBOLS[c & 7]
                row += s if c & 8 else s.upper()
            rows.append(row + (str(gap) if gap else ""))
        return "/".join(rows) + (" w" if rec[90] else " b")


# ---- helpers for tests / benches (not part of the cchess surface) -------------------------

def start_record() -> np.ndarray:
    b = ctypes.create_string_buffer(96)
    lib().xq_start(b)
    return np.frombuffer(b.raw, dtype=np.uint8).copy()


def perft(record: np.ndarray, depth: int) -> int:
    raw = np.ascontiguousarray(record, dtype=np.uint8)
    return int(lib().xq_perft(raw.ctypes.data, depth))


def collect_leaves(record: np.ndarray, depth: int, cap: int) -> np.ndarray:
    """All positions exactly ``depth`` plies from ``record`` in generation order, (n,96) uint8."""
    raw = np.ascontiguousarray(record, dtype=np.uint8)
    out = np.zeros((cap, 96), dtype=np.uint8)
    n = lib().xq_collect_leaves(raw.ctypes.data, depth, out.ctypes.data, cap)
    if n > cap:
        raise ValueError(f"cap {cap} < {n} leaves")
    return out[:n]


def batch_movegen_encode(boards: np.ndarray, want_planes: bool = True):
    """Oracle for ccz_movegen_encode: (ids[n,128] i16, counts[n] i16, flags[n] u8, planes[n,10710] u16|None)."""
    boards = np.ascontiguousarray(boards, dtype=np.uint8)
    n = boards.shape[0]
    ids = np.empty((n, 128), dtype=np.int16)
    counts = np.empty(n, dtype=np.int16)
    flags = np.empty(n, dtype=np.uint8)
    planes = np.empty((n, 10710), dtype=np.uint16) if want_planes else None
    lib().xq_batch_movegen_encode(boards.ctypes.data, n, ids.ctypes.data, counts.ctypes.data,
                                  flags.ctypes.data, planes.ctypes.data if want_planes else None)
    return ids, counts, flags, planes


def action_table():
    """(id_of[90,90] i16, from_of[2086] u8, to_of[2086] u8) built by the C restatement of tools.py:172-272."""
    id_of = np.empty(8100, dtype=np.int16)
    fr = np.empty(2086, dtype=np.uint8)
    to = np.empty(2086, dtype=np.uint8)
    n = lib().xq_build_action_table(id_of.ctypes.data, fr.ctypes.data, to.ctypes.data)
    assert n == 2086
    return id_of.reshape(90, 90), fr, to
