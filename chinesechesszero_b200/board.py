"""Device-backed stand-in for ``cchess.Board`` / ``cchess.Move`` on the self-play path.

The reference drives its game loop through a ``cchess.Board`` object (game.py:148,201,208;
mcts.py:111,116,151; net.py:155-156; tools.py:92).  ``Board`` here exposes the same duck-typed
surface for ONE position, but the state is a 96-byte board record + 128-entry key window in device
memory and every rule query is answered by the CUDA kernels through the C ABI
(``ccz_movegen_encode``, ``ccz_board_push``).  It exists so that the reference-shaped façades
(``mcts.MCTS_AI``, ``game.Game``) can be used one game at a time; the throughput path is
``selfplay.SelfPlayEngine`` which keeps thousands of such records in lockstep.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, tools

RED = True
BLACK = False


class Move:
    __slots__ = ("from_square", "to_square")

    def __init__(self, from_square: int, to_square: int):
        self.from_square, self.to_square = int(from_square), int(to_square)

    @classmethod
    def from_uci(cls, uci: str) -> "Move":
        return cls(tools.parse_square(uci[0:2]), tools.parse_square(uci[2:4]))

    def uci(self) -> str:
        return tools.square_name(self.from_square) + tools.square_name(self.to_square)

    def action_id(self) -> int:
        return int(tools.ID_OF[self.from_square, self.to_square])

    def __eq__(self, other):
        return isinstance(other, Move) and (self.from_square, self.to_square) == (other.from_square, other.to_square)

    def __hash__(self):
        return hash((self.from_square, self.to_square))

    def __repr__(self):
        return f"Move.from_uci({self.uci()!r})"


class Piece:
    __slots__ = ("piece_type", "color")

    def __init__(self, piece_type: int, color: bool):
        self.piece_type, self.color = piece_type, color


class Outcome:
    __slots__ = ("termination", "winner")

    def __init__(self, termination: str, winner):
        self.termination, self.winner = termination, winner


class Board:
    def __init__(self, record=None, device="cuda"):
        self.device = torch.device(device)
        if record is None:
            self._board = _lib.boards_start(1, self.device)
        else:
            rec = np.ascontiguousarray(record, dtype=np.uint8).reshape(1, _lib.BOARD_BYTES)
            self._board = torch.from_numpy(rec.copy()).to(self.device)
        self._keys = _lib.board_keys_init(self._board)
        self.move_stack: list[Move] = []
        self._cache = None

    # ---- state -------------------------------------------------------------------------------
    def record(self) -> np.ndarray:
        return self._board[0].cpu().numpy()

    def copy(self) -> "Board":
        b = Board.__new__(Board)
        b.device = self.device
        b._board, b._keys = self._board.clone(), self._keys.clone()
        b.move_stack = list(self.move_stack)
        b._cache = self._cache
        return b

    @property
    def turn(self) -> bool:
        return bool(self.record()[90])

    @property
    def halfmove_clock(self) -> int:
        return int(self.record()[91])

    def piece_at(self, square: int):
        c = int(self.record()[square])
        return Piece(c & 7, not (c & 8)) if c else None

    # ---- rules (all through K1) --------------------------------------------------------------
    def _gen(self):
        if self._cache is None:
            ids, counts, flags, _ = _lib.movegen_encode(self._board, planes=False)
            n = int(counts[0])
            self._cache = (ids[0, :n].cpu().numpy().astype(np.int64), int(flags[0]))
        return self._cache

    def legal_ids(self) -> np.ndarray:
        return self._gen()[0]

    @property
    def legal_moves(self):
        return [Move(int(tools.FROM_OF[i]), int(tools.TO_OF[i])) for i in self.legal_ids()]

    def flags(self) -> int:
        return self._gen()[1]

    def push(self, move) -> None:
        """board.push (game.py:201); accepts a Move or an action id."""
        mid = move.action_id() if isinstance(move, Move) else int(move)
        if mid < 0 or mid >= _lib.N_ACTIONS:
            raise ValueError(f"not an action: {move}")
        ids = torch.tensor([mid], dtype=torch.int16, device=self.device)
        _lib.board_push(self._board, ids, self._keys)
        self.move_stack.append(Move(int(tools.FROM_OF[mid]), int(tools.TO_OF[mid])))
        self._cache = None

    def peek(self):
        return self.move_stack[-1] if self.move_stack else None

    def is_check(self) -> bool:
        return bool(self.flags() & _lib.FLAG_CHECK)

    def is_checkmate(self) -> bool:
        fl = self.flags()
        return bool(fl & _lib.FLAG_CHECK) and bool(fl & _lib.FLAG_NOMOVES)

    def is_stalemate(self) -> bool:
        fl = self.flags()
        return not (fl & _lib.FLAG_CHECK) and bool(fl & _lib.FLAG_NOMOVES)

    def is_insufficient_material(self) -> bool:
        return bool(self.flags() & _lib.FLAG_INSUFFICIENT)

    def is_fourfold_repetition(self) -> bool:
        return bool(self.flags() & _lib.FLAG_FOURFOLD)

    def is_sixty_moves(self) -> bool:
        return bool(self.flags() & _lib.FLAG_SIXTY)

    def is_game_over(self) -> bool:
        return tools.is_game_over_flags(self.flags())

    def outcome(self):
        fl = self.flags()
        if not tools.is_game_over_flags(fl):
            return None
        return Outcome("over", tools.outcome_winner_flags(fl, self.turn))


def is_tie(board) -> bool:
    """tools.is_tie (tools.py:109-123) for a device Board."""
    return tools.is_tie_flags(board.flags())
