"""Host-side mirror of the reference's ``tools.py`` for the self-play path.

Same names and meaning as /root/reference/tools.py:
  move_id2move_action, move_action2move_id   (tools.py:172-272)   2086-entry action table
  flip(uci)                                  (tools.py:133-164)   file mirror of a UCI move
  softmax(x)                                 (tools.py:126-129)
  decode_board(board)                        (tools.py:74-106)    two int8 (7,10,9) one-hot arrays
  is_tie(board)                              (tools.py:109-123)
The action table itself comes from the CUDA library's host-side builder (ccz_action_table) so
that the Python names and the device tables can never diverge.  ``board`` arguments here are
96-byte board records (numpy uint8) -- the packed device representation that replaces
cchess.Board -- or anything exposing ``.record()``.
"""
from __future__ import annotations

import numpy as np

from . import _lib

_FILES = "abcdefghi"


def square_name(sq: int) -> str:
    return f"{_FILES[sq % 9]}{sq // 9}"


def parse_square(s: str) -> int:
    return _FILES.index(s[0]) + 9 * int(s[1])


def _build_tables():
    id_of, fr, to = _lib.host_action_table()
    id2act = {i: square_name(int(fr[i])) + square_name(int(to[i])) for i in range(_lib.N_ACTIONS)}
    act2id = {a: i for i, a in id2act.items()}
    return id_of, fr, to, id2act, act2id


ID_OF, FROM_OF, TO_OF, move_id2move_action, move_action2move_id = _build_tables()


def get_all_legal_moves():
    """Same return value as the reference function of this name (tools.py:172)."""
    return dict(move_id2move_action), dict(move_action2move_id)


def flip(string: str) -> str:
    """Mirror the files of a UCI move string (tools.py:133-164)."""
    return "".join(_FILES[8 - _FILES.index(ch)] if i in (0, 2) else ch for i, ch in enumerate(string[:4]))


FLIP_MAP = np.array([move_action2move_id[flip(move_id2move_action[i])] for i in range(_lib.N_ACTIONS)],
                    dtype=np.int64)  # collect.py:117-122


def softmax(x):
    probs = np.exp(x - np.max(x))
    probs /= np.sum(probs)
    return probs


def _record(board) -> np.ndarray:
    rec = board.record() if hasattr(board, "record") else board
    return np.asarray(rec, dtype=np.uint8).reshape(-1)


def decode_board(board):
    """(red_state, black_state): int8 (7,10,9) one-hot planes, channel = piece_type-1 (tools.py:74-106)."""
    sq = _record(board)[:90].astype(np.int64)
    red = np.zeros((7, 90), dtype=np.int8)
    black = np.zeros((7, 90), dtype=np.int8)
    occ = np.nonzero(sq)[0]
    codes = sq[occ]
    is_black = (codes & 8) != 0
    red[(codes[~is_black] & 7) - 1, occ[~is_black]] = 1
    black[(codes[is_black] & 7) - 1, occ[is_black]] = 1
    return red.reshape(7, 10, 9), black.reshape(7, 10, 9)


def is_tie_flags(flags) -> bool:
    """is_tie (tools.py:109-123) from the flag byte written by ccz_movegen_encode."""
    return bool(int(flags) & _lib.FLAG_TIE_MASK)


def is_game_over_flags(flags) -> bool:
    """board.is_game_over(): checkmate / stalemate / insufficient / fourfold / sixty."""
    return bool(int(flags) & (_lib.FLAG_TIE_MASK | _lib.FLAG_NOMOVES))


def outcome_winner_flags(flags, turn_red: bool):
    """board.outcome().winner in cchess order: True = RED, False = BLACK,
    None = draw or not over.  checkmate -> not turn; insufficient -> draw; stalemate -> not turn;
    fourfold / sixty -> draw."""
    fl = int(flags)
    if (fl & _lib.FLAG_CHECK) and (fl & _lib.FLAG_NOMOVES):
        return not turn_red
    if fl & _lib.FLAG_INSUFFICIENT:
        return None
    if fl & _lib.FLAG_NOMOVES:
        return not turn_red
    return None


_LEVELS = {"DEBUG": 1, "INFO": 2, "WARNING": 3, "ERROR": 4, "CRITICAL": 5}


def log(message: str, level: str = "INFO", log_path: str | None = None):
    """Same call signature and behaviour as the reference logger (tools.py:12-71): every message is
    appended to ``<log_path or ./logs>/<script>.log``; the console shows levels >= parameters.LOG_LEVEL."""
    import os
    import sys
    import time

    from . import parameters

    lvl = (level or "INFO").upper()
    rank = _LEVELS.get(lvl, 2)
    script = os.path.splitext(os.path.basename(sys.argv[0] or "app"))[0] or "app"
    target_dir = log_path or os.path.join(os.getcwd(), "logs")
    try:
        os.makedirs(target_dir, exist_ok=True)
        with open(os.path.join(target_dir, f"{script}.log"), "a", encoding="utf-8") as f:
            f.write(f"{time.strftime('%Y-%m-%d %H:%M:%S')} | {lvl:<8} | {message}\n")
    except OSError:
        pass
    try:
        threshold = max(1, min(5, int(parameters.LOG_LEVEL)))
    except (TypeError, ValueError):
        threshold = 2
    if rank >= threshold:
        print(f"[{lvl}] {message}", file=sys.stderr)
