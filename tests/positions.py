"""Deterministic position sets for parity tests (built with the oracle; test infrastructure)."""
from __future__ import annotations

import numpy as np

from oracle import cchess_shim as cs


def record_from_fen(fen: str, clock: int = 0, rep: int = 0) -> np.ndarray:
    """FEN placement (rank 9 first) + side ('w' = RED) -> 96-byte board record."""
    sym = {"p": 1, "c": 2, "r": 3, "n": 4, "b": 5, "a": 6, "k": 7}
    placement, side = fen.split()[:2]
    rec = np.zeros(96, dtype=np.uint8)
    rows = placement.split("/")
    assert len(rows) == 10
    for i, row in enumerate(rows):
        rank, f = 9 - i, 0
        for ch in row:
            if ch.isdigit():
                f += int(ch)
            else:
                rec[rank * 9 + f] = sym[ch.lower()] | (8 if ch.islower() else 0)
                f += 1
        assert f == 9, row
    rec[90] = 1 if side == "w" else 0
    rec[91], rec[92] = clock, rep
    return rec


def perft_leaves(depth: int) -> np.ndarray:
    n = cs.perft(cs.start_record(), depth) if depth > 0 else 1
    return cs.collect_leaves(cs.start_record(), depth, n)


def random_playout_positions(n_games: int, max_plies: int, seed: int, every: int = 1) -> np.ndarray:
    """Positions visited by uniformly random legal play from the start (captures, checks, endgames,
    growing half-move clocks and real repetition counts included)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_games):
        b = cs.Board()
        for ply in range(max_plies):
            if ply % every == 0:
                out.append(b.record())
            moves = b.legal_moves
            if not moves or b.is_game_over():
                out.append(b.record())
                break
            b.push(moves[int(rng.integers(len(moves)))])
    return np.stack(out)


# hand-made edge cases: (name, fen, clock, rep)
EDGE_CASES = [
    ("start", "rnbakabnr/9/1c5c1/p1p1p1p1p/9/9/P1P1P1P1P/1C5C1/9/RNBAKABNR w", 0, 0),
    ("start_black", "rnbakabnr/9/1c5c1/p1p1p1p1p/9/9/P1P1P1P1P/1C5C1/9/RNBAKABNR b", 0, 0),
    # bare kings on different files: insufficient material
    ("bare_kings", "3k5/9/9/9/9/9/9/9/9/5K3 w", 0, 0),
    # kings + advisors/elephants only: insufficient
    ("no_attackers", "2bak4/4a4/4b4/9/9/9/9/4B4/4A4/3AK1B2 b", 7, 0),
    # flying general: red rook pinned on the e-file between the kings
    ("flying_pin", "4k4/9/9/9/9/9/9/9/4R4/4K4 w", 0, 0),
    # checkmate: black king d9 facing open d-file rook with e-file covered by a second rook
    ("mate_two_rooks", "3k5/9/9/9/9/9/9/9/9/3RRK3 b", 0, 0),
    # stalemate-like: black king boxed in by red pawns and a rook, no other black pieces
    ("stalemate_box", "3k5/9/3P1R3/9/9/9/9/9/9/4K4 b", 0, 0),
    # cannon check over one screen and the screen's pinned status
    ("cannon_screen", "4k4/4a4/9/9/9/9/9/4C4/9/3K5 b", 0, 0),
    # double cannons: screen piece may not leave
    ("cannon_pin", "3ak4/9/4n4/9/9/9/4C4/9/9/4K4 b", 3, 0),
    # hobbled horse check / unhobbled
    ("horse_leg", "4k4/9/3N5/9/9/9/9/9/9/3K5 b", 0, 0),
    ("horse_leg_blocked", "4k4/3p5/3N5/9/9/9/9/9/9/3K5 b", 0, 0),
    # pawns across the river attack sideways
    ("pawn_fwd_block", "4k4/4P4/9/9/9/9/9/9/9/4K4 b", 0, 0),
    ("pawn_side_check2", "3Pk4/9/9/9/9/9/9/9/9/5K3 b", 0, 0),
    # elephants: eye blocking and river
    ("elephant_eye", "4k4/9/9/9/2b6/2B6/3P5/9/9/3K5 w", 0, 0),
    # sixty-move and fourfold flags
    ("sixty", "rnbakabnr/9/1c5c1/p1p1p1p1p/9/9/P1P1P1P1P/1C5C1/9/RNBAKABNR w", 120, 0),
    ("sixty_minus_one", "rnbakabnr/9/1c5c1/p1p1p1p1p/9/9/P1P1P1P1P/1C5C1/9/RNBAKABNR b", 119, 2),
    ("fourfold", "rnbakabnr/9/1c5c1/p1p1p1p1p/9/9/P1P1P1P1P/1C5C1/9/RNBAKABNR w", 12, 3),
    # sixty-move clock reached but no legal move: stalemate wins over sixty (needs a legal move)
    ("sixty_nomoves", "3k5/9/3P1R3/9/9/9/9/9/9/4K4 b", 121, 0),
    # many pieces with maximal mobility
    ("open_board", "4k4/9/9/R7R/1C5C1/1N5N1/9/9/9/3K5 w", 0, 0),
    ("open_board_b", "4k4/9/r7r/1c5c1/1n5n1/9/9/9/9/3K5 b", 0, 0),
]


def edge_case_records() -> np.ndarray:
    return np.stack([record_from_fen(f, c, r) for _, f, c, r in EDGE_CASES])
