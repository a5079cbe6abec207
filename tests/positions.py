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


# Published Xiangqi perft results beyond the start position (the perft suite circulated with the chessprogramming
# wiki's "Chinese Chess Perft Results": mid- and end-game positions with pins, cannon screens, blocked horses and
# elephants, flying-general lines).  FEN (rank 9 first, 'w' = RED to move) -> perft(1..5).  All 50 values are
# reproduced by the plain-C oracle and by K1 + K2 on the device; they pin the LEGAL-MOVE SETS (not the order).
PERFT_SUITE = {
    "r1ba1a3/4kn3/2n1b4/pNp1p1p1p/4c4/6P2/P1P2R2P/1CcC5/9/2BAKAB2 w": (38, 1128, 43929, 1339047, 53112976),
    "1cbak4/9/n2a5/2p1p3p/5cp2/2n2N3/6PCP/3AB4/2C6/3A1K1N1 w": (7, 281, 8620, 326201, 10369923),
    "5a3/3k5/3aR4/9/5r3/5n3/9/3A1A3/5K3/2BC2B2 w": (25, 424, 9850, 202884, 4739553),
    "CRN1k1b2/3ca4/4ba3/9/2nr5/9/9/4B4/4A4/4KA3 w": (28, 516, 14808, 395483, 11842230),
    "R1N1k1b2/9/3aba3/9/2nr5/2B6/9/4B4/4A4/4KA3 w": (21, 364, 7626, 162837, 3500505),
    "C1nNk4/9/9/9/9/9/n1pp5/B3C4/9/3A1K3 w": (28, 222, 6241, 64971, 1914306),
    "4ka3/4a4/9/9/4N4/p8/9/4C3c/7n1/2BK5 w": (23, 345, 8124, 149272, 3513104),
    "2b1ka3/9/b3N4/4n4/9/9/9/4C4/2p6/2BK5 w": (21, 195, 3883, 48060, 933096),
    "1C2ka3/9/C1Nab1n2/p3p3p/6p2/9/P3P3P/3AB4/3p2c2/c1BAK4 w": (30, 830, 22787, 649866, 17920736),
    "CnN1k1b2/c3a4/4ba3/9/2nr5/9/9/4C4/4A4/4KA3 w": (19, 583, 11714, 376467, 8148177),
}
