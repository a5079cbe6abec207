"""Pin the CPU oracle (oracle/xq_oracle.c + cchess_shim) with external known answers.

The reference ships no tests or golden vectors for its cchess dependency, so
the pins are rule-determined facts: published Xiangqi perft values from the start position and
hand-analysed positions.
"""
import json
import os

import numpy as np
import pytest

from oracle import cchess_shim as cs
from tests import positions

PERFT = {1: 44, 2: 1920, 3: 79666, 4: 3290240}


@pytest.mark.parametrize("depth", [1, 2, 3, 4])
def test_perft_known_values(depth):
    assert cs.perft(cs.start_record(), depth) == PERFT[depth]


@pytest.mark.slow
def test_perft5():
    assert cs.perft(cs.start_record(), 5) == 133312995


def test_start_position_moves_and_order():
    b = cs.Board()
    ucis = [m.uci() for m in b.legal_moves]
    assert len(ucis) == 44 and len(set(ucis)) == 44
    # generation order policy: non-pawns by from-square descending, destinations descending, then pawns
    assert ucis[:5] == ["h2h9", "h2h6", "h2h5", "h2h4", "h2h3"]
    assert ucis[-5:] == ["i3i4", "g3g4", "e3e4", "c3c4", "a3a4"]
    assert b.fen() == "rnbakabnr/9/1c5c1/p1p1p1p1p/9/9/P1P1P1P1P/1C5C1/9/RNBAKABNR w"


def test_action_table_matches_reference_golden(golden_dir):
    with open(os.path.join(golden_dir, "action_table.json")) as f:
        gold = json.load(f)
    id_of, fr, to = cs.action_table()
    names = [cs.Move(int(fr[i]), int(to[i])).uci() for i in range(2086)]
    assert names == gold["move_id2move_action"]
    assert names[0] == "a0a1" and names[2038] == "d0e1" and names[2054] == "a2c0" and names[2085] == "i7g9"
    assert int((id_of >= 0).sum()) == 2086


def _flags(name):
    for n, fen, clock, rep in positions.EDGE_CASES:
        if n == name:
            rec = positions.record_from_fen(fen, clock, rep)
            ids, counts, flags, _ = cs.batch_movegen_encode(rec[None], want_planes=False)
            return int(counts[0]), int(flags[0]), ids[0]
    raise KeyError(name)


def test_edge_case_flags():
    n, fl, _ = _flags("bare_kings")
    assert fl & cs.FLAG_INSUFFICIENT and n > 0
    n, fl, _ = _flags("no_attackers")
    assert fl & cs.FLAG_INSUFFICIENT
    n, fl, _ = _flags("mate_two_rooks")
    assert n == 0 and fl & cs.FLAG_CHECK and fl & cs.FLAG_NOMOVES
    n, fl, _ = _flags("stalemate_box")
    assert n == 0 and not (fl & cs.FLAG_CHECK) and fl & cs.FLAG_NOMOVES
    n, fl, _ = _flags("sixty")
    assert fl & cs.FLAG_SIXTY and n == 44
    n, fl, _ = _flags("sixty_minus_one")
    assert not (fl & cs.FLAG_SIXTY) and not (fl & cs.FLAG_FOURFOLD)
    n, fl, _ = _flags("fourfold")
    assert fl & cs.FLAG_FOURFOLD
    n, fl, _ = _flags("sixty_nomoves")
    assert n == 0 and not (fl & cs.FLAG_SIXTY)
    n, fl, _ = _flags("pawn_fwd_block")  # red pawn e8 checks the king on e9 and shields the kings
    assert fl & cs.FLAG_CHECK and n > 0
    n, fl, _ = _flags("pawn_side_check2")
    assert fl & cs.FLAG_CHECK
    n, fl, _ = _flags("horse_leg")
    assert fl & cs.FLAG_CHECK
    n, fl, _ = _flags("horse_leg_blocked")
    assert not (fl & cs.FLAG_CHECK)
    n, fl, _ = _flags("cannon_screen")
    assert fl & cs.FLAG_CHECK


def test_flying_general_pins_rook():
    rec = positions.record_from_fen("4k4/9/9/9/9/9/9/9/4R4/4K4 w")
    b = cs.Board.from_record(rec)
    ucis = {m.uci() for m in b.legal_moves}
    # the rook may slide on the e-file (and capture nothing), never sideways; the king may step aside
    assert all(u[0] == "e" and u[2] == "e" for u in ucis if u.startswith("e1"))
    assert "e0d0" in ucis and "e0f0" in ucis
    assert len([u for u in ucis if u.startswith("e1")]) == 8


def test_repetition_counts_follow_reversible_window():
    b = cs.Board()
    shuffle = ["b0c2", "b9c7", "c2b0", "c7b9"]  # knights out and back: position repeats every 4 plies
    for cycle in range(3):
        for u in shuffle:
            b.push(cs.Move.from_uci(u))
        assert b.record()[92] == cycle + 1
        assert b.is_fourfold_repetition() == (cycle + 1 >= 3)
    assert b.halfmove_clock == 12 and b.is_game_over() and b.outcome().winner is None
    # a capture clears the window
    b2 = cs.Board()
    for u in ["b0c2", "b9c7", "c2b0", "c7b9", "h2h9"]:  # cannon takes knight
        b2.push(cs.Move.from_uci(u))
    assert b2.halfmove_clock == 0 and b2.record()[92] == 0
    b2.pop()
    assert b2.halfmove_clock == 4 and b2.record()[92] == 1


def test_planes_layout_matches_policy_value_fn():
    """net.py:160-177: 7 zero states + current; play 7 = red, 15 = black, 16 = turn."""
    rec = cs.start_record()
    _, _, _, planes = cs.batch_movegen_encode(rec[None])
    p = planes[0].reshape(17, 7, 10, 9)
    assert (p[[0, 1, 2, 3, 4, 5, 6, 8, 9, 10, 11, 12, 13, 14]] == 0).all()
    assert (p[16] == 0x3F80).all()
    assert p[7, 6, 0, 4] == 0x3F80 and p[15, 6, 9, 4] == 0x3F80  # kings, channel = KING-1
    assert p[7, 0, 3, 0] == 0x3F80 and p[15, 0, 6, 8] == 0x3F80  # pawns
    assert int((p[7] != 0).sum()) == 16 and int((p[15] != 0).sum()) == 16
    rec_b = rec.copy()
    rec_b[90] = 0
    _, _, _, planes_b = cs.batch_movegen_encode(rec_b[None])
    assert (planes_b[0].reshape(17, 7, 10, 9)[16] == 0).all()


@pytest.mark.parametrize("fen", list(__import__("tests.positions", fromlist=["PERFT_SUITE"]).PERFT_SUITE))
def test_published_perft_suite(fen):
    """Ten published mid-/end-game perft positions, depths 1-5 (50 known answers): pins the oracle's legal-move
    sets on positions with pins, cannon screens, blocked horses / elephants and flying-general lines."""
    from tests.positions import PERFT_SUITE, record_from_fen

    rec = record_from_fen(fen)
    expect = PERFT_SUITE[fen]
    for depth, want in enumerate(expect, start=1):
        if want > 20_000_000:  # keep the CPU suite short: the two largest depth-5 counts run in the GPU test
            continue
        assert cs.perft(rec, depth) == want, (fen, depth)
