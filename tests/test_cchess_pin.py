"""The `cchess` residue (generation order, clock convention, draw predicates, outcome):
pinned by ``tests/golden/cchess_pin.json`` when somebody with a real ``cchess`` has run
``scripts/pin_cchess.py``; until then those tests skip ("parity unpinned").  The machinery itself --
policy family, inference, oracle <-> K1 agreement under any policy -- is tested here without cchess."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import cchess_shim as cs
from tests import positions

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PIN = os.path.join(ROOT, "tests", "golden", "cchess_pin.json")

ODD_POLICY = {"class_rank": {"p": 0, "c": 2, "r": 2, "n": 1, "b": 1, "a": 1, "k": 3},
              "from_descending": 0, "to_descending": 1, "capture_mode": 2, "check_king_first": 1}


@pytest.fixture
def default_policy_after():
    yield
    cs.set_order_policy(None)
    try:
        from chinesechesszero_b200 import _lib

        _lib.set_order_policy(None)
    except Exception:  # noqa: BLE001 - library not built
        pass


def _shim_board(entry):
    if entry["kind"] == "playout":
        b = cs.Board()
        for u in entry["moves"]:
            b.push(cs.Move.from_uci(u))
        return b
    return cs.Board.from_record(positions.record_from_fen(entry["fen"], clock=entry["clock"]))


def _load_pin(path=PIN):
    if not os.path.exists(path):
        pytest.skip("tests/golden/cchess_pin.json absent: run scripts/pin_cchess.py where `import cchess` works")
    with open(path) as f:
        return json.load(f)


def _check_oracle_against(pin):
    assert pin["order_policy"] is not None, "no policy of the family reproduces cchess's order: the generation order needs code"
    cs.set_order_policy(pin["order_policy"])
    for e in pin["entries"]:
        b = _shim_board(e)
        tag = e.get("name") or e.get("fen") or " ".join(e["moves"][-4:])
        assert [m.uci() for m in b.legal_moves] == e["legal"], ("order", tag)
        assert b.turn == e["turn"] and b.halfmove_clock == e["halfmove_clock"], ("clock", tag)
        assert b.is_game_over() == e["is_game_over"], ("is_game_over", tag)
        if e["is_game_over"]:
            out = b.outcome()
            assert (None if out is None else out.winner) == e["winner"], ("winner", tag)
        assert b.is_insufficient_material() == e["insufficient"], ("insufficient", tag)
        assert b.is_fourfold_repetition() == e["fourfold"], ("fourfold", tag)
        assert b.is_sixty_moves() == e["sixty"], ("sixty", tag)
        if e["is_check"] is not None:
            assert b.is_check() == e["is_check"], ("check", tag)


def _check_k1_against(pin):
    import torch

    from chinesechesszero_b200 import _lib, tools

    _lib.set_order_policy(pin["order_policy"])
    recs = np.stack([_shim_board(e).record() for e in pin["entries"]])
    ids, counts, flags, _ = _lib.movegen_encode(torch.from_numpy(recs).cuda(), planes=False)
    ids, counts, flags = ids.cpu().numpy(), counts.cpu().numpy(), flags.cpu().numpy()
    for i, e in enumerate(pin["entries"]):
        got = [tools.move_id2move_action[int(a)] for a in ids[i, : counts[i]]]
        assert got == e["legal"], e.get("name") or e.get("fen") or e["moves"][-4:]
        over = bool(flags[i] & _lib.FLAG_NOMOVES)
        tie = bool(flags[i] & _lib.FLAG_TIE_MASK)
        assert (over or tie) == (e["is_game_over"] or e["insufficient"] or e["fourfold"] or e["sixty"])
        assert bool(flags[i] & _lib.FLAG_INSUFFICIENT) == e["insufficient"]
        assert bool(flags[i] & _lib.FLAG_FOURFOLD) == e["fourfold"] and bool(flags[i] & _lib.FLAG_SIXTY) == e["sixty"]


# ---- the real pin (skips until the fixture exists) ---------------------------------------------------
def test_oracle_equals_cchess_pin(default_policy_after):
    pin = _load_pin()
    assert pin["source"] == "cchess", "the committed fixture must come from the real package"
    _check_oracle_against(pin)


@pytest.mark.gpu
def test_k1_equals_cchess_pin(default_policy_after):
    pin = _load_pin()
    _check_k1_against(pin)


# ---- the machinery, without cchess -------------------------------------------------------------------
@pytest.fixture(scope="module")
def synthetic_pin(tmp_path_factory):
    """scripts/pin_cchess.py run against the shim generating with an unusual policy."""
    out = str(tmp_path_factory.mktemp("pin") / "pin.json")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "pin_cchess.py"), "--use-shim", "--shim-policy",
                        json.dumps(ODD_POLICY), "--games", "8", "--max-plies", "80", "--out", out],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    with open(out) as f:
        return json.load(f)


def test_policy_is_recovered_from_a_dump(synthetic_pin, default_policy_after):
    pin = synthetic_pin
    assert pin["source"] == "shim" and pin["move_sets_equal_oracle"] and pin["n_entries"] > 300
    assert pin["order_policy"] is not None
    # the recovered policy reproduces every dumped order (class numbers may differ from ODD_POLICY's, their order may not)
    _check_oracle_against(pin)
    # and it is not the default: the default order fails on this dump
    cs.set_order_policy(None)
    assert any([m.uci() for m in _shim_board(e).legal_moves] != e["legal"] for e in pin["entries"] if len(e["legal"]) > 1)


def test_policy_only_permutes_the_legal_set(default_policy_after):
    recs = positions.random_playout_positions(4, 120, seed=5)
    cs.set_order_policy(None)
    base = [[m.uci() for m in cs.Board.from_record(r).legal_moves] for r in recs]
    cs.set_order_policy(ODD_POLICY)
    assert cs.get_order_policy() == ODD_POLICY
    seen_in_check = [0]
    for r, want in zip(recs, base):
        b = cs.Board.from_record(r)
        got = [m.uci() for m in b.legal_moves]
        assert sorted(got) == sorted(want)
        # pawns first (class 0), kings last (class 3) -- except in check, where the king's moves lead
        types = [int(r[cs.Move.from_uci(u).from_square]) & 7 for u in got]
        if b.is_check():
            seen_in_check[0] += 1
            ranks = [0 if t == 7 else 1 + {1: 0, 2: 2, 3: 2, 4: 1, 5: 1, 6: 1}[t] for t in types]
        else:
            ranks = [{1: 0, 2: 2, 3: 2, 4: 1, 5: 1, 6: 1, 7: 3}[t] for t in types]
        assert ranks == sorted(ranks)
    assert seen_in_check[0] > 0  # the evasion ordering was exercised
    with pytest.raises(ValueError):
        cs.set_order_policy(dict(ODD_POLICY, capture_mode=3))


@pytest.mark.gpu
def test_k1_follows_the_policy_like_the_oracle(synthetic_pin, default_policy_after):
    """K1's SORTED instantiation == the oracle under the same policy on the dumped positions plus the perft-3
    leaves (order, counts, flags); back on the default policy K1 is the native kernel again."""
    import torch

    from chinesechesszero_b200 import _lib

    _check_k1_against(synthetic_pin)
    pol = synthetic_pin["order_policy"]
    recs = positions.perft_leaves(3)[::7]
    cs.set_order_policy(pol)
    _lib.set_order_policy(pol)
    assert _lib.get_order_policy() == pol
    o_ids, o_counts, o_flags, _ = cs.batch_movegen_encode(recs, want_planes=False)
    ids, counts, flags, _ = _lib.movegen_encode(torch.from_numpy(recs).cuda(), planes=False)
    assert np.array_equal(ids.cpu().numpy(), o_ids) and np.array_equal(counts.cpu().numpy(), o_counts)
    assert np.array_equal(flags.cpu().numpy(), o_flags)
    cs.set_order_policy(None)
    _lib.set_order_policy(None)
    d_ids, _, _, _ = cs.batch_movegen_encode(recs, want_planes=False)
    ids2, _, _, _ = _lib.movegen_encode(torch.from_numpy(recs).cuda(), planes=False)
    assert np.array_equal(ids2.cpu().numpy(), d_ids) and not np.array_equal(d_ids, o_ids)
