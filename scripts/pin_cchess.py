"""pin_cchess -- pin the `cchess` residue the moment the real package is importable.

The reference takes its move ORDER from ``board.legal_moves`` (net.py:154-157) and its terminal rules
from ``board.is_game_over() / outcome() / is_insufficient_material() / is_fourfold_repetition() /
is_sixty_moves()`` (tools.py:109-123, mcts.py:116-126).  cchess (windshadow233/python-chinese-chess) is
neither vendored nor pinned by the reference and is not installable where this repository is built, so
those behaviours are recalled and "parity unpinned".  Run this script once in any
environment where ``import cchess`` works:

    python scripts/pin_cchess.py                      # writes tests/golden/cchess_pin.json

It dumps, for the hand-made edge cases, the ten published perft positions and a few thousand positions of
seeded random play: the ordered legal moves (UCI), side to move, half-move clock, is_check, is_game_over,
outcome winner, and the three draw predicates; then it searches the generation-order policy family
(``ccz_order_policy``: class rank per piece type, from / to direction, capture placement) for the policy
that reproduces every dumped order with the oracle, and stores it under "order_policy" (null if none fits:
then the order needs code, and the test says so).  tests/test_cchess_pin.py compares the oracle (CPU) and K1
(GPU) with the file and skips while it is absent; the product loads the pinned order from
``chinesechesszero_b200/order_policy.json``, written alongside.  ``--use-shim`` runs the same dump against the oracle's own
shim (a self-test of this machinery; such a file pins nothing and is marked "source": "shim").
"""
from __future__ import annotations

import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import cchess_shim as cs  # noqa: E402
from tests import positions  # noqa: E402

SYMS = "pcrnbak"


def _board_from_fen(cchess, fen: str, clock: int):
    """FEN placement + side -> a real cchess board with the given half-move clock."""
    if cchess is cs:  # --use-shim self-test
        return cs.Board.from_record(positions.record_from_fen(fen, clock=clock))
    full = f"{fen} - - {clock} 1"
    try:
        b = cchess.Board(full)
    except Exception:  # noqa: BLE001 - older / different constructor: build it square by square
        b = cchess.Board()
        b.clear() if hasattr(b, "clear") else None
        rec = positions.record_from_fen(fen)
        for sq in range(90):
            c = int(rec[sq])
            if c:
                b.set_piece_at(sq, cchess.Piece(c & 7, cchess.RED if not (c & 8) else cchess.BLACK))
        b.turn = cchess.RED if rec[90] else cchess.BLACK
    try:
        b.halfmove_clock = clock
    except Exception:  # noqa: BLE001
        pass
    return b


def _describe(cchess, board) -> dict:
    legal = [cchess.Move.uci(m) for m in board.legal_moves]
    over = bool(board.is_game_over())
    winner = None
    if over:
        out = board.outcome()
        winner = None if out is None or out.winner is None else bool(out.winner)
    rec = [0] * 90
    for sq in range(90):
        p = board.piece_at(sq)
        if p:
            rec[sq] = int(p.piece_type) | (0 if p.color == cchess.RED else 8)
    checkers = getattr(board, "is_check", None)
    return {
        "squares": rec, "turn": bool(board.turn), "halfmove_clock": int(board.halfmove_clock), "legal": legal,
        "is_check": bool(checkers()) if callable(checkers) else None, "is_game_over": over, "winner": winner,
        "insufficient": bool(board.is_insufficient_material()), "fourfold": bool(board.is_fourfold_repetition()),
        "sixty": bool(board.is_sixty_moves()),
    }


def dump(cchess, n_random_games: int, max_plies: int, seed: int) -> list[dict]:
    entries = []
    for name, fen, clock, _rep in positions.EDGE_CASES:
        e = _describe(cchess, _board_from_fen(cchess, fen, clock))
        e.update(kind="edge", name=name, fen=fen, clock=clock)
        entries.append(e)
    for fen in positions.PERFT_SUITE:
        e = _describe(cchess, _board_from_fen(cchess, fen, 0))
        e.update(kind="perft", fen=fen, clock=0)
        entries.append(e)
    rng = np.random.default_rng(seed)
    for g in range(n_random_games):
        board = cchess.Board()
        line = []
        for ply in range(max_plies):
            e = _describe(cchess, board)
            e.update(kind="playout", moves=list(line))   # replayable from the start: history-dependent rules included
            entries.append(e)
            if e["is_game_over"] or not e["legal"]:
                break
            mv = e["legal"][int(rng.integers(len(e["legal"])))]
            board.push(cchess.Move.from_uci(mv))
            line.append(mv)
    return entries


def shim_board(entry):
    """The oracle's board for a dumped entry."""
    if entry["kind"] == "playout":
        b = cs.Board()
        for u in entry["moves"]:
            b.push(cs.Move.from_uci(u))
        return b
    return cs.Board.from_record(positions.record_from_fen(entry["fen"], clock=entry["clock"]))


def infer_policy(entries) -> tuple[dict | None, dict]:
    """Search the policy family for one that reproduces every dumped order with the oracle."""
    # class ranks from the data: type a is in an earlier class than b if a's moves precede b's wherever both occur
    before = np.ones((8, 8), dtype=bool)
    seen = np.zeros((8, 8), dtype=bool)
    for e in entries:
        if e.get("is_check"):
            continue  # evasions may be generated king-first (check_king_first): classes are read off quiet positions
        sq = e["squares"]
        types = [sq[cs.Move.from_uci(u).from_square] & 7 for u in e["legal"]]
        first, last = {}, {}
        for i, t in enumerate(types):
            first.setdefault(t, i)
            last[t] = i
        for a, b in itertools.permutations(first, 2):
            seen[a, b] = True
            if last[a] > first[b]:
                before[a, b] = False
    ranks = {}
    for t in range(1, 8):
        ranks[SYMS[t - 1]] = int(sum(1 for u in range(1, 8) if u != t and seen[u, t] and before[u, t] and not before[t, u]))
    # compress to consecutive class numbers
    order = sorted(set(ranks.values()))
    inferred = {k: order.index(v) for k, v in ranks.items()}
    candidates = [inferred, cs.DEFAULT_ORDER_POLICY["class_rank"], {s: 0 for s in SYMS}]
    report = {"inferred_class_rank": inferred, "tried": 0}
    boards = [(shim_board(e).record(), e["legal"]) for e in entries if e["legal"]]
    try:
        for cr in candidates:
            for fd, td, cm, kf in itertools.product((1, 0), (1, 0), (0, 1, 2), (0, 1)):
                pol = {"class_rank": cr, "from_descending": fd, "to_descending": td, "capture_mode": cm, "check_king_first": kf}
                cs.set_order_policy(pol)
                report["tried"] += 1
                if all([m.uci() for m in cs.Board.from_record(rec).legal_moves] == legal for rec, legal in boards):
                    return pol, report
    finally:
        cs.set_order_policy(None)
    return None, report


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "cchess_pin.json"))
    ap.add_argument("--games", type=int, default=40)
    ap.add_argument("--max-plies", type=int, default=120)
    ap.add_argument("--seed", type=int, default=20260101)
    ap.add_argument("--use-shim", action="store_true", help="self-test: dump the oracle's shim instead of the real cchess")
    ap.add_argument("--shim-policy", default=None, help="self-test: JSON order policy the dumped shim generates with")
    args = ap.parse_args()
    if args.use_shim:
        cchess, source, version = cs, "shim", "oracle/cchess_shim.py"
        if args.shim_policy:
            cs.set_order_policy(json.loads(args.shim_policy))
    else:
        import cchess  # the real package: pip install from github.com/windshadow233/python-chinese-chess

        source, version = "cchess", getattr(cchess, "__version__", "unknown")
    entries = dump(cchess, args.games, args.max_plies, args.seed)
    cs.set_order_policy(None)
    policy, report = infer_policy(entries)
    sets_equal = all(sorted(e["legal"]) == sorted(m.uci() for m in shim_board(e).legal_moves) for e in entries)
    out = {"source": source, "cchess_version": version, "seed": args.seed, "n_entries": len(entries),
           "order_policy": policy, "policy_search": report, "move_sets_equal_oracle": sets_equal, "entries": entries}
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    if source == "cchess" and policy is not None:
        # the product picks the pinned order up at load time (chinesechesszero_b200/_lib.py)
        with open(os.path.join(ROOT, "chinesechesszero_b200", "order_policy.json"), "w") as f:
            json.dump(policy, f)
    print(f"{args.out}: {len(entries)} positions from {source}; move sets equal the oracle: {sets_equal}; "
          f"order policy: {policy if policy else 'NONE FITS (the generation order needs code)'}")


if __name__ == "__main__":
    main()
