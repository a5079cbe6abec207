"""Generate the committed golden fixtures under tests/golden/ from the UNMODIFIED reference.

Runs only in the authoring container (needs /root/reference).  The reference modules are
imported as they are, with oracle.cchess_shim standing in for the absent `cchess` package
(oracle/load_reference.py).  Usage:  python scripts/make_golden.py [table|mcts|net|game|all]
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from oracle import load_reference  # noqa: E402


def make_table():
    """tools.py:172-272 action table and tools.py:133-164 flip(), straight from the reference."""
    tools = load_reference.load("tools")
    ids = [tools.move_id2move_action[i] for i in range(len(tools.move_id2move_action))]
    assert all(tools.move_action2move_id[a] == i for i, a in enumerate(ids))
    flip = [tools.move_action2move_id[tools.flip(a)] for a in ids]
    with open(os.path.join(GOLDEN, "action_table.json"), "w") as f:
        json.dump({"source": "reference tools.py get_all_legal_moves()/flip()", "move_id2move_action": ids,
                   "flip_map": flip}, f)
    print("action_table.json:", len(ids), "actions")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    os.makedirs(GOLDEN, exist_ok=True)
    if what in ("table", "all"):
        make_table()
    if what in ("mcts", "all"):
        from scripts import golden_mcts
        golden_mcts.main(GOLDEN)
    if what in ("game", "all"):
        from scripts import golden_game
        golden_game.main(GOLDEN)
    if what in ("train", "all"):
        from scripts import golden_train
        golden_train.main(GOLDEN)
    if what in ("net", "all"):
        from scripts import golden_net
        golden_net.main(GOLDEN)
