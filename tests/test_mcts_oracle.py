"""Pin oracle/mcts_oracle.FlatMCTS (the checker used on the GPU box) to the UNMODIFIED reference
mcts.py via the committed golden file tests/golden/mcts_reference.json: per-move root actions in
generation order, visit counts, child Q values (fp32 bit patterns) and visit-softmax probabilities
must be identical."""
import json
import os

import numpy as np
import pytest

from oracle import cchess_shim as cs
from oracle import mcts_oracle


def load_golden(golden_dir):
    with open(os.path.join(golden_dir, "mcts_reference.json")) as f:
        return json.load(f)


def root_board(sc):
    board = cs.Board.from_record(np.array(sc["root_record"], dtype=np.uint8))
    for u in sc["pre"]:
        board.push(cs.Move.from_uci(u))
    return board


def scenario_names(golden_dir=os.path.join(os.path.dirname(__file__), "golden")):
    return [sc["name"] for sc in load_golden(golden_dir)["scenarios"]]


@pytest.mark.parametrize("name", scenario_names())
def test_flat_oracle_equals_reference(name, golden_dir):
    sc = next(s for s in load_golden(golden_dir)["scenarios"] if s["name"] == name)
    board = root_board(sc)
    search = mcts_oracle.FlatMCTS(mcts_oracle.make_policy(sc["kind"]), c_puct=sc["c_puct"], n_playout=sc["n_playout"])
    for mv in sc["moves"]:
        acts, probs = search.get_move_probs(board, temp=1.0)
        a, visits, q = search.root_children()
        assert list(acts) == mv["acts"]
        assert visits == mv["visits"]
        assert [int(np.float32(x).view(np.uint32)) for x in q] == mv["q_bits"]
        assert [float(p).hex() for p in probs] == mv["probs_hex"]
        assert search.N[search.root] == mv["root_visits"]
        search.update_with_move(mv["chosen"])
        board.push(mcts_oracle.move_from_id(mv["chosen"]))


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference only in the authoring container")
def test_golden_is_reproducible_from_reference(golden_dir):
    """Re-run the unmodified reference on one scenario and compare with the committed file."""
    from scripts import golden_mcts

    gold = load_golden(golden_dir)["scenarios"][0]
    again = golden_mcts.run_reference(golden_mcts.SCENARIOS[0])
    assert again["moves"] == gold["moves"]
