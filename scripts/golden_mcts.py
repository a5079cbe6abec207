"""Golden vectors for the search path, produced by the UNMODIFIED reference mcts.py
(/root/reference/mcts.py: Node, MCTS.playout/get_move_probs/update_with_move, MCTS_AI.get_action)
driven by the cchess shim board and the deterministic stand-in policies of oracle/mcts_oracle.py.
Run in the authoring container only:  python scripts/make_golden.py mcts
"""
from __future__ import annotations

import json
import os

import numpy as np

from oracle import cchess_shim as cs
from oracle import load_reference, mcts_oracle
from tests import positions

SCENARIOS = [
    # name, root FEN (None = start), clock, pre-moves (uci), policy kind, n_playout, n_moves, temp
    dict(name="start_hash", fen=None, clock=0, pre=[], kind="hash", n_playout=120, n_moves=6),
    dict(name="start_uniform", fen=None, clock=0, pre=[], kind="uniform", n_playout=150, n_moves=4),
    dict(name="mate_in_one", fen="3k5/9/9/9/9/9/9/9/4R4/R4K3 w", clock=0, pre=[], kind="hash", n_playout=200,
         n_moves=3),
    dict(name="sixty_in_tree", fen=None, clock=117, pre=[], kind="hash", n_playout=160, n_moves=3),
    dict(name="repetition_in_tree", fen=None, clock=0,
         pre=["b0c2", "b9c7", "c2b0", "c7b9", "b0c2", "b9c7", "c2b0", "c7b9"], kind="uniform", n_playout=200,
         n_moves=3),
    dict(name="endgame_hash", fen="4k4/4a4/9/9/4p4/9/9/4C4/4A4/3K1R3 w", clock=5, pre=[], kind="hash",
         n_playout=250, n_moves=5),
]


def _root_board(sc):
    if sc["fen"] is None:
        rec = cs.start_record()
    else:
        rec = positions.record_from_fen(sc["fen"])
    rec[91] = sc["clock"]
    board = cs.Board.from_record(rec)
    for u in sc["pre"]:
        board.push(cs.Move.from_uci(u))
    return rec, board


def _f32_bits(x) -> int:
    return int(np.asarray(x, dtype=np.float32).reshape(-1)[0].view(np.uint32))


def run_reference(sc):
    ref_tools, ref_mcts = load_reference.load("tools", "mcts")
    base = mcts_oracle.make_policy(sc["kind"])

    def policy_value_fn(board, red_states=None, black_states=None):
        ids, probs, value = base(board)
        # same return types as PolicyValueNet.policy_value_fn (net.py:202-205)
        return zip(ids, probs), np.array([[value]], dtype=np.float32)

    rec, board = _root_board(sc)
    search = ref_mcts.MCTS(policy_value_fn, c_puct=5, n_playout=sc["n_playout"])
    moves = []
    for _ in range(sc["n_moves"]):
        if board.is_game_over() or ref_tools.is_tie(board):
            break
        acts, probs = search.get_move_probs(board, temp=1.0)
        kids = list(search.root.children.items())
        assert tuple(a for a, _ in kids) == tuple(acts)
        visits = [n.visits for _, n in kids]
        qbits = [_f32_bits(n.value) for _, n in kids]
        chosen = int(acts[int(np.argmax(visits))])  # first maximum
        moves.append(dict(acts=[int(a) for a in acts], visits=[int(v) for v in visits], q_bits=qbits,
                          probs_hex=[float(p).hex() for p in probs], root_visits=int(search.root.visits),
                          chosen=chosen))
        search.update_with_move(chosen)
        board.push(cs.Move.from_uci(ref_tools.move_id2move_action[chosen]))
    return dict(name=sc["name"], root_record=rec.tolist(), pre=sc["pre"], kind=sc["kind"],
                n_playout=sc["n_playout"], c_puct=5, moves=moves)


def run_get_action():
    """MCTS_AI.get_action (mcts.py:203-233) with the global NumPy RNG seeded: move + 2086-vector."""
    ref_mcts = load_reference.load("mcts")
    base = mcts_oracle.make_policy("hash")

    def policy_value_fn(board, red_states=None, black_states=None):
        ids, probs, value = base(board)
        return zip(ids, probs), np.array([[value]], dtype=np.float32)

    out = []
    for selfplay in (True, False):
        ai = ref_mcts.MCTS_AI(policy_value_fn, c_puct=5, n_playout=60, is_selfplay=selfplay)
        board = cs.Board()
        np.random.seed(2024)
        steps = []
        for i in range(3):
            temp = 1.0 if selfplay else 1e-3
            move, probs = ai.get_action(board, temp=temp, return_prob=True)
            nz = np.nonzero(probs)[0]
            steps.append(dict(move=int(move), nz=[int(i) for i in nz], probs_hex=[float(probs[i]).hex() for i in nz]))
            board.push(mcts_oracle.move_from_id(int(move)))
        out.append(dict(is_selfplay=selfplay, seed=2024, n_playout=60, steps=steps))
    return out


def main(golden_dir):
    data = dict(source="unmodified /root/reference/mcts.py driven by oracle.cchess_shim + oracle.mcts_oracle policies",
                scenarios=[run_reference(sc) for sc in SCENARIOS], get_action=run_get_action())
    path = os.path.join(golden_dir, "mcts_reference.json")
    with open(path, "w") as f:
        json.dump(data, f)
    for sc in data["scenarios"]:
        print(sc["name"], "moves:", [m["chosen"] for m in sc["moves"]], "root visits:", [m["root_visits"] for m in sc["moves"]])
    print("wrote", path, os.path.getsize(path), "bytes")
