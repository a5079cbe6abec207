"""Scratch: which pool configuration makes the golden scenario diverge, and at which move."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import cchess_shim as cs
from tests.test_mcts_gpu import fake_evaluator, root_children
from tests.test_mcts_oracle import load_golden
from tests.arena_util import walk_tree, pool_accounting
from chinesechesszero_b200.search import LockstepSearch

gold = load_golden(os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
for name in ("start_hash",):
    sc = next(s for s in gold["scenarios"] if s["name"] == name)
    print(name, "n_playout", sc["n_playout"], "moves", len(sc["moves"]), "pre", sc["pre"])
    for shift, npg in ((11, 65536), (7, 1 << 17), (11, 4096), (7, 8192), (8, 1 << 17), (9, 1 << 17)):
        G = 2
        s = LockstepSearch(n_games=G, nodes_per_game=npg, page_shift=shift, c_puct=float(sc["c_puct"]))
        s.set_roots(np.tile(np.array(sc["root_record"], dtype=np.uint8), (G, 1)))
        ev = fake_evaluator(sc["kind"])
        res = "ok"
        for k, mv in enumerate(sc["moves"]):
            s.run(ev, sc["n_playout"])
            acts, visits, qbits, rootn = root_children(s, 0)
            if not (acts == mv["acts"] and visits == mv["visits"] and qbits == mv["q_bits"]):
                bad = [i for i in range(len(visits)) if visits[i] != mv["visits"][i]]
                res = f"MISMATCH at move {k}: acts_ok={acts == mv['acts']} n_bad_visits={len(bad)} first={bad[:3]} rootn={rootn} vs {mv['root_visits']}"
                break
            s.advance(np.full(G, mv["chosen"], dtype=np.int16))
        print(f"  shift={shift} nodes_per_game={npg} grown={s.pool_grown} pages={s.arena.n_pages}: {res}", s.pool_stats())
