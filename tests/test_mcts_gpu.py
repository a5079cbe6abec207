"""K3/K4/K5/K7 parity: the CUDA arena search (through the C ABI) vs the UNMODIFIED reference
mcts.py (golden file) and vs oracle.mcts_oracle.FlatMCTS run here, given identical policy outputs:
root actions in generation order, visit counts, child Q (fp32 bits), visit-softmax probabilities and
chosen moves must match exactly, across consecutive moves with tree reuse."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import cchess_shim as cs
from oracle import mcts_oracle
from tests import positions
from tests.test_mcts_oracle import load_golden, root_board, scenario_names

pytestmark = pytest.mark.gpu


def fake_evaluator(kind):
    """Device-facing wrapper of the deterministic stand-in policy (probabilities, bit-exact path)."""
    from chinesechesszero_b200 import _lib

    def evaluator(planes, leaf_boards):
        recs = leaf_boards.cpu().numpy()
        pol = np.empty((recs.shape[0], 2086), dtype=np.float32)
        val = np.empty((recs.shape[0],), dtype=np.float32)
        for i, rec in enumerate(recs):
            pol[i], val[i] = mcts_oracle.fake_policy_arrays(rec, kind)
        return torch.from_numpy(pol).cuda(), _lib.POLICY_PROBS, torch.from_numpy(val).cuda()

    return evaluator


def root_children(search, g):
    a = search.arena
    root = int(a.root[g])
    kids = a.children(root)
    return (kids["move"].cpu().numpy().tolist(), kids["visits"].cpu().numpy().tolist(),
            kids["value"].cpu().numpy().view(np.uint32).tolist(), int(a.nodes[root, 0]))


def make_search(sc, n_games):
    from chinesechesszero_b200.search import LockstepSearch

    search = LockstepSearch(n_games=n_games, node_cap=65536, c_puct=float(sc["c_puct"]))
    rec = np.array(sc["root_record"], dtype=np.uint8)
    search.set_roots(np.tile(rec, (n_games, 1)))
    id_of = cs.action_table()[0]
    for u in sc["pre"]:
        m = cs.Move.from_uci(u)
        search.advance(np.full(n_games, id_of[m.from_square, m.to_square], dtype=np.int16))
    return search


@pytest.mark.parametrize("name", scenario_names())
def test_device_search_equals_reference_golden(name, golden_dir):
    from chinesechesszero_b200.search import visit_softmax

    sc = next(s for s in load_golden(golden_dir)["scenarios"] if s["name"] == name)
    G = 3
    search = make_search(sc, G)
    ev = fake_evaluator(sc["kind"])
    for mv in sc["moves"]:
        search.run(ev, sc["n_playout"])
        search.check_status()
        acts_t, visits_t, counts_t = search.root_visits()
        torch.cuda.synchronize()
        for g in range(G):
            acts, visits, qbits, rootn = root_children(search, g)
            assert acts == mv["acts"], (name, g)
            assert visits == mv["visits"], (name, g)
            assert qbits == mv["q_bits"], (name, g)
            assert rootn == mv["root_visits"]
            n = int(counts_t[g])
            assert acts_t[g, :n].tolist() == mv["acts"] and visits_t[g, :n].tolist() == mv["visits"]
            assert (acts_t[g, n:] == -1).all()
            probs = visit_softmax(np.array(visits), 1.0)
            assert [float(p).hex() for p in probs] == mv["probs_hex"]
            assert int(acts[int(np.argmax(visits))]) == mv["chosen"]
        search.advance(np.full(G, mv["chosen"], dtype=np.int16))
    # root board after the played moves equals the oracle board (squares, turn, clock, repetition)
    board = root_board(sc)
    for mv in sc["moves"]:
        board.push(mcts_oracle.move_from_id(mv["chosen"]))
    assert np.array_equal(search.root_boards[0].cpu().numpy(), board.record())


def test_mixed_roots_in_lockstep_vs_flat_oracle():
    """Different positions in different game slots of one lockstep batch, checked per slot against
    the CPU oracle search; includes terminal roots' neighbours, high clocks and endgames."""
    from chinesechesszero_b200.search import LockstepSearch

    recs = np.concatenate([positions.edge_case_records(),
                           positions.random_playout_positions(6, 200, seed=77, every=23)])
    keep = []
    for r in recs:  # searchable roots only (the game loop never searches a finished game)
        b = cs.Board.from_record(r)
        if not b.is_game_over():
            keep.append(r)
    recs = np.stack(keep)
    G, n_playout = recs.shape[0], 90
    search = LockstepSearch(n_games=G, node_cap=16384)
    search.set_roots(recs)
    search.run(fake_evaluator("hash"), n_playout)
    search.check_status()
    torch.cuda.synchronize()
    for g in range(G):
        o = mcts_oracle.FlatMCTS(mcts_oracle.make_policy("hash"), c_puct=5, n_playout=n_playout)
        o.get_move_probs(cs.Board.from_record(recs[g]), temp=1.0)
        acts, visits, q = o.root_children()
        d_acts, d_visits, d_q, _ = root_children(search, g)
        assert d_acts == acts and d_visits == visits, g
        assert d_q == [int(np.float32(x).view(np.uint32)) for x in q], g


def test_reset_and_new_game_slots():
    from chinesechesszero_b200.search import LockstepSearch

    search = LockstepSearch(n_games=4, node_cap=8192)
    ev = fake_evaluator("hash")
    search.run(ev, 40)
    first = [root_children(search, g) for g in range(4)]
    assert all(f == first[0] for f in first)
    # slot 1 restarts (-1), slot 2 keeps its position but drops the tree (-2), others play a move
    mv = first[0][0][0]
    search.advance(np.array([mv, -1, -2, mv], dtype=np.int16))
    rb = search.root_boards.cpu().numpy()
    assert np.array_equal(rb[1], cs.start_record()) and np.array_equal(rb[2], cs.start_record())
    assert not np.array_equal(rb[0], cs.start_record()) and np.array_equal(rb[0], rb[3])
    search.run(ev, 40)
    again = [root_children(search, g) for g in range(4)]
    assert again[1] == first[1] and again[2] == first[2]  # fresh trees reproduce the first search
    assert again[0] == again[3]


def test_logits_policy_matches_probs_policy():
    """CCZ_POLICY_LOGITS (softmax fused into the gather) vs CCZ_POLICY_PROBS fed with
    torch.softmax of the same logits: priors agree to fp32 rounding (<=1e-6 abs)."""
    from chinesechesszero_b200 import _lib
    from chinesechesszero_b200.search import LockstepSearch

    torch.manual_seed(0)
    G = 8
    logits = torch.randn(G, 2086, device="cuda") * 3
    vals = torch.zeros(G, device="cuda")
    priors = []
    for kind, pol in ((_lib.POLICY_LOGITS, logits), (_lib.POLICY_PROBS, torch.softmax(logits, dim=1))):
        s = LockstepSearch(n_games=G, node_cap=4096)
        s.step(lambda planes, boards: (pol, kind, vals))
        a = s.arena
        priors.append(torch.stack([a.children(int(a.root[g]))["prior"] for g in range(G)]).cpu())
        assert priors[-1].shape == (G, 44)
    assert torch.allclose(priors[0], priors[1], atol=1e-6, rtol=1e-5)
    assert float(priors[0].sum()) > 0


def test_pool_exhaustion_is_per_game_and_counted():
    """Raw C-ABI use without the reserve guard on a pool with no spare page: the leaf whose child run does
    not fit stays unexpanded (status bit on THAT game only, cumulative counter), every other game and the
    next searches of the same game carry on; LockstepSearch.check_status names the game."""
    from chinesechesszero_b200 import _lib
    from chinesechesszero_b200.search import LockstepSearch

    s = LockstepSearch(n_games=2, nodes_per_game=256, page_shift=7)  # 4 pages of 128 nodes
    assert s.arena.n_pages == 4
    ev = fake_evaluator("hash")
    # two playouts per game fit (root run of 44 + one more run in the game's own page + one popped page)
    for _ in range(2):
        s.step(ev)
    assert s.pool_stats()["expand_failed"] == 0
    for _ in range(6):  # 2 spare pages only: both games run dry within a few expansions
        s.step(ev)
    st = s.pool_stats()
    assert st["expand_failed"] >= 1 and st["free_pages"] == 0 and st["min_free_pages"] == 0
    status = s.arena.status.cpu().numpy()
    assert (status & _lib.STATUS_EXPAND_FAILED).any()
    with pytest.raises(_lib.CczError, match="pool exhausted"):
        s.check_status()
    # visits stay consistent: every playout was backed up to the root even when its leaf was not expanded
    assert [int(s.arena.nodes[int(s.arena.root[g]), 0]) for g in range(2)] == [8, 8]
    # a new game in the slot gives the pages back and clears the sticky bit; the counter stays
    s.reset()
    assert int(s.arena.status.abs().sum()) == 0
    st2 = s.pool_stats()
    assert st2["free_pages"] == 2 and st2["expand_failed"] == st["expand_failed"]
    s.run(ev, 2)  # through the guard: 2 playouts need 3 pages per game -> the pool grows instead of failing
    assert s.pool_stats()["expand_failed"] == st["expand_failed"] and s.pool_grown == 1
    s.check_status()
