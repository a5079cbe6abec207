"""The pooled MCTS arena (ccz_arena, csrc/ccz_mcts.cuh): a tree can take as many pages as it needs --
the reference's tree has no capacity (mcts.py:31-39,168-178).  Covered here:
  * paging is invisible: 128-node pages (every expansion crosses a page) give the same visits / Q bits
    as the unmodified reference over multi-move scenarios with tree reuse;
  * structural invariants and page accounting after every advance;
  * growth (ccz_mcts_migrate): the migrated trees are node-for-node the old ones and the search goes on
    bit-identically;
  * the device-side guard (ccz_mcts_reserve): per-game, counted, recoverable;
  * the BASELINE-size soak: 4096 games x 400 playouts x 60 moves, noise on, through SelfPlayEngine --
    the run that died with `node_cap=65536` in round 1."""
import numpy as np
import pytest
import torch

from oracle import cchess_shim as cs
from tests import positions
from tests.arena_util import game_pages, pool_accounting, walk_tree
from tests.test_mcts_gpu import fake_evaluator, root_children
from tests.test_mcts_oracle import load_golden, scenario_names

pytestmark = pytest.mark.gpu


def board_hash_evaluator(n_games, seed=0, scale=1.0):
    """Device-resident stand-in net: logits and value are functions of the leaf position (a hash of the
    board record rotates a per-slot random logit row), so different positions get different priors and
    values without a forward pass.  Capture-safe, no host work."""
    from chinesechesszero_b200 import _lib

    g = torch.Generator(device="cuda").manual_seed(seed)
    base = torch.randn(n_games, 2086, device="cuda", generator=g) * scale
    w = torch.randint(1, 1 << 20, (96,), device="cuda", generator=g, dtype=torch.int64)
    col = torch.arange(2086, device="cuda").view(1, -1)

    def evaluator(planes, leaf_boards):
        h = (leaf_boards.to(torch.int64) * w).sum(1)
        pol = base.gather(1, (col + (h % 2086).view(-1, 1)) % 2086)
        val = ((h % 2001).to(torch.float32) / 1000.0 - 1.0) * 0.8
        return pol, _lib.POLICY_LOGITS, val

    evaluator.needs_planes = False
    return evaluator


@pytest.mark.parametrize("nodes_per_game", [1 << 17, 2048])
@pytest.mark.parametrize("name", scenario_names())
def test_tiny_pages_equal_reference_golden(name, nodes_per_game, golden_dir):
    """page_shift=7: a page holds one or two child runs, so almost every expansion pops a page and every
    advance compacts across many pages -- results must not change.  With 2048 nodes per game the pool is
    far too small and grows (ccz_mcts_migrate) before the first and between later searches."""
    from chinesechesszero_b200.search import LockstepSearch

    sc = next(s for s in load_golden(golden_dir)["scenarios"] if s["name"] == name)
    G = 2
    search = LockstepSearch(n_games=G, nodes_per_game=nodes_per_game, page_shift=7, c_puct=float(sc["c_puct"]))
    rec = np.array(sc["root_record"], dtype=np.uint8)
    search.set_roots(np.tile(rec, (G, 1)))
    id_of = cs.action_table()[0]
    for u in sc["pre"]:
        m = cs.Move.from_uci(u)
        search.advance(np.full(G, id_of[m.from_square, m.to_square], dtype=np.int16))
    ev = fake_evaluator(sc["kind"])
    for mv in sc["moves"]:
        search.run(ev, sc["n_playout"])
        search.check_status()
        for g in range(G):
            acts, visits, qbits, rootn = root_children(search, g)
            assert acts == mv["acts"] and visits == mv["visits"] and qbits == mv["q_bits"], (name, g)
            assert rootn == mv["root_visits"]
            walk_tree(search.arena, g)
        search.advance(np.full(G, mv["chosen"], dtype=np.int16))
        for g in range(G):
            walk_tree(search.arena, g)
        pool_accounting(search.arena)
    st = search.pool_stats()
    assert st["expand_failed"] == 0 and st["trees_dropped"] == 0
    assert (st["pool_grown"] >= 1) == (nodes_per_game == 2048)


def test_migrate_keeps_trees_and_search_continues_bit_identically():
    from chinesechesszero_b200.search import LockstepSearch

    G, P = 6, 120
    ev = board_hash_evaluator(G, seed=3)

    def run(grow_after_first):
        s = LockstepSearch(n_games=G, nodes_per_game=1 << 15, page_shift=8)
        s.run(ev, P)
        acts, visits, _ = s.root_visits()
        chosen = acts.gather(1, visits.argmax(1, keepdim=True).long()).view(-1).contiguous()
        s.advance(chosen)
        if grow_after_first:
            before = [walk_tree(s.arena, g) for g in range(G)]
            n0 = s.arena.n_pages
            assert s._grow(3 * n0) and s.arena.n_pages == 3 * n0 and s.pool_grown == 1
            for g in range(G):
                after = walk_tree(s.arena, g)
                for key in ("visits", "value", "prior", "move", "n_child", "depth"):
                    assert np.array_equal(before[g][key], after[key]), (g, key)
            pool_accounting(s.arena)
        s.run(ev, P)
        s.check_status()
        return [root_children(s, g) for g in range(G)], s.root_boards.cpu().numpy()

    plain, b0 = run(False)
    grown, b1 = run(True)
    assert plain == grown and np.array_equal(b0, b1)


def test_reserve_guard_drops_only_the_largest_trees():
    """A pool that cannot grow (max_pool_nodes) and runs short: the guard drops the sub-trees of the games
    above their share -- flagged per game, counted -- the other games keep theirs and every game keeps its
    position; the search that follows cannot fail."""
    from chinesechesszero_b200 import _lib
    from chinesechesszero_b200.search import LockstepSearch

    G, P, shift = 8, 40, 7
    w = _lib.search_pages(P, shift)            # 41 pages per game per search
    pool_pages = G * (w + 1) + 8               # just above the geometric floor
    s = LockstepSearch(n_games=G, nodes_per_game=(pool_pages << shift) // G, page_shift=shift,
                       max_pool_nodes=pool_pages << shift)
    assert s.arena.n_pages == pool_pages
    ev = board_hash_evaluator(G, seed=5)
    s.run(ev, P)
    assert s.pool_stats()["trees_dropped"] == 0
    # games 0..3 keep their whole tree (CCZ_ADVANCE_KEEP: ~20 pages each), games 4..7 restart (1 page each)
    keep_before = [walk_tree(s.arena, g) for g in range(4)]
    s.advance(np.array([_lib.ADVANCE_KEEP] * 4 + [_lib.ADVANCE_NEW_GAME] * 4, dtype=np.int16))
    for g in range(4):
        after = walk_tree(s.arena, g)
        for key in ("visits", "value", "prior", "move", "n_child", "depth"):
            assert np.array_equal(keep_before[g][key], after[key]), (g, key)
    boards_before = s.root_boards.cpu().numpy().copy()
    pages_before = [len(game_pages(s.arena, g)) for g in range(G)]
    free_before = s.pool_stats()["free_pages"]
    assert free_before < G * w                 # short: the guard has to act
    s.run(ev, P, may_sync=False)               # device-side guard alone (the resident path) + search
    st = s.pool_stats()
    status = s.arena.status.cpu().numpy()
    share = pool_pages // G - w
    expect_dropped = [g for g in range(G) if pages_before[g] > max(share, 1)]
    assert expect_dropped and all(g < 4 for g in expect_dropped)
    assert [g for g in range(G) if status[g] & _lib.STATUS_TREE_DROPPED] == expect_dropped
    assert st["trees_dropped"] == len(expect_dropped) and st["expand_failed"] == 0
    s.check_status()                           # dropped trees do not raise
    # positions untouched, searches consistent: a dropped game restarted from a fresh root (P-1 child visits)
    assert np.array_equal(s.root_boards.cpu().numpy(), boards_before)
    _, v, _ = s.root_visits()
    for g in expect_dropped:
        assert int(v[g].sum()) == P - 1
    for g in range(G):
        walk_tree(s.arena, g)
    pool_accounting(s.arena)


def test_forced_reply_roots_keep_their_subtrees():
    """Roots with one to four legal replies (lone king against rooks, two of them in check):
    update_with_move keeps most of the tree, the case that overflowed the fixed slabs of round 1.  Twelve
    consecutive moves from a pool that starts far too small; the pool grows with the trees (host path),
    nothing is dropped, and the visit bookkeeping of tree reuse holds: the children of a reused root carry
    kept-1 visits into the next search (mcts.py:168-178)."""
    from chinesechesszero_b200 import _lib
    from chinesechesszero_b200.search import LockstepSearch

    fens = ["3k5/9/9/9/9/9/9/9/9/3R1K3 b", "4k4/9/9/9/9/9/9/9/4R4/3K5 b", "3k5/9/9/9/9/9/9/9/3R5/R4K3 b"]
    recs = np.stack([positions.record_from_fen(f) for f in fens] * 2)
    for r in recs:
        assert 1 <= len(cs.Board.from_record(r).legal_moves) <= 4
    G, P = recs.shape[0], 200
    ev = board_hash_evaluator(G, seed=9)
    s = LockstepSearch(n_games=G, nodes_per_game=1 << 13, page_shift=8)
    s.set_roots(recs)
    kept = np.zeros(G, dtype=np.int64)
    peak = 0
    for move in range(12):
        s.run(ev, P)
        s.check_status()
        acts, visits, counts = (t.cpu().numpy() for t in s.root_visits())
        expect = np.where(kept > 0, kept - 1 + P, P - 1)
        assert np.array_equal(visits.sum(1), expect), move
        peak = max(peak, int(s.arena.n_nodes.max()))
        pick = visits.argmax(1)
        kept = visits[np.arange(G), pick].astype(np.int64)
        s.advance(acts[np.arange(G), pick].astype(np.int16))
        _, _, flags, _ = _lib.movegen_encode(s.root_boards, planes=False)
        over = ((flags & (_lib.FLAG_TIE_MASK | _lib.FLAG_NOMOVES)) != 0).to(torch.uint8)
        s.reset(over)  # finished games restart from the start position
        kept[over.cpu().numpy() != 0] = 0
        for g in range(G):
            walk_tree(s.arena, g)
        pool_accounting(s.arena)
    st = s.pool_stats()
    assert st["expand_failed"] == 0 and st["trees_dropped"] == 0 and st["pool_grown"] >= 1
    assert peak > P * 20  # kept sub-trees accumulated well beyond one search's growth from a 1-4 reply root


def test_soak_4096_games_400_playouts_60_moves():
    """BASELINE configs[2] geometry for 60 lockstep moves with noise, device-resident stand-in net (no
    forward pass: seconds of GPU time).  Round 1 died here after 26 moves.  Asserts: no pool event, exact
    visit bookkeeping in every slot every move (P-1 on a fresh root, kept-1+P on a reused one), finished
    slots restart, live-node statistics reported."""
    from chinesechesszero_b200 import _lib
    from chinesechesszero_b200.selfplay import SelfPlayEngine

    G, P, MOVES = 4096, 400, 60
    eng = SelfPlayEngine(board_hash_evaluator(G, seed=1234, scale=0.3), n_games=G, n_playout=P, seed=1234,
                         max_game_moves=48)
    s = eng.search
    kept = np.zeros(G, dtype=np.int64)          # visits of the child each slot advanced to (0 = fresh root)
    peak_nodes = np.zeros(G, dtype=np.int64)
    games = 0
    for move in range(MOVES):
        # replicate play_move() but look at the visit sums before the move is chosen
        s.run(eng.evaluator, P, ctl=eng._h_ctl.numpy() if eng._ctl_valid else None)
        _, visits_d, _ = s.root_visits()
        vs = visits_d.sum(1).cpu().numpy()
        expect = np.where(kept > 0, kept - 1 + P, P - 1)
        assert np.array_equal(vs, expect), (move, np.nonzero(vs != expect)[0][:8])
        peak_nodes = np.maximum(peak_nodes, s.arena.n_nodes.cpu().numpy())
        # the engine's own move: play_move() re-runs nothing (n_playout=0) and chooses, advances, refills
        eng.n_playout = 0
        finished = eng.play_move()
        eng.n_playout = P
        games += len(finished)
        # visits of the chosen child = what the next root starts with
        acts = eng._h_acts.numpy()
        vis = eng._h_visits.numpy()
        chosen = eng._h_chosen.numpy()
        pick = (acts == chosen[:, None]).argmax(1)
        kept = vis[np.arange(G), pick].astype(np.int64)
        for rec in finished:
            kept[rec.slot] = 0
        assert int(s.arena.status.abs().sum()) == 0, move
    st = s.pool_stats()
    assert st["expand_failed"] == 0 and st["trees_dropped"] == 0
    assert games >= G  # max_game_moves=48 < 60: every slot finished at least once and was refilled
    q = np.percentile(peak_nodes, [50, 99, 100]).astype(int).tolist()
    print(f"soak: peak live nodes per game median/p99/max = {q}, pool {st['n_pages']} pages x {st['page_nodes']}, "
          f"min free {st['min_free_pages']}, grown {st['pool_grown']}x")
    for g in (0, 150, 1053, 4095):
        walk_tree(s.arena, g)
    pool_accounting(s.arena)
