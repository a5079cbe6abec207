"""BASELINE-size runs checked through size-independent properties (the oracle is too slow there):
K1 on 1,048,576 positions (configs[1]) and the arena kernels on 4096 lockstep games (configs[2])."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_movegen_encode_one_million_positions_properties():
    from chinesechesszero_b200 import _lib, positions

    boards = positions.bench_positions(1 << 20, seed=0)
    n = boards.shape[0]
    assert n == 1 << 20
    ids, counts, flags, planes = _lib.movegen_encode(boards)
    c = counts.to(torch.int64)
    assert int(c.min()) >= 1 and int(c.max()) <= 119
    # ids valid, distinct and -1 padded exactly after the count
    col = torch.arange(128, device="cuda").view(1, -1)
    valid = col < c.view(-1, 1)
    assert bool(((ids >= 0) == valid).all()) and int(ids.max()) < 2086
    sorted_ids, _ = torch.sort(ids.to(torch.int32).masked_fill(~valid, 1 << 20), dim=1)
    dup = (sorted_ids[:, 1:] == sorted_ids[:, :-1]) & (sorted_ids[:, 1:] < (1 << 20))
    assert not bool(dup.any())
    # every move starts on a square holding a piece of the side to move
    from_of = torch.from_numpy(_lib.host_action_table()[1].astype(np.int64)).cuda()
    fr = from_of[ids.clamp_min(0).to(torch.int64)]
    code = torch.gather(boards[:, :90].to(torch.int64), 1, fr)
    red = boards[:, 90:91] != 0
    own = (code != 0) & (((code & 8) == 0) == red)
    assert bool((own | ~valid).all())
    # planes: ones exactly at the pieces (plays 7 / 15) and, for RED to move, the whole turn play
    p = planes.view(n, 17, 630)
    zero_plays = [i for i in range(16) if i not in (7, 15)]
    assert float(p[:, zero_plays].float().abs().sum()) == 0.0
    pieces = (boards[:, :90] != 0).sum(1)
    assert bool(((p[:, 7].float().sum(1) + p[:, 15].float().sum(1)) == pieces).all())
    turn_sum = p[:, 16].float().sum(1)
    assert bool((turn_sum == 630.0 * boards[:, 90].float()).all())
    assert bool(((planes == 0) | (planes == 1)).all())
    # flags: perft-3/4 positions from the start are never drawn; NOMOVES would need count 0
    assert int((flags & (_lib.FLAG_TIE_MASK | _lib.FLAG_NOMOVES)).sum()) == 0


def test_lockstep_search_4096_games_tree_invariants():
    from chinesechesszero_b200 import _lib
    from chinesechesszero_b200.search import LockstepSearch
    from tests.arena_util import pool_accounting, walk_tree

    G, P = 4096, 64
    torch.manual_seed(0)
    logits = torch.randn(G, 2086, device="cuda")
    values = torch.tanh(torch.randn(G, device="cuda") * 0.5)
    s = LockstepSearch(G, nodes_per_game=8192)
    ev = lambda planes, boards: (logits, _lib.POLICY_LOGITS, values)
    s.run(ev, P)
    s.check_status()
    a = s.arena
    acts, visits, counts = s.root_visits()
    # first search of a game: the root is expanded by playout 1, children receive P-1 visits (mcts.py:94, B.4)
    assert bool((visits.sum(1) == P - 1).all()) and bool((counts == 44).all())
    root = a.root.to(torch.int64)
    assert bool((a.nodes[root, 0] == P).all())
    nn = a.n_nodes.to(torch.int64)
    assert int(nn.min()) >= 45 and int(nn.max()) <= 1 + P * 119
    # whole trees: child runs inside one owned page, parent links, N(node) = 1 + sum N(children), node counts
    for g0 in (0, 17, 4095):
        t = walk_tree(a, g0)
        assert np.abs(t["value"]).max() <= 1.0 + 1e-6
    pool_accounting(a)
    # advancing every game keeps exactly the chosen sub-tree
    choice = visits.argmax(1, keepdim=True)
    chosen = acts.gather(1, choice.long()).view(-1).contiguous()
    kept_visits = visits.gather(1, choice.long()).view(-1)
    before = walk_tree(a, 17, check=False)
    s.advance(chosen)
    root2 = a.root.to(torch.int64)
    assert bool((a.nodes[root2, 0] == kept_visits).all()) and bool((a.links[root2, 0] == -1).all())
    after = walk_tree(a, 17)
    # the kept sub-tree is the chosen child's, node for node in breadth-first order
    kid = int(np.nonzero((before["depth"] == 1) & (before["move"] == int(chosen[17])))[0][0])
    sub = [kid]
    index_pos = {int(ix): i for i, ix in enumerate(before["index"])}
    h = 0
    while h < len(sub):
        i = sub[h]
        fc, nc = int(before["first_child"][i]), int(before["n_child"][i])
        sub.extend(index_pos[fc + k] for k in range(nc))
        h += 1
    for key in ("visits", "value", "prior", "move", "n_child"):
        assert np.array_equal(before[key][sub], after[key]), key
    st = pool_accounting(a)
    assert st["expand_failed"] == 0 and st["trees_dropped"] == 0
    s.run(ev, 8)
    s.check_status()
    _, v2, _ = s.root_visits()
    assert bool((v2.sum(1) == torch.clamp(kept_visits - 1, min=0) + 8).all())
    walk_tree(a, 17)
    pool_accounting(a)


@pytest.mark.parametrize("tap", [(1, 1), (0, 0), (2, 2), (0, 2)])
def test_conv_full_batch_shifted_identity_is_exact(tap):
    """K9 at the bench size (4096 boards = 2880 M-tiles): with a weight that is the identity on one filter tap
    the convolution is a pure board shift with zero padding, so y = relu(shift(x) + bias + skip) must hold
    EXACTLY in every tile, row and board (fp32 sum of exactly representable terms, one rounding to bf16)."""
    from chinesechesszero_b200 import _lib

    r, s = tap
    n = 4096
    cl = torch.channels_last
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(n, 256, 10, 9, device="cuda", generator=g).to(torch.bfloat16).contiguous(memory_format=cl)
    skip = torch.randn(n, 256, 10, 9, device="cuda", generator=g).to(torch.bfloat16).contiguous(memory_format=cl)
    bias = (torch.randn(256, device="cuda", generator=g) * 0.5).to(torch.bfloat16).float()
    w = torch.zeros(256, 256, 3, 3, device="cuda", dtype=torch.bfloat16)
    w[torch.arange(256), torch.arange(256), r, s] = 1.0
    w = w.contiguous(memory_format=cl)
    # cross-correlation: y[h, w] += x[h + r - 1, w + s - 1]
    shifted = torch.zeros_like(x, dtype=torch.float32)
    hs, ws = r - 1, s - 1
    dst_h = slice(max(0, -hs), 10 - max(0, hs))
    src_h = slice(max(0, hs), 10 - max(0, -hs))
    dst_w = slice(max(0, -ws), 9 - max(0, ws))
    src_w = slice(max(0, ws), 9 - max(0, -ws))
    shifted[:, :, dst_h, dst_w] = x[:, :, src_h, src_w].float()
    for sk in (None, skip):
        ref = shifted + bias.view(1, -1, 1, 1)
        if sk is not None:
            ref = ref + sk.float()
        ref = torch.relu(ref).to(torch.bfloat16)
        y = _lib.conv3x3_c256(x, w, bias, sk)
        assert torch.equal(y, ref.contiguous(memory_format=cl))
