"""C-ABI behaviour on the device: status codes and ccz_last_error for bad arguments (no crash, no
silent fallback), the host-facing PolicyValueNet.policy_value_fn, and the device-resident move path."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import cchess_shim as cs
from oracle import net_oracle
from tests.test_mcts_gpu import fake_evaluator

pytestmark = pytest.mark.gpu


def test_bad_arguments_return_status_and_message():
    from chinesechesszero_b200 import _lib

    lib = _lib.load()
    boards = _lib.boards_start(4)
    ids = torch.empty((4, 128), dtype=torch.int16, device="cuda")
    counts = torch.empty(4, dtype=torch.int16, device="cuda")
    flags = torch.empty(4, dtype=torch.uint8, device="cuda")
    s = _lib.stream_ptr()
    assert lib.ccz_movegen_encode(boards.data_ptr(), -1, ids.data_ptr(), counts.data_ptr(), flags.data_ptr(), None, s) < 0
    assert b"n < 0" in lib.ccz_last_error()
    assert lib.ccz_movegen_encode(None, 4, ids.data_ptr(), counts.data_ptr(), flags.data_ptr(), None, s) < 0
    assert b"NULL" in lib.ccz_last_error()
    # misaligned board pointer
    assert lib.ccz_movegen_encode(boards.data_ptr() + 1, 1, ids.data_ptr(), counts.data_ptr(), flags.data_ptr(), None, s) < 0
    assert b"aligned" in lib.ccz_last_error()
    # n == 0 is a no-op
    assert lib.ccz_movegen_encode(None, 0, None, None, None, None, s) == 0
    with pytest.raises(_lib.CczError):
        _lib.movegen_encode(torch.zeros((2, 96), dtype=torch.uint8))  # host tensor: device pointers only
    a = _lib.Arena(2, 8, page_shift=7)
    bad = _lib.ArenaStruct.from_buffer_copy(a.struct)
    bad.d_links = None
    assert lib.ccz_mcts_reset(ctypes.byref(bad), None, s) < 0 and b"NULL" in lib.ccz_last_error()
    bad = _lib.ArenaStruct.from_buffer_copy(a.struct)
    bad.page_shift = 6  # a page must hold one child run
    assert lib.ccz_mcts_reset(ctypes.byref(bad), None, s) < 0 and b"geometry" in lib.ccz_last_error()
    bad = _lib.ArenaStruct.from_buffer_copy(a.struct)
    bad.n_pages = 1  # fewer pages than games
    assert lib.ccz_mcts_pool_init(ctypes.byref(bad), s) < 0 and b"geometry" in lib.ccz_last_error()
    b2 = _lib.Arena(3, 8, page_shift=7)
    chosen = torch.zeros(2, dtype=torch.int16, device="cuda")
    assert lib.ccz_mcts_migrate(a.ref, b2.ref, s) < 0 and b"n_games" in lib.ccz_last_error()
    assert lib.ccz_mcts_migrate(a.ref, a.ref, s) < 0 and b"distinct" in lib.ccz_last_error()
    assert lib.ccz_mcts_advance(a.ref, None, s) < 0 and b"NULL" in lib.ccz_last_error()
    assert lib.ccz_mcts_reserve(a.ref, 0, s) < 0 and lib.ccz_mcts_reserve(a.ref, 4, s) < 0  # 2 games x 5 pages > 8
    assert b"pool cannot hold" in lib.ccz_last_error()
    assert lib.ccz_mcts_reserve(a.ref, 3, s) == 0
    assert lib.ccz_mcts_expand_backup(a.ref, chosen.data_ptr(), chosen.data_ptr(), 7, chosen.data_ptr(),
                                      chosen.data_ptr(), chosen.data_ptr(), chosen.data_ptr(), s) < 0
    torch.cuda.synchronize()


def test_board_push_without_keys_and_invalid_ids():
    from chinesechesszero_b200 import _lib

    boards = _lib.boards_start(3)
    mv = torch.tensor([cs.action_table()[0][19, 22], -1, 5000], dtype=torch.int16, device="cuda")  # b2e2, skip, skip
    _lib.board_push(boards, mv, None)
    out = boards.cpu().numpy()
    ref = cs.Board()
    ref.push(cs.Move.from_uci("b2e2"))
    assert np.array_equal(out[0][:92], ref.record()[:92])
    assert np.array_equal(out[1], cs.start_record()) and np.array_equal(out[2], cs.start_record())


def test_policy_value_fn_matches_fp32_oracle():
    from chinesechesszero_b200.net import PolicyValueNet

    torch.manual_seed(0)
    pv = PolicyValueNet(num_channels=64, resblocks_num=4)
    net_oracle.perturb_(pv.policy_value_net.state_dict(), seed=3)
    board = cs.Board()
    for u in ["h2e2", "h9g7", "e2e6"]:  # cannon takes the e6 pawn with check
        board.push(cs.Move.from_uci(u))
    act_probs, value = pv.policy_value_fn(board.record())
    ids, probs = zip(*act_probs)
    id_of = cs.action_table()[0]
    assert list(ids) == [int(id_of[m.from_square, m.to_square]) for m in board.legal_moves]
    sd = {k: v.detach().cpu().float() for k, v in pv.policy_value_net.state_dict().items()}
    logp, v = net_oracle.forward(sd, net_oracle.search_planes(board.record()[None]))
    ref = np.exp(logp.numpy().reshape(-1))[list(ids)]
    assert np.abs(np.array(probs) - ref).max() < 1e-2 and abs(float(value[0, 0]) - float(v[0, 0])) < 1e-2
    assert value.shape == (1, 1)


def test_resident_move_path_matches_host_path_in_deterministic_mode():
    from chinesechesszero_b200.selfplay import SelfPlayEngine

    ev = fake_evaluator("hash")
    a = SelfPlayEngine(ev, n_games=5, n_playout=30, deterministic=True, node_cap=8192)
    b = SelfPlayEngine(ev, n_games=5, n_playout=30, deterministic=True, node_cap=8192)
    for _ in range(4):
        a.play_move()
        b.play_move_resident()
        assert torch.equal(a.search.root_boards, b.search.root_boards)
    assert int(b._res["finished"]) == 0 and a.total_moves == b.total_moves == 20
    # the resident path keeps every move's sample in its device ring: same data the host path recorded
    d = b.drain_resident()
    assert d["boards"].shape == (4, 5, 96) and d["pi"].shape == (4, 5, 128)
    for m in range(4):
        for g in range(5):
            f = a._frames[m]
            n = int(f["counts"][g])
            assert np.array_equal(d["boards"][m, g], f["boards"][g]) and int(d["counts"][m, g]) == n
            assert np.array_equal(d["acts"][m, g, :n], f["acts"][g, :n]) and int(d["moves"][m, g]) == int(f["moves"][g])
            assert np.allclose(d["pi"][m, g, :n], f["probs"][g, :n], rtol=0, atol=1e-12)
    assert not d["over"].any() and b.drain_resident()["boards"].shape[0] == 0


def test_cuda_graph_replay_equals_eager_search():
    """The captured lockstep step (K3 -> K1 -> bf16 forward -> K4/K5) replayed from CUDA graphs gives
    the same trees as eager launches, across advances (the graph holds the pool's pointers, not the trees)."""
    from chinesechesszero_b200.net import BatchedEvaluator, Net
    from chinesechesszero_b200.search import LockstepSearch

    torch.manual_seed(0)
    ev = BatchedEvaluator(Net(num_channels=32, resblocks_num=2).cuda().eval())
    outs = []
    for graphs in (False, True):
        s = LockstepSearch(n_games=16, node_cap=8192)
        if graphs:
            s.enable_graphs(ev)
        res = []
        for move in range(3):
            s.run(ev, 25)
            s.check_status()
            acts, visits, counts = s.root_visits()
            res.append((acts.clone(), visits.clone(), s.arena.nodes[s.arena.root.long(), 1].clone()))  # root Q bits
            chosen = acts.gather(1, visits.argmax(1, keepdim=True).long()).view(-1).contiguous()
            s.advance(chosen)
        outs.append(res)
    for (a0, v0, q0), (a1, v1, q1) in zip(*outs):
        assert torch.equal(a0, a1) and torch.equal(v0, v1) and torch.equal(q0, q1)


def test_graph_replay_follows_refreshed_weights():
    """after a training update BatchedEvaluator.refresh() re-folds the weights IN PLACE, so the
    CUDA graph captured before the update replays with the new weights (equal to an eager search over a
    fresh evaluator of the updated net), not with freed storage."""
    from chinesechesszero_b200.net import BatchedEvaluator, Net
    from chinesechesszero_b200.search import LockstepSearch
    from oracle import net_oracle

    def root_q(s):  # fp32 bit patterns of the root children's Q: sensitive to every net output of the search
        a = s.arena
        return torch.stack([a.children(int(a.root[g]))["value"].view(torch.int32) for g in range(a.n_games)])

    torch.manual_seed(2)
    net = Net(num_channels=32, resblocks_num=2).cuda().eval()
    ev = BatchedEvaluator(net)
    s = LockstepSearch(n_games=8, nodes_per_game=8192)
    s.enable_graphs(ev)
    s.run(ev, 100)                     # captures the graph
    assert s._graphs.get(0) is not None
    old_q = root_q(s)
    net_oracle.perturb_(net.state_dict(), seed=5)   # "one training step"
    ev.refresh(net)
    torch.cuda.empty_cache()
    junk = [torch.randn(1 << 20, device="cuda") for _ in range(8)]  # would land in freed weight blocks
    s.reset()
    s.run(ev, 100)                     # graph replay with refreshed weights
    got = root_q(s)
    ref = LockstepSearch(n_games=8, nodes_per_game=8192)
    ref.run(BatchedEvaluator(net), 100)   # eager, fresh evaluator
    assert torch.equal(got, root_q(ref))
    assert not torch.equal(got, old_q)
    del junk
