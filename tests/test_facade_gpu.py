"""Reference-shaped façades on the GPU: board.Board vs the shim, MCTS_AI.get_action vs the unmodified
reference (seeded global NumPy RNG), Game.start_self_play + K8 replay packing vs the unmodified
reference game.py / collect.py output (sha256 of the saved arrays)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import cchess_shim as cs
from oracle import mcts_oracle, replay_oracle
from tests.test_game_oracle import load as load_game_golden
from tests.test_game_oracle import sha
from tests.test_mcts_gpu import fake_evaluator
from tests.test_mcts_oracle import load_golden

pytestmark = pytest.mark.gpu


def device_hash_evaluator():
    ev = fake_evaluator("hash")
    ev.device_evaluator = True
    return ev


def test_board_facade_matches_shim():
    from chinesechesszero_b200.board import Board

    rng = np.random.default_rng(3)
    for _ in range(3):
        d, s = Board(), cs.Board()
        for ply in range(150):
            assert [m.uci() for m in d.legal_moves] == [m.uci() for m in s.legal_moves]
            assert d.flags() == s.flags()
            assert np.array_equal(d.record(), s.record())
            assert d.turn == s.turn and d.is_game_over() == s.is_game_over()
            if s.is_game_over():
                assert d.outcome().winner == s.outcome().winner
                break
            mv = s.legal_moves[int(rng.integers(len(s.legal_moves)))]
            s.push(mv)
            d.push(type(d.legal_moves[0]).from_uci(mv.uci()))
    # repetition through the key window: knights out and back three times
    d = Board()
    for cycle in range(3):
        for u in ["b0c2", "b9c7", "c2b0", "c7b9"]:
            d.push(type(d.legal_moves[0]).from_uci(u))
        assert d.record()[92] == cycle + 1
    assert d.is_fourfold_repetition() and d.is_game_over() and d.outcome().winner is None
    c = d.copy()
    c.push(c.legal_moves[0])
    assert d.is_fourfold_repetition() and len(d.move_stack) == 12 and len(c.move_stack) == 13


def test_get_action_matches_reference_golden(golden_dir):
    from chinesechesszero_b200.board import Board
    from chinesechesszero_b200.mcts import MCTS_AI

    for case in load_golden(golden_dir)["get_action"]:
        ai = MCTS_AI(device_hash_evaluator(), c_puct=5, n_playout=case["n_playout"], is_selfplay=case["is_selfplay"])
        board = Board()
        np.random.seed(case["seed"])
        for step in case["steps"]:
            temp = 1.0 if case["is_selfplay"] else 1e-3
            move, probs = ai.get_action(board, temp=temp, return_prob=True)
            nz = np.nonzero(probs)[0]
            assert int(move) == step["move"]
            assert nz.tolist() == step["nz"]
            assert [float(probs[i]).hex() for i in nz] == step["probs_hex"]
            board.push(int(move))


def test_game_facade_and_replay_pack_match_reference(golden_dir):
    from chinesechesszero_b200 import replay
    from chinesechesszero_b200.board import Board
    from chinesechesszero_b200.game import Game
    from chinesechesszero_b200.mcts import MCTS_AI
    from chinesechesszero_b200.selfplay import GameRecord

    gold = load_game_golden(golden_dir)
    np.random.seed(gold["seed"])
    ai = MCTS_AI(device_hash_evaluator(), c_puct=5, n_playout=gold["n_playout"], is_selfplay=True)
    game = Game(Board())
    play_data = game.start_self_play(ai)
    ucis = [m.uci() for m in game.board.move_stack]
    assert ucis == gold["moves"]
    z = np.array([d[3] for d in play_data])
    assert z.tolist() == gold["z"]
    probs = np.stack([d[2] for d in play_data])
    assert sha(probs) == gold["probs_sha"]
    # final aliased histories (game.py:234-237) -> the reference's preprocess + flip via the oracle
    assert all(d[0] is play_data[0][0] and d[1] is play_data[0][1] for d in play_data)
    # K8: rebuild the game as a GameRecord and pack it in reference mode
    b = Board()
    boards, turns = [], []
    for u in ucis:
        boards.append(b.record())
        turns.append(b.turn)
        b.push(type(b.legal_moves[0]).from_uci(u))
    acts = [np.nonzero(p)[0].astype(np.int16) for p in probs]
    rec = GameRecord(boards=np.stack(boards), acts=acts, probs=[p[a] for p, a in zip(probs, acts)],
                     turns=np.array(turns), moves=np.zeros(len(ucis), np.int16), winner=None, z=z, final_flags=0)
    states, mcts_probs, winners = replay.pack_game(rec, "reference")
    assert list(states.shape) == gold["states_shape"] and str(states.dtype) == gold["states_dtype"]
    assert sha(states) == gold["states_sha"]
    assert sha(mcts_probs) == gold["mcts_probs_sha"]
    assert winners.tolist() == gold["winners"]
    # per_move mode against the NumPy restatement
    s2, p2, w2 = replay.pack_game(rec, "per_move")
    o_s, o_p, o_w = replay_oracle.pack_reference(rec.boards, probs, turns, z, "per_move")
    assert np.array_equal(s2, o_s) and np.array_equal(p2, o_p) and np.array_equal(w2, o_w)


def test_collect_pipeline_writes_npy_triple(tmp_path):
    from chinesechesszero_b200.collect import CollectPipeline

    pipe = CollectPipeline(n_games=8, n_playout=6, data_dir=str(tmp_path), max_game_moves=5, node_cap=4096,
                           net_kwargs=dict(num_channels=32, resblocks_num=2))
    n = pipe.run(max_games=8)
    assert n >= 8
    states = np.load(tmp_path / "states.npy")
    mcts = np.load(tmp_path / "mcts.npy")
    winners = np.load(tmp_path / "winners.npy")
    meta = json.load(open(tmp_path / "meta.json"))
    assert states.dtype == np.float16 and mcts.dtype == np.float64 and winners.dtype == np.float32
    assert states.shape[1:] == (17, 7, 10, 9) and mcts.shape[1] == 2086
    assert states.shape[0] == mcts.shape[0] == winners.shape[0] == meta["total_count"] == n * 5 * 2
    assert np.allclose(mcts.sum(axis=1), 1.0, atol=1e-9)
    assert (states[:, 16] == 1).all()  # reference mode: turn plane always ones
    # the same games in the reference's data.h5 layout (collect.py:146-167)
    from chinesechesszero_b200 import h5lite

    with h5lite.H5Reader(str(tmp_path / "data.h5")) as r:
        assert int(r.root_attrs()["iters"]) == n
        rows = 0
        for k in range(n):
            d = r.read_group(f"game_{k}")
            t2 = d["states"].shape[0]
            assert np.array_equal(d["states"], states[rows:rows + t2])
            assert np.array_equal(d["mcts_probs"], mcts[rows:rows + t2])
            assert np.array_equal(d["winners"].astype(np.float32), winners[rows:rows + t2])
            rows += t2
        assert rows == states.shape[0]
    # a second pipeline on the same directory continues the game counter (collect.py:39-45)
    pipe2 = CollectPipeline(n_games=8, n_playout=6, data_dir=str(tmp_path), max_game_moves=5, node_cap=4096,
                            net_kwargs=dict(num_channels=32, resblocks_num=2), write_npy=False)
    assert pipe2.iters == n
    pipe2.run(max_games=8)
    with h5lite.H5Reader(str(tmp_path / "data.h5")) as r:
        assert int(r.root_attrs()["iters"]) == pipe2.iters >= n + 8


def test_mcts_ai_with_real_net_uses_graphs_and_reports_progress():
    from chinesechesszero_b200.board import Board
    from chinesechesszero_b200.mcts import MCTS_AI
    from chinesechesszero_b200.net import PolicyValueNet

    torch.manual_seed(0)
    pv = PolicyValueNet(num_channels=32, resblocks_num=2)
    ai = MCTS_AI(pv.policy_value_fn, c_puct=5, n_playout=200, is_selfplay=True)
    assert ai.mcts._search._graphs is not None
    board = Board()
    seen = []
    np.random.seed(0)
    for _ in range(3):
        move, probs = ai.get_action(board, temp=1.0, return_prob=True, on_playout=seen.append)
        assert abs(probs.sum() - 1.0) < 1e-9 and probs[int(move)] > 0
        assert int(move) in board.legal_ids().tolist()
        board.push(int(move))
    assert sum(seen) == 600 and max(seen) == 2


def test_selfplay_train_loop_refreshes_the_evaluator(tmp_path):
    from chinesechesszero_b200.loop import SelfPlayTrainLoop

    torch.manual_seed(0)
    loop = SelfPlayTrainLoop(n_games=8, n_playout=6, data_dir=str(tmp_path / "data"), model_dir=str(tmp_path / "models"),
                             batch_size=16, games_per_iteration=8, max_game_moves=4, node_cap=4096,
                             net_kwargs=dict(num_channels=32, resblocks_num=2))
    ev = loop.collect.policy_value_net.evaluator()
    w0 = ev.stem[0].clone()
    r1 = loop.iterate()
    assert r1["games"] >= 8 and r1["samples"] == r1["games"] * 4 * 2 and np.isfinite(r1["loss"])
    assert not torch.equal(w0, ev.stem[0])          # trained weights were folded back into the evaluator
    r2 = loop.iterate()
    assert r2["games"] >= 16 and r2["samples"] > r1["samples"]
    loop.close()
    assert (tmp_path / "models" / "current_policy.pkl").exists() and (tmp_path / "data" / "data.h5").exists()
    # the checkpoint is a plain state_dict with the reference's keys (net.py:208-209)
    sd = torch.load(tmp_path / "models" / "current_policy.pkl", map_location="cpu")
    assert "conv_block.weight" in sd and "policy_fc.bias" in sd
