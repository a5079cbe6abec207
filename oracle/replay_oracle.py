"""replay_oracle -- CPU ORACLE (TEST INFRASTRUCTURE ONLY): NumPy restatement of the reference's
sample post-processing, for checking K8 (ccz_replay_pack) and chinesechesszero_b200.replay.

  Game history lists            game.py:23-44      -> history_states
  CollectPipeline.preprocess    collect.py:64-112  -> preprocess
  CollectPipeline.flip_data     collect.py:114-131 -> flip_data
"""
from __future__ import annotations

import numpy as np

from . import cchess_shim as cs


def decode(rec):
    red = np.zeros(630, dtype=np.int8)
    black = np.zeros(630, dtype=np.int8)
    raw = np.ascontiguousarray(rec, dtype=np.uint8)
    cs.lib().xq_decode_board(raw.ctypes.data, red.ctypes.data, black.ctypes.data)
    return red.reshape(7, 10, 9), black.reshape(7, 10, 9)


def history_states(boards, upto):
    """The two 8-deep lists as game.py keeps them after `upto`+1 calls of update_states_history():
    8 copies of the initial position, then each searched position inserted at the front."""
    r0, b0 = decode(boards[0])
    red = [r0.copy() for _ in range(8)]
    black = [b0.copy() for _ in range(8)]
    for i in range(upto + 1):
        r, b = decode(boards[i])
        red.pop(); red.insert(0, r)
        black.pop(); black.insert(0, b)
    return red, black


def preprocess(red_states, black_states, turn_is_red, mcts_prob):
    """collect.py:74-112 for one sample: (states (17,7,10,9) float16, prob float64)."""
    current_player = (np.ones if turn_is_red else np.zeros)((1, 7, 10, 9), dtype=np.float16)
    states = np.concatenate((red_states, black_states), axis=0)
    states = np.concatenate((states, current_player), axis=0)
    prob_sum = np.sum(mcts_prob)
    if abs(prob_sum - 1.0) > 1e-6:
        mcts_prob = mcts_prob / prob_sum
    return states, mcts_prob


def flip_map():
    """collect.py:117-122 via the C action table."""
    id_of, fr, to = cs.action_table()
    out = np.empty(2086, dtype=np.int64)
    for i in range(2086):
        f, t = int(fr[i]), int(to[i])
        mf, mt = (8 - f % 9) + 9 * (f // 9), (8 - t % 9) + 9 * (t // 9)
        out[i] = id_of[mf, mt]
    return out


def flip_data(data):
    """collect.py:114-131: data + mirrored data."""
    fm = flip_map()
    flipped = []
    for states, prob, winner in data:
        flipped.append((np.stack([np.flip(s, axis=2) for s in states]), prob[fm], winner))
    return data + flipped


def pack_reference(boards, dense_probs, turns, z, states_mode):
    """Full per-game output in the reference layout: (states (2T,17,7,10,9) f16, probs (2T,2086) f64, z (2T,))."""
    t = len(boards)
    data = []
    for i in range(t):
        upto = t - 1 if states_mode == "reference" else i
        red, black = history_states(boards, upto)
        turn = True if states_mode == "reference" else bool(turns[i])
        states, prob = preprocess(red, black, turn, dense_probs[i])
        data.append((states, prob, z[i]))
    data = flip_data(data)
    return (np.stack([d[0] for d in data]).astype(np.float16), np.stack([d[1] for d in data]).astype(np.float64),
            np.array([d[2] for d in data], dtype=np.float64))
