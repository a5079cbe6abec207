"""collect_oracle -- CPU ORACLE / reference-arm port (TEST INFRASTRUCTURE + bench.py baselines ONLY).

The reference's self-play data path restated end to end on the CPU so that it can be timed on
the GPU box's host cores (where /root/reference does not exist):

  CollectPipeline.collect_data -> Game.start_self_play   collect.py:133-142, game.py:133-237
  MCTS_AI.get_action / MCTS.get_move_probs / playout     mcts.py:101-166, 203-233
  PolicyValueNet.policy_value_fn, CPU branch             net.py:151-205 (batch-1 fp32 forward)

Board = oracle.cchess_shim (C-backed, so cheaper than the pure-Python cchess the reference uses:
the baseline errs in the reference's favour), search = oracle.mcts_oracle.FlatMCTS (pinned
bit-exact to the unmodified mcts.py), net = oracle.net_oracle.forward (pinned to net.Net).
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import cchess_shim as cs
from . import mcts_oracle, net_oracle

EPS, ALPHA = 0.25, 0.2  # parameters.py:10-12


class PortedSelfPlay:
    def __init__(self, state_dict=None, n_playout=400, c_puct=5, seed=0, num_channels=256, resblocks_num=40,
                 threads=None):
        if threads:
            torch.set_num_threads(threads)
        torch.manual_seed(seed)
        np.random.seed(seed)
        if state_dict is None:
            state_dict = random_state_dict(num_channels, resblocks_num)
        self.sd = state_dict
        self.n_evals = 0
        self.search = mcts_oracle.FlatMCTS(self.policy_value_fn, c_puct=c_puct, n_playout=n_playout)
        self.new_game()

    def new_game(self):
        self.board = cs.Board()
        self.search.reset()
        self.move_count = 0
        self.samples = []

    # net.py:151-205, CPU branch: legal ids, search-time planes, batch-1 fp32 forward, exp, gather
    def policy_value_fn(self, board):
        ids = mcts_oracle.legal_ids(board)
        x = net_oracle.search_planes(board.record()[None])
        logp, v = net_oracle.forward(self.sd, x)
        self.n_evals += 1
        probs = np.exp(logp.numpy().flatten())
        return ids, probs[ids], v.numpy().reshape(-1)[0]

    # game.py:155-237 (one iteration) + mcts.py:203-224 (self-play branch)
    def play_move(self):
        self.move_count += 1
        temp = 1.0 if self.move_count <= 30 else max(0.1, 1.0 * 0.5)
        acts, probs = self.search.get_move_probs(self.board, temp)
        move_probs = np.zeros(2086)
        move_probs[list(acts)] = probs
        move = int(np.random.choice(acts, p=(1 - EPS) * probs + EPS * np.random.dirichlet(ALPHA * np.ones(len(probs)))))
        self.search.update_with_move(move)
        move_probs = move_probs / np.sum(move_probs)
        self.samples.append((move_probs, self.board.turn))
        self.board.push(mcts_oracle.move_from_id(move))
        tie = (self.board.is_insufficient_material() or self.board.is_fourfold_repetition()
               or self.board.is_sixty_moves())
        if self.board.is_game_over() or tie:
            self.new_game()
            return move, True
        return move, False


def random_state_dict(num_channels=256, resblocks_num=40):
    """Random-init weights with the reference's state_dict keys and default PyTorch init (what
    collect.py:51-56 falls back to when no model file exists), built from plain torch layers."""
    from torch import nn

    sd = {}

    def add(prefix, mod):
        for k, v in mod.state_dict().items():
            sd[f"{prefix}.{k}"] = v

    add("conv_block", nn.Conv2d(119, num_channels, 3, padding=1))
    add("conv_block_bn", nn.BatchNorm2d(num_channels))
    for i in range(resblocks_num):
        add(f"res_blocks.{i}.conv1", nn.Conv2d(num_channels, num_channels, 3, padding=1))
        add(f"res_blocks.{i}.conv1_bn", nn.BatchNorm2d(num_channels))
        add(f"res_blocks.{i}.conv2", nn.Conv2d(num_channels, num_channels, 3, padding=1))
        add(f"res_blocks.{i}.conv2_bn", nn.BatchNorm2d(num_channels))
    add("policy_conv", nn.Conv2d(num_channels, 17, 1))
    add("policy_bn", nn.BatchNorm2d(17))
    add("policy_fc", nn.Linear(17 * 90, 2086))
    add("value_conv", nn.Conv2d(num_channels, 7, 1))
    add("value_bn", nn.BatchNorm2d(7))
    add("value_fc1", nn.Linear(7 * 90, 256))
    add("value_fc2", nn.Linear(256, 1))
    return sd


def time_moves(n_moves: int, n_playout: int = 400, warmup_playouts: int = 2, **kw):
    """(moves_per_sec, seconds, evals) for the first n_moves of one self-play game."""
    sp = PortedSelfPlay(n_playout=n_playout, **kw)
    for _ in range(warmup_playouts):  # page the weights in, untimed
        sp.policy_value_fn(sp.board)
    sp.n_evals = 0
    t0 = time.perf_counter()
    for _ in range(n_moves):
        sp.play_move()
    dt = time.perf_counter() - t0
    return n_moves / dt, dt, sp.n_evals
