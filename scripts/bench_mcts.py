"""Timing of the MCTS pool kernels (K3 select, K1 on leaves, K4+K5 expand/backup, K6, K7a/K7b advance)
at the lockstep batch size with a device-resident stand-in net (position-dependent logits / values, no
forward pass), so that nothing but our kernels runs.  Plain: CUDA-event timings per kernel and the
algorithmic bytes K3 moves (sum over levels of 16 B per child + 8 B link + board / keys).
Under ncu (`CCZ_NCU=1` shortens the run): the counters for profiles/."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from chinesechesszero_b200 import _lib
from chinesechesszero_b200.search import LockstepSearch
from tests.test_pool_gpu import board_hash_evaluator

G = int(os.environ.get("CCZ_G", "4096"))
P = int(os.environ.get("CCZ_P", "400"))
MOVES = int(os.environ.get("CCZ_MOVES", "3"))
NCU = os.environ.get("CCZ_NCU") == "1"
s = LockstepSearch(G, nodes_per_game=int(os.environ.get("CCZ_NODES", str(8 * 41 * P))))
ev = board_hash_evaluator(G, seed=1, scale=0.3)


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    return e0, e1


out = {"games": G, "playouts": P}
for move in range(MOVES):
    s.ensure_capacity(P, s.arena.pool_ctl.cpu().numpy())
    _lib.mcts_reserve(s.arena, s.search_pages(P))
    ev_sel, ev_mg, ev_exp = [], [], []
    depth_sum = 0.0
    steps = P if not NCU else min(P, 40)
    for step in range(steps):
        ev_sel.append(timed(lambda: _lib.mcts_select(s.arena, s.c_puct, s.leaf_boards, s.leaf_nodes)))
        ev_mg.append(timed(lambda: _lib.movegen_encode(s.leaf_boards, planes=False,
                                                       out=(s.move_ids, s.counts, s.flags, None))))
        pol, kind, val = ev(None, s.leaf_boards)
        ev_exp.append(timed(lambda: s.expand_backup(pol, kind, val)))
    torch.cuda.synchronize()
    last = slice(max(0, steps - 100), steps)  # the deep end of the search
    t_sel = [a.elapsed_time(b) for a, b in ev_sel]
    t_mg = [a.elapsed_time(b) for a, b in ev_mg]
    t_exp = [a.elapsed_time(b) for a, b in ev_exp]
    acts, visits, counts = s.root_visits()
    chosen = acts.gather(1, visits.argmax(1, keepdim=True).long()).view(-1).contiguous()
    nodes_before = s.arena.n_nodes.float().mean().item()
    a, b = timed(lambda: s.advance(chosen))
    torch.cuda.synchronize()
    out[f"move{move}"] = {
        "select_us_mean": sum(t_sel) / len(t_sel) * 1e3, "select_us_last100": sum(t_sel[last]) / len(t_sel[last]) * 1e3,
        "movegen_us": sum(t_mg) / len(t_mg) * 1e3,
        "expand_backup_us_mean": sum(t_exp) / len(t_exp) * 1e3, "expand_backup_us_last100": sum(t_exp[last]) / len(t_exp[last]) * 1e3,
        "advance_us": a.elapsed_time(b) * 1e3, "nodes_per_game_before_advance": nodes_before,
        "kept_nodes_per_game": s.arena.n_nodes.float().mean().item(),
    }
out["pool"] = s.pool_stats()
print(json.dumps(out))
