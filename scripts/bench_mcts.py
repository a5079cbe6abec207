"""Timing of the MCTS arena kernels (K3 select, K1 on leaves, K4+K5 expand/backup, K6, K7) at the
lockstep batch size, with a device-resident stand-in policy (fixed random logits) so that nothing
but our kernels runs.  Used plain for CUDA-event timings and under ncu for the counters."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from chinesechesszero_b200 import _lib
from chinesechesszero_b200.search import LockstepSearch

G = int(os.environ.get("CCZ_G", "4096"))
P = int(os.environ.get("CCZ_P", "400"))
torch.manual_seed(0)
logits = torch.randn(G, 2086, device="cuda")
values = torch.tanh(torch.randn(G, device="cuda") * 0.3)
s = LockstepSearch(G, node_cap=32768)


def ev(planes, boards):
    return logits, _lib.POLICY_LOGITS, values


def timed(fn, n=1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {"games": G, "playouts": P}
for move in range(2):
    t_sel = t_mg = t_exp = 0.0
    for step in range(P):
        t_sel += timed(lambda: _lib.mcts_select(s.arena, s.c_puct, s.leaf_boards, s.leaf_nodes))
        t_mg += timed(lambda: _lib.movegen_encode(s.leaf_boards, out=(s.move_ids, s.counts, s.flags, s.planes)))
        t_exp += timed(lambda: s.expand_backup(logits, _lib.POLICY_LOGITS, values))
    s.check_status()
    acts, visits, counts = s.root_visits()
    chosen = acts.gather(1, visits.argmax(1, keepdim=True).long()).view(-1).contiguous()
    t_adv = timed(lambda: s.advance(chosen))
    nodes = s.arena.n_nodes.float().mean().item()
    out[f"move{move}"] = {"select_us": t_sel / P * 1e3, "movegen_encode_us": t_mg / P * 1e3,
                          "expand_backup_us": t_exp / P * 1e3, "advance_us": t_adv * 1e3,
                          "kept_nodes_per_game_after_advance": nodes}
print(json.dumps(out))
