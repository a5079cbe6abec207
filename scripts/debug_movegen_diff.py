import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chinesechesszero_b200 import _lib, tools
from oracle import cchess_shim as cs
from tests import positions
recs = positions.perft_leaves(3)
ids, counts, flags, _ = _lib.movegen_encode(torch.from_numpy(recs).cuda(), planes=False)
ids, counts, flags = ids.cpu().numpy(), counts.cpu().numpy(), flags.cpu().numpy()
o_ids, o_counts, o_flags, _ = cs.batch_movegen_encode(recs, want_planes=False)
bad = np.nonzero((ids != o_ids).any(axis=1) | (counts != o_counts) | (flags != o_flags))[0]
print("n bad", len(bad))
for i in bad[:6]:
    b = cs.Board.from_record(recs[i])
    dev = [tools.move_id2move_action[int(x)] for x in ids[i, :counts[i]]]
    ora = [tools.move_id2move_action[int(x)] for x in o_ids[i, :o_counts[i]]]
    print(i, b.fen(), "flags", flags[i], o_flags[i])
    print("  only device:", [m for m in dev if m not in ora], " only oracle:", [m for m in ora if m not in dev])
for i in bad[:2]:
    dev = [tools.move_id2move_action[int(x)] if x >= 0 else "??" for x in ids[i, :counts[i]]]
    ora = [tools.move_id2move_action[int(x)] for x in o_ids[i, :o_counts[i]]]
    print("dev", counts[i], dev)
    print("ora", o_counts[i], ora)
