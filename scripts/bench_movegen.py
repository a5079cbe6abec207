"""Quick K1 timing: ccz_movegen_encode on ~1M perft positions (config 2 of BASELINE.json)."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from chinesechesszero_b200 import _lib
from oracle import cchess_shim as cs


def build_positions(n_target=1 << 20, seed=0):
    l3 = cs.collect_leaves(cs.start_record(), 3, 79666)
    l4 = cs.collect_leaves(cs.start_record(), 4, 3290240)
    rng = np.random.default_rng(seed)
    pick = np.sort(rng.choice(l4.shape[0], size=n_target - l3.shape[0], replace=False))
    return np.concatenate([l3, l4[pick]])


def main():
    recs = build_positions()
    n = recs.shape[0]
    boards = torch.from_numpy(recs).cuda()
    ids = torch.empty((n, 128), dtype=torch.int16, device="cuda")
    counts = torch.empty((n,), dtype=torch.int16, device="cuda")
    flags = torch.empty((n,), dtype=torch.uint8, device="cuda")
    planes = torch.empty((n, 17, 7, 10, 9), dtype=torch.bfloat16, device="cuda")
    out = (ids, counts, flags, planes)
    for variant, o in (("planes", out), ("noplanes", (ids, counts, flags, None))):
        for _ in range(3):
            _lib.movegen_encode(boards, planes=o[3] is not None, out=o)
        torch.cuda.synchronize()
        times = []
        for _ in range(10):
            if o[3] is not None:
                planes.view(torch.int16).fill_(-1)
            ids.fill_(-1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.movegen_encode(boards, planes=o[3] is not None, out=o)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        nl = float(counts.float().mean())
        per_pos = (21520 + 2 * nl) if o[3] is not None else (100 + 2 * nl)
        best, med = min(times), sorted(times)[len(times) // 2]
        print(json.dumps({"variant": variant, "n": n, "best_ms": best, "median_ms": med,
                          "mpos_per_s": n / best / 1e3, "GBps_best": n * per_pos / best / 1e6,
                          "frac_of_6454": n * per_pos / best / 1e6 / 6454.0, "mean_legal": nl}))


if __name__ == "__main__":
    main()
