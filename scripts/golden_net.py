"""Golden vectors for the network forward from the UNMODIFIED reference net.Net
(/root/reference/net.py) on CPU fp32:  python scripts/make_golden.py net

Two weight sets, both reproducible without shipping 200 MB of weights:
  "seed0"     torch.manual_seed(0); Net()          (default init, what collect.py:51-56 falls back to)
  "perturbed" the same followed by oracle.net_oracle.perturb_(state_dict, seed=1)
Inputs: the search-time planes of 6 fixed positions.  Outputs: log-probs (6,2086) and values (6,1).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import load_reference, net_oracle
from tests import positions


def golden_positions() -> np.ndarray:
    recs = positions.random_playout_positions(2, 80, seed=42, every=17)
    return np.concatenate([positions.edge_case_records()[:2], recs[:4]])


def main(golden_dir):
    ref_net = load_reference.load("net")
    recs = golden_positions()
    x = net_oracle.search_planes(recs)
    out = {"records": recs}
    torch.set_num_threads(8)
    for name in ("seed0", "perturbed"):
        torch.manual_seed(0)
        net = ref_net.Net()
        if name == "perturbed":
            net_oracle.perturb_(net.state_dict(), seed=1)
        net.eval()
        with torch.no_grad():
            logp, v = net(x)
        out[f"{name}_logp"] = logp.numpy()
        out[f"{name}_value"] = v.numpy()
        print(name, "logp range", float(logp.min()), float(logp.max()), "values", v.flatten().tolist())
    path = os.path.join(golden_dir, "net_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
