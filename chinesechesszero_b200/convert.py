"""``convert.py`` of the reference (convert.py:21-105): ``data.h5`` -> ``states.npy`` / ``mcts.npy`` /
``winners.npy`` / ``meta.json``, the form ``train.py:95-100`` memory-maps.  Same function name, defaults,
output names, dtypes (winners become float32, convert.py:67) and meta keys; the file is read with the
in-tree HDF5 reader (``h5lite``, memory-mapped) because h5py / libhdf5 do not exist in this image; the arrays
are written through ``numpy.lib.format.open_memmap``, so one game at a time is resident on either side."""
from __future__ import annotations

import json
import os

import numpy as np

from . import h5lite
from .parameters import DATA_DIR
from .tools import log


def _derive_out_paths(out_dir: str):
    root = out_dir if out_dir else "."
    return tuple(os.path.join(root, n) for n in ("states.npy", "mcts.npy", "winners.npy", "meta.json"))


def convert_h5_to_npy(h5_path: str | None = None, out_dir: str | None = None) -> int:
    """Returns the number of samples written (the reference returns None and logs it)."""
    if h5_path is None:
        h5_path = os.path.join(DATA_DIR, "data.h5")
    if out_dir is None:
        out_dir = DATA_DIR
    os.makedirs(out_dir, exist_ok=True)
    out_states, out_mcts, out_winners, out_meta = _derive_out_paths(out_dir)
    log(f"Start converting {h5_path}")
    with h5lite.H5Reader(h5_path) as r:
        games_count = int(r.root_attrs().get("iters", 0))
        links = r.root_links()
        # first pass (convert.py:44-49): sample count, from the data-space messages only
        steps, sample_state_shape, sample_mcts_shape = [], None, None
        for k in range(games_count):
            addr = links.get(f"game_{k}")
            if addr is None:
                steps.append(0)
                continue
            sub = r.links(addr)
            if "states" not in sub:
                steps.append(0)
                continue
            shape = r.dataset_shape(sub["states"])
            steps.append(int(shape[0]))
            if sample_state_shape is None:
                sample_state_shape = tuple(shape[1:])
                sample_mcts_shape = tuple(r.dataset_shape(sub["mcts_probs"])[1:])
        total = int(sum(steps))
        log(f"Total games: {games_count}")
        log(f"Total steps: {total}")
        if sample_state_shape is None:
            sample_state_shape, sample_mcts_shape = (17, 7, 10, 9), (2086,)
        states = np.lib.format.open_memmap(out_states, mode="w+", dtype=np.float16, shape=(total,) + sample_state_shape)
        mcts = np.lib.format.open_memmap(out_mcts, mode="w+", dtype=np.float64, shape=(total,) + sample_mcts_shape)
        winners = np.lib.format.open_memmap(out_winners, mode="w+", dtype=np.float32, shape=(total,))
        cur = 0
        for k in range(games_count):  # second pass (convert.py:66-82)
            if not steps[k]:
                continue
            g = r.read_group(f"game_{k}")
            n = steps[k]
            states[cur:cur + n] = g["states"]
            mcts[cur:cur + n] = g["mcts_probs"]
            winners[cur:cur + n] = g["winners"]
            cur += n
        meta = {
            "total_count": total,
            "states_shape": list(states.shape), "states_dtype": str(states.dtype),
            "mcts_shape": list(mcts.shape), "mcts_dtype": str(mcts.dtype),
            "winners_shape": list(winners.shape), "winners_dtype": str(winners.dtype),
        }
        for a in (states, mcts, winners):
            a.flush()
    with open(out_meta, "w", encoding="utf-8") as f:
        json.dump(meta, f, ensure_ascii=False, indent=2)
    log(f"Saved {total} samples to {out_dir}")
    return total


if __name__ == "__main__":
    convert_h5_to_npy()
