"""Replay container round trip (collect.py:146-167 layout) through the self-written minimal HDF5
writer/reader: group per game, three datasets, gzip, root attribute ``iters``, append mode."""
import struct

import numpy as np
import pytest

from chinesechesszero_b200 import h5lite


def _game(rng, t):
    states = (rng.random((2 * t, 17, 7, 10, 9)) > 0.97).astype(np.float16)
    probs = rng.random((2 * t, 2086))
    probs /= probs.sum(1, keepdims=True)
    winners = rng.choice([-1.0, 0.0, 1.0], size=2 * t)
    return states, probs, winners


def test_roundtrip_append_and_iters(tmp_path):
    path = str(tmp_path / "data" / "data.h5")
    rng = np.random.default_rng(0)
    games = [_game(rng, t) for t in (3, 5, 2)]
    w = h5lite.H5ReplayWriter(path)
    for g in games:
        w.add(*g)
    w.close()
    # the reference reopens the file in "a" mode for every game (collect.py:146)
    more = [_game(rng, 1 + i % 3) for i in range(40)]
    w = h5lite.H5ReplayWriter(path, flush_every=7)
    assert w.iters == 3
    for g in more:
        w.add(*g)
    w.close()
    with h5lite.H5Reader(path) as r:
        assert int(r.root_attrs()["iters"]) == 43
        links = r.root_links()
        assert sorted(links) == sorted(f"game_{i}" for i in range(43))
        for i, g in enumerate(games + more):   # convert.py:66-81 access pattern
            d = r.read_group(f"game_{i}")
            assert d["states"].dtype == np.float16 and d["mcts_probs"].dtype == np.float64
            assert d["winners"].dtype == np.float64
            assert np.array_equal(d["states"], g[0]) and np.array_equal(d["mcts_probs"], g[1])
            assert np.array_equal(d["winners"], g[2])


def test_structure_follows_the_format_spec(tmp_path):
    path = str(tmp_path / "s.h5")
    rng = np.random.default_rng(1)
    with h5lite.H5Writer(path, "w") as w:
        for i in range(300):   # > 2*LEAF_K*2*INTERNAL_K = 256 links => a two-level group B-tree
            w.create_group(f"game_{i}", {"winners": rng.random(4)})
        w.attrs["iters"] = np.int64(300)
    buf = open(path, "rb").read()
    assert buf[:8] == b"\x89HDF\r\n\x1a\n" and buf[8] == 0
    eof = struct.unpack_from("<Q", buf, 40)[0]
    assert eof == len(buf)
    with h5lite.H5Reader(path) as r:
        links = r.root_links()
        assert len(links) == 300
        # root symbol-table message -> B-tree root of level 1
        (mtype, data), = [m for m in r.messages(r.root_header) if m[0] == 0x0011]
        btree, heap = struct.unpack_from("<QQ", data, 0)
        assert buf[btree:btree + 4] == b"TREE" and buf[btree + 4] == 0 and buf[btree + 5] == 1
        assert buf[heap:heap + 4] == b"HEAP"
        # every SNOD is sorted by name and the names are globally sorted left to right
        names = []

        def walk(addr):
            level, used = buf[addr + 5], struct.unpack_from("<H", buf, addr + 6)[0]
            for i in range(used):
                child = struct.unpack_from("<Q", buf, addr + 24 + 8 + 16 * i)[0]
                if level:
                    walk(child)
                else:
                    n = struct.unpack_from("<H", buf, child + 6)[0]
                    assert 1 <= n <= 8
                    for k in range(n):
                        off = struct.unpack_from("<Q", buf, child + 8 + 40 * k)[0]
                        names.append(r._heap_string(heap, off))

        walk(btree)
        assert names == sorted(names) and len(names) == 300
        assert int(r.root_attrs()["iters"]) == 300


def test_uncompressed_and_empty(tmp_path):
    path = str(tmp_path / "u.h5")
    w = h5lite.H5ReplayWriter(path, gzip_level=None)
    s, p, z = _game(np.random.default_rng(2), 2)
    w.add(s, p, z)
    w.close()
    with h5lite.H5Reader(path) as r:
        d = r.read_group("game_0")
        assert np.array_equal(d["states"], s) and np.array_equal(d["winners"], z)


def test_merge_rank_shards(tmp_path):
    from chinesechesszero_b200 import distributed
    from chinesechesszero_b200.collect import merge_h5_shards

    rng = np.random.default_rng(5)
    world, per_rank = 2, 3
    truth = {}
    paths = []
    for rank in range(world):
        p = str(tmp_path / f"rank{rank}" / "data.h5")
        w = h5lite.H5ReplayWriter(p)
        for i in range(per_rank):
            g = _game(rng, 2)
            k = distributed.global_game_index(i, rank, world)
            truth[k] = g
            w.add(*g, index=k)
        w.close()
        paths.append(p)
    merged = str(tmp_path / "data.h5")
    assert merge_h5_shards(paths, merged) == world * per_rank
    with h5lite.H5Reader(merged) as r:
        assert int(r.root_attrs()["iters"]) == 6
        for k in range(6):
            d = r.read_group(f"game_{k}")
            assert np.array_equal(d["states"], truth[k][0]) and np.array_equal(d["mcts_probs"], truth[k][1])


def test_convert_h5_to_npy_matches_the_reference_layout(tmp_path):
    """convert.py:21-105: data.h5 -> states.npy / mcts.npy / winners.npy / meta.json (winners float32), games in
    index order, a missing game index skipped."""
    import json

    from chinesechesszero_b200 import h5lite
    from chinesechesszero_b200.convert import convert_h5_to_npy

    rng = np.random.default_rng(0)
    games = []
    path = str(tmp_path / "data.h5")
    w = h5lite.H5ReplayWriter(path, gzip_level=4)
    for t in (3, 1, 5):
        st = (rng.random((2 * t, 17, 7, 10, 9)) > 0.9).astype(np.float16)
        pi = rng.random((2 * t, 2086))
        z = rng.choice([-1.0, 0.0, 1.0], size=2 * t)
        w.add(st, pi, z)
        games.append((st, pi, z))
    w.close()
    out = str(tmp_path / "npy")
    n = convert_h5_to_npy(path, out)
    assert n == 18
    states, mcts, winners = (np.load(f"{out}/{k}.npy") for k in ("states", "mcts", "winners"))
    assert states.dtype == np.float16 and mcts.dtype == np.float64 and winners.dtype == np.float32
    assert np.array_equal(states, np.concatenate([g[0] for g in games]))
    assert np.array_equal(mcts, np.concatenate([g[1] for g in games]))
    assert np.array_equal(winners, np.concatenate([g[2] for g in games]).astype(np.float32))
    meta = json.load(open(f"{out}/meta.json"))
    assert meta["total_count"] == 18 and meta["states_shape"] == [18, 17, 7, 10, 9] and meta["winners_dtype"] == "float32"
    # what train.py:95-100 does with the result
    from chinesechesszero_b200.train import NpyReplayDataset

    assert len(NpyReplayDataset(out)) == 18
