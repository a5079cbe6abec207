"""Host side of the replay output (collect.py:64-169, convert.py:85-99): the streaming npy writer, the
h5 index flush policy, the asynchronous writer thread, and the shard game numbering across a restart."""
import json
import os
import time

import numpy as np
import pytest

from chinesechesszero_b200 import distributed, h5lite, replay


def _game(rng, t):
    states = (rng.random((2 * t, 17, 7, 10, 9)) > 0.97).astype(np.float16)
    probs = rng.random((2 * t, 2086))
    probs /= probs.sum(1, keepdims=True)
    winners = rng.choice([-1.0, 0.0, 1.0], size=2 * t)
    return states, probs, winners


def test_npy_writer_streams_and_appends(tmp_path):
    """Rows go to disk as they arrive (no list of games in RAM), np.load sees a valid file after every
    flush, and reopening the directory appends (collect -> train -> collect ... in loop.py)."""
    rng = np.random.default_rng(0)
    d = str(tmp_path / "npy")
    w = replay.NpyReplayWriter(d)
    assert np.load(os.path.join(d, "states.npy")).shape == (0, 17, 7, 10, 9)
    games = [_game(rng, t) for t in (3, 1, 4)]
    for g in games[:2]:
        w.add(*g)
    assert not hasattr(w, "_states")  # nothing accumulates on the host
    assert w.flush() == 8
    mm = np.load(os.path.join(d, "states.npy"), mmap_mode="r")  # the trainer's memory map stays valid ...
    assert mm.shape == (8, 17, 7, 10, 9)
    w.add(*games[2])                                             # ... while the writer appends behind it
    w.close()
    assert np.array_equal(np.asarray(mm), np.concatenate([g[0] for g in games[:2]]))
    w = replay.NpyReplayWriter(d)
    assert w.rows == 16
    extra = _game(rng, 2)
    w.add(*extra)
    w.close()
    games.append(extra)
    states, mcts, winners = (np.load(os.path.join(d, f"{k}.npy")) for k in ("states", "mcts", "winners"))
    assert states.dtype == np.float16 and mcts.dtype == np.float64 and winners.dtype == np.float32
    assert np.array_equal(states, np.concatenate([g[0] for g in games]))
    assert np.array_equal(mcts, np.concatenate([g[1] for g in games]))
    assert np.array_equal(winners, np.concatenate([g[2] for g in games]).astype(np.float32))
    meta = json.load(open(os.path.join(d, "meta.json")))
    assert meta["total_count"] == 20 and meta["states_shape"] == [20, 17, 7, 10, 9] and meta["winners_dtype"] == "float32"
    from chinesechesszero_b200.train import NpyReplayDataset

    assert len(NpyReplayDataset(d)) == 20
    with pytest.raises(ValueError):
        replay.NpyReplayWriter(d).add(extra[0][:, :16], extra[1], extra[2])


def test_npy_writer_does_not_append_to_foreign_files(tmp_path):
    d = str(tmp_path / "npy")
    os.makedirs(d)
    np.save(os.path.join(d, "states.npy"), np.zeros((2, 17, 7, 10, 9), np.float16))  # np.save header: not appendable
    np.save(os.path.join(d, "mcts.npy"), np.zeros((2, 2086)))
    np.save(os.path.join(d, "winners.npy"), np.zeros((2,), np.float32))
    w = replay.NpyReplayWriter(d)
    assert w.rows == 0
    w.close()


def test_h5_index_flush_is_geometric(tmp_path):
    """flushing the root index after every game costs O(games^2) bytes.  The default policy
    flushes after max(64, games/16) pending games (or flush_seconds) and on close."""
    rng = np.random.default_rng(2)
    g = _game(rng, 1)
    sizes = {}
    for name, kw in (("every", dict(flush_every=1)), ("default", {})):
        path = str(tmp_path / f"{name}.h5")
        w = h5lite.H5ReplayWriter(path, gzip_level=1, flush_seconds=1e9, **kw)
        for _ in range(400):
            w.add(*g)
        flushes_before_close = w.flushes
        w.close()
        sizes[name] = (os.path.getsize(path), flushes_before_close)
        with h5lite.H5Reader(path) as r:
            assert int(r.root_attrs()["iters"]) == 400 and len(r.root_links()) == 400
    assert sizes["every"][1] == 400 and sizes["default"][1] == 400 // 64
    # 400 games: ~65 B x 400^2 / 2 = 5 MB of orphaned index copies with a flush per game, < 0.2 MB by default
    assert sizes["every"][0] - sizes["default"][0] > 4_000_000
    # time-based flush: a slow trickle of games is indexed without waiting for 64 of them
    path = str(tmp_path / "t.h5")
    w = h5lite.H5ReplayWriter(path, gzip_level=1, flush_seconds=0.05)
    w.add(*g)
    time.sleep(0.08)
    w.add(*g)
    assert w.flushes == 1
    with h5lite.H5Reader(path) as r:
        assert len(r.root_links()) == 2
    w.close()


class _Rec:
    def __init__(self, z):
        self.z = z


def _chunk(rng, lens):
    n = sum(lens)
    states = (rng.random((2 * n, 17, 7, 10, 9)) > 0.97).astype(np.float16)
    pi = rng.random((2 * n, 2086))
    spans, off = [], 0
    for t in lens:
        spans.append((_Rec(rng.choice([-1.0, 0.0, 1.0], size=t)), off, t))
        off += t
    released = []

    class Pool:
        def put(self, b):
            released.append(b)

    return replay.PackedChunk(Pool(), object(), states, pi, n, spans), released


def test_async_writer_cuts_chunks_into_games_in_order(tmp_path):
    rng = np.random.default_rng(3)
    h5 = h5lite.H5ReplayWriter(str(tmp_path / "data.h5"), gzip_level=1)
    npy = replay.NpyReplayWriter(str(tmp_path))
    w = replay.AsyncReplayWriter(h5, npy)
    truth = []
    for lens in ((2, 3), (1,), (4, 1, 2)):
        chunk, released = _chunk(rng, lens)
        for k in range(len(lens)):
            truth.append(chunk.game_arrays(k))
        w.submit(chunk)
    w.drain()
    assert w.games == 6 and w.samples == 2 * 13 and released  # staging buffers handed back
    with h5lite.H5Reader(str(tmp_path / "data.h5")) as r:      # drain() leaves a consistent, indexed file
        assert int(r.root_attrs()["iters"]) == 6
        for i, (st, pi, z) in enumerate(truth):
            d = r.read_group(f"game_{i}")
            assert np.array_equal(d["states"], st) and np.array_equal(d["mcts_probs"], pi) and np.array_equal(d["winners"], z)
    assert np.array_equal(np.load(str(tmp_path / "mcts.npy")), np.concatenate([t[1] for t in truth]))
    # the per-game layout: T samples then their T mirrored rows (collect.py:131)
    st0 = truth[0][0]
    assert st0.shape[0] == 4
    w.close()
    h5.close()
    npy.close()


def test_async_writer_reports_errors_on_the_submitting_thread(tmp_path):
    class Broken:
        def add(self, *a, **k):
            raise OSError("disk full")

    rng = np.random.default_rng(4)
    w = replay.AsyncReplayWriter(Broken(), None)
    chunk, released = _chunk(rng, (1,))
    w.submit(chunk)
    with pytest.raises(OSError):
        w.drain()
    assert released  # the buffer is released even when the write failed
    w.close()


def test_shard_numbering_survives_a_restart(tmp_path):
    """a restarted multi-GPU collection must continue its shard's numbering (rank r owns
    r, r + world, ...) instead of starting again at game_{rank} and colliding with an existing group."""
    from chinesechesszero_b200.collect import CollectPipeline

    rng = np.random.default_rng(6)
    world = 2
    for rank in range(world):
        for session in range(2):
            pipe = CollectPipeline(n_games=2, data_dir=str(tmp_path), rank=rank, world=world, write_npy=False,
                                   async_writer=False)
            assert pipe.local_games == 3 * session
            for i in range(3):  # what collect_data does with a finished game, without a GPU
                k = distributed.global_game_index(pipe.local_games, rank, world)
                pipe.h5.add(*_game(rng, 1), index=k)
                pipe.local_games += 1
            pipe.close()
    for rank in range(world):
        with h5lite.H5Reader(str(tmp_path / f"rank{rank}" / "data.h5")) as r:
            assert sorted(int(n.split("_")[1]) for n in r.root_links()) == [rank + world * i for i in range(6)]


def test_writers_keep_host_memory_flat(tmp_path):
    """the round-1 npy writer held every game in RAM until flush (70 GB at configs[2] sizes).  Both
    writers stream: after 3,000 games the Python heap has grown by the h5 link table only (a few hundred KB), not
    by the 100+ MB of rows that went to disk."""
    import tracemalloc

    rng = np.random.default_rng(7)
    g = _game(rng, 1)                       # 2 rows = 76 KB raw per game
    h5 = h5lite.H5ReplayWriter(str(tmp_path / "data.h5"), gzip_level=1)
    npy = replay.NpyReplayWriter(str(tmp_path))
    for _ in range(200):                    # warm the allocator
        h5.add(*g)
        npy.add(*g)
    tracemalloc.start()
    base = tracemalloc.get_traced_memory()[0]
    for _ in range(3000):
        h5.add(*g)
        npy.add(*g)
    grown = tracemalloc.get_traced_memory()[0] - base
    peak = tracemalloc.get_traced_memory()[1] - base
    tracemalloc.stop()
    written = 3000 * sum(a.nbytes for a in (g[0], g[1], g[2].astype(np.float32)))
    assert written > 200e6
    assert grown < 2e6 and peak < 8e6, (grown, peak)
    h5.close()
    npy.close()
    assert np.load(str(tmp_path / "winners.npy"), mmap_mode="r").shape == (6400,)


def test_packer_chunking_keeps_games_whole():
    """ReplayPacker._chunks (host logic, no GPU): runs of WHOLE games with at most max_samples samples, order kept,
    a game longer than max_samples alone in its chunk."""
    from types import SimpleNamespace

    class Rec:
        def __init__(self, t):
            self.t = t

        def __len__(self):
            return self.t

    lens = [3, 5, 2, 9, 1, 1, 1, 40, 4, 4, 8]
    recs = [Rec(t) for t in lens]
    runs = list(replay.ReplayPacker._chunks(SimpleNamespace(max_samples=10), recs))
    assert [r for run in runs for r in run] == recs                      # nothing lost, order kept
    assert all(sum(len(r) for r in run) <= 10 or len(run) == 1 for run in runs)
    assert [[len(r) for r in run] for run in runs] == [[3, 5, 2], [9, 1], [1, 1], [40], [4, 4], [8]]
    assert list(replay.ReplayPacker._chunks(SimpleNamespace(max_samples=10), [])) == []
