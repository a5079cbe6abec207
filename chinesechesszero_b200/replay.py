"""Replay output of the self-play path.

Turns finished games (``selfplay.GameRecord``) into the arrays the reference stores per game
(collect.py:64-169): ``states (2T,17,7,10,9) float16``, ``mcts_probs (2T,2086) float64``,
``winners (2T,) float64`` -- first T rows the game's samples, next T their file-mirrored twins
(collect.py:131) -- with the densification done on the device by K8 (``ccz_replay_pack``).

``states_mode``
  "reference": byte-for-byte what the reference saves: every sample of a game
               carries the FINAL 8-deep history (game.py:234-237 aliasing) and the turn plane is
               all ones (collect.py:28,78-81);
  "per_move":  sample i carries the history as of move i (slot 0 = the position searched, most
               recent first, padded with the initial position) and the true side-to-move plane.

``NpyReplayWriter`` appends to the ``states.npy / mcts.npy / winners.npy / meta.json`` layout that
the reference's trainer actually reads (convert.py:85-99, train.py:95-100).
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import _lib


def history_boards(rec, states_mode: str = "per_move"):
    """(hist [T,8,96] uint8, turn_plane [T] uint8) for a GameRecord."""
    boards = rec.boards
    t = boards.shape[0]
    idx = np.arange(t)[:, None] - np.arange(8)[None, :]          # slot k = position searched at move i-k
    hist = boards[np.clip(idx, 0, None)]                          # before move 0: the initial position
    if states_mode == "reference":
        # history after the last update_states_history(): slot 0 = position before the last move
        hist = np.broadcast_to(hist[t - 1], (t, 8, _lib.BOARD_BYTES)).copy()
        turn = np.ones(t, dtype=np.uint8)
    elif states_mode == "per_move":
        turn = rec.turns.astype(np.uint8)
    else:
        raise ValueError(f"unknown states_mode {states_mode!r}")
    return np.ascontiguousarray(hist), turn


def pack_game(rec, states_mode: str = "per_move", device="cuda"):
    """GameRecord -> (states f16 (2T,17,7,10,9), mcts_probs f64 (2T,2086), winners f64 (2T,)) NumPy,
    computed by ccz_replay_pack on the device."""
    t = len(rec)
    hist, turn = history_boards(rec, states_mode)
    acts = np.full((t, _lib.MAX_MOVES), -1, dtype=np.int16)
    probs = np.zeros((t, _lib.MAX_MOVES), dtype=np.float64)
    counts = np.zeros(t, dtype=np.int16)
    for i, (a, p) in enumerate(zip(rec.acts, rec.probs)):
        counts[i] = len(a)
        acts[i, : len(a)] = a
        probs[i, : len(a)] = p
    dev = torch.device(device)
    states, pi = _lib.replay_pack(torch.from_numpy(hist).to(dev), torch.from_numpy(turn).to(dev),
                                  torch.from_numpy(acts).to(dev), torch.from_numpy(probs).to(dev),
                                  torch.from_numpy(counts).to(dev))
    winners = np.concatenate([rec.z, rec.z]).astype(np.float64)
    return states.cpu().numpy(), pi.cpu().numpy(), winners


class ReplayPacker:
    """K8 over ALL the games that finish with one lockstep move: one upload of the sparse samples, one
    launch, one read-back into pinned memory per chunk of ``max_samples`` samples (collect.py:141-142 does
    preprocess + flip per game on the host).  ``pack`` yields ``PackedChunk``s; a chunk's arrays live in
    one of ``n_buffers`` pinned staging buffers and stay valid until ``chunk.release()`` -- the consumer
    (``AsyncReplayWriter``) calls it, so a slow writer throttles the producer instead of growing memory."""

    def __init__(self, device="cuda", states_mode: str = "per_move", max_samples: int = 4096, n_buffers: int = 3):
        import queue

        self.device = torch.device(device)
        self.states_mode = states_mode
        self.max_samples = int(max_samples)
        self._free = queue.Queue()
        m = self.max_samples
        for _ in range(n_buffers):
            self._free.put(dict(
                states=torch.empty((2 * m, 17, 7, 10, 9), dtype=torch.float16, pin_memory=True),
                pi=torch.empty((2 * m, _lib.N_ACTIONS), dtype=torch.float64, pin_memory=True)))
        self._in = dict(hist=torch.empty((m, 8, _lib.BOARD_BYTES), dtype=torch.uint8, pin_memory=True),
                        turn=torch.empty((m,), dtype=torch.uint8, pin_memory=True),
                        acts=torch.empty((m, _lib.MAX_MOVES), dtype=torch.int16, pin_memory=True),
                        probs=torch.empty((m, _lib.MAX_MOVES), dtype=torch.float64, pin_memory=True),
                        counts=torch.empty((m,), dtype=torch.int16, pin_memory=True))
        self._d_states = torch.empty((2 * m, 17, 7, 10, 9), dtype=torch.float16, device=self.device)
        self._d_pi = torch.empty((2 * m, _lib.N_ACTIONS), dtype=torch.float64, device=self.device)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.launches = 0

    def _chunks(self, records):
        """Split the records into runs of whole games with at most max_samples samples (a longer game
        gets a chunk of its own, packed piecewise by ``pack_game``)."""
        run, n = [], 0
        for rec in records:
            t = len(rec)
            if run and n + t > self.max_samples:
                yield run
                run, n = [], 0
            run.append(rec)
            n += t
        if run:
            yield run

    def pack(self, records):
        lib = _lib.load()
        for run in self._chunks(records):
            n = sum(len(r) for r in run)
            if n > self.max_samples:  # one very long game: the unbatched path
                rec = run[0]
                states, pi, winners = pack_game(rec, self.states_mode, device=self.device)
                t = len(rec)
                yield PackedChunk(None, None, states, pi, t, [(rec, 0, t)])
                continue
            h = self._in
            hist, turn = h["hist"].numpy(), h["turn"].numpy()
            acts, probs, counts = h["acts"].numpy(), h["probs"].numpy(), h["counts"].numpy()
            acts[:n] = -1
            probs[:n] = 0.0
            spans, off = [], 0
            for rec in run:
                t = len(rec)
                hist[off:off + t], turn[off:off + t] = history_boards(rec, self.states_mode)
                for i, (a, p) in enumerate(zip(rec.acts, rec.probs)):
                    counts[off + i] = len(a)
                    acts[off + i, : len(a)] = a
                    probs[off + i, : len(a)] = p
                spans.append((rec, off, t))
                off += t
            dev = {k: v[:n].to(self.device, non_blocking=True) for k, v in h.items()}
            self.h2d_bytes += sum(v[:n].numel() * v.element_size() for v in h.values())
            d_states, d_pi = self._d_states[: 2 * n], self._d_pi[: 2 * n]
            with torch.cuda.device(self.device):
                _lib.check(lib.ccz_replay_pack(dev["hist"].data_ptr(), dev["turn"].data_ptr(), dev["acts"].data_ptr(),
                                               dev["probs"].data_ptr(), dev["counts"].data_ptr(), n, d_states.data_ptr(),
                                               d_pi.data_ptr(), _lib.stream_ptr(self.device)), "ccz_replay_pack")
            self.launches += 2
            buf = self._free.get()  # blocks while every staging buffer is still being written out
            buf["states"][: 2 * n].copy_(d_states, non_blocking=True)
            buf["pi"][: 2 * n].copy_(d_pi, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            self.d2h_bytes += 2 * n * (_lib.PLANE_ELEMS * 2 + _lib.N_ACTIONS * 8)
            yield PackedChunk(self._free, buf, buf["states"].numpy(), buf["pi"].numpy(), n, spans)


class PackedChunk:
    """Rows [0,n) of ``states`` / ``pi`` are the samples of the chunk's games back to back, rows [n,2n)
    their mirrored twins; ``spans`` = (record, first row, T) per game."""

    def __init__(self, pool, buf, states, pi, n, spans):
        self._pool, self._buf = pool, buf
        self.states, self.pi, self.n, self.spans = states, pi, n, spans

    def game_arrays(self, k: int):
        """(states (2T,17,7,10,9) f16, mcts_probs (2T,2086) f64, winners (2T,) f64) of game k, in the
        reference's per-game layout: T samples then their T mirrored twins (collect.py:131)."""
        rec, off, t = self.spans[k]
        n = self.n
        states = np.concatenate([self.states[off:off + t], self.states[n + off:n + off + t]])
        pi = np.concatenate([self.pi[off:off + t], self.pi[n + off:n + off + t]])
        return states, pi, np.concatenate([rec.z, rec.z]).astype(np.float64)

    def release(self):
        if self._pool is not None and self._buf is not None:
            self._pool.put(self._buf)
            self._buf = None


class AsyncReplayWriter:
    """Compression and file I/O off the thread that launches kernels.  ``submit(chunk, indices)`` queues a
    packed chunk; the worker cuts it into games and appends each to the sinks (``H5ReplayWriter``,
    ``NpyReplayWriter``) in submission order, then releases the chunk's staging buffer.  zlib and the file
    writes release the GIL.  Errors surface on the next ``submit`` / ``close``."""

    def __init__(self, h5=None, npy=None, max_pending: int = 8):
        import queue
        import threading

        self.h5, self.npy = h5, npy
        self._q = queue.Queue(maxsize=max_pending)
        self._err = None
        self.games = 0
        self.samples = 0
        self.raw_bytes = 0
        self.busy_seconds = 0.0
        self._thread = threading.Thread(target=self._work, name="ccz-replay-writer", daemon=True)
        self._thread.start()

    def _work(self):
        import time

        while True:
            item = self._q.get()
            if item is None:
                break
            chunk, indices = item
            t0 = time.perf_counter()
            try:
                if self._err is None:
                    for k in range(len(chunk.spans)):
                        states, pi, winners = chunk.game_arrays(k)
                        if self.h5 is not None:
                            self.h5.add(states, pi, winners, index=None if indices is None else indices[k])
                        if self.npy is not None:
                            self.npy.add(states, pi, winners)
                        self.games += 1
                        self.samples += states.shape[0]
                        self.raw_bytes += states.nbytes + pi.nbytes + winners.nbytes
            except BaseException as e:  # noqa: BLE001 - re-raised on the submitting thread
                self._err = e
            finally:
                chunk.release()
                self.busy_seconds += time.perf_counter() - t0
                self._q.task_done()

    def _check(self):
        if self._err is not None:
            err, self._err = self._err, None
            raise err

    def submit(self, chunk: PackedChunk, indices=None) -> None:
        self._check()
        self._q.put((chunk, indices))

    def drain(self) -> None:
        """Wait until everything submitted so far is on disk (index flushed)."""
        self._q.join()
        self._check()
        if self.h5 is not None:
            self.h5.flush()
        if self.npy is not None:
            self.npy.flush()

    def close(self) -> None:
        self._q.join()
        self._q.put(None)
        self._thread.join()
        self._check()


class NpyReplayWriter:
    """Streams packed games into the npy triple + meta.json (convert.py:85-99 dtypes: states float16,
    mcts float64, winners float32): rows are appended to the three files as they arrive and the
    fixed-width ``.npy`` headers are rewritten in place by ``flush()``, so host memory stays flat no
    matter how many games an iteration produces.  Opening an existing triple appends to it."""

    HEADER_BYTES = 256
    FILES = (("states.npy", np.float16, (17, 7, 10, 9)), ("mcts.npy", np.float64, (_lib.N_ACTIONS,)),
             ("winners.npy", np.float32, ()))

    def __init__(self, out_dir: str, mode: str = "a"):
        self.out_dir = out_dir
        os.makedirs(out_dir, exist_ok=True)
        self.games = 0
        self.rows = 0
        self._f = []
        rows = []
        for name, dt, tail in self.FILES:
            path = os.path.join(out_dir, name)
            n = self._existing_rows(path, np.dtype(dt), tail) if mode == "a" else None
            if n is None:
                f = open(path, "w+b")
                f.write(b"\0" * self.HEADER_BYTES)
                n = 0
            else:
                f = open(path, "r+b")
                f.seek(0, os.SEEK_END)
            rows.append(n)
            self._f.append(f)
        if len(set(rows)) != 1:
            raise ValueError(f"{out_dir}: states/mcts/winners.npy hold different numbers of rows {rows}")
        self.rows = rows[0]
        self.flush()

    @classmethod
    def _existing_rows(cls, path, dt, tail):
        """Rows of an npy file written by this class (fixed-width header); None if absent / foreign."""
        if not os.path.exists(path) or os.path.getsize(path) < cls.HEADER_BYTES:
            return None
        try:
            arr = np.load(path, mmap_mode="r")
        except Exception:  # noqa: BLE001 - unreadable header: start over
            return None
        if arr.dtype != dt or tuple(arr.shape[1:]) != tuple(tail):
            return None
        with open(path, "rb") as f:
            f.seek(8)
            hlen = int.from_bytes(f.read(2), "little")
        if 10 + hlen != cls.HEADER_BYTES:
            return None  # written by np.save: not appendable in place
        row_bytes = int(np.prod(tail, dtype=np.int64)) * dt.itemsize if tail else dt.itemsize
        return (os.path.getsize(path) - cls.HEADER_BYTES) // row_bytes

    def add(self, states, mcts_probs, winners):
        arrs = (states, mcts_probs, winners)
        n = int(np.asarray(winners).shape[0])
        for f, (name, dt, tail), a in zip(self._f, self.FILES, arrs):
            a = np.ascontiguousarray(a, dtype=dt)
            if a.shape != (n, *tail):
                raise ValueError(f"{name}: expected rows of shape {tail}, got {a.shape}")
            f.write(a.data)
        self.rows += n
        self.games += 1

    def _header(self, dt, shape) -> bytes:
        d = "{'descr': '%s', 'fortran_order': False, 'shape': %s, }" % (np.dtype(dt).str, repr(tuple(int(x) for x in shape)))
        pad = self.HEADER_BYTES - 10 - len(d) - 1
        if pad < 0:
            raise ValueError("npy header does not fit")
        return b"\x93NUMPY\x01\x00" + (self.HEADER_BYTES - 10).to_bytes(2, "little") + (d + " " * pad + "\n").encode("latin1")

    def flush(self):
        shapes = []
        for f, (name, dt, tail) in zip(self._f, self.FILES):
            f.flush()
            pos = f.tell()
            f.seek(0)
            f.write(self._header(dt, (self.rows, *tail)))
            f.flush()
            f.seek(pos)
            shapes.append([self.rows, *tail])
        meta = {  # same keys as convert.py:89-97
            "total_count": int(self.rows),
            "states_shape": shapes[0], "states_dtype": "float16",
            "mcts_shape": shapes[1], "mcts_dtype": "float64",
            "winners_shape": shapes[2], "winners_dtype": "float32",
        }
        with open(os.path.join(self.out_dir, "meta.json"), "w", encoding="utf-8") as f:
            json.dump(meta, f, ensure_ascii=False, indent=2)
        return self.rows

    def close(self):
        self.flush()
        for f in self._f:
            f.close()
        self._f = []
