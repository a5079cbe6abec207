"""h5lite -- a minimal HDF5 writer/reader for the reference's replay container ``data/data.h5``.

The reference appends one group per game with h5py (collect.py:146-167):
    /                      attribute  iters (int64 scalar) = number of games
    /game_{k}/states       (2T,17,7,10,9) float16   gzip
    /game_{k}/mcts_probs   (2T,2086)      float64   gzip
    /game_{k}/winners      (2T,)          float64
and reads it back with ``h5f.get(f"game_{i}")[name][:]`` (convert.py:38-81).  Neither h5py nor
libhdf5 exists in this image, so this module writes that layout directly from the HDF5 File Format
Specification using only the oldest, universally readable structures: version-0 superblock,
version-1 object headers, symbol-table groups (v1 B-tree + local heap + SNOD), contiguous or
single-chunk deflate datasets (v1 chunk B-tree, filter pipeline v1), version-1 attribute messages.
Appending rewrites only the root group's index (new copies at the end of the file; the superblock
is updated last, so a crash leaves the previous consistent state).

STATUS: round-trips through the reader below (which also walks multi-chunk B-trees, header
continuation blocks and v2 dataspaces as h5py writes them); it has NOT been opened with libhdf5,
which is unavailable here.
"""
from __future__ import annotations

import os
import mmap
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
LEAF_K, INTERNAL_K, CHUNK_K = 4, 16, 32  # library defaults (group leaf / group internal / chunk B-tree)
FREE_NULL = 1                            # H5HL_FREE_NULL: end of a local heap's free list


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


# ---- message encoders -----------------------------------------------------------------------

def _dtype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f":
        prec = dt.itemsize * 8
        exp_size, mant_size, bias = {2: (5, 10, 15), 4: (8, 23, 127), 8: (11, 52, 1023)}[dt.itemsize]
        head = bytes([0x11, 0x20, prec - 1, 0x00]) + struct.pack("<I", dt.itemsize)
        props = struct.pack("<HHBBBBI", 0, prec, mant_size, exp_size, 0, mant_size, bias)
        return head + props
    if dt.kind in "iu":
        head = bytes([0x10, 0x08 if dt.kind == "i" else 0x00, 0x00, 0x00]) + struct.pack("<I", dt.itemsize)
        return head + struct.pack("<HH", 0, dt.itemsize * 8)
    raise TypeError(f"unsupported dtype {dt}")


def _dataspace_msg(shape) -> bytes:
    return bytes([1, len(shape), 0, 0, 0, 0, 0, 0]) + b"".join(struct.pack("<Q", int(d)) for d in shape)


def _message(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _object_header(messages: list[bytes]) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


def _attribute_msg(name: str, value) -> bytes:
    arr = np.asarray(value)
    nm = name.encode() + b"\0"
    dt, ds = _dtype_msg(arr.dtype), _dataspace_msg(arr.shape)
    head = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds))
    return head + _pad8(nm) + _pad8(dt) + _pad8(ds) + arr.tobytes()


class RawDataset:
    """A dataset as stored: shape, dtype and either the raw bytes (``gzip_level`` None) or the single deflate
    chunk this module writes (``gzip_level`` = the level recorded in the filter message)."""

    def __init__(self, shape, dtype, data: bytes, gzip_level: int | None):
        self.shape, self.dtype, self.data, self.gzip_level = tuple(shape), np.dtype(dtype), data, gzip_level
        self.ndim = len(self.shape)


class H5Writer:
    """Create / append ``game_{k}`` groups; ``mode`` "w" truncates, "a" keeps an existing file."""

    def __init__(self, path: str, mode: str = "a"):
        self.path = path
        self.groups: dict[str, int] = {}   # top-level name -> object header address
        self.attrs: dict[str, object] = {}
        if mode == "a" and os.path.exists(path) and os.path.getsize(path) > 0:
            with H5Reader(path) as r:
                self.groups = dict(r.root_links())
                self.attrs = dict(r.root_attrs())
            self.f = open(path, "r+b")
            self.f.seek(0, os.SEEK_END)
            self.eof = self.f.tell()
        else:
            self.f = open(path, "w+b")
            self.f.write(b"\0" * 96)
            self.eof = 96
        self._dirty = True

    # ---- low level -----------------------------------------------------------------------
    def _append(self, data: bytes) -> int:
        addr = self.eof + (-self.eof % 8)
        self.f.seek(addr)
        self.f.write(data)
        self.eof = addr + len(data)
        return addr

    def _write_dataset(self, arr, gzip_level: int | None) -> int:
        """``arr``: an array, or a ``RawDataset`` (shape, dtype and the stored bytes of a dataset read with
        ``H5Reader.raw_dataset``: merging shards copies the deflate streams instead of recompressing them)."""
        if isinstance(arr, RawDataset):
            raw = comp = arr.data
            gzip_level = arr.gzip_level
        else:
            arr = np.ascontiguousarray(arr)
            raw = arr.tobytes()
            comp = None
        msgs = [_message(0x0001, _dataspace_msg(arr.shape)), _message(0x0003, _dtype_msg(arr.dtype), 1)]
        if gzip_level is None or int(np.prod(arr.shape)) == 0:
            addr = self._append(raw) if raw else UNDEF
            msgs.append(_message(0x0005, bytes([2, 1, 2, 0])))
            msgs.append(_message(0x0008, bytes([3, 1]) + struct.pack("<QQ", addr, len(raw))))
        else:
            if comp is None:
                comp = zlib.compress(raw, gzip_level)
            chunk_addr = self._append(comp)
            rank = arr.ndim
            key_size = 8 + 8 * (rank + 1)
            key0 = struct.pack("<II", len(comp), 0) + b"".join(struct.pack("<Q", 0) for _ in range(rank + 1))
            key1 = struct.pack("<II", 0, 0) + struct.pack("<Q", int(arr.shape[0])) + b"\0" * (8 * rank)
            node = b"TREE" + struct.pack("<BBHQQ", 1, 0, 1, UNDEF, UNDEF) + key0 + struct.pack("<Q", chunk_addr) + key1
            node_size = 24 + (2 * CHUNK_K + 1) * key_size + 2 * CHUNK_K * 8
            btree = self._append(node + b"\0" * (node_size - len(node)))
            msgs.append(_message(0x0005, bytes([2, 3, 2, 0])))
            pipeline = struct.pack("<BB6x", 1, 1) + struct.pack("<HHHH", 1, 0, 1, 1) + struct.pack("<I4x", gzip_level)
            msgs.append(_message(0x000B, pipeline))
            dims = b"".join(struct.pack("<I", int(d)) for d in arr.shape) + struct.pack("<I", np.dtype(arr.dtype).itemsize)
            msgs.append(_message(0x0008, bytes([3, 2, rank + 1]) + struct.pack("<Q", btree) + dims))
        return self._append(_object_header(msgs))

    def _write_group_index(self, links: dict[str, int], extra_msgs=()) -> tuple[int, int, int]:
        """Heap + SNODs + B-tree + object header for a symbol-table group. Returns
        (header address, btree address, heap address)."""
        names = sorted(links)
        # local heap: "" at offset 0, then the names, then one free block
        offs, data = {}, bytearray(b"\0" * 8)
        for n in names:
            offs[n] = len(data)
            data += _pad8(n.encode() + b"\0")
        free_off = len(data)
        free_size = max(32, -(len(data) + 32) % 64 + 32)
        data += struct.pack("<QQ", FREE_NULL, free_size) + b"\0" * (free_size - 16)
        data_addr = self._append(bytes(data))
        heap = self._append(b"HEAP" + struct.pack("<B3xQQQ", 0, len(data), free_off, data_addr))
        # leaves: symbol table nodes of <= 2*LEAF_K entries
        cap = 2 * LEAF_K
        level: list[tuple[int, int]] = []   # (address, heap offset of the largest name below)
        for i in range(0, max(len(names), 1), cap):
            part = names[i:i + cap]
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
            for n in part:
                body += struct.pack("<QQII16x", offs[n], links[n], 0, 0)
            body += b"\0" * (8 + cap * 40 - len(body))
            level.append((self._append(body), offs[part[-1]] if part else 0))
        # B-tree levels bottom-up
        depth = 0
        fan = 2 * INTERNAL_K
        node_size = 24 + (fan + 1) * 8 + fan * 8
        while True:
            groups = [level[i:i + fan] for i in range(0, len(level), fan)]
            addrs = []
            base = self.eof + (-self.eof % 8)
            for gi in range(len(groups)):
                addrs.append(base + gi * node_size)
            nxt = []
            for gi, grp in enumerate(groups):
                left = addrs[gi - 1] if gi > 0 else UNDEF
                right = addrs[gi + 1] if gi + 1 < len(groups) else UNDEF
                first_key = 0 if gi == 0 else groups[gi - 1][-1][1]
                body = b"TREE" + struct.pack("<BBHQQ", 0, depth, len(grp), left, right) + struct.pack("<Q", first_key)
                for child, maxkey in grp:
                    body += struct.pack("<QQ", child, maxkey)
                got = self._append(body + b"\0" * (node_size - len(body)))
                assert got == addrs[gi]
                nxt.append((got, grp[-1][1]))
            level = nxt
            depth += 1
            if len(level) == 1:
                break
        btree = level[0][0]
        msgs = [_message(0x0011, struct.pack("<QQ", btree, heap))] + list(extra_msgs)
        return self._append(_object_header(msgs)), btree, heap

    # ---- public --------------------------------------------------------------------------
    def create_group(self, name: str, datasets: dict[str, np.ndarray], gzip: dict[str, int | None] | None = None):
        """One group with its datasets (the reference's per-game unit, collect.py:148-163)."""
        if name in self.groups:
            raise ValueError(f"group {name!r} exists")
        gzip = gzip or {}
        links = {ds: self._write_dataset(arr, gzip.get(ds)) for ds, arr in datasets.items()}
        self.groups[name], _, _ = self._write_group_index(links)
        self._dirty = True

    def flush(self):
        if not self._dirty:
            return
        attr_msgs = [_message(0x000C, _attribute_msg(k, v)) for k, v in sorted(self.attrs.items())]
        header, btree, heap = self._write_group_index(self.groups, attr_msgs)
        eof = self.eof + (-self.eof % 8)
        sb = SIGNATURE + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, header, 1, 0) + struct.pack("<QQ", btree, heap)
        assert len(sb) == 96
        self.f.seek(self.eof)
        self.f.write(b"\0" * (eof - self.eof))
        self.eof = eof
        self.f.flush()
        self.f.seek(0)
        self.f.write(sb)
        self.f.flush()
        self._dirty = False

    def close(self):
        self.flush()
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


# ---- reader ---------------------------------------------------------------------------------

class H5Reader:
    def __init__(self, path: str):
        self.f = open(path, "rb")
        try:  # map instead of read: replay files grow to tens of GB, only the touched pages become resident
            self.buf = mmap.mmap(self.f.fileno(), 0, access=mmap.ACCESS_READ)
        except ValueError:  # empty file
            self.buf = self.f.read()
        if self.buf[:8] != SIGNATURE:
            raise ValueError("not an HDF5 file")
        if self.buf[8] != 0 or self.buf[13] != 8 or self.buf[14] != 8:
            raise ValueError("only superblock v0 with 8-byte offsets is supported")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", self.buf, 16)
        self.eof = struct.unpack_from("<Q", self.buf, 40)[0]
        self.root_header = struct.unpack_from("<Q", self.buf, 56 + 8)[0]

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        if isinstance(self.buf, mmap.mmap):
            self.buf.close()
        self.f.close()

    # ---- object headers ------------------------------------------------------------------
    def messages(self, addr: int):
        b = self.buf
        version, _, nmsgs, _, size = struct.unpack_from("<BBHII", b, addr)
        if version != 1:
            raise ValueError("only version-1 object headers are supported")
        blocks, out = [(addr + 16, size)], []
        while blocks and len(out) < nmsgs:
            pos, length = blocks.pop(0)
            end = pos + length
            while pos + 8 <= end and len(out) < nmsgs:
                mtype, msize, flags = struct.unpack_from("<HHB", b, pos)
                data = b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x0010:
                    blocks.append(struct.unpack_from("<QQ", data, 0))
                out.append((mtype, data))
        return out

    def _heap_string(self, heap_addr: int, off: int) -> str:
        assert self.buf[heap_addr:heap_addr + 4] == b"HEAP"
        data_addr = struct.unpack_from("<Q", self.buf, heap_addr + 24)[0]
        end = self.buf.find(b"\0", data_addr + off)
        return self.buf[data_addr + off:end].decode()

    def _walk_group_btree(self, addr: int, heap: int, out: dict):
        b = self.buf
        assert b[addr:addr + 4] == b"TREE" and b[addr + 4] == 0
        level, used = b[addr + 5], struct.unpack_from("<H", b, addr + 6)[0]
        pos = addr + 24
        for _ in range(used):
            child = struct.unpack_from("<Q", b, pos + 8)[0]
            pos += 16
            if level > 0:
                self._walk_group_btree(child, heap, out)
            else:
                assert b[child:child + 4] == b"SNOD"
                n = struct.unpack_from("<H", b, child + 6)[0]
                for i in range(n):
                    name_off, obj = struct.unpack_from("<QQ", b, child + 8 + 40 * i)
                    out[self._heap_string(heap, name_off)] = obj

    def links(self, header_addr: int) -> dict[str, int]:
        out: dict[str, int] = {}
        for mtype, data in self.messages(header_addr):
            if mtype == 0x0011:
                btree, heap = struct.unpack_from("<QQ", data, 0)
                self._walk_group_btree(btree, heap, out)
        return out

    def root_links(self):
        if getattr(self, "_root_links", None) is None:   # the file is read-only while this reader is open
            self._root_links = self.links(self.root_header)
        return self._root_links

    def raw_dataset(self, header_addr: int) -> RawDataset | None:
        """The stored form of a contiguous or single-chunk-deflate dataset (what this module writes); None for
        any other layout (the caller falls back to ``dataset``)."""
        shape = dt = layout = None
        level, filters = None, 0
        for mtype, data in self.messages(header_addr):
            if mtype == 0x0001:
                shape = self._parse_dataspace(data)
            elif mtype == 0x0003:
                dt = self._parse_dtype(data)
            elif mtype == 0x0008:
                layout = data
            elif mtype == 0x000B:
                filters = data[1]
                if data[0] == 1 and data[1] == 1 and struct.unpack_from("<H", data, 8)[0] == 1:
                    level = struct.unpack_from("<I", data, 16)[0]
        if layout is None or layout[0] != 3:
            return None
        if layout[1] == 1 and filters == 0:
            addr, size = struct.unpack_from("<QQ", layout, 2)
            return RawDataset(shape, dt, b"" if addr == UNDEF else bytes(self.buf[addr:addr + size]), None)
        if layout[1] == 2 and filters == 1 and level is not None:
            rank1 = layout[2]
            btree = struct.unpack_from("<Q", layout, 3)[0]
            cdims = struct.unpack_from(f"<{rank1}I", layout, 11)[:-1]
            chunks: list = []
            if btree != UNDEF:
                self._chunks(btree, rank1, chunks)
            if tuple(cdims) == tuple(shape) and len(chunks) == 1 and chunks[0][2] == 0:
                _, size, _, addr = chunks[0]
                return RawDataset(shape, dt, bytes(self.buf[addr:addr + size]), int(level))
        return None

    def read_group_raw(self, name: str) -> dict:
        """{dataset name: RawDataset | ndarray} of a top-level group, without decompressing where possible."""
        out = {}
        for k, a in self.links(self.root_links()[name]).items():
            raw = self.raw_dataset(a)
            out[k] = raw if raw is not None else self.dataset(a)
        return out

    @staticmethod
    def _parse_dtype(data: bytes) -> np.dtype:
        cls, size = data[0] & 0x0F, struct.unpack_from("<I", data, 4)[0]
        if cls == 1:
            return np.dtype(f"<f{size}")
        if cls == 0:
            return np.dtype(("<i" if data[1] & 0x08 else "<u") + str(size))
        raise TypeError(f"unsupported datatype class {cls}")

    @staticmethod
    def _parse_dataspace(data: bytes):
        version, rank, flags = data[0], data[1], data[2]
        off = 8 if version == 1 else 4
        return tuple(struct.unpack_from("<Q", data, off + 8 * i)[0] for i in range(rank))

    def dataset_shape(self, header_addr: int) -> tuple:
        """Shape of a dataset from its data-space message, without reading the data."""
        for mtype, data in self.messages(header_addr):
            if mtype == 0x0001:
                return self._parse_dataspace(data)
        raise ValueError("object has no data-space message")

    def attrs(self, header_addr: int) -> dict:
        out = {}
        for mtype, data in self.messages(header_addr):
            if mtype != 0x000C or data[0] != 1:
                continue
            nlen, dlen, slen = struct.unpack_from("<HHH", data, 2)
            pos = 8
            name = data[pos:pos + nlen].split(b"\0")[0].decode()
            pos += nlen + (-nlen % 8)
            dt = self._parse_dtype(data[pos:pos + dlen])
            pos += dlen + (-dlen % 8)
            shape = self._parse_dataspace(data[pos:pos + slen])
            pos += slen + (-slen % 8)
            n = int(np.prod(shape)) if shape else 1
            val = np.frombuffer(data, dtype=dt, count=n, offset=pos).reshape(shape)
            out[name] = val[()] if shape == () else val.copy()
        return out

    def root_attrs(self):
        return self.attrs(self.root_header)

    def _chunks(self, addr: int, rank1: int, out: list):
        b = self.buf
        assert b[addr:addr + 4] == b"TREE" and b[addr + 4] == 1
        level, used = b[addr + 5], struct.unpack_from("<H", b, addr + 6)[0]
        key_size = 8 + 8 * rank1
        pos = addr + 24
        for _ in range(used):
            size, mask = struct.unpack_from("<II", b, pos)
            offs = struct.unpack_from(f"<{rank1}Q", b, pos + 8)
            child = struct.unpack_from("<Q", b, pos + key_size)[0]
            pos += key_size + 8
            if level > 0:
                self._chunks(child, rank1, out)
            else:
                out.append((offs[:-1], size, mask, child))

    def dataset(self, header_addr: int) -> np.ndarray:
        shape = dt = layout = None
        filters = []
        for mtype, data in self.messages(header_addr):
            if mtype == 0x0001:
                shape = self._parse_dataspace(data)
            elif mtype == 0x0003:
                dt = self._parse_dtype(data)
            elif mtype == 0x0008:
                layout = data
            elif mtype == 0x000B:
                version, nf = data[0], data[1]
                pos = 8 if version == 1 else 2
                for _ in range(nf):
                    fid, nlen, _, ncd = struct.unpack_from("<HHHH", data, pos)
                    pos += 8
                    if version == 1 or fid >= 256:
                        pos += nlen + (-nlen % 8 if version == 1 else 0)
                    pos += 4 * ncd
                    if version == 1 and ncd % 2:
                        pos += 4
                    filters.append(fid)
        if layout[0] != 3:
            raise ValueError("only data layout message v3 is supported")
        n = int(np.prod(shape)) if shape else 1
        if layout[1] == 1:
            addr, size = struct.unpack_from("<QQ", layout, 2)
            if n == 0 or addr == UNDEF:
                return np.zeros(shape, dtype=dt)
            return np.frombuffer(self.buf, dtype=dt, count=n, offset=addr).reshape(shape).copy()
        if layout[1] == 0:
            size = struct.unpack_from("<H", layout, 2)[0]
            return np.frombuffer(layout, dtype=dt, count=n, offset=4).reshape(shape).copy()
        rank1 = layout[2]
        btree = struct.unpack_from("<Q", layout, 3)[0]
        cdims = struct.unpack_from(f"<{rank1}I", layout, 11)[:-1]
        out = np.zeros(shape, dtype=dt)
        chunks: list = []
        if btree != UNDEF:
            self._chunks(btree, rank1, chunks)
        for offs, size, mask, addr in chunks:
            raw = self.buf[addr:addr + size]
            for k, fid in reversed(list(enumerate(filters))):
                if mask & (1 << k):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:  # shuffle
                    a = np.frombuffer(raw, dtype=np.uint8).reshape(dt.itemsize, -1)
                    raw = a.T.tobytes()
                else:
                    raise ValueError(f"unsupported filter {fid}")
            chunk = np.frombuffer(raw, dtype=dt, count=int(np.prod(cdims))).reshape(cdims)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
            out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
        return out

    def read_group(self, name: str) -> dict[str, np.ndarray]:
        links = self.root_links()
        return {k: self.dataset(a) for k, a in self.links(links[name]).items()}


class H5ReplayWriter:
    """The reference's ``collect_data`` persistence (collect.py:146-169) on top of H5Writer.

    The game data is appended as soon as ``add`` is called; the root group's INDEX (heap, symbol nodes,
    B-tree, header: ~65 B per game, re-emitted whole because the v1 structures are written densely) and
    the ``iters`` attribute are rewritten by ``flush``.  ``flush_every`` = N flushes after every N games
    (1 = the reference's behaviour, one consistent file per game, at O(games^2) dead index bytes);
    the default None flushes when ``max(64, games / 16)`` games are pending or ``flush_seconds`` have
    passed, which keeps the dead index below ~1 KB per game, and always on ``close``.  A crash loses at
    most the games since the last flush (their data is in the file but not linked)."""

    def __init__(self, path: str, gzip_level: int | None = 4, flush_every: int | None = None,
                 flush_seconds: float = 30.0):
        import time

        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        self.w = H5Writer(path, "a")
        self.gzip_level = gzip_level
        self.flush_every = flush_every
        self.flush_seconds = float(flush_seconds)
        self.iters = int(self.w.attrs.get("iters", 0))
        self._pending = 0
        self._clock = time.monotonic
        self._last_flush = self._clock()
        self.flushes = 0

    @property
    def n_groups(self) -> int:
        return len(self.w.groups)

    def add(self, states, mcts_probs, winners, index: int | None = None) -> int:
        k = self.iters if index is None else int(index)
        self.w.create_group(
            f"game_{k}",
            {"states": np.asarray(states, dtype=np.float16), "mcts_probs": np.asarray(mcts_probs, dtype=np.float64),
             "winners": np.asarray(winners, dtype=np.float64)},
            gzip={"states": self.gzip_level, "mcts_probs": self.gzip_level, "winners": None})
        self.iters = max(self.iters, k + 1) if index is not None else self.iters + 1
        self.w.attrs["iters"] = np.int64(self.iters)
        self._pending += 1
        if self.flush_every is not None:
            due = self._pending >= self.flush_every
        else:
            due = (self._pending >= max(64, len(self.w.groups) // 16)
                   or self._clock() - self._last_flush >= self.flush_seconds)
        if due:
            self.flush()
        return self.iters

    def add_raw(self, datasets: dict, index: int | None = None) -> int:
        """Append a game whose datasets are already in stored form (``H5Reader.read_group_raw``)."""
        k = self.iters if index is None else int(index)
        gz = {n: (d.gzip_level if isinstance(d, RawDataset) else (None if n == "winners" else self.gzip_level))
              for n, d in datasets.items()}
        self.w.create_group(f"game_{k}", datasets, gzip=gz)
        self.iters = max(self.iters, k + 1) if index is not None else self.iters + 1
        self.w.attrs["iters"] = np.int64(self.iters)
        self._pending += 1
        if self._pending >= max(64, len(self.w.groups) // 16) and self.flush_every is None or \
                (self.flush_every is not None and self._pending >= self.flush_every):
            self.flush()
        return self.iters

    def flush(self):
        if self._pending or self.w._dirty:
            self.w.flush()
            self.flushes += 1
        self._pending = 0
        self._last_flush = self._clock()
        return self.iters

    def close(self):
        self.flush()
        self.w.close()
