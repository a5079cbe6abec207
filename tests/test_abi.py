"""The C-ABI library loads and exports every symbol include/ccz_b200.h declares (no GPU needed)."""
import ctypes
import json
import os
import re

import numpy as np

from oracle import cchess_shim as cs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ccz_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ccz_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    from chinesechesszero_b200 import _lib, build

    build.build()
    lib = ctypes.CDLL(_lib.library_path())
    names = _declared_symbols()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/ccz_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTS)


def test_host_action_table_equals_oracle_and_reference_golden(golden_dir):
    from chinesechesszero_b200 import tools

    with open(os.path.join(golden_dir, "action_table.json")) as f:
        gold = json.load(f)
    assert [tools.move_id2move_action[i] for i in range(2086)] == gold["move_id2move_action"]
    assert tools.FLIP_MAP.tolist() == gold["flip_map"]
    id_of, fr, to = cs.action_table()
    assert np.array_equal(id_of, tools.ID_OF) and np.array_equal(fr, tools.FROM_OF) and np.array_equal(to, tools.TO_OF)
    assert tools.flip("a0a1") == "i0i1" and tools.flip("d9e8") == "f9e8"


def test_decode_board_matches_oracle():
    from chinesechesszero_b200 import tools
    from tests import positions

    recs = positions.random_playout_positions(3, 60, seed=5, every=7)
    for rec in recs:
        red = np.zeros(630, dtype=np.int8)
        black = np.zeros(630, dtype=np.int8)
        raw = np.ascontiguousarray(rec)
        cs.lib().xq_decode_board(raw.ctypes.data, red.ctypes.data, black.ctypes.data)
        r2, b2 = tools.decode_board(rec)
        assert np.array_equal(red.reshape(7, 10, 9), r2) and np.array_equal(black.reshape(7, 10, 9), b2)


def test_struct_layout_matches_header():
    from chinesechesszero_b200 import _lib

    assert ctypes.sizeof(_lib.ArenaStruct) == 16 + 16 * 8


def test_header_is_valid_c_and_matches_ctypes_struct(tmp_path):
    """include/ccz_b200.h compiles as plain C (it is the drop-in boundary) and sizeof(ccz_arena)
    equals the ctypes mirror."""
    import subprocess

    from chinesechesszero_b200 import _lib

    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include "ccz_b200.h"\n'
                   'int main(void){printf("%zu %d %d %d %zu %zu\\n", sizeof(ccz_arena), CCZ_BOARD_BYTES, CCZ_N_ACTIONS, '
                   'CCZ_FLAG_TIE_MASK, sizeof(ccz_node), sizeof(ccz_link));return 0;}\n')
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) == ctypes.sizeof(_lib.ArenaStruct)
    assert [int(x) for x in out[1:]] == [_lib.BOARD_BYTES, _lib.N_ACTIONS, _lib.FLAG_TIE_MASK, 16, 8]
    assert _lib.NODE_BYTES == 24


def test_conv_work_plan_covers_every_tile_and_channel_exactly_once():
    """K9's persistent schedule (host logic of ccz_conv3x3_c256, csrc/ccz_conv.cuh): whole tiles round-robin over the
    resident clusters, the last partial round sliced by output channels.  Every (tile, channel) must be produced by
    exactly one work item, slices may only appear in the last round, and no cluster gets more than one slice."""
    from chinesechesszero_b200 import _lib

    for variant, clusters in ((0, 74), (2, 74), (1, 148), (34, 34), (66, 16), (6, 74)):
        for n in (1, 2, 3, 37, 90, 128, 300, 777, 1024, 2048, 4096, 8192, 12345):
            p = _lib.conv3x3_plan(n, variant, clusters)
            assert p["n_tiles"] == -(-n * 90 // p["rows_per_tile"])
            split = 1 << p["split_log2"]
            assert p["n_full"] <= p["n_tiles"] and p["n_items"] == p["n_full"] + (p["n_tiles"] - p["n_full"]) * split
            cover = {}
            for i in range(p["n_items"]):
                if i < p["n_full"]:
                    tile, lo, width = i, 0, 256
                else:
                    j = i - p["n_full"]
                    width = 256 >> p["split_log2"]
                    tile, lo = p["n_full"] + (j >> p["split_log2"]), (j & (split - 1)) * width
                for c in range(lo, lo + width, 64):
                    assert (tile, c) not in cover
                    cover[(tile, c)] = i
            assert len(cover) == p["n_tiles"] * 4 and {t for t, _ in cover} == set(range(p["n_tiles"]))
            assert 1 <= p["clusters"] <= min(clusters, p["n_items"])
            if split > 1:
                assert variant != 6                                    # bit 2 switches the slicing off
                assert p["n_full"] % clusters == 0                     # whole rounds first
                assert p["n_items"] - p["n_full"] <= clusters          # the sliced round fits one pass
            rounds = -(-p["n_items"] // p["clusters"])
            assert rounds == -(-p["n_tiles"] // clusters) or split > 1


def test_split_evaluator_routes_each_half_to_its_net():
    """evaluate.SplitEvaluator (host logic of the evaluation match): leaves [0, split) go to the first evaluator,
    the rest to the second; swapped() exchanges them; planes may be None (evaluators that read the boards)."""
    import torch

    from chinesechesszero_b200 import _lib
    from chinesechesszero_b200.evaluate import SplitEvaluator

    def make(tag, needs_planes):
        def ev(planes, boards):
            assert (planes is None) == (not needs_planes) or planes is not None
            n = boards.shape[0]
            return torch.full((n, 2086), float(tag)), _lib.POLICY_LOGITS, torch.full((n,), float(tag))
        ev.needs_planes = needs_planes
        return ev

    a, b = make(1, False), make(2, False)
    boards = torch.zeros((6, 96), dtype=torch.uint8)
    s = SplitEvaluator(a, b, 2)
    assert s.needs_planes is False
    pol, kind, val = s(None, boards)
    assert kind == _lib.POLICY_LOGITS and val.tolist() == [1, 1, 2, 2, 2, 2] and pol.shape == (6, 2086)
    assert s.swapped()(None, boards)[2].tolist() == [2, 2, 1, 1, 1, 1]
    assert SplitEvaluator(a, make(3, True), 3).needs_planes is True


def test_pool_geometry_helpers():
    """Host arithmetic of the page pool (no GPU): the worst-case pages of a search and the policy struct encoding."""
    import ctypes

    from chinesechesszero_b200 import _lib

    # 2048-node pages hold >= 17 child runs of <= 119 nodes; + the partly used page the search starts in
    assert _lib.search_pages(400, 11) == -(-400 // 17) + 1 == 25
    assert _lib.search_pages(800, 11) == 49 and _lib.search_pages(1, 11) == 2
    assert _lib.search_pages(40, 7) == 41          # 128-node pages: one run per page in the worst case
    assert _lib.NODE_BYTES == 24 and _lib.MAX_CHILDREN == 119
    # the order policy round-trips through the library's host state
    odd = {"class_rank": {"p": 0, "c": 2, "r": 2, "n": 1, "b": 1, "a": 1, "k": 3}, "from_descending": 0,
           "to_descending": 1, "capture_mode": 2, "check_king_first": 1}
    try:
        _lib.set_order_policy(odd)
        assert _lib.get_order_policy() == odd
        bad = ctypes.create_string_buffer(bytes([0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 3, 0]), 12)  # capture_mode 3
        assert _lib.load().ccz_set_order_policy(bad) < 0 and b"capture_mode" in _lib.load().ccz_last_error()
        assert _lib.get_order_policy() == odd      # a rejected policy changes nothing
    finally:
        _lib.set_order_policy(None)
    assert _lib.get_order_policy() == _lib.DEFAULT_ORDER_POLICY
