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

    assert ctypes.sizeof(_lib.ArenaStruct) == 8 + 12 * 8


def test_header_is_valid_c_and_matches_ctypes_struct(tmp_path):
    """include/ccz_b200.h compiles as plain C (it is the drop-in boundary) and sizeof(ccz_arena)
    equals the ctypes mirror."""
    import subprocess

    from chinesechesszero_b200 import _lib

    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include "ccz_b200.h"\n'
                   'int main(void){printf("%zu %d %d %d\\n", sizeof(ccz_arena), CCZ_BOARD_BYTES, CCZ_N_ACTIONS, '
                   'CCZ_FLAG_TIE_MASK);return 0;}\n')
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) == ctypes.sizeof(_lib.ArenaStruct)
    assert [int(x) for x in out[1:]] == [_lib.BOARD_BYTES, _lib.N_ACTIONS, _lib.FLAG_TIE_MASK]
