"""load_reference -- CPU ORACLE helper (TEST INFRASTRUCTURE ONLY).

Imports the UNMODIFIED reference modules (tools, mcts, net, game, collect) from /root/reference
after installing stand-ins for the packages it needs but the container lacks:
``cchess`` -> oracle.cchess_shim, ``cchess.svg``, ``IPython.display``, ``h5py``, ``frontend``.

/root/reference only exists in the authoring container.  Parity tests never depend on it at run
time (they use the golden fixtures under tests/golden/ that scripts/make_golden.py wrote from it); the
reference arm of bench.py uses the git-ignored snapshot ``oracle/_ref`` (oracle/make_ref.py) on the GPU box.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_SNAPSHOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")  # written by oracle/make_ref.py


def _find_reference() -> str:
    """/root/reference in the authoring container; on the GPU box the git-ignored snapshot oracle/_ref."""
    env = os.environ.get("CCZ_REFERENCE_DIR")
    if env:
        return env
    if os.path.isfile("/root/reference/mcts.py"):
        return "/root/reference"
    return _SNAPSHOT


REFERENCE_DIR = _find_reference()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "mcts.py"))


def install_stubs() -> None:
    from . import cchess_shim

    if "cchess" not in sys.modules or sys.modules["cchess"] is not cchess_shim:
        sys.modules["cchess"] = cchess_shim
    svg = types.ModuleType("cchess.svg")
    svg.board = lambda *a, **k: "<svg/>"
    sys.modules["cchess.svg"] = svg
    cchess_shim.svg = svg
    if "IPython" not in sys.modules:
        ip = types.ModuleType("IPython")
        disp = types.ModuleType("IPython.display")
        disp.display = lambda *a, **k: None
        disp.SVG = lambda x: x
        ip.display = disp
        sys.modules["IPython"] = ip
        sys.modules["IPython.display"] = disp
    if "h5py" not in sys.modules:
        try:
            importlib.import_module("h5py")
        except Exception:
            sys.modules["h5py"] = types.ModuleType("h5py")


def load(*names: str):
    """Return the named unmodified reference modules, e.g. ``tools, mcts = load("tools", "mcts")``."""
    if not available():
        raise FileNotFoundError(f"reference not present at {REFERENCE_DIR}")
    install_stubs()
    if REFERENCE_DIR not in sys.path:
        sys.path.append(REFERENCE_DIR)
    cwd = os.getcwd()
    mods = []
    for n in names:
        key = f"_ccz_ref_{n}"
        if key in sys.modules:
            mods.append(sys.modules[key])
            continue
        # the reference modules import each other by bare name (tools, parameters, ...)
        mods.append(importlib.import_module(n))
    os.chdir(cwd)
    return mods[0] if len(mods) == 1 else tuple(mods)
