"""make_ref -- recipe for ``oracle/_ref``: a snapshot of the UNMODIFIED reference modules on the
self-play path, so that ``bench.py --impl reference`` can run the reference itself (not a port) on
the GPU box, where /root/reference does not exist.  TEST / BASELINE INFRASTRUCTURE ONLY.

``oracle/_ref/`` is listed in .gitignore (the reference's sources never enter the history) and is not
gpurun-ignored (it travels with the snapshot like the built .so files).  Run by
``__graft_entry__.build()`` whenever /root/reference is present; a no-op elsewhere.
"""
from __future__ import annotations

import filecmp
import os
import shutil

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("CCZ_REFERENCE_SRC", "/root/reference")
OUT = os.path.join(_HERE, "_ref")
# the import closure of collect.py -> game.py -> mcts.py -> net.py -> tools.py (frontend.py is imported by game.py)
MODULES = ("parameters.py", "tools.py", "mcts.py", "net.py", "game.py", "collect.py", "frontend.py")


def ensure_ref() -> str | None:
    """Refresh the snapshot when the reference is present; return the directory (None if there is none)."""
    if os.path.isfile(os.path.join(SRC, "mcts.py")):
        os.makedirs(OUT, exist_ok=True)
        for name in MODULES:
            src, dst = os.path.join(SRC, name), os.path.join(OUT, name)
            if not os.path.exists(dst) or not filecmp.cmp(src, dst, shallow=False):
                shutil.copyfile(src, dst)
    return OUT if os.path.isfile(os.path.join(OUT, "mcts.py")) else None


if __name__ == "__main__":
    print(ensure_ref())
