"""Compile the plain-C oracle (TEST INFRASTRUCTURE ONLY) into oracle/_build/libxq_oracle.so.

`__graft_entry__.build()` calls :func:`ensure_built`; the resulting .so is git-ignored but
travels to the GPU box with the snapshot.  gcc is the only requirement.
"""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "xq_oracle.c")
OUT_DIR = os.path.join(_HERE, "_build")
OUT = os.path.join(OUT_DIR, "libxq_oracle.so")


def ensure_built(force: bool = False) -> str:
    stale = (not os.path.exists(OUT)) or os.path.getmtime(OUT) < os.path.getmtime(SRC)
    if force or stale:
        os.makedirs(OUT_DIR, exist_ok=True)
        tmp = OUT + f".tmp{os.getpid()}"
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-Wall", "-o", tmp, SRC], check=True)
        os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    print(ensure_built(force=True))
