"""Search / training constants of the self-play path.

The VALUES are part of the drop-in contract (they are the reference's, parameters.py:8-28) and are
exported under the reference's names; ``override`` lets a launcher change them in one place before
the pipelines are constructed (e.g. ``override(PLAYOUT=400)`` for the 400-playout benchmark setting).
"""
from __future__ import annotations

_CONTRACT = {
    # name: (value, meaning, reference line)
    "C_PUCT": (5, "PUCT exploration constant: Q + C_PUCT * P * sqrt(N_parent) / (1 + N)", 8),
    "EPS": (0.25, "weight of the Dirichlet noise in the self-play move choice", 10),
    "ALPHA": (0.2, "Dirichlet concentration", 12),
    "PLAYOUT": (1600, "playouts per move", 14),
    "DATA_DIR": ("data", "replay directory (data.h5, states/mcts/winners.npy)", 16),
    "MODEL_DIR": ("models", "checkpoint directory (current_policy.pkl)", 18),
    "BATCH_SIZE": (2048, "training batch size", 20),
    "EPOCHS": (10, "unused by the reference's trainer", 22),
    "KL_TARG": (0.02, "KL target of the adaptive learning-rate multiplier", 24),
    "CHECK_FREQ": (10, "numbered checkpoint every CHECK_FREQ iterations", 26),
    "LOG_LEVEL": (1, "console log threshold (1 DEBUG .. 5 CRITICAL)", 28),
}

globals().update({name: spec[0] for name, spec in _CONTRACT.items()})
__all__ = sorted(_CONTRACT)


def describe() -> dict:
    """name -> (value, meaning, 'parameters.py:<line>') for documentation and logs."""
    return {n: (globals()[n], m, f"parameters.py:{ln}") for n, (_, m, ln) in _CONTRACT.items()}


def override(**values) -> None:
    """Change constants by name; unknown names are an error."""
    for name, value in values.items():
        if name not in _CONTRACT:
            raise KeyError(f"unknown parameter {name!r}")
        globals()[name] = value
