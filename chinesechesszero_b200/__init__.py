"""chinesechesszero_b200 -- B200-native drop-in for the self-play data path of
Symb0x76/ChineseChessZero (collect.py -> game.py -> mcts.py -> net.py -> cchess).

CUDA kernels for sm_100a live in ``csrc/`` behind the C ABI declared in ``include/ccz_b200.h``;
the modules here mirror the reference's host-side interface for that path.
"""
__version__ = "0.1.0"
