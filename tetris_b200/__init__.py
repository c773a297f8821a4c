"""tetris_b200 -- B200-native batched Tetris environment.

Python surface of the reference (s0phia-/tetris: game.Tetris, state.State, tetromino.*, utils.*) over
device-resident bitboards stepped by hand-written sm_100a CUDA kernels (csrc/), reached through the C ABI in
include/tetris_b200.h.  `BatchedTetris` is the tensor-in/tensor-out API for millions of lockstep envs.
"""
from . import _lib
from ._lib import build  # noqa: F401

PIECE_NAMES = ("Straight", "RCorner", "LCorner", "Square", "SnakeR", "SnakeL", "T", "ThreeL", "ThreeLine")
PIECE_SETS = {0: (7, 8), 1: (0, 1, 2, 3, 4, 5, 6)}
FEATURE_NAMES = ("rows_with_holes", "column_transitions", "holes", "landing_height", "cumulative_wells",
                 "row_transitions", "eroded", "hole_depth")
# Thiery & Scherrer BCTS weights as hard-coded in Tetris.fitness (game.py:111-118)
BCTS_WEIGHTS = (-24.04, -19.77, -13.08, -12.63, -10.49, -9.22, 6.6, -1.61)


def __getattr__(name):
    if name == "BatchedTetris":
        from .batched import BatchedTetris
        return BatchedTetris
    if name == "HostRollout":
        from .batched import HostRollout
        return HostRollout
    if name == "Tetris":
        from .game import Tetris
        return Tetris
    if name in ("game", "state", "tetromino", "utils", "batched", "distributed"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
