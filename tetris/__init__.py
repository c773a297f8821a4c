"""Package name the reference's own imports use (`from tetris import state`, `from tetris.game import Tetris`,
game.py:3-5, tetromino.py:2).  Everything resolves to the B200-native implementation in `tetris_b200`; the
directory can also be put on sys.path directly, as the reference's checkout is used (`from game import Tetris`,
example_play.py:1).
"""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

for _name in ("utils", "state", "tetromino", "game"):
    _mod = importlib.import_module("tetris_b200." + _name)
    sys.modules[__name__ + "." + _name] = _mod
    globals()[_name] = _mod

from tetris_b200 import BatchedTetris  # noqa: E402,F401  (batched API, not in the reference)
