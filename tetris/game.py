"""`game` as a top-level module (the reference's checkout is used with its directory on sys.path)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tetris_b200.game import *  # noqa: E402,F401,F403
from tetris_b200 import game as _impl  # noqa: E402

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
