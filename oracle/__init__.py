"""CPU oracle for the tetris hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this package.  See tetris_oracle.c.
"""
from .oracle import *  # noqa: F401,F403
