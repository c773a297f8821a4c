"""Small fixed workload for ncu: every kernel of the hot path, twice (first launch = warm-up).  The rollouts run on
boards after random play (tall, many terminal placements), K1 / K2 on boards after greedy play (as in bench.py).

    python profiles/prof_run.py [n_env] [rollout_steps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tetris_b200 import BatchedTetris

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
env = BatchedTetris(10, 20, n, piece_set=1, seed=0x5EED)
env.rollout(30, "random")
for _ in range(2):
    env.rollout(T, "greedy")
for _ in range(2):
    env.rollout(T, "random")
env.rollout(64, "greedy")          # back to greedy-play boards: K1 / K2 below see what bench.py's roofline leg sees
feats = valid = count = None
for _ in range(2):
    feats, valid, count = env.get_after_states()
for _ in range(2):
    env.step(torch.zeros(n, dtype=torch.int32, device="cuda"), auto_reset=True, check=False)
# steady-state greedy play (what bench.py's timed region sees): greedy launches 3 (warm-up to the steady state), 4 and 5
env.rollout(600, "greedy")
for _ in range(2):
    env.rollout(T, "greedy")
torch.cuda.synchronize()
print("ok", env.stats_dict())
