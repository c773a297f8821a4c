"""Small workload touching every kernel and board shape, for compute-sanitizer (memcheck / racecheck):

    compute-sanitizer --tool memcheck python profiles/sanitize_run.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tetris_b200 import BatchedTetris

for (C, R) in ((10, 20), (10, 10), (6, 12), (8, 16), (4, 4)):
    for ps in (0, 1):
        n = 777
        env = BatchedTetris(C, R, n, piece_set=ps, seed=5, feature_directions=[-1, -1, -1, -1, -1, -1, 1, -1] if ps else None)
        env.rollout(12, "random")
        env.rollout(6, "greedy")
        for inc in (False, True):
            f, v, c = env.get_after_states(include_terminal=inc)
        a = (torch.arange(n, device="cuda") % c.long()).int()
        env.step(a, auto_reset=True)
        env.reset_masked(np.arange(n) % 3 == 0)
        r, h, p = env.export_boards()
        env.import_boards(r.cpu().numpy().view(np.uint16), piece=p)
torch.cuda.synchronize()
print("sanitize_run ok")
