"""HostRollout.play timing: chunks x stream priorities (2^20 envs, 32 placements per play)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tetris_b200 import HostRollout
E, T, K = 1 << 20, 32, 15
res = {}
for chunks in (1, 2, 4, 8, 16):
    for prio in (False, True):
        hr = HostRollout(10, 20, E, chunks=chunks, piece_set=1, seed=0x5EED, prioritized=prio)
        for sub in hr.envs:
            sub.rollout(30, "random"); sub.rollout(64, "greedy")
        hr.pull()
        for _ in range(3):
            hr.play(T, "greedy")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(K):
            hr.play(T, "greedy")
        dt = time.perf_counter() - t0
        res["chunks%d_prio%d" % (chunks, prio)] = E * T * K / dt
        del hr
print(json.dumps(res))
