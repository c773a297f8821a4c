"""Tile quantisation of the tile kernels: per-env time of K1 (3 CTAs of 256 envs per SM = 444 resident tiles) and of the
greedy rollout K3 (same residency) at env counts that fill the last round of tiles fully or partially.

    python profiles/tail_effect.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tetris_b200 import BatchedTetris

out = []
for n in (1 << 20, 444 * 9 * 256, 444 * 10 * 256, 444 * 18 * 256):
    env = BatchedTetris(10, 20, n, piece_set=1, seed=0x5EED)
    env.rollout(30, "random"); env.rollout(64, "greedy")
    f = torch.empty((n, env.a_max, 8), dtype=torch.float32, device="cuda")
    v = torch.empty(n, dtype=torch.int64, device="cuda"); c = torch.empty(n, dtype=torch.int32, device="cuda")
    for _ in range(3):
        env.get_after_states(out=(f, v, c))
    torch.cuda.synchronize()

    def timed(fn, reps):
        ts = []
        for _ in range(reps):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
        return sorted(ts)[len(ts) // 2]
    k1 = timed(lambda: env.get_after_states(out=(f, v, c)), 7)
    env.rollout(32, "greedy")
    k3 = timed(lambda: env.rollout(32, "greedy"), 5)
    out.append({"envs": n, "tiles_per_resident_cta": n / 256 / 444, "k1_ms": k1, "k1_ns_per_env": 1e6 * k1 / n,
                "k3_ms_per_32": k3, "k3_placements_per_s": n * 32 / (k3 * 1e-3)})
    del f, env
print(json.dumps(out))
