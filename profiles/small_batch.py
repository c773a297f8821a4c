"""Small-batch A/B: K1 / K3 device time at n_env below the SM count x 256, per tile configuration (TB_K1_CFG / TB_K3_CFG
forced through the environment; unset = the library's own choice).  Usage: python profiles/small_batch.py"""
import json
import os
import subprocess
import sys

CHILD = r'''
import os, sys, json, torch
from tetris_b200 import BatchedTetris
res = {}
for n in (1024, 4096, 8192, 16384, 32768, 65536):
    env = BatchedTetris(10, 20, n, piece_set=1, seed=0x5EED)
    env.rollout(30, "random"); env.rollout(32, "greedy")
    f = torch.empty((n, env.a_max, 8), dtype=torch.float32, device="cuda"); v = torch.empty(n, dtype=torch.int64, device="cuda"); c = torch.empty(n, dtype=torch.int32, device="cuda")
    for _ in range(5): env.get_after_states(out=(f, v, c))
    torch.cuda.synchronize()
    def med(fn, reps=15):
        ts = []
        for _ in range(reps):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
        return sorted(ts)[len(ts) // 2]
    k1 = med(lambda: env.get_after_states(out=(f, v, c)))
    k3 = med(lambda: env.rollout(32, "greedy"), 7)
    res[n] = {"k1_us": 1e3 * k1, "k3_us_per_32": 1e3 * k3}
print(json.dumps(res))
'''
out = {}
for name, envs in (("default", {}), ("cfg0", {"TB_K1_CFG": "0", "TB_K3_CFG": "0"}), ("k1cfg3_k3cfg2", {"TB_K1_CFG": "3", "TB_K3_CFG": "2"}),
                   ("cfg4_32x128", {"TB_K1_CFG": "4", "TB_K3_CFG": "4"}), ("cfg5_64x256", {"TB_K1_CFG": "5", "TB_K3_CFG": "5"})):
    e = dict(os.environ); e.update(envs)
    r = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out[name] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else r.stderr[-400:]
print(json.dumps(out, indent=1))
