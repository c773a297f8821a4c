"""K2 (k_step) and random-rollout device times at 2^20 envs on greedy-play and on random-play boards."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tetris_b200 import BatchedTetris
n = 1 << 20
res = {}
for name, prep in (("greedy_boards", lambda e: (e.rollout(30, "random"), e.rollout(64, "greedy"))), ("random_boards", lambda e: e.rollout(60, "random"))):
    env = BatchedTetris(10, 20, n, piece_set=1, seed=0x5EED)
    prep(env)
    a = torch.zeros(n, dtype=torch.int32, device="cuda")
    ts = []
    for i in range(8):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); env.step(a, auto_reset=True, check=False); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    res["k2_ms_" + name] = sorted(ts[2:])[len(ts[2:]) // 2]
    ts = []
    for i in range(4):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); env.rollout(32, "random"); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    res["random_rollout32_ms_after_" + name] = sorted(ts[1:])[1]
print(json.dumps(res))
