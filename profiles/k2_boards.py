"""tb_step device time on three kinds of boards (2^20 envs, 10x20): steady-state greedy play (what bench.py's K2 leg sees),
boards shortly after random play, and tall random-play boards.  State restored before every repetition, L2 flushed.

    [TB_SO_PATH=...] python profiles/k2_boards.py [k2_cfg ...]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tetris_b200 import BatchedTetris, _lib

n = 1 << 20
cfgs = [int(x) for x in sys.argv[1:]] or [-1]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
a0 = torch.zeros(n, dtype=torch.int32, device="cuda")
boards = {}
env = BatchedTetris(10, 20, n, piece_set=1, seed=0x5EED)
env.rollout(30, "random"); boards["tall_random_play"] = env.state.clone()
env.rollout(64, "greedy"); boards["after_64_greedy"] = env.state.clone()
env.rollout(700, "greedy"); boards["steady_greedy"] = env.state.clone()
out = {}
for cfg in cfgs:
    try:
        _lib.set_tuning("k2_cfg", cfg)
    except Exception:
        pass
    for name, st in boards.items():
        ts = []
        for rep in range(8):
            env.state.copy_(st)
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); env.step(a0, auto_reset=True, check=False); e.record()
            torch.cuda.synchronize()
            if rep >= 2:
                ts.append(s.elapsed_time(e))
        ts.sort()
        out["cfg%d_%s_ms" % (cfg, name)] = round(ts[len(ts) // 2], 4)
print(json.dumps(out))
