#!/bin/bash
# quick GPU check: parity tests + bench line (no extras) + K1/K3 timing under the tuning configs
[ -n "$TB_SO_PATH" ] || python -m pytest tests/test_cuda_parity.py -x -q 2>&1 | tail -5
for cfg in ${K1CFGS:-0 1 2}; do TB_K1_CFG=$cfg python - <<'PY'
import os, torch
from tetris_b200 import BatchedTetris
env = BatchedTetris(10, 20, 1 << 20, piece_set=1, seed=0x5EED)
env.rollout(30, "random"); env.rollout(64, "greedy")
f = torch.empty((1 << 20, env.a_max, 8), dtype=torch.float32, device="cuda"); v = torch.empty(1 << 20, dtype=torch.int64, device="cuda"); c = torch.empty(1 << 20, dtype=torch.int32, device="cuda")
for _ in range(3): env.get_after_states(out=(f, v, c))
torch.cuda.synchronize()
ts = []
for _ in range(7):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); env.get_after_states(out=(f, v, c)); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
print("K1 cfg", os.environ.get("TB_K1_CFG"), "ms", sorted(ts)[3], "afterstates/s %.3e" % (float(c.sum()) / (sorted(ts)[3] * 1e-3)))
PY
done
for cfg in ${K3CFGS:-0 1}; do TB_K3_CFG=$cfg python - <<'PY'
import os, torch
from tetris_b200 import BatchedTetris
env = BatchedTetris(10, 20, 1 << 20, piece_set=1, seed=0x5EED)
env.rollout(30, "random"); env.rollout(64, "greedy")
torch.cuda.synchronize()
ts = []
for _ in range(5):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); env.rollout(32, "greedy"); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
t = sorted(ts)[2]
print("K3 cfg", os.environ.get("TB_K3_CFG"), "ms/32 steps", t, "placements/s %.3e" % ((1 << 20) * 32 / (t * 1e-3)))
ts = []
for _ in range(3):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); env.rollout(32, "random"); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
t = sorted(ts)[1]
print("K3 random ms/32 steps", t, "placements/s %.3e" % ((1 << 20) * 32 / (t * 1e-3)))
PY
done
