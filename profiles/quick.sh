#!/bin/bash
# quick GPU check: parity tests + bench line (no extras) + K1 timing under both occupancy settings
python -m pytest tests/test_cuda_parity.py -x -q 2>&1 | tail -5
python bench.py --no-extras 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench placements/s', d['value'], 'ms/step', d['ms_per_step'])"
for mb in 3 2; do TB_K1_MINB=$mb python - <<'PY'
import os, torch
from tetris_b200 import BatchedTetris
env = BatchedTetris(10, 20, 1 << 20, piece_set=1, seed=0x5EED)
env.rollout(30, "random"); env.rollout(64, "greedy")
out = None
f = torch.empty((1 << 20, env.a_max, 8), dtype=torch.float32, device="cuda"); v = torch.empty(1 << 20, dtype=torch.int64, device="cuda"); c = torch.empty(1 << 20, dtype=torch.int32, device="cuda")
for _ in range(3): env.get_after_states(out=(f, v, c))
torch.cuda.synchronize()
ts = []
for _ in range(5):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); env.get_after_states(out=(f, v, c)); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
print("K1 minb", os.environ.get("TB_K1_MINB"), "ms", sorted(ts)[2], "afterstates/s %.3e" % (float(c.sum()) / (sorted(ts)[2] * 1e-3)))
PY
done
