import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, sys
from tetris_b200 import BatchedTetris
for n in (1 << 20, 444 * 9 * 256, 444 * 10 * 256, 444 * 18 * 256):
    env = BatchedTetris(10, 20, n, piece_set=1, seed=0x5EED)
    env.rollout(30, "random"); env.rollout(64, "greedy")
    f = torch.empty((n, env.a_max, 8), dtype=torch.float32, device="cuda"); v = torch.empty(n, dtype=torch.int64, device="cuda"); c = torch.empty(n, dtype=torch.int32, device="cuda")
    for _ in range(3): env.get_after_states(out=(f, v, c))
    torch.cuda.synchronize(); ts = []
    for _ in range(7):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); env.get_after_states(out=(f, v, c)); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    t = sorted(ts)[3]
    print("n", n, "tiles/CTA %.2f" % (n / 256 / 444), "ms %.4f" % t, "afterstates/s %.3e" % (float(c.sum()) / (t * 1e-3)))
    del f, env
