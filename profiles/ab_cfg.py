"""A/B of tile configurations on one GPU: K1 (tb_afterstates) and K3 (tb_rollout greedy) device times at 2^20 envs on
greedy-play boards, per configuration (tb_set_tuning), L2 flushed between launches.

    python profiles/ab_cfg.py [--k1 0,6,2] [--k3 0,6] [--envs N]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tetris_b200 import BatchedTetris, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--k1", default="0,6,2")
ap.add_argument("--k3", default="0,6")
ap.add_argument("--k2", default="0")
ap.add_argument("--k3r", default="0")
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--board", default="10x20")
a = ap.parse_args()
C, R = (int(x) for x in a.board.split("x"))
n = a.envs
env = BatchedTetris(C, R, n, piece_set=1, seed=0x5EED)
env.rollout(30, "random")
env.rollout(64, "greedy")
saved = env.state.clone()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
feats = torch.empty((n, env.a_max, 8), dtype=torch.float32, device="cuda")
valid = torch.empty(n, dtype=torch.int64, device="cuda")
count = torch.empty(n, dtype=torch.int32, device="cuda")
out = {"envs": n, "board": a.board}


def timed(fn, reps):
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


for cfg in [int(x) for x in a.k1.split(",") if x != ""]:
    _lib.set_tuning("k1_cfg", cfg)
    for _ in range(3):
        env.get_after_states(out=(feats, valid, count))
    ms = timed(lambda: env.get_after_states(out=(feats, valid, count)), 9)
    out["k1_cfg%d_ms" % cfg] = ms
    out["k1_cfg%d_checksum" % cfg] = int(count.sum().item())
_lib.set_tuning("k1_cfg", -1)
try:                                            # compact int16 output (TB_FLAG_FEATS_I16), default configuration
    feats16 = torch.empty((n, env.a_max, 8), dtype=torch.int16, device="cuda")
    for _ in range(3):
        env.get_after_states(out=(feats16, valid, count), compact=True)
    out["k1_i16_ms"] = timed(lambda: env.get_after_states(out=(feats16, valid, count), compact=True), 9)
except TypeError:
    pass
for cfg in [int(x) for x in a.k3.split(",") if x != ""]:
    _lib.set_tuning("k3_cfg", cfg)
    env.state.copy_(saved)
    env.rollout(32, "greedy")
    ms = timed(lambda: env.rollout(32, "greedy"), 7)
    out["k3_cfg%d_ms_per_32" % cfg] = ms
    out["k3_cfg%d_placements_per_s" % cfg] = n * 32 / (ms * 1e-3)
_lib.set_tuning("k3_cfg", -1)
def tune(name, v):
    try:
        _lib.set_tuning(name, v)
        return True
    except Exception:                       # older library in an A/B run: no such knob
        return False


a0 = torch.zeros(n, dtype=torch.int32, device="cuda")
envr = BatchedTetris(C, R, n, piece_set=1, seed=7)
envr.rollout(30, "random")
saved_r = envr.state.clone()
for cfg in [int(x) for x in a.k2.split(",") if x != ""]:
    if not tune("k2_cfg", cfg) and cfg != 0:
        continue
    for name, ev, sv_ in (("", env, saved), ("_tall", envr, saved_r)):   # greedy-play boards; tall (random-play) boards
        ts = []
        for rep in range(7):
            ev.state.copy_(sv_)                  # the same boards every repetition (a step changes them)
            flush.zero_()
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record(); ev.step(a0, auto_reset=True, check=False); e_.record()
            torch.cuda.synchronize()
            if rep >= 2:
                ts.append(s_.elapsed_time(e_))
        ts.sort()
        out["k2_cfg%d%s_ms" % (cfg, name)] = ts[len(ts) // 2]
    envr.state.copy_(saved_r)
tune("k2_cfg", -1)
for cfg in [int(x) for x in a.k3r.split(",") if x != ""]:
    if not tune("k3r_cfg", cfg) and cfg != 0:
        continue
    envr.state.copy_(saved_r)
    envr.rollout(32, "random")
    out["k3r_cfg%d_ms_per_32" % cfg] = timed(lambda: envr.rollout(32, "random"), 5)
tune("k3r_cfg", -1)
print(json.dumps(out))
