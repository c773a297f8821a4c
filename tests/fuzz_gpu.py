"""Manual stress run (not collected by pytest): many seeds x shapes x piece sets, fused rollouts interleaved with
afterstate / step comparisons against the oracle.  python tests/fuzz_gpu.py [n_rounds]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import oracle as orc
from tetris_b200 import BatchedTetris


def one(Cc, R, ps, seed, n):
    rng = np.random.default_rng(seed)
    off = int(rng.integers(0, 1 << 40))
    dirs = rng.choice([-1.0, 1.0], size=8).astype(np.float32) if seed % 3 == 0 else None
    env = BatchedTetris(Cc, R, n, piece_set=ps, seed=seed, env_offset=off, feature_directions=dirs)
    ob = orc.Batch(Cc, R, n, piece_set=ps, seed=seed, env_offset=off)
    ob.reset()
    d = np.ones(8, np.float32) if dirs is None else dirs
    for rnd in range(4):
        tg, tr = int(rng.integers(1, 30)), int(rng.integers(1, 30))
        w = (orc.BCTS_WEIGHTS * rng.uniform(0.5, 1.5, 8).astype(np.float32)) if seed % 2 else orc.BCTS_WEIGHTS
        env.rollout(tg, "greedy", w); ob.rollout(tg, 1, w, threads=8)
        env.rollout(tr, "random"); ob.rollout(tr, 0, threads=8)
        assert np.array_equal(env.rows(), ob.rows()), ("rows", Cc, R, ps, seed, rnd)
        assert np.array_equal(env.piece, ob.piece)
        inc = bool(rnd & 1)
        feats, valid, count = env.get_after_states(include_terminal=inc)
        of, ov, oc, on = ob.afterstates()
        assert np.array_equal(valid.cpu().numpy().view(np.uint64), ov) and np.array_equal(count.cpu().numpy(), oc)
        mask = np.arange(env.a_max)[None, :] < on[:, None]
        if not inc:
            mask &= ((ov[:, None] >> np.arange(env.a_max, dtype=np.uint64)) & np.uint64(1)).astype(bool)
        assert np.array_equal(feats.cpu().numpy()[mask], (of * d)[mask]), ("feats", Cc, R, ps, seed, rnd)
        a = (rng.integers(0, 1 << 30, n) % np.maximum(oc, 1)).astype(np.int32)
        obs, rew, done, lines = env.step(a, auto_reset=True)
        oobs, orew, odone, olines = ob.step(a, auto_reset=True)
        assert np.array_equal(obs.cpu().numpy(), oobs * d) and np.array_equal(rew.cpu().numpy(), orew)
        assert np.array_equal(done.cpu().numpy(), odone)
    return n * 4


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    t0, tot = time.time(), 0
    shapes = [(10, 20), (10, 10), (6, 12), (8, 16), (4, 4)]
    for i in range(rounds):
        Cc, R = shapes[i % len(shapes)]
        tot += one(Cc, R, (i // len(shapes)) % 2, 1000 + i, [257, 1000, 3001, 513][i % 4])
    print("fuzz ok: %d rounds, %d env-rounds, %.1f s" % (rounds, tot, time.time() - t0))


if __name__ == "__main__":
    main()
