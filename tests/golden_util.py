"""Helpers shared by the golden-fixture tests (oracle on CPU, CUDA path on the GPU box)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRACES = ("7p_10x20_random", "2p_10x10_random_dir", "7p_6x12_random", "7p_10x20_greedy", "7p_6x12_greedy")


def load(name):
    """Fixture as a plain dict (NpzFile would re-decompress on every access)."""
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def rows_to_rep(rows, C):
    """u16 row masks [..., N] -> 0/1 cells [..., N, C]."""
    rows = np.asarray(rows, np.uint16)
    return ((rows[..., None] >> np.arange(C, dtype=np.uint16)) & 1).astype(np.uint8)


def rep_to_rows(rep):
    rep = np.asarray(rep)
    w = (1 << np.arange(rep.shape[-1])).astype(np.uint32)
    return (rep.astype(np.uint32) * w).sum(axis=-1).astype(np.uint16)


def feat2(f):
    """Features doubled -> exact integers (the fixtures store them this way)."""
    v = np.asarray(f, np.float64) * 2
    r = np.rint(v)
    assert np.array_equal(r, v), "feature is not a half-integer"
    return r.astype(np.int64)


def rollout_configs(g):
    """The part-A configs of rollouts.npz as dicts (tests/golden/make_golden.py gen_rollouts)."""
    out = []
    for k in range(int(g["n_configs"])):
        pre = "c%d_" % k
        c = {key[len(pre):]: g[key] for key in g if key.startswith(pre)}
        for key in ("C", "R", "piece_set", "length", "n_forks", "seed2"):
            c[key] = int(c[key])
        c["policy"] = str(c["policy"])
        out.append(c)
    return out


def replay_rollout_config(c, make_batch):
    """Replay one config through an engine with the oracle.Batch API (rep / heights / piece views or import, and
    rollout_values(..., piece_tape=)): returns (ret_sum, valid bits) to compare with the fixture."""
    C, R = c["C"], c["R"]
    P = len(c["piece"])
    b = make_batch(C, R, P, c["piece_set"])
    b.load(c["rows"][:, :R + 4], c["piece"])
    ret, valid = b.rollout_values(c["length"], c["n_forks"], 1 if c["policy"] == "greedy" else 0, seed2=c["seed2"],
                                  piece_tape=c["tape"])
    a_max = c["valid"].shape[1]
    bits = ((np.asarray(valid, np.uint64)[:, None] >> np.arange(a_max, dtype=np.uint64)) & np.uint64(1)).astype(bool)
    return np.asarray(ret), bits
