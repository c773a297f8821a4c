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
