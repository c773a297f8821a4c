"""CPU check of the bit-parallel core (tetris_b200/csrc/tb_core.cuh, compiled with g++ by
tests/hostcheck) against the cell-by-cell oracle: every piece x slot on random boards, fast
(incremental) and slow (from-scratch) evaluation paths, transposes, RNG/bag, fitness."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from golden_util import load, rep_to_rows, rows_to_rep
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
SO = os.path.join(HERE, "hostcheck", "libhostcheck.so")
CORE = os.path.join(os.path.dirname(HERE), "tetris_b200", "csrc", "tb_core.cuh")
SHAPES = [(10, 20), (10, 10), (6, 12), (8, 16), (4, 4), (16, 27), (12, 24), (7, 9)]


@pytest.fixture(scope="module")
def hc():
    if (not os.path.exists(SO)) or max(os.path.getmtime(SRC), os.path.getmtime(CORE)) > os.path.getmtime(SO):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                               "-fno-strict-aliasing", "-Wno-unknown-pragmas", "-o", SO, SRC])
    L = C.CDLL(SO)
    L.hc_afterstates.restype = C.c_int
    L.hc_afterstates.argtypes = [C.c_int] * 3 + [C.c_void_p, C.c_int] + [C.c_void_p] * 6
    L.hc_transpose.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 3
    L.hc_rng.restype = C.c_uint32
    L.hc_rng.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
    L.hc_bag_draw.restype = C.c_int
    L.hc_bag_draw.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    L.hc_fitness.restype = C.c_float
    L.hc_fitness.argtypes = [C.c_void_p, C.c_void_p]
    L.hc_num_slots.restype = C.c_int
    return L


def hc_afterstates(L, Cc, R, piece, rows, mode):
    N = R + 4
    A = 64
    feats = np.zeros((A, 8), np.float32)
    term = np.zeros(A, np.uint8)
    ncl = np.zeros(A, np.int32)
    rows_out = np.zeros((A, N), np.uint16)
    anchor = np.zeros(A, np.int32)
    fast = np.zeros(A, np.uint8)
    rows = np.ascontiguousarray(rows, np.uint16)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    n = L.hc_afterstates(Cc, R, piece, p(rows), mode, p(feats), p(term), p(ncl), p(rows_out), p(anchor), p(fast))
    assert n >= 0, "hostcheck failed (%d)" % n
    return n, feats[:n], term[:n].astype(bool), ncl[:n], rows_out[:n], anchor[:n], fast[:n].astype(bool)


def random_board(rng, Cc, R, fill):
    N = R + 4
    rep = np.zeros((N, Cc), np.uint8)
    hmax = rng.integers(0, R + 1)
    for c in range(Cc):
        h = rng.integers(0, hmax + 1)
        if h:
            rep[:h, c] = rng.random(h) < fill
            rep[h - 1, c] = 1
    return rep


def tall_dense_board(rng, Cc, R):
    """Dense stack reaching the top rows but for a 1-4 column gap: the top rows are near-full, so placements that poke
    above row R are legal exactly when they complete enough of them (the rescue path of valid_slots)."""
    rep = np.zeros((R + 4, Cc), np.uint8)
    top = int(rng.integers(max(R - 4, 1), R + 1))
    rep[:top] = 1
    gw = int(rng.integers(1, min(4, Cc - 1) + 1))
    g0 = int(rng.integers(0, Cc - gw + 1))
    for c in range(g0, g0 + gw):
        rep[int(rng.integers(max(top - 5, 0), top)):, c] = 0
    for _ in range(int(rng.integers(0, 4))):                 # a few holes below the surface (never a full row: the gap)
        r, c = int(rng.integers(0, top)), int(rng.integers(0, Cc))
        if rep[r + 1:, c].any():
            rep[r, c] = 0
    return rep


def played_boards(Cc, R, n_env, steps, seed):
    b = orc.Batch(Cc, R, n_env, piece_set=1, seed=seed)
    b.reset()
    out = []
    for _ in range(steps):
        b.rollout(3, 0)
        out.append(b.rep.copy())
    return np.concatenate(out)


def check_board(L, Cc, R, rep, stats):
    rows = rep_to_rows(rep)
    for piece in range(9):
        ref = orc.afterstates(Cc, R, piece, rep)
        for mode in (0, 1):
            n, feats, term, ncl, rows_out, anchor, fast = hc_afterstates(L, Cc, R, piece, rows, mode)
            assert n == ref["n"]
            assert np.array_equal(feats, ref["features"]), (piece, mode, rep_to_rows(rep).tolist())
            assert np.array_equal(term, ref["terminal"])
            assert np.array_equal(ncl, ref["n_cleared"])
            assert np.array_equal(anchor, ref["anchor_row"])
            assert np.array_equal(rows_out, rep_to_rows(ref["rep"]))
            if mode == 0:
                stats[0] += n
                stats[1] += int(fast.sum())


@pytest.mark.parametrize("shape", SHAPES)
def test_core_vs_oracle(hc, shape):
    Cc, R = shape
    rng = np.random.default_rng(100 + Cc * 31 + R)
    stats = [0, 0]
    boards = [np.zeros((R + 4, Cc), np.uint8)]
    for fill in (0.4, 0.7, 0.9, 0.97, 1.0):
        boards += [random_board(rng, Cc, R, fill) for _ in range(40)]
    boards += list(played_boards(Cc, R, 24, 12, seed=5))
    boards += [tall_dense_board(rng, Cc, R) for _ in range(120)]
    for rep in boards:
        check_board(hc, Cc, R, rep, stats)
    assert stats[0] > 5000
    assert stats[1] > 0.5 * stats[0]          # the incremental path actually carries most placements


@pytest.mark.parametrize("name,stride", [("afterstates", 3), ("afterstates_dense", 2)])
def test_core_vs_golden(hc, name, stride):
    """Straight against the reference-generated fixtures as well (the dense one: near-full stacks, multi-line clears,
    overflow rescued by a clear, on five shapes)."""
    g = load(name)
    for i in range(0, len(g["piece"]), stride):
        Cc, R = (int(x) for x in g["shape"][i])
        N = R + 4
        n, feats, term, ncl, rows_out, anchor, fast = hc_afterstates(hc, Cc, R, int(g["piece"][i]), g["rows"][i][:N], 0)
        s = int(g["start"][i])
        sl = slice(s, s + n)
        assert n == g["count"][i]
        assert np.array_equal(np.rint(feats * 2).astype(np.int64), g["a_feat2"][sl])
        assert np.array_equal(term, g["a_terminal"][sl])
        assert np.array_equal(rows_out, g["a_rows"][sl][:, :N])


@pytest.mark.parametrize("shape", SHAPES)
def test_transpose(hc, shape):
    Cc, R = shape
    N = R + 4
    rng = np.random.default_rng(3)
    for _ in range(200):
        rep = (rng.random((N, Cc)) < 0.5).astype(np.uint8)
        rows = rep_to_rows(rep)
        back = np.zeros(N, np.uint16)
        cols = np.zeros(Cc, np.uint32)
        hc.hc_transpose(Cc, R, rows.ctypes.data_as(C.c_void_p), back.ctypes.data_as(C.c_void_p),
                        cols.ctypes.data_as(C.c_void_p))
        assert np.array_equal(back, rows)
        want = (rep.astype(np.uint32) * (1 << np.arange(N, dtype=np.uint32))[:, None]).sum(axis=0)
        assert np.array_equal(cols, want.astype(np.uint32))


def test_rng_and_bag(hc):
    rng = np.random.default_rng(0)
    for _ in range(2000):
        seed, env = int(rng.integers(0, 2 ** 63)), int(rng.integers(0, 2 ** 40))
        ctr, stream = int(rng.integers(0, 2 ** 32)), int(rng.integers(0, 4))
        assert hc.hc_rng(seed, env, ctr, stream) == orc.rng(seed, env, ctr, stream)
    # bag draws: every 7 consecutive draws are a permutation, and match the oracle's batch draws
    b = orc.Batch(10, 20, 5, piece_set=1, seed=99, env_offset=1000)
    b.reset()
    for e in range(5):
        bag, draws = C.c_uint32(0), C.c_uint32(0)
        seq = [hc.hc_bag_draw(7, 99, 1000 + e, C.byref(bag), C.byref(draws)) for _ in range(21)]
        assert sorted(seq[:7]) == sorted(seq[7:14]) == sorted(seq[14:]) == list(range(7))
        assert seq[0] == b.piece[e]


def test_fitness(hc):
    g = load("fitness")
    f = (g["feat2"].astype(np.float32) * np.float32(0.5))
    w = orc.BCTS_WEIGHTS
    for i in range(len(f)):
        fi = np.ascontiguousarray(f[i])
        got = np.float32(hc.hc_fitness(fi.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p)))
        assert got == g["fitness"][i]


def test_num_slots(hc):
    for Cc in (4, 6, 7, 8, 10, 12, 16):
        for p in range(9):
            assert hc.hc_num_slots(p, Cc) == orc.num_afterstates(p, Cc)


def test_compile_time_table_images(hc):
    """The constexpr images of the run table and of the decoded orientation descriptors (copied into shared memory
    by the tile kernels) equal run_tab_entry / decode_ori word for word, for every compiled board height."""
    hc.hc_table_images.restype = C.c_int
    assert hc.hc_table_images() == 0


def test_bit_helpers(hc):
    """interleave16, the byte-permute selectors of the legality test (poke_sel), piece_cells4 and hole_depth_of against
    plain restatements (the kernels' results are compared with the oracle elsewhere; this pins the helpers themselves)."""
    hc.hc_bit_helpers.restype = C.c_int
    assert hc.hc_bit_helpers() == 0
