"""Pin the C oracle (oracle/tetris_oracle.c) to fixtures recorded from the live Python
reference (tests/golden/make_golden.py).  CPU only."""
import hashlib

import numpy as np
import pytest

from golden_util import TRACES, feat2, load, rep_to_rows, replay_rollout_config, rollout_configs, rows_to_rep
from oracle import oracle as orc


def test_afterstate_counts():
    # A per piece at C=10 and C=6
    assert [orc.num_afterstates(p, 10) for p in range(9)] == [17, 34, 34, 9, 17, 17, 34, 36, 18]
    assert [orc.num_afterstates(p, 6) for p in range(9)] == [9, 18, 18, 5, 9, 9, 18, 20, 10]


def test_empty_board_features():
    # empty 10x20 board
    assert orc.reset_state_features(10, 20).tolist() == [0, 10, 0, 1, 0, 40, 0, 0]


@pytest.mark.parametrize("name,at_least", [("afterstates", 30000), ("afterstates_dense", 80000)])
def test_afterstates_fixture(name, at_least):
    """`afterstates`: played and arbitrary boards of 10x20 / 10x10 / 6x12.  `afterstates_dense`: near-full stacks of
    those three plus 8x16, 4x4, 16x27, 12x24 and 7x9 -- over a hundred four-line clears, about a thousand placements that
    poke above row R and are rescued by the clear, hundreds that clear and are terminal all the same (tests/golden/make_golden.py prints the census)."""
    g = load(name)
    n_checked = 0
    for i in range(len(g["piece"])):
        C, R = (int(x) for x in g["shape"][i])
        N = R + 4
        rep = rows_to_rep(g["rows"][i][:N], C)
        out = orc.afterstates(C, R, int(g["piece"][i]), rep, g["heights"][i][:C].astype(np.int32))
        s, n = int(g["start"][i]), int(g["count"][i])
        assert out["n"] == n
        sl = slice(s, s + n)
        assert np.array_equal(feat2(out["features"]), g["a_feat2"][sl])
        assert np.array_equal(out["terminal"], g["a_terminal"][sl])
        assert np.array_equal(out["n_cleared"], g["a_n_cleared"][sl])
        assert np.array_equal(rep_to_rows(out["rep"]), g["a_rows"][sl][:, :N])
        assert np.array_equal(out["heights"], g["a_heights"][sl][:, :C])
        assert np.array_equal(out["anchor_col"], g["a_anchor"][sl][:, 0])
        assert np.array_equal(out["anchor_row"], g["a_anchor"][sl][:, 1])
        assert np.array_equal(out["is_full"], g["a_is_full"][sl])
        n_checked += n
    assert n_checked == len(g["a_terminal"]) > at_least
    if name == "afterstates_dense":
        ncl = g["a_n_cleared"]
        assert (ncl == 4).sum() >= 50 and (ncl == 3).sum() >= 50 and (g["a_terminal"] & (ncl > 0)).sum() >= 50
        assert {tuple(x) for x in g["shape"].tolist()} == {(10, 20), (10, 10), (6, 12), (8, 16), (4, 4), (16, 27), (12, 24), (7, 9)}


def replay_trace(g, make_batch, rng_mode):
    """Drive an engine with the fixture's piece/action tapes and compare after every step.

    make_batch(C, R, n_env, piece_set, seed) must return an object with the oracle.Batch API.
    rng_mode: pieces come from the engine's own bag RNG (must reproduce the tape) instead of the tape.
    """
    C, R, ps, seed = int(g["C"]), int(g["R"]), int(g["piece_set"]), int(g["seed"])
    T, n = g["action"].shape
    b = make_batch(C, R, n, ps, seed)
    b.reset(None if rng_mode else g["piece"][0].astype(np.uint8))
    for t in range(T):
        assert np.array_equal(b.piece, g["piece"][t]), t
        feats, valid, count, n_all = b.afterstates()
        assert np.array_equal(count, g["n_valid"][t]), t
        assert np.array_equal(n_all, g["n_all"][t]), t
        a_max = g["feat2"].shape[2]
        vbits = ((valid[:, None] >> np.arange(a_max, dtype=np.uint64)) & np.uint64(1)).astype(bool)
        assert np.array_equal(vbits, g["valid"][t]), t
        for e in range(n):
            na = int(n_all[e])
            assert np.array_equal(feat2(feats[e, :na]), g["feat2"][t, e, :na]), (t, e)
        if str(g["policy"]) == "greedy":
            # the recorded action must be the first arg-max of the float32 fitness over valid afterstates
            for e in range(n):
                fv = [orc.fitness(feats[e, s], orc.BCTS_WEIGHTS) for s in range(int(n_all[e])) if vbits[e, s]]
                assert int(np.argmax(np.array(fv, np.float32))) == int(g["action"][t, e]), (t, e)
        tape = None if rng_mode else g["next_piece"][t].astype(np.uint8)
        obs, reward, done, lines = b.step(g["action"][t].astype(np.int32), tape=tape)
        assert np.array_equal(feat2(obs), g["obs2"][t]), t
        assert np.array_equal(reward, g["reward"][t]), t
        assert np.array_equal(done, g["done"][t]), t
        assert np.array_equal(lines, g["lines"][t]), t
        assert np.array_equal(b.rows(), g["rows"][t]), t
        assert np.array_equal(b.heights, g["heights"][t]), t
        assert np.array_equal(b.piece, g["next_piece"][t]), t
        if done.any():
            rp = g["reset_piece"][t]
            b.reset_masked(done, None if rng_mode else np.where(rp >= 0, rp, 0).astype(np.uint8))
            assert np.array_equal(b.piece[done], rp[done]), t


def _oracle_batch(C, R, n, ps, seed):
    return orc.Batch(C, R, n, piece_set=ps, seed=seed)


@pytest.mark.parametrize("name", TRACES)
def test_trace_tape(name):
    replay_trace(load("trace_" + name), _oracle_batch, rng_mode=False)


@pytest.mark.parametrize("name", TRACES)
def test_trace_rng(name):
    """The oracle's own per-env bag RNG reproduces the recorded piece tape."""
    replay_trace(load("trace_" + name), _oracle_batch, rng_mode=True)


def test_known_answer():
    """SURVEY.md Appendix C.3 (reference sampler, np.random.seed(0)) replayed from the piece tape."""
    g = load("known_answer")
    d = np.array([-1, -1, -1, -1, -1, -1, 1, -1], np.float64)
    b = orc.Batch(10, 10, 1, piece_set=0)
    b.reset(np.array([g["first_piece"]], np.uint8))
    hsh = hashlib.sha256()
    total = 0
    for t in range(len(g["action"])):
        feats, valid, count, n_all = b.afterstates()
        vf = np.array([feats[0, s] for s in range(int(n_all[0])) if (int(valid[0]) >> s) & 1], np.float32) * d
        assert len(vf) == g["n_valid"][t]
        i = int(np.argmax(vf.sum(axis=1)))
        assert i == g["action"][t]
        obs, reward, done, lines = b.step(np.array([i], np.int32), tape=np.array([g["next_piece"][t]], np.uint8))
        o = obs[0] * d
        assert np.array_equal(o, g["obs"][t]) and np.array_equal(np.signbit(o), np.signbit(g["obs"][t]))
        assert reward[0] == g["reward"][t] and lines[0] == g["lines"][t] and bool(done[0]) == bool(g["done"][t])
        hsh.update(b.rep[0].astype(np.uint8).tobytes())
        hsh.update(np.asarray(o, np.float64).tobytes())
        total += int(reward[0])
        if done[0]:
            b.reset_masked(done, np.array([g["reset_piece"][t]], np.uint8))
    assert total == -71
    assert hsh.hexdigest() == str(g["sha256"])
    assert np.array_equal(b.rows()[0], g["final_rows"])
    assert b.rows()[0][:2].tolist() == [807, 519]


def test_fitness_fixture():
    g = load("fitness")
    C, R = int(g["C"]), int(g["R"])
    for i in range(len(g["piece"])):
        out = orc.afterstates(C, R, int(g["piece"][i]), rows_to_rep(g["rows"][i], C))
        s, n = int(g["start"][i]), int(g["count"][i])
        fv = np.array([orc.fitness(f, orc.BCTS_WEIGHTS) for f in out["features"]], np.float32)
        assert np.array_equal(fv, g["fitness"][s:s + n])          # bit-exact float32
        pol = (fv == fv.max()).astype(float)
        assert np.array_equal(pol / pol.sum(), g["best_policy"][s:s + n])


class _OracleRolloutBatch:
    def __init__(self, C, R, n, ps):
        self.b = orc.Batch(C, R, n, piece_set=ps, seed=0)
        self.b.reset()                                     # one draw per env, like a freshly constructed game
        self.C = C

    def load(self, rows, piece):
        rep = rows_to_rep(rows, self.C)
        self.b.rep[:] = rep
        self.b.heights[:] = [orc.calc_lowest_free_rows(r) for r in rep]
        self.b.piece[:] = piece

    def rollout_values(self, *a, **kw):
        return self.b.rollout_values(*a, **kw)


def test_rollouts_fixture():
    """SURVEY 8f-1 pinned to the reference: Tetris.single_rollout (game.py:129-148) for every legal action and fork
    of recorded parent states, with the forks' recorded piece tapes and a deterministic policy (greedy BCTS, or the
    per-fork counter RNG): sums of returns per enumeration slot and the legal-slot masks are the reference's."""
    g = load("rollouts")
    n_ret = 0
    for c in rollout_configs(g):
        ret, bits = replay_rollout_config(c, _OracleRolloutBatch)
        assert np.array_equal(bits, c["valid"]), (c["C"], c["R"], c["policy"])
        assert np.array_equal(ret, c["ret_sum"]), (c["C"], c["R"], c["policy"])
        n_ret += int(bits.sum()) * c["n_forks"]
    assert n_ret > 2000
