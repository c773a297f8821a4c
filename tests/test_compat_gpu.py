"""GPU: the reference-facing single-game API (tetris.game.Tetris, tetris.state.State, tetris.tetromino.*) driven the
way a policy written against the reference drives it, checked against fixtures recorded from the live reference."""
import hashlib
import os
import sys

import numpy as np
import pytest

from golden_util import feat2, load, rep_to_rows, rows_to_rep

pytestmark = pytest.mark.gpu
DIRS = np.array([-1, -1, -1, -1, -1, -1, 1, -1])


def _pid(piece):
    return piece.piece_id


class TapeSampler:
    """Replaces env.tetromino_sampler (a plain attribute, game.py:50) to replay a recorded piece tape."""

    def __init__(self, pieces, tape):
        self.pieces, self.tape = pieces, iter(tape)

    def next_tetromino(self):
        return self.pieces[int(next(self.tape))]


def test_known_answer_seed0():
    """np.random.seed(0), Tetris(10, 10, directions), corrected example_play flow."""
    from tetris.game import Tetris
    g = load("known_answer")
    np.random.seed(0)
    env = Tetris(10, 10, feature_directions=DIRS)
    env.reset()
    hsh, total = hashlib.sha256(), 0
    for t in range(100):
        assert _pid(env.current_tetromino) == g["piece"][t]
        feats, none = env.get_after_states()
        assert none is None and feats.dtype == np.float64 and feats.shape == (g["n_valid"][t], 8)
        i = int(np.argmax(feats.sum(axis=1)))
        assert i == g["action"][t]
        obs, rew, done, lines = env.step(i)
        assert obs.dtype == np.float64 and isinstance(rew, int) and isinstance(done, bool) and isinstance(lines, int)
        assert np.array_equal(obs, g["obs"][t]) and np.array_equal(np.signbit(obs), np.signbit(g["obs"][t]))
        assert (rew, lines, done) == (g["reward"][t], g["lines"][t], bool(g["done"][t]))
        assert env.current_state.representation.dtype == np.int64
        hsh.update(env.current_state.representation.astype(np.uint8).tobytes())
        hsh.update(np.asarray(obs, np.float64).tobytes())
        total += rew
        if done:
            env.reset()
    assert total == -71 and hsh.hexdigest() == str(g["sha256"])
    assert np.array_equal(rep_to_rows(env.current_state.representation), g["final_rows"])
    assert np.array_equal(env.current_state.lowest_free_rows, g["final_heights"])


@pytest.mark.parametrize("name,envs", [("7p_10x20_random", (0, 7)), ("2p_10x10_random_dir", (3,)), ("7p_6x12_greedy", (1,))])
def test_trace_through_tetris_class(name, envs):
    from tetris import tetromino
    from tetris.game import Tetris
    g = load("trace_" + name)
    C, R = int(g["C"]), int(g["R"])
    dirs = g["directions"] if len(g["directions"]) else None
    T = min(g["action"].shape[0], 60)
    for e in envs:
        tape = [g["piece"][0, e]]
        for t in range(g["action"].shape[0]):
            tape.append(g["next_piece"][t, e])
            if g["done"][t, e]:
                tape.append(g["reset_piece"][t, e])
        env = Tetris(C, R, feature_directions=dirs)
        env.tetrominos = [cls("bcts", 8, C) for cls in tetromino.PIECE_CLASSES]      # indexable by global id
        env.tetromino_sampler = TapeSampler(env.tetrominos, tape)
        env.reset()
        for t in range(T):
            assert _pid(env.current_tetromino) == g["piece"][t, e]
            feats, allf = env.get_after_states(include_terminal=True)
            na = int(g["n_all"][t, e])
            assert len(feats) == g["n_valid"][t, e] and len(allf) == na
            und = allf * dirs if dirs is not None else allf
            assert np.array_equal(feat2(und), g["feat2"][t, e, :na])
            assert np.array_equal(feat2(feats * dirs if dirs is not None else feats), g["feat2"][t, e, :na][g["valid"][t, e, :na]])
            obs, rew, done, lines = env.step(int(g["action"][t, e]))
            assert np.array_equal(feat2(obs * dirs if dirs is not None else obs), g["obs2"][t, e])
            assert obs.dtype == (np.float64 if dirs is not None else np.float32)
            assert (rew, done, lines) == (g["reward"][t, e], bool(g["done"][t, e]), g["lines"][t, e])
            assert np.array_equal(rep_to_rows(env.current_state.representation), g["rows"][t, e])
            assert np.array_equal(env.current_state.lowest_free_rows, g["heights"][t, e])
            if done:
                env.reset()


def test_piece_get_after_states_state_objects():
    """tetromino.X.get_after_states(State) -> [State]: every public State attribute against the fixture."""
    from tetris import state, tetromino
    g = load("afterstates")
    idx = np.arange(0, len(g["piece"]), 7)
    for i in idx:
        C, R = (int(x) for x in g["shape"][i])
        N = R + 4
        base = state.State(rows_to_rep(g["rows"][i][:N], C).astype(np.int64))
        assert np.array_equal(base.lowest_free_rows, g["heights"][i][:C]) and base.lowest_free_rows.dtype == np.int64
        piece = tetromino.PIECE_CLASSES[int(g["piece"][i])]("bcts", 8, C)
        kids = piece.get_after_states(base)
        s, n = int(g["start"][i]), int(g["count"][i])
        assert len(kids) == n
        for k, ch in enumerate(kids):
            j = s + k
            assert ch.representation.shape == (N, C) and ch.representation.dtype == np.int64
            assert np.array_equal(rep_to_rows(ch.representation), g["a_rows"][j][:N])
            assert np.array_equal(ch.lowest_free_rows, g["a_heights"][j][:C])
            assert (ch.anchor_col, ch.anchor_row) == tuple(g["a_anchor"][j])
            assert ch.n_cleared_lines == g["a_n_cleared"][j] and ch.terminal_state == bool(g["a_terminal"][j])
            assert ch.reward == (0 if ch.terminal_state else ch.n_cleared_lines)
            nchg = len(ch.cleared_rows_relative_to_anchor)
            assert np.array_equal(ch.cleared_rows_relative_to_anchor, g["a_is_full"][j][:nchg])
            assert not g["a_is_full"][j][nchg:].any()
            assert ch.num_rows == N and ch.num_columns == C and ch.n_legal_rows == R and ch.value_estimate == 0.0
            assert ch.features is None                               # lazy cache like the reference
            f = ch.get_features()
            assert f.dtype == np.float32 and ch.features is f and ch.get_features() is f
            assert np.array_equal(feat2(f), g["a_feat2"][j])
            fd = ch.get_features(direct_by=DIRS)
            assert fd.dtype == np.float64 and np.array_equal(fd, f * DIRS)


def test_state_constructor_and_helpers():
    from oracle import oracle as orc
    from tetris import state
    rng = np.random.default_rng(9)
    for (C, R) in ((10, 20), (10, 10), (6, 12)):
        N = R + 4
        for k in range(12):
            rep = (rng.random((N, C)) < rng.random()).astype(np.int64)
            rep[R:] = 0
            if k % 3 == 0:
                rep[0] = 1                                            # a full bottom row: cleared by the constructor
            o = orc.board_features(C, R, rep)
            s = state.State(rep.copy())
            assert np.array_equal(s.get_features(), o["features"])
            assert np.array_equal(s.lowest_free_rows, o["heights"]) and np.array_equal(s.representation, o["rep"])
            assert s.n_cleared_lines == o["n_cleared"] and s.terminal_state == o["terminal"]
            assert s.anchor_row == 0 and s.anchor_col == 0 and s.landing_height_bonus == 0.0
            assert np.array_equal(state.calc_lowest_free_rows(o["rep"]), o["heights"])
            six = state.get_feature_values_jitted(s.lowest_free_rows, s.representation, R, C)
            assert [float(v) for v in six] == [float(o["features"][i]) for i in (0, 1, 2, 4, 5, 7)] or s.n_cleared_lines
            is_full, ncl, rep2, h2 = state.clear_lines_jitted(np.arange(0, 1), rep.copy(), None, C)
            assert ncl == o["n_cleared"] and np.array_equal(rep2, o["rep"]) and np.array_equal(h2, o["heights"])
    s = state.State(np.zeros((24, 10), np.int64))
    assert s.get_features().tolist() == [0, 10, 0, 1, 0, 40, 0, 0]
    assert str(s.get_features(direct_by=DIRS)[0]) == "-0.0"
    assert repr(s).count("\n") == 21 and len(s.print_board_to_string().split("\n")[1]) == 22
    with pytest.raises(ValueError):
        state.State(np.zeros((24, 10), np.int64), feature_type="other").get_features()   # state.py:95


def test_errors_like_the_reference():
    from tetris.game import Tetris
    env = Tetris(10, 10)
    with pytest.raises(AttributeError):
        env.step(0)                                                   # step before get_after_states (game.py:83)
    feats, _ = env.get_after_states()
    with pytest.raises(IndexError):
        env.step(len(feats))
    with pytest.raises(ValueError):
        Tetris(10, 10, feature_type="other").get_after_states()


def test_fitness_and_best_policy():
    from tetris import tetromino
    from tetris.game import Tetris
    g = load("fitness")
    C, R = int(g["C"]), int(g["R"])
    env = Tetris(C, R)
    env.tetrominos = tetromino.standard_set(C)
    for i in range(0, len(g["piece"]), 5):
        from tetris import state
        env.current_state = state.State(rows_to_rep(g["rows"][i], C).astype(np.int64))
        env.current_tetromino = env.tetrominos[int(g["piece"][i])]
        s, n = int(g["start"][i]), int(g["count"][i])
        kids = env.current_tetromino.get_after_states(env.current_state)
        fv = np.array([env.fitness(k) for k in kids])
        assert fv.dtype == np.float32 and np.array_equal(fv, g["fitness"][s:s + n])
        assert np.array_equal(env.get_best_policy(), g["best_policy"][s:s + n])


class ListSampler:
    def __init__(self, pieces):
        self.pieces, self.tape, self.pos = pieces, [], 0

    def load(self, tape):
        self.tape, self.pos = [int(x) for x in tape], 0

    def next_tetromino(self):
        p = self.pieces[self.tape[self.pos]]
        self.pos += 1
        return p


def _greedy_policy_function(state, feats):
    w = np.array([-24.04, -19.77, -13.08, -12.63, -10.49, -9.22, 6.6, -1.61], np.float32)
    f = np.asarray(feats).astype(np.float32)
    acc = f[:, 0] * w[0]
    for i in range(1, 8):
        acc = acc + f[:, i] * w[i]
    return int(np.argmax(acc))


def test_single_rollout_vs_reference():
    """Tetris.single_rollout through the compatibility class == the reference's (tests/golden/rollouts.npz part A,
    greedy configs): same parents, actions, per-fork piece tapes and policy_function -> same returns."""
    from golden_util import rollout_configs
    from tetris import state, tetromino
    from tetris.game import Tetris
    g = load("rollouts")
    checked = 0
    for c in rollout_configs(g):
        if c["policy"] != "greedy":
            continue
        C, R = c["C"], c["R"]
        env = Tetris(C, R)
        env.tetrominos = [cls("bcts", 8, C) for cls in tetromino.PIECE_CLASSES]
        sampler = ListSampler(env.tetrominos)
        env.tetromino_sampler = sampler
        for p in range(0, len(c["piece"]), 4):
            parent = state.State(rows_to_rep(c["rows"][p][:R + 4], C).astype(np.int64))
            piece = env.tetrominos[int(c["piece"][p])]
            legal = np.nonzero(c["valid"][p])[0]
            for a, s in enumerate(legal):
                total = 0
                for f in range(c["n_forks"]):
                    sampler.load(c["tape"][p, s, f])
                    env.current_state, env.current_tetromino = parent, piece
                    env.get_after_states()
                    total += env.single_rollout(a, _greedy_policy_function, c["length"])
                    assert env.current_state is parent and env.current_tetromino is piece       # game.py:147-148
                assert total == c["ret_sum"][p, s], (C, R, p, s)
                checked += 1
    assert checked > 100


def test_perform_rollouts_as_shipped():
    """perform_rollouts called the way the reference ships it (one get_after_states before it): `afterstates` is left
    stale by every rollout (game.py:140,147-148), so later rollouts index the previous rollout's last list -- and the
    reference raises IndexError when that list is too short.  The drop-in class reproduces the recorded returns call
    by call and fails at the same call (tests/golden/rollouts.npz part B)."""
    from tetris import state, tetromino
    from tetris.game import Tetris
    g = load("rollouts")
    for k in range(int(g["n_stale"])):
        pre = "s%d_" % k
        C, R, length, n = (int(g[pre + x]) for x in ("C", "R", "length", "n"))
        env = Tetris(C, R)
        env.tetrominos = [cls("bcts", 8, C) for cls in tetromino.PIECE_CLASSES]
        sampler = ListSampler(env.tetrominos)
        env.tetromino_sampler = sampler
        sampler.load(g[pre + "tape"])
        env.current_state = state.State(rows_to_rep(g[pre + "rows"], C).astype(np.int64))
        env.current_tetromino = env.tetrominos[int(g[pre + "piece"])]
        env.get_after_states()
        calls, err_at = [], -1
        for a in range(int(g[pre + "n_actions"])):
            for i in range(n):
                try:
                    calls.append(env.single_rollout(a, _greedy_policy_function, length))
                except IndexError:
                    err_at = len(calls)
                    break
            if err_at >= 0:
                break
        assert calls == g[pre + "returns"].tolist() and err_at == int(g[pre + "error_at"]), (k, calls, err_at)
        assert sampler.pos == len(g[pre + "tape"])                     # same number of pieces consumed


def test_rollout_helpers():
    from tetris.game import Tetris
    np.random.seed(3)
    env = Tetris(10, 10)
    feats, _ = env.get_after_states()
    before = (env.current_state, env.current_tetromino)
    policy = lambda st, f: int(np.argmax(f.sum(axis=1)))
    acts, rets = env.perform_rollouts(list(range(2)), policy, length=1, n=2)     # length 1: no policy step, nothing stale
    assert acts == [0, 1] and rets == [0.0, 0.0]
    assert (env.current_state, env.current_tetromino) == before      # restored (game.py:147-148)


def test_example_play_flow():
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tetris")
    sys.path.insert(0, d)
    try:
        import example_play
        assert example_play.main(100, quiet=True, seed=0) == -71
    finally:
        sys.path.remove(d)


def test_state_rejects_inconsistent_heights():
    """State(representation, lowest_free_rows=...) (state.py:22-25): heights that match the board are accepted (what
    every reference call site passes), heights that do not are rejected instead of silently ignored."""
    from tetris import state
    rep = np.zeros((14, 10), np.int64)
    rep[:3, 2] = 1
    rep[0, 5] = 1
    h = np.zeros(10, np.int64); h[2] = 3; h[5] = 1
    s = state.State(rep, lowest_free_rows=h)
    assert np.array_equal(s.lowest_free_rows, h)
    bad = h.copy(); bad[5] = 2
    with pytest.raises(ValueError, match="lowest_free_rows"):
        state.State(rep, lowest_free_rows=bad)
