"""GPU parity tests: the CUDA path (through the C ABI, via tetris_b200.BatchedTetris) against the golden
fixtures recorded from the reference and against the CPU oracle on the same seeded inputs.  Bit-exact."""
import ctypes as C

import numpy as np
import pytest

from golden_util import TRACES, feat2, load, rep_to_rows, replay_rollout_config, rollout_configs, rows_to_rep
from test_oracle_golden import replay_trace

pytestmark = pytest.mark.gpu

SHAPES = [(10, 20), (10, 10), (6, 12), (8, 16), (4, 4), (16, 27), (12, 24), (7, 9)]   # 7x9: loaded as a shape plugin


def _torch():
    import torch
    return torch


class CudaBatch:
    """tetris_b200.BatchedTetris behind the oracle.Batch interface the trace replayer drives."""

    def __init__(self, Cc, R, n, piece_set=1, seed=0, env_offset=0):
        from tetris_b200 import BatchedTetris, _lib
        self.env = BatchedTetris(Cc, R, n, piece_set=piece_set, seed=seed, env_offset=env_offset)
        self.n, self.C, self.R = n, Cc, R
        self._lib = _lib.lib()

    @property
    def piece(self):
        return self.env.piece.astype(np.int32)

    @property
    def heights(self):
        return self.env.heights.astype(np.int32)

    def rows(self):
        return self.env.rows()

    def reset(self, tape=None):
        self.env.reset(tape)

    def reset_masked(self, mask, tape=None):
        self.env.reset_masked(np.asarray(mask, np.uint8), tape)

    def afterstates(self):
        feats, valid, count = self.env.get_after_states(include_terminal=True)
        n_all = np.array([self._lib.tb_num_slots(int(p), self.C) for p in self.piece], np.int32)
        f = feats.cpu().numpy()
        for e in range(self.n):
            f[e, n_all[e]:] = 0
        return f, valid.cpu().numpy().view(np.uint64), count.cpu().numpy(), n_all

    def step(self, actions, tape=None, auto_reset=False, action_is_slot=False):
        obs, reward, done, lines = self.env.step(actions, tape=tape, auto_reset=auto_reset,
                                                 action_is_slot=action_is_slot)
        return obs.cpu().numpy(), reward.cpu().numpy(), done.cpu().numpy(), lines.cpu().numpy()


def _cuda_batch(Cc, R, n, ps, seed):
    return CudaBatch(Cc, R, n, ps, seed)


def test_library_loaded():
    from tetris_b200 import _lib
    assert _lib.lib().tb_version() >= 100


@pytest.mark.parametrize("name", TRACES)
def test_trace_tape(name):
    replay_trace(load("trace_" + name), _cuda_batch, rng_mode=False)


@pytest.mark.parametrize("name", TRACES)
def test_trace_rng(name):
    replay_trace(load("trace_" + name), _cuda_batch, rng_mode=True)


@pytest.mark.parametrize("name", ["afterstates", "afterstates_dense"])
def test_afterstates_fixture(name):
    """Every board x piece of the reference-generated fixtures, through tb_afterstates and tb_afterstates_export.
    `afterstates_dense` holds near-full stacks (four-line clears, overflow rescued by a clear) on five shapes."""
    torch = _torch()
    from tetris_b200 import BatchedTetris, _lib
    g = load(name)
    shapes = g["shape"].astype(int)
    for (Cc, R) in sorted({tuple(x) for x in shapes.tolist()}):
        idx = np.nonzero((shapes[:, 0] == Cc) & (shapes[:, 1] == R))[0]
        N = R + 4
        for ps in (0, 1):
            sel = np.array([i for i in idx if (g["piece"][i] >= 7) == (ps == 0)])
            env = BatchedTetris(Cc, R, len(sel), piece_set=ps)
            env.import_boards(g["rows"][sel][:, :N], piece=g["piece"][sel].astype(np.uint8))
            feats, valid, count = env.get_after_states(include_terminal=True)
            feats, valid, count = feats.cpu().numpy(), valid.cpu().numpy(), count.cpu().numpy()
            A = env.a_max
            rows_o = torch.empty((len(sel), A, N), dtype=torch.int16, device="cuda")
            h_o = torch.empty((len(sel), A, Cc), dtype=torch.uint8, device="cuda")
            info_o = torch.empty((len(sel), A, 4), dtype=torch.int32, device="cuda")
            f_o = torch.empty((len(sel), A, 8), dtype=torch.float32, device="cuda")
            p = lambda t: C.c_void_p(t.data_ptr())
            _lib.check(_lib.lib().tb_afterstates_export(C.c_void_p(env.state.data_ptr()), Cc, R, len(sel), p(f_o),
                                                        p(rows_o), p(h_o), p(info_o), A, None))
            torch.cuda.synchronize()
            rows_o = rows_o.cpu().numpy().view(np.uint16)
            h_o, info_o, f_o = h_o.cpu().numpy(), info_o.cpu().numpy(), f_o.cpu().numpy()
            for k, i in enumerate(sel):
                s, n = int(g["start"][i]), int(g["count"][i])
                sl = slice(s, s + n)
                assert np.array_equal(feat2(feats[k, :n]), g["a_feat2"][sl]), (Cc, R, i)
                assert np.array_equal(feat2(f_o[k, :n]), g["a_feat2"][sl])
                term = g["a_terminal"][sl]
                vb = ((int(valid[k]) >> np.arange(n)) & 1).astype(bool)
                assert np.array_equal(vb, ~term)
                assert count[k] == (~term).sum()
                assert np.array_equal(rows_o[k, :n], g["a_rows"][sl][:, :N])
                assert np.array_equal(h_o[k, :n], g["a_heights"][sl][:, :Cc])
                assert np.array_equal(info_o[k, :n, 0], g["a_anchor"][sl][:, 1])
                assert np.array_equal(info_o[k, :n, 3], g["a_anchor"][sl][:, 0])
                assert np.array_equal(info_o[k, :n, 2].astype(bool), term)
                ncl = np.array([bin(int(m) & 0xffffffff).count("1") for m in info_o[k, :n, 1]])
                assert np.array_equal(ncl, g["a_n_cleared"][sl])


def _sync_oracle_from_device(env, ob):
    """Copy the device envs' boards/pieces into an oracle batch."""
    rows, heights, piece = env.export_boards()
    ob.rep[:] = rows_to_rep(rows.cpu().numpy().view(np.uint16), env.num_columns)
    ob.heights[:] = heights.cpu().numpy()
    ob.piece[:] = piece.cpu().numpy()


def _compare_state(env, ob):
    assert np.array_equal(env.rows(), ob.rows())
    assert np.array_equal(env.heights, ob.heights)
    assert np.array_equal(env.piece, ob.piece)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("piece_set", [0, 1])
def test_lockstep_vs_oracle(shape, piece_set):
    """BASELINE config 2: 4096 lockstep envs, random legal placements, device RNG, compared with the oracle after
    every step (boards, heights, pieces, legal counts, feature matrices, obs, reward, done, lines)."""
    from oracle import oracle as orc
    from tetris_b200 import BatchedTetris
    Cc, R = shape
    n = 4096 if shape == (10, 20) else (512 if Cc > 10 else 1024)
    steps = 200 if shape == (10, 20) else (120 if R > 20 else 80)
    seed = 0x5EED
    env = BatchedTetris(Cc, R, n, piece_set=piece_set, seed=seed)
    ob = orc.Batch(Cc, R, n, piece_set=piece_set, seed=seed)
    ob.reset()
    _compare_state(env, ob)
    u = np.random.RandomState(0).randint(0, 2 ** 31 - 1, size=(steps, n))
    n_done = 0
    for t in range(steps):
        with_terminal = (t % 2 == 0)                       # alternate game.py:74-78 / the default of game.py:69
        feats, valid, count = env.get_after_states(include_terminal=with_terminal)
        of, ov, oc, on = ob.afterstates()
        assert np.array_equal(count.cpu().numpy(), oc), t
        assert np.array_equal(valid.cpu().numpy().view(np.uint64), ov), t
        f = feats.cpu().numpy()
        mask = np.arange(env.a_max)[None, :] < on[:, None]
        if not with_terminal:                              # only the legal afterstates' rows are defined
            mask &= ((ov[:, None] >> np.arange(env.a_max, dtype=np.uint64)) & np.uint64(1)).astype(bool)
        assert np.array_equal(f[mask], of[mask]), t
        a = (u[t] % oc).astype(np.int32)
        obs, rew, done, lines = env.step(a, auto_reset=True)
        oobs, orew, odone, olines = ob.step(a, auto_reset=True)
        assert np.array_equal(obs.cpu().numpy(), oobs), t
        assert np.array_equal(rew.cpu().numpy(), orew), t
        assert np.array_equal(done.cpu().numpy(), odone), t
        assert np.array_equal(lines.cpu().numpy(), olines), t
        _compare_state(env, ob)
        n_done += int(odone.sum())
    assert n_done > 0


@pytest.mark.parametrize("policy", ["random", "greedy"])
@pytest.mark.parametrize("shape", [(10, 20), (10, 10), (6, 12), (4, 4), (16, 27), (7, 9)])
def test_rollout_vs_oracle(shape, policy):
    """Fused rollout kernels (in-kernel policy, RNG, game-over, auto-reset) against the oracle's loop: identical
    final boards/pieces/counters and identical episode statistics."""
    from oracle import oracle as orc
    from tetris_b200 import BatchedTetris
    Cc, R = shape
    n, seed = (2048 if Cc <= 10 else 512) + 17, 77     # not a multiple of 32: exercises the ragged last warp tile
    T = (40 if policy == "random" else 25) * (3 if R > 20 else 1)
    env = BatchedTetris(Cc, R, n, piece_set=1, seed=seed, env_offset=5)
    ob = orc.Batch(Cc, R, n, piece_set=1, seed=seed, env_offset=5)
    ob.reset()
    total = np.zeros(16, np.int64)
    for rnd in range(3):
        env.rollout(T, policy)
        st = ob.rollout(T, 0 if policy == "random" else 1, threads=8)
        mx = np.maximum(total[10:12], st[10:12])
        total += st
        total[10:12] = mx                                  # the two maxima combine by max, the rest by sum
        _compare_state(env, ob)
        assert np.array_equal(env.stats.cpu().numpy(), total), rnd
    assert total[0] == 3 * T * n
    if policy == "random" or shape not in ((10, 20), (16, 27)):
        assert total[1] > 0                                # episodes ended and were auto-reset


def test_sharding_invariance():
    """per-env RNG is keyed by the GLOBAL env id, so splitting the env range into shards (as the
    multi-GPU launcher does) leaves every env's trajectory and the summed statistics unchanged."""
    from tetris_b200 import BatchedTetris
    n, seed, T = 3000, 5, 30
    whole = BatchedTetris(10, 10, n, piece_set=1, seed=seed)
    whole.rollout(T, "greedy")
    whole.rollout(T, "random")
    parts, off = [], 0
    for cnt in (1000, 1500, 500):
        p = BatchedTetris(10, 10, cnt, piece_set=1, seed=seed, env_offset=off)
        p.rollout(T, "greedy")
        p.rollout(T, "random")
        parts.append(p)
        off += cnt
    assert np.array_equal(np.concatenate([p.rows() for p in parts]), whole.rows())
    assert np.array_equal(np.concatenate([p.piece for p in parts]), whole.piece)
    s = sum(p.stats.cpu().numpy() for p in parts)
    s[10:12] = np.max([p.stats.cpu().numpy()[10:12] for p in parts], axis=0)
    assert np.array_equal(s, whole.stats.cpu().numpy())


def test_step_errors_and_slots():
    from tetris_b200 import BatchedTetris
    env = BatchedTetris(10, 20, 64, piece_set=1, seed=1)
    with pytest.raises(IndexError):
        env.step(np.full(64, 40, np.int32))            # game.py:83 IndexError
    with pytest.raises(IndexError):
        env.step(np.full(64, -1, np.int32))
    before = env.rows().copy()
    assert not before.any()                             # failed steps leave the envs untouched
    # slot-indexed actions == rank-indexed actions while every slot is legal
    a = np.arange(64, dtype=np.int32) % 9
    env2 = BatchedTetris(10, 20, 64, piece_set=1, seed=1)
    o1 = env.step(a)
    o2 = env2.step(a, action_is_slot=True)
    for x, y in zip(o1, o2):
        assert np.array_equal(x.cpu().numpy(), y.cpu().numpy())
    assert np.array_equal(env.rows(), env2.rows())


def test_million_envs_properties():
    """BASELINE config 3 at full size: 1M envs x every rotation x column placement.  Too many for the oracle,
    so: (1) bit-exact parity on a strided 4096-env sample, (2) size-independent invariants on all of them."""
    torch = _torch()
    from oracle import oracle as orc
    from tetris_b200 import BatchedTetris, _lib
    n, seed = 1 << 20, 2024
    env = BatchedTetris(10, 20, n, piece_set=1, seed=seed)
    env.rollout(30, "random")                            # realistic boards: ~30 random placements each
    feats, valid, count = env.get_after_states()
    rows, heights, piece = env.export_boards()
    piece_np = piece.cpu().numpy()
    n_all = torch.as_tensor(np.array([_lib.lib().tb_num_slots(p, 10) for p in range(9)], np.int64), device="cuda")[piece.long()]
    # invariants on all 1M envs
    assert bool((count.long() == torch.stack([(valid >> s) & 1 for s in range(env.a_max)]).sum(0)).all())
    assert bool((count.long() <= n_all).all()) and bool((count > 0).all())
    slot = torch.arange(env.a_max, device="cuda")[None, :]
    live = ((valid[:, None] >> slot) & 1).bool()         # rows of legal afterstates (terminal rows are not written)
    assert bool((live <= (slot < n_all[:, None])).all())
    f = feats[live]
    assert bool((f[:, 1] >= 10).all())                   # column_transitions >= C
    assert bool((f[:, 0] <= f[:, 2]).all())              # rows_with_holes <= holes
    assert bool(((f * 2) == (f * 2).round()).all())      # half-integers only
    assert bool((f[:, 6] >= 0).all()) and bool((f[:, 3] >= 1).all())
    # sample parity
    sel = np.arange(0, n, n // 4096)[:4096]
    ob = orc.Batch(10, 20, len(sel), piece_set=1)
    ob.rep[:] = rows_to_rep(rows.cpu().numpy().view(np.uint16)[sel], 10)
    ob.heights[:] = heights.cpu().numpy()[sel]
    ob.piece[:] = piece_np[sel]
    of, ov, oc, on = ob.afterstates()
    fs = feats[torch.as_tensor(sel, device="cuda")].cpu().numpy()
    mask = ((ov[:, None] >> np.arange(env.a_max, dtype=np.uint64)) & np.uint64(1)).astype(bool)
    assert np.array_equal(fs[mask], of[mask])
    assert np.array_equal(valid.cpu().numpy().view(np.uint64)[sel], ov)
    assert np.array_equal(count.cpu().numpy()[sel], oc)


def test_eval_states():
    """tb_eval_states == State(representation).get_features() (state.py:5-38) on arbitrary boards."""
    torch = _torch()
    from oracle import oracle as orc
    from tetris_b200 import _lib
    rng = np.random.default_rng(5)
    for (Cc, R) in SHAPES:
        N = R + 4
        reps = (rng.random((200, N, Cc)) < rng.random((200, 1, 1))).astype(np.uint8)
        reps[:, R:, :] = 0
        reps[:50, 0, :] = 1                               # full bottom row: cleared by State.__init__ (changed_lines=[0])
        rows = torch.as_tensor(rep_to_rows(reps).view(np.int16), device="cuda")
        rows_o = torch.empty_like(rows)
        h_o = torch.empty((200, Cc), dtype=torch.uint8, device="cuda")
        info = torch.empty((200, 4), dtype=torch.int32, device="cuda")
        f = torch.empty((200, 8), dtype=torch.float32, device="cuda")
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(_lib.lib().tb_eval_states(Cc, R, 200, p(rows), None, p(rows_o), p(h_o), p(info), p(f), None))
        torch.cuda.synchronize()
        for i in range(200):
            o = orc.board_features(Cc, R, reps[i])
            assert np.array_equal(f[i].cpu().numpy(), o["features"]), (Cc, R, i)
            assert np.array_equal(h_o[i].cpu().numpy(), o["heights"])
            assert np.array_equal(rows_o[i].cpu().numpy().view(np.uint16), rep_to_rows(o["rep"]))
            assert int(info[i, 0]) == o["n_cleared"] and bool(info[i, 2]) == o["terminal"]


@pytest.mark.parametrize("policy", ["greedy", "random"])
@pytest.mark.parametrize("shape", [(10, 20), (10, 10), (6, 12)])
def test_rollout_values_vs_oracle(shape, policy):
    """Tetris.perform_rollouts (game.py:129-160) for every env x action on the device -- forks, follow-up
    policy steps, -1 on game over -- against the oracle's loop with the same fork RNG convention.  Bit-exact sums."""
    from oracle import oracle as orc
    from tetris_b200 import BatchedTetris
    Cc, R = shape
    n, seed = 300, 21
    env = BatchedTetris(Cc, R, n, piece_set=1, seed=seed, env_offset=3)
    ob = orc.Batch(Cc, R, n, piece_set=1, seed=seed, env_offset=3)
    ob.reset()
    T0 = 14 if shape == (10, 20) else 9
    env.rollout(T0, "random"); ob.rollout(T0, 0)          # boards tall enough that some forks end
    _compare_state(env, ob)
    before = (env.rows().copy(), env.piece.copy())
    length, forks = 4, 3
    mean, valid = env.rollout_values(length=length, n=forks, policy=policy, seed=1234)
    want, wvalid = ob.rollout_values(length, forks, 0 if policy == "random" else 1, seed2=1234,
                                     child_offset=3 * env.a_max * forks)
    assert np.array_equal(valid.cpu().numpy().view(np.uint64), wvalid)
    got = np.rint(mean.cpu().numpy() * forks).astype(np.int64)
    assert np.array_equal(got, want)
    assert (want < 0).any() and (want.min() >= -forks * max(1, length - 1))
    assert np.array_equal(env.rows(), before[0]) and np.array_equal(env.piece, before[1])   # the envs are not stepped
    # length 1: only the action itself -> 0, or -1 where the next piece cannot be placed (game.py:133-137)
    m1, _ = env.rollout_values(length=1, n=2, policy=policy, seed=5)
    w1, _ = ob.rollout_values(1, 2, 0, seed2=5, child_offset=3 * env.a_max * 2)
    assert np.array_equal(np.rint(m1.cpu().numpy() * 2).astype(np.int64), w1)


def test_action_probabilities_vs_numpy():
    """utils.compute_action_probabilities / grad_of_log_action_probabilities (utils.py:26-38) batched on
    the device, against the reference's NumPy formulas per env.  float64; tolerance 1e-12 absolute on probabilities
    (exp implementations differ in the last ulp), 1e-10 on the gradient."""
    torch = _torch()
    from tetris_b200 import BatchedTetris, utils
    n = 2000
    env = BatchedTetris(10, 20, n, piece_set=1, seed=8)
    env.rollout(25, "random")
    feats, valid, count = env.get_after_states()
    w = np.array([-2.4, -1.9, -1.3, -1.2, -1.0, -0.9, 0.6, -0.1])
    vm = valid.cpu().numpy().view(np.uint64)
    bits = ((vm[:, None] >> np.arange(env.a_max, dtype=np.uint64)) & np.uint64(1)).astype(bool)
    rng = np.random.default_rng(0)
    slots = np.array([rng.choice(np.nonzero(b)[0]) for b in bits], np.int32)
    for temp in (1.0, 0.25, 7.5):
        probs, grad = env.action_probabilities(feats, valid, w, temperature=temp, actions=torch.as_tensor(slots))
        probs, grad, f = probs.cpu().numpy(), grad.cpu().numpy(), feats.cpu().numpy()
        assert not probs[~bits].any()
        for e in range(0, n, 7):
            fe = f[e][bits[e]].astype(np.float64)
            want = utils.compute_action_probabilities(fe, w, temp)
            assert np.allclose(probs[e][bits[e]], want, rtol=0, atol=1e-12), (e, temp)
            k = int(np.nonzero(np.nonzero(bits[e])[0] == slots[e])[0][0])
            assert np.allclose(grad[e], utils.grad_of_log_action_probabilities(fe, want, k), rtol=0, atol=1e-10)
        assert np.allclose(probs.sum(axis=1), 1.0, atol=1e-12)


def test_action_probabilities_vs_reference_fixture():
    """tb_action_probabilities against vectors recorded from the reference's own utils.compute_action_probabilities /
    grad_of_log_action_probabilities (utils.py:26-38; tests/golden/learner.npz): arbitrary legal-slot subsets, three
    temperatures.  float64 on the device; tolerance 1e-12 absolute on probabilities, 1e-10 on the gradient (the device's
    exp differs from NumPy's in the last ulp; sums run in slot order on both sides)."""
    torch = _torch()
    from golden_util import load
    from tetris_b200 import BatchedTetris
    g = load("learner")
    n, a_max = g["feats"].shape[:2]
    env = BatchedTetris(10, 20, n, piece_set=1, seed=1)
    assert env.a_max == a_max
    feats = torch.as_tensor(g["feats"]).cuda()
    valid = torch.as_tensor(g["valid"].view(np.int64)).cuda()
    bits = ((g["valid"][:, None] >> np.arange(a_max, dtype=np.uint64)) & np.uint64(1)).astype(bool)
    for ti, t in enumerate(g["temps"]):
        probs, grad = env.action_probabilities(feats, valid, g["weights"], temperature=float(t),
                                               actions=torch.as_tensor(g["actions"]))
        probs, grad = probs.cpu().numpy(), grad.cpu().numpy()
        assert not probs[~bits].any()
        assert np.allclose(probs, g["probs"][ti], rtol=0, atol=1e-12), t
        assert np.allclose(grad, g["grads"][ti], rtol=0, atol=1e-10), t


@pytest.mark.parametrize("n", [1, 31, 255, 256, 257, 700])
def test_ragged_sizes(n):
    """Env counts around the warp / CTA-tile boundaries (tile = 256 envs): K1, K2, K3 against the oracle, with feature
    directions and include_terminal, on a small board where game overs and line clears are frequent."""
    from oracle import oracle as orc
    from tetris_b200 import BatchedTetris
    dirs = np.array([-1, -1, -1, -1, -1, -1, 1, -1], np.float32)
    for (Cc, R, ps) in ((6, 12, 1), (10, 10, 0)):
        env = BatchedTetris(Cc, R, n, piece_set=ps, seed=n, feature_directions=dirs)
        ob = orc.Batch(Cc, R, n, piece_set=ps, seed=n)
        ob.reset()
        for rnd in range(3):
            env.rollout(7, "greedy"); ob.rollout(7, 1)
            env.rollout(9, "random"); ob.rollout(9, 0)
            _compare_state(env, ob)
            feats, valid, count = env.get_after_states(include_terminal=True)
            of, ov, oc, on = ob.afterstates()
            mask = np.arange(env.a_max)[None, :] < on[:, None]
            assert np.array_equal(feats.cpu().numpy()[mask], (of * dirs)[mask])
            assert np.array_equal(np.signbit(feats.cpu().numpy()[mask]), np.signbit((of * dirs)[mask]))   # -0.0 where 0 * -1
            assert np.array_equal(valid.cpu().numpy().view(np.uint64), ov) and np.array_equal(count.cpu().numpy(), oc)
            a = (np.arange(n) * 7 % np.maximum(oc, 1)).astype(np.int32)
            obs, rew, done, lines = env.step(a, auto_reset=True)
            oobs, orew, odone, olines = ob.step(a, auto_reset=True)
            assert np.array_equal(obs.cpu().numpy(), oobs * dirs) and np.array_equal(rew.cpu().numpy(), orew)
            assert np.array_equal(done.cpu().numpy(), odone) and np.array_equal(lines.cpu().numpy(), olines)


@pytest.mark.parametrize("shape,steps", [((10, 20), 1200), ((10, 10), 1500), ((4, 4), 1500)])
def test_long_greedy_rollout_subset(shape, steps):
    """Long fused rollouts (many episodes per env on the small boards, thousands of line clears, queue overflows on 4x4)
    on 20,000 envs; envs are independent and their RNG is keyed by the global env id, so the oracle replays only the
    first 192 envs and must end on identical boards, pieces and episode counters."""
    from oracle import oracle as orc
    from tetris_b200 import BatchedTetris
    Cc, R = shape
    n, m, seed = 20000, 192, 4242
    env = BatchedTetris(Cc, R, n, piece_set=1, seed=seed)
    ob = orc.Batch(Cc, R, m, piece_set=1, seed=seed)
    ob.reset()
    for chunk in (steps // 3, steps - steps // 3):
        env.rollout(chunk, "greedy")
        ob.rollout(chunk, 1, threads=8)
    rows, heights, piece = env.export_boards(0, m)
    assert np.array_equal(rows.cpu().numpy().view(np.uint16), ob.rows())
    assert np.array_equal(heights.cpu().numpy(), ob.heights) and np.array_equal(piece.cpu().numpy(), ob.piece)
    assert int(env.stats[0]) == n * steps


def test_fuzz_against_oracle():
    """tests/fuzz_gpu.py, 40 rounds: random seeds, env offsets, shapes, piece sets, policy weights and feature
    directions; rollouts interleaved with afterstate / step comparisons.  (1500 rounds were run by hand.)"""
    import fuzz_gpu
    shapes = [(10, 20), (10, 10), (6, 12), (8, 16), (4, 4)]
    for i in range(40):
        Cc, R = shapes[i % len(shapes)]
        fuzz_gpu.one(Cc, R, (i // len(shapes)) % 2, 7000 + i, [257, 1000, 3001, 513][i % 4])


def test_cuda_graph_lockstep_equals_eager():
    """BatchedTetris.capture_lockstep: the captured get_after_states -> policy -> step iteration replays to exactly
    what the eager calls produce (same boards, rewards, dones) over many steps with game overs and auto-reset."""
    torch = _torch()
    from tetris_b200 import BatchedTetris
    n = 4096
    pol = lambda f, v, c: (torch.arange(n, device="cuda", dtype=torch.int64) * 7919 % c.long().clamp(min=1)).int()
    a = BatchedTetris(10, 10, n, piece_set=1, seed=31)
    b = BatchedTetris(10, 10, n, piece_set=1, seed=31)
    replay, out = a.capture_lockstep(pol)
    dones = 0
    for t in range(120):
        replay()
        f, v, c = b.get_after_states()
        obs, rew, done, lines = b.step(pol(f, v, c), auto_reset=True)
        assert torch.equal(out["reward"], rew) and torch.equal(out["done"], done) and torch.equal(out["obs"], obs)
        dones += int(done.sum())
    assert np.array_equal(a.rows(), b.rows()) and np.array_equal(a.piece, b.piece) and dones > 0


def test_batched_fitness():
    torch = _torch()
    from oracle import oracle as orc
    from tetris_b200 import BatchedTetris
    env = BatchedTetris(10, 20, 500, piece_set=1, seed=2)
    env.rollout(20, "random")
    feats, valid, count = env.get_after_states(include_terminal=True)
    s = env.fitness(feats).cpu().numpy()
    f = feats.cpu().numpy()
    for e in range(0, 500, 37):
        for k in range(9):
            assert s[e, k] == orc.fitness(f[e, k], orc.BCTS_WEIGHTS)


@pytest.mark.parametrize("chunks", [1, 3, 4])
def test_host_rollout_equals_resident(chunks):
    """HostRollout (host-resident boards, chunked over CUDA streams so copies overlap the rollout kernel) ends in the
    same boards, pieces and statistics as the device-resident whole job, whatever the chunking, and its host
    buffers agree with the oracle's replay of the same seeded games."""
    from oracle import oracle as orc
    from tetris_b200 import BatchedTetris, HostRollout
    n, seed, T = 2051, 9, 25
    whole = BatchedTetris(10, 10, n, piece_set=1, seed=seed)
    host = HostRollout(10, 10, n, chunks=chunks, piece_set=1, seed=seed)
    ob = orc.Batch(10, 10, n, piece_set=1, seed=seed)
    ob.reset()
    assert np.array_equal(host.h_piece.numpy(), whole.piece)
    for policy, pid in (("random", 0), ("greedy", 1), ("greedy", 1)):
        whole.rollout(T, policy)
        ob.rollout(T, pid)
        st = host.play(T, policy)
        assert np.array_equal(host.h_rows.numpy().view(np.uint16), whole.rows())
        assert np.array_equal(host.h_heights.numpy(), whole.heights)
        assert np.array_equal(host.h_piece.numpy(), whole.piece)
        assert np.array_equal(st.numpy(), whole.stats.cpu().numpy())
        assert np.array_equal(host.h_rows.numpy().view(np.uint16), ob.rows())
        assert np.array_equal(host.h_piece.numpy().astype(np.int32), np.asarray(ob.piece, np.int32))
    assert int(st[1]) > 0                                  # episodes ended and were auto-reset on the way


@pytest.mark.parametrize("k1cfg,k3cfg", [(0, 0), (2, 0), (3, 2), (3, 3), (4, 4), (5, 5)])
def test_tile_configurations(k1cfg, k3cfg):
    """tb_afterstates / tb_rollout pick a CTA tile configuration by batch size (throughput: 256-env tiles; small
    batches: 128-env tiles, or 32 / 64 envs with several warps per env group).  Every configuration must give the
    oracle's results: forced here through the library's tuning variables on a batch with ragged tiles, on a board where
    line clears and game overs are frequent and on the headline board."""
    from oracle import oracle as orc
    from tetris_b200 import BatchedTetris
    from tetris_b200 import _lib
    _lib.set_tuning("k1_cfg", k1cfg)
    _lib.set_tuning("k3_cfg", k3cfg)
    try:
        _tile_configuration_body()
    finally:
        _lib.set_tuning("k1_cfg", -1)
        _lib.set_tuning("k3_cfg", -1)


def _tile_configuration_body():
    from oracle import oracle as orc
    from tetris_b200 import BatchedTetris
    for (Cc, R, n, T) in ((6, 12, 1337, 40), (10, 20, 777, 60)):
        env = BatchedTetris(Cc, R, n, piece_set=1, seed=77)
        ob = orc.Batch(Cc, R, n, piece_set=1, seed=77)
        ob.reset()
        env.rollout(25, "random"); s0 = ob.rollout(25, 0, threads=8)
        env.rollout(T, "greedy"); s1 = ob.rollout(T, 1, threads=8)
        _compare_state(env, ob)
        total = s0 + s1
        total[10:12] = np.maximum(s0[10:12], s1[10:12])    # the two maxima combine by max, the rest by sum
        assert np.array_equal(env.stats.cpu().numpy(), total)
        feats, valid, count = env.get_after_states(include_terminal=True)
        of, ov, oc, on = ob.afterstates()
        mask = np.arange(env.a_max)[None, :] < on[:, None]
        assert np.array_equal(feats.cpu().numpy()[mask], of[mask])
        assert np.array_equal(valid.cpu().numpy().view(np.uint64), ov) and np.array_equal(count.cpu().numpy(), oc)


def _oracle_window(Cc, R, first, count, seed, warm, T):
    """The oracle's replay of envs [first, first + count) of a bigger job (RNG keyed by the global env id)."""
    from oracle import oracle as orc
    ob = orc.Batch(Cc, R, count, piece_set=1, seed=seed, env_offset=first)
    ob.reset()
    ob.rollout(warm, 0, threads=8)
    ob.rollout(T, 1, threads=8)
    return ob


@pytest.mark.parametrize("k1cfg,k3cfg", [(0, 0), (3, 2)])
def test_multi_tile_per_cta(k1cfg, k3cfg):
    """A CTA of the tile kernels loops over several tiles once there are more tiles than the grid holds (above
    ~1.2 M envs per GPU for K3); the loop carries state from tile to tile (piece-count parity, shared-memory records).
    Forced here on a small batch by capping the grid through the ABI's tuning hook: 3 CTAs for 11 / 21 tiles, i.e.
    3-7 tiles per CTA, K1 and K3, against the oracle."""
    from oracle import oracle as orc
    from tetris_b200 import BatchedTetris, _lib
    _lib.set_tuning("k1_cfg", k1cfg); _lib.set_tuning("k3_cfg", k3cfg); _lib.set_tuning("max_ctas", 3)
    try:
        for (Cc, R, n, T) in ((6, 12, 2570, 40), (10, 20, 2570, 50)):
            env = BatchedTetris(Cc, R, n, piece_set=1, seed=123)
            ob = orc.Batch(Cc, R, n, piece_set=1, seed=123)
            ob.reset()
            env.rollout(25, "random"); s0 = ob.rollout(25, 0, threads=8)
            env.rollout(T, "greedy"); s1 = ob.rollout(T, 1, threads=8)
            _compare_state(env, ob)
            total = s0 + s1
            total[10:12] = np.maximum(s0[10:12], s1[10:12])
            assert np.array_equal(env.stats.cpu().numpy(), total)
            feats, valid, count = env.get_after_states(include_terminal=True)
            of, ov, oc, on = ob.afterstates()
            mask = np.arange(env.a_max)[None, :] < on[:, None]
            assert np.array_equal(feats.cpu().numpy()[mask], of[mask])
            assert np.array_equal(valid.cpu().numpy().view(np.uint64), ov) and np.array_equal(count.cpu().numpy(), oc)
    finally:
        _lib.set_tuning("k1_cfg", -1); _lib.set_tuning("k3_cfg", -1); _lib.set_tuning("max_ctas", 0)


def test_two_million_env_greedy_rollout():
    """2^21 envs on one GPU: more 256-env tiles (8192) than K3's grid holds (16 x 148 x 2 = 4736 CTAs), so CTAs take a
    second tile with the shipped configuration -- no tuning hook.  Too many for the oracle: it replays the first and
    the last 192 envs (their RNG is keyed by the global env id) and must end on identical boards, pieces and counters;
    the statistics must satisfy the integer invariants of a complete run."""
    from tetris_b200 import BatchedTetris, _lib
    n, m, seed, warm, T = 1 << 21, 192, 777, 20, 24
    assert _lib.get_tuning("max_ctas") == 0 and _lib.get_tuning("k3_cfg") == -1
    env = BatchedTetris(10, 20, n, piece_set=1, seed=seed)
    env.rollout(warm, "random")
    env.stats.zero_()
    env.rollout(T, "greedy")
    st = env.stats_dict()
    assert st["placements"] == n * T and sum(st["lines%d" % i] for i in range(5)) == n * T
    assert st["reward"] == st["lines"] - st["placements"] - 100 * st["episodes"]
    for first in (0, n - m):
        ob = _oracle_window(10, 20, first, m, seed, warm, T)
        rows, heights, piece = env.export_boards(first, m)
        assert np.array_equal(rows.cpu().numpy().view(np.uint16), ob.rows()), first
        assert np.array_equal(heights.cpu().numpy(), ob.heights) and np.array_equal(piece.cpu().numpy(), ob.piece)
    # K1 over the same 2^21 envs (8192 tiles over 8 x 148 x 3 = 3552 CTAs): sampled parity at both ends
    feats, valid, count = env.get_after_states()
    for first in (0, n - m):
        ob = _oracle_window(10, 20, first, m, seed, warm, T)
        of, ov, oc, on = ob.afterstates()
        mask = ((ov[:, None] >> np.arange(env.a_max, dtype=np.uint64)) & np.uint64(1)).astype(bool)
        assert np.array_equal(feats[first:first + m].cpu().numpy()[mask], of[mask])
        assert np.array_equal(valid[first:first + m].cpu().numpy().view(np.uint64), ov)


def test_step_validates_before_mutation_and_defines_outputs():
    """Tetris.step raises IndexError before anything changes (game.py:83).  check=True: one bad action among many ->
    IndexError naming the env, and NO env was stepped.  check=False (the asynchronous path): the offending env is left
    untouched and reports zero obs / reward / lines, every other env is stepped normally."""
    torch = _torch()
    from oracle import oracle as orc
    from tetris_b200 import BatchedTetris
    n = 300
    env = BatchedTetris(10, 10, n, piece_set=1, seed=4)
    ob = orc.Batch(10, 10, n, piece_set=1, seed=4)
    ob.reset()
    env.rollout(12, "random"); ob.rollout(12, 0)
    before = env.rows().copy()
    a = np.zeros(n, np.int32)
    a[137] = 60
    with pytest.raises(IndexError, match="env 137"):
        env.step(a)
    assert np.array_equal(env.rows(), before) and np.array_equal(env.piece, ob.piece)
    obs, rew, done, lines = env.step(torch.as_tensor(a), check=False)
    a_ok = a.copy(); a_ok[137] = 0
    oobs, orew, odone, olines = ob.step(a_ok)
    keep = np.arange(n) != 137
    assert np.array_equal(obs.cpu().numpy()[keep], oobs[keep]) and np.array_equal(rew.cpu().numpy()[keep], orew[keep])
    assert np.array_equal(done.cpu().numpy()[keep], odone[keep]) and np.array_equal(lines.cpu().numpy()[keep], olines[keep])
    assert not obs[137].any() and int(rew[137]) == 0 and int(lines[137]) == 0 and not bool(done[137])
    assert np.array_equal(env.rows()[137], before[137])
    assert np.array_equal(env.rows()[keep], ob.rows()[keep])


def test_bad_piece_ids_and_terminal_imports_are_inert():
    """Piece ids index shared-memory tables in the kernels: ids that name no piece must not get there.  Host arrays
    are rejected; ids in device tensors make the env inert (no afterstates, not stepped).  A board imported with a
    cell at or above row R is a terminal state (state.py:111-117): marked finished."""
    torch = _torch()
    from tetris_b200 import BatchedTetris
    env = BatchedTetris(10, 20, 64, piece_set=1, seed=1)
    with pytest.raises(ValueError):
        env.reset(tape=np.full(64, 9, np.uint8))
    tape = torch.full((64,), 3, dtype=torch.uint8, device="cuda")
    tape[5] = 200
    env.reset(tape=tape)
    feats, valid, count = env.get_after_states()
    assert int(count[5]) == 0 and int(valid[5]) == 0 and bool((count[torch.arange(64) != 5] == 9).all())
    st0 = env.stats.clone()
    env.rollout(10, "greedy"); env.rollout(10, "random")
    assert not env.rows()[5].any()                                     # never stepped
    assert int(env.stats[0] - st0[0]) == 63 * 20
    # terminal board
    rows = np.zeros((1, 24), np.uint16)
    rows[0, :21] = 1                                                   # column 0 filled up to row 20 = R
    env2 = BatchedTetris(10, 20, 4, piece_set=1, seed=1)
    env2.import_boards(rows, piece=np.array([0], np.uint8), first=2)
    assert env2.piece[2] == 0xFF
    f, v, c = env2.get_after_states()
    assert int(c[2]) == 0 and bool((c[[0, 1, 3]] > 0).all())


def test_combine_stats_kernel():
    """tb_combine_stats (the local half of reduce_stats) == distributed.combine_stats."""
    torch = _torch()
    from tetris_b200 import _lib, distributed as D
    g = torch.Generator().manual_seed(3)
    parts = torch.randint(0, 1 << 40, (8, len(_lib.STATS)), generator=g, dtype=torch.int64)
    dev = parts.cuda()
    out = torch.empty(len(_lib.STATS), dtype=torch.int64, device="cuda")
    _lib.check(_lib.lib().tb_combine_stats(C.c_void_p(dev.data_ptr()), 8, C.c_void_p(out.data_ptr()), None))
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), D.combine_stats(list(parts)))


class _CudaRolloutBatch:
    def __init__(self, Cc, R, n, ps):
        from tetris_b200 import BatchedTetris
        self.env = BatchedTetris(Cc, R, n, piece_set=ps, seed=0)

    def load(self, rows, piece):
        self.env.import_boards(rows, piece=piece)

    def rollout_values(self, length, n_forks, policy, seed2=0, piece_tape=None):
        mean, valid = self.env.rollout_values(length=length, n=n_forks, policy=("random", "greedy")[policy], seed=seed2,
                                              piece_tape=piece_tape)
        return np.rint(mean.cpu().numpy() * n_forks).astype(np.int64), valid.cpu().numpy().view(np.uint64)


def test_rollouts_fixture():
    """tb_rollout_values against the REFERENCE's Tetris.single_rollout (game.py:129-148): recorded parent states, every
    legal action x fork, the forks' recorded piece tapes, greedy and counter-RNG policies (tests/golden/rollouts.npz)."""
    g = load("rollouts")
    for c in rollout_configs(g):
        ret, bits = replay_rollout_config(c, _CudaRolloutBatch)
        assert np.array_equal(bits, c["valid"]), (c["C"], c["R"], c["policy"])
        assert np.array_equal(ret, c["ret_sum"]), (c["C"], c["R"], c["policy"])


def test_compact_features():
    """TB_FLAG_FEATS_I16: the int16 output is exactly 2 x the float32 features (x directions), for legal rows and --
    with include_terminal -- for terminal rows; masks and counts are unchanged.  Small and headline boards."""
    torch = _torch()
    from tetris_b200 import BatchedTetris, _lib
    for (Cc, R, n) in ((10, 20, 5000), (6, 12, 1500)):
        env = BatchedTetris(Cc, R, n, piece_set=1, seed=5)
        env.rollout(25, "random")
        n_all = torch.as_tensor([_lib.lib().tb_num_slots(p, Cc) for p in range(9)], device="cuda")[env.export_boards()[2].long()]
        slot = torch.arange(env.a_max, device="cuda")[None, :]
        for dirs in (None, [-1, -1, -1, -1, -1, -1, 1, -1]):
            env.set_directions(dirs)
            for term in (False, True):
                f, v, c = env.get_after_states(include_terminal=term)
                h, v2, c2 = env.get_after_states(include_terminal=term, compact=True)
                assert h.dtype == torch.int16 and h.shape == f.shape
                assert torch.equal(v, v2) and torch.equal(c, c2)
                mask = (slot < n_all[:, None]) if term else ((v[:, None] >> slot) & 1).bool()
                assert torch.equal(h[mask].float() * 0.5, f[mask]), (Cc, R, dirs, term)
                assert int(mask.sum()) > 5 * n
        env.set_directions(None)


def test_render_any_env_and_wide_board(capsys):
    """BatchedTetris.render(env) / board_string(env): utils.print_board_to_string (utils.py:179-191, the 4 buffer rows
    included) for any env of a batch, on a 16-column board; and the reference-shaped Tetris class on a size outside the
    built-in list (game.py:21-31 takes any num_columns / num_rows)."""
    from tetris_b200 import BatchedTetris
    env = BatchedTetris(16, 27, 50, piece_set=1, seed=3)
    env.rollout(40, "random")
    rep = env.representation(17)
    assert rep.shape == (31, 16) and rep.dtype == np.int64 and rep.any()
    text = env.board_string(17)
    lines = text.strip("\n").split("\n")
    assert len(lines) == 31 and all(len(ln) == 2 + 2 * 16 for ln in lines)
    assert [[ch != " " for ch in ln[1:-1][::2]] for ln in lines] == (rep[::-1] != 0).tolist()
    env.render(17)
    out = capsys.readouterr().out
    assert text.strip("\n") in out and any(name in out for name in ("Straight", "RCorner", "LCorner", "Square", "SnakeR", "SnakeL", "T"))
    assert np.array_equal(env.heights[17], [(np.nonzero(rep[:, c])[0].max() + 1) if rep[:, c].any() else 0 for c in range(16)])
    from tetris.game import Tetris
    np.random.seed(1)
    g = Tetris(7, 9)
    total = 0
    for _ in range(30):
        feats, _none = g.get_after_states()
        obs, rew, done, lines_ = g.step(int(np.argmax(feats.sum(axis=1))))
        total += rew
        if done:
            g.reset()
    assert g.current_state.representation.shape == (13, 7)


@pytest.mark.parametrize("k2cfg", [0, 1, 2])
def test_step_kernel_configurations(k2cfg):
    """tb_step picks its CTA shape by batch size (256 envs x 4 CTAs per SM; 128 x 8 for small batches); every shape must
    give the oracle's results.  Forced through the tuning hook on boards where many envs are within 4 rows of the top
    (the pooled legality test has work) and game overs are frequent; plus one batch big enough to take the default
    path without the hook."""
    from oracle import oracle as orc
    from tetris_b200 import BatchedTetris, _lib
    _lib.set_tuning("k2_cfg", k2cfg)
    try:
        cases = [(6, 12, 3001, 30), (10, 20, 1500, 40)] + ([(10, 10, 40000, 12)] if k2cfg == 0 else [])
        for (Cc, R, n, steps) in cases:
            if n == 40000:
                _lib.set_tuning("k2_cfg", -1)
            env = BatchedTetris(Cc, R, n, piece_set=1, seed=19)
            ob = orc.Batch(Cc, R, n, piece_set=1, seed=19)
            ob.reset()
            env.rollout(20, "random"); ob.rollout(20, 0, threads=8)
            u = np.random.RandomState(3).randint(0, 2 ** 31 - 1, size=(steps, n))
            n_done = 0
            for t in range(steps):
                _f, _v, count = env.get_after_states()
                a = (u[t] % count.cpu().numpy()).astype(np.int32)
                obs, rew, done, lines = env.step(a, auto_reset=True)
                oobs, orew, odone, olines = ob.step(a, auto_reset=True)
                assert np.array_equal(obs.cpu().numpy(), oobs) and np.array_equal(rew.cpu().numpy(), orew), (Cc, R, t)
                assert np.array_equal(done.cpu().numpy(), odone) and np.array_equal(lines.cpu().numpy(), olines), (Cc, R, t)
                n_done += int(odone.sum())
            _compare_state(env, ob)
            assert n_done > 0
    finally:
        _lib.set_tuning("k2_cfg", -1)
