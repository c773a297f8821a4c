#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the LIVE Python reference.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so every pin
is produced here by importing the unmodified reference (SURVEY.md Appendix C.1 recipe:
expose the checkout as package ``tetris``, stub matplotlib) and recording its outputs on
seeded inputs.  The reference cannot travel to the GPU box, so the vectors are committed
next to this script; tests only ever read the .npz files.

Fixtures:
  afterstates.npz   per shape x piece x board: every afterstate's features, terminal flag,
                    lines cleared, board, heights, anchor, cleared-row flags
  trace_*.npz       lockstep traces of N reference Tetris objects driven by a piece tape
                    (injected through the replaceable ``tetromino_sampler`` attribute,
                    game.py:50) and an action rule, recorded after every step
  known_answer.npz  the corrected example_play.py flow under np.random.seed(0) with the
                    reference's own sampler
  fitness.npz       Tetris.fitness / get_best_policy outputs (game.py:102-120)
"""
import hashlib
import importlib
import importlib.machinery
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("TETRIS_REFERENCE", "/root/reference")

PIECE_NAMES = ("Straight", "RCorner", "LCorner", "Square", "SnakeR", "SnakeL", "T", "ThreeL", "ThreeLine")
SETS = {0: (7, 8), 1: (0, 1, 2, 3, 4, 5, 6)}
DIRECTIONS = np.array([-1, -1, -1, -1, -1, -1, 1, -1])


def load_reference(path=REF):
    """Import the reference as a private package (Appendix C.1) and return its modules."""
    for n in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(n, types.ModuleType(n))
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "tetris" or k.startswith("tetris.")}
    spec = importlib.machinery.ModuleSpec("tetris", None, is_package=True)
    spec.submodule_search_locations = [path]
    sys.modules["tetris"] = importlib.util.module_from_spec(spec)
    game = importlib.import_module("tetris.game")
    mods = dict(game=game, state=sys.modules["tetris.state"], tetromino=sys.modules["tetris.tetromino"],
                utils=sys.modules["tetris.utils"])
    for k in [k for k in sys.modules if k == "tetris" or k.startswith("tetris.")]:
        sys.modules["ref_" + k] = sys.modules.pop(k)
    sys.modules.update(saved)
    return mods


def make_pieces(ref, C):
    t = ref["tetromino"]
    return [getattr(t, n)("bcts", 8, C) for n in PIECE_NAMES]


def piece_id(piece):
    return PIECE_NAMES.index(type(piece).__name__)


class TapeSampler:
    """Drop-in for TetrominoSampler (tetromino.py:12-22): pieces come from a callback."""

    def __init__(self, pieces, next_id):
        self.pieces, self.next_id = pieces, next_id

    def next_tetromino(self):
        return self.pieces[self.next_id()]


def rows_of(rep):
    rep = np.asarray(rep)
    w = (1 << np.arange(rep.shape[1])).astype(np.int64)
    return (rep.astype(np.int64) * w).sum(axis=1).astype(np.uint16)


def feat2(f):
    """Features doubled -> exact small integers (landing height is a half-integer)."""
    v = np.asarray(f, np.float64) * 2
    r = np.rint(v)
    assert np.all(r == v) and np.all(np.abs(r) < 32768)
    return r.astype(np.int16)


# ---------------------------------------------------------------------------------------------
# bag RNG of the new framework restated in Python (oracle/tetris_oracle.c bag_draw); used only to
# produce piece tapes that the RNG-mode tests can reproduce from (seed, env id).
# ---------------------------------------------------------------------------------------------
M64 = (1 << 64) - 1


def mix64(z):
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def rng32(seed, env, ctr, stream):
    k = mix64((seed + 0x9E3779B97F4A7C15 * (env + 1)) & M64)
    return mix64(k ^ ((stream << 32) | ctr)) >> 32


class Bag:
    def __init__(self, n_set, seed, env):
        self.n, self.seed, self.env, self.bag, self.draws = n_set, seed, env, 0, 0

    def draw(self):
        if self.bag == 0:
            self.bag = (1 << self.n) - 1
        k = bin(self.bag).count("1")
        j = (rng32(self.seed, self.env, self.draws, 0) * k) >> 32
        self.draws += 1
        b = self.bag
        for _ in range(j):
            b &= b - 1
        idx = (b & -b).bit_length() - 1
        self.bag &= ~(1 << idx)
        return idx


# ---------------------------------------------------------------------------------------------
def random_board(rng, C, R, fill):
    """Arbitrary (not necessarily reachable) board with all heights <= R and row 0 not full."""
    N = R + 4
    rep = np.zeros((N, C), np.int64)
    hmax = rng.integers(0, R + 1)
    for c in range(C):
        h = rng.integers(0, hmax + 1)
        if h:
            rep[:h, c] = rng.random(h) < fill
            rep[h - 1, c] = 1
    if rep[0].all():
        rep[0, rng.integers(0, C)] = 0
        # keep heights consistent: a column emptied at row 0 only is still fine
    return rep


def gen_afterstates(ref):
    st = ref["state"]
    rng = np.random.default_rng(1234)
    out = {k: [] for k in ("shape", "piece", "rows", "heights", "start", "count")}
    per = {k: [] for k in ("feat2", "terminal", "n_cleared", "rows", "heights", "anchor", "is_full")}
    total = 0
    for (C, R) in ((10, 20), (10, 10), (6, 12)):
        pieces = make_pieces(ref, C)
        boards = []
        # (a) boards reached by random play of the reference itself (7-piece set)
        env = ref["game"].Tetris(C, R)
        env.tetrominos = [pieces[i] for i in SETS[1]]
        env.tetromino_sampler = TapeSampler(env.tetrominos, lambda: int(rng.integers(0, 7)))
        env.reset()
        while len(boards) < 40:
            feats, _ = env.get_after_states()
            _, _, done, _ = env.step(int(rng.integers(0, len(feats))))
            if done:
                env.reset()
            elif rng.random() < 0.5:
                boards.append(env.current_state.representation.copy())
        # (b) arbitrary boards (holes, overhangs, occasionally pre-filled rows)
        for fill in (0.5, 0.75, 0.9, 0.97):
            for _ in range(6):
                boards.append(random_board(rng, C, R, fill))
        boards.append(np.zeros((R + 4, C), np.int64))
        for rep in boards:
            h = st.calc_lowest_free_rows(rep)
            assert h.max() <= R
            for pid, piece in enumerate(pieces):
                base = st.State(representation=rep.copy(), lowest_free_rows=h.copy())
                assert not rep[0].all()
                children = piece.get_after_states(base)
                out["shape"].append((C, R)); out["piece"].append(pid)
                out["rows"].append(np.pad(rows_of(rep), (0, 24 - (R + 4))))
                out["heights"].append(np.pad(h, (0, 10 - C)))
                out["start"].append(total); out["count"].append(len(children))
                for ch in children:
                    per["feat2"].append(feat2(ch.get_features()))
                    per["terminal"].append(ch.terminal_state)
                    per["n_cleared"].append(ch.n_cleared_lines)
                    per["rows"].append(np.pad(rows_of(ch.representation), (0, 24 - (R + 4))))
                    per["heights"].append(np.pad(ch.lowest_free_rows, (0, 10 - C)))
                    per["anchor"].append((ch.anchor_col, ch.anchor_row))
                    per["is_full"].append(np.pad(np.asarray(ch.cleared_rows_relative_to_anchor, bool),
                                                 (0, 4 - len(ch.cleared_rows_relative_to_anchor))))
                total += len(children)
    np.savez_compressed(
        os.path.join(HERE, "afterstates.npz"),
        shape=np.array(out["shape"], np.int16), piece=np.array(out["piece"], np.int8),
        rows=np.array(out["rows"], np.uint16), heights=np.array(out["heights"], np.int8),
        start=np.array(out["start"], np.int32), count=np.array(out["count"], np.int16),
        a_feat2=np.array(per["feat2"], np.int16), a_terminal=np.array(per["terminal"], bool),
        a_n_cleared=np.array(per["n_cleared"], np.int8), a_rows=np.array(per["rows"], np.uint16),
        a_heights=np.array(per["heights"], np.int8), a_anchor=np.array(per["anchor"], np.int8),
        a_is_full=np.array(per["is_full"], bool))
    print("afterstates.npz:", len(out["piece"]), "boards x pieces,", total, "afterstates,",
          int(np.sum(per["terminal"])), "terminal,", int(np.sum(np.array(per["n_cleared"]) > 0)), "with clears")


def dense_board(rng, C, R, kind):
    """Near-full stacks (line clears of every multiplicity, overflow rescued by a clear): H rows that are
    full but for one hole each, topped by d rows that are full but for a gap of gw columns at the same place -- the
    piece that fits the gap clears up to d lines.  kind 'tall' puts the stack top within 3 rows of R, so most
    placements poke above row R and are terminal unless they clear enough lines."""
    N = R + 4
    rep = np.zeros((N, C), np.int64)
    d = int(rng.choice([1, 2, 3, 4], p=[0.2, 0.2, 0.2, 0.4]))
    d = min(d, R)
    if kind == "tall":
        H = int(rng.integers(max(d, R - 3), R + 1))
    else:
        H = int(rng.integers(d, max(d, R - 2) + 1))
    gw = int(rng.choice([1, 2, 3], p=[0.6, 0.3, 0.1]))
    g = int(rng.integers(0, C - gw + 1))
    rep[:H] = 1
    for r in range(H - d):                       # lower rows: one hole each, anywhere (never full)
        rep[r, rng.integers(0, C)] = 0
    for r in range(H - d, H):
        rep[r, g:g + gw] = 0
        if rng.random() < 0.1:                   # now and then a second gap: that row does not clear
            rep[r, rng.integers(0, C)] = 0
    # ragged top: a few extra cells above the stack, outside the gap (heights stay <= R)
    for c in range(C):
        if not (g <= c < g + gw) and rng.random() < 0.3:
            extra = int(rng.integers(1, 3))
            rep[H:min(R, H + extra), c] = 1
    # the gap columns keep whatever is below the gap; sometimes dig the gap deeper (a hole column) or roof part of it
    if gw == 2 and rng.random() < 0.3 and H - d > 0:
        rep[H - d - 1, g] = 0
    return rep


def gen_afterstates_dense(ref):
    """afterstates_dense.npz: > 100 k afterstates on near-full boards of eight shapes (the three of BASELINE.json,
    8x16, 4x4, and 16x27 / 12x24 / 7x9 beyond the usual sizes), all nine pieces -- hundreds of multi-line clears incl. four-line clears, terminal placements
    rescued by a clear, terminal-with-clear.  Rows are padded to 32, columns to 16."""
    st = ref["state"]
    rng = np.random.default_rng(4321)
    out = {k: [] for k in ("shape", "piece", "rows", "heights", "start", "count")}
    per = {k: [] for k in ("feat2", "terminal", "n_cleared", "rows", "heights", "anchor", "is_full")}
    total = 0
    PADN, PADC = 32, 16
    for (C, R), n_boards in (((10, 20), 150), ((10, 10), 110), ((8, 16), 110), ((6, 12), 130), ((4, 4), 120),
                             ((16, 27), 40), ((12, 24), 50), ((7, 9), 60)):
        pieces = make_pieces(ref, C)
        boards = [dense_board(rng, C, R, "tall" if k % 2 else "mid") for k in range(n_boards)]
        for rep in boards:
            h = st.calc_lowest_free_rows(rep)
            assert h.max() <= R
            for pid, piece in enumerate(pieces):
                base = st.State(representation=rep.copy(), lowest_free_rows=h.copy())
                children = piece.get_after_states(base)
                out["shape"].append((C, R)); out["piece"].append(pid)
                out["rows"].append(np.pad(rows_of(rep), (0, PADN - (R + 4))))
                out["heights"].append(np.pad(h, (0, PADC - C)))
                out["start"].append(total); out["count"].append(len(children))
                for ch in children:
                    per["feat2"].append(feat2(ch.get_features()))
                    per["terminal"].append(ch.terminal_state)
                    per["n_cleared"].append(ch.n_cleared_lines)
                    per["rows"].append(np.pad(rows_of(ch.representation), (0, PADN - (R + 4))))
                    per["heights"].append(np.pad(ch.lowest_free_rows, (0, PADC - C)))
                    per["anchor"].append((ch.anchor_col, ch.anchor_row))
                    per["is_full"].append(np.pad(np.asarray(ch.cleared_rows_relative_to_anchor, bool),
                                                 (0, 4 - len(ch.cleared_rows_relative_to_anchor))))
                total += len(children)
    ncl = np.array(per["n_cleared"])
    term = np.array(per["terminal"])
    R_of = np.repeat(np.array(out["shape"])[:, 1], np.array(out["count"]))
    rescued = ~term & (np.array(per["heights"]).max(axis=1) + ncl > R_of)     # poked above row R before clearing
    np.savez_compressed(
        os.path.join(HERE, "afterstates_dense.npz"),
        shape=np.array(out["shape"], np.int16), piece=np.array(out["piece"], np.int8),
        rows=np.array(out["rows"], np.uint16), heights=np.array(out["heights"], np.int8),
        start=np.array(out["start"], np.int32), count=np.array(out["count"], np.int16),
        a_feat2=np.array(per["feat2"], np.int16), a_terminal=term,
        a_n_cleared=ncl.astype(np.int8), a_rows=np.array(per["rows"], np.uint16),
        a_heights=np.array(per["heights"], np.int8), a_anchor=np.array(per["anchor"], np.int8),
        a_is_full=np.array(per["is_full"], bool))
    print("afterstates_dense.npz:", len(out["piece"]), "boards x pieces,", total, "afterstates,", int(term.sum()),
          "terminal; clears 1/2/3/4:", [int((ncl == k).sum()) for k in (1, 2, 3, 4)],
          "; terminal with clear:", int((term & (ncl > 0)).sum()),
          "; overflow rescued by a clear:", int(rescued.sum()))


def gen_trace(ref, name, C, R, piece_set, n_env, n_steps, policy, directions, seed):
    """policy: 'random' (a = u_t mod n_valid) or 'greedy' (first argmax of Tetris.fitness, float32)."""
    pieces = make_pieces(ref, C)
    ids = SETS[piece_set]
    a_max = max(len(pieces[i].get_after_states(
        ref["state"].State(np.zeros((R + 4, C), np.int_), np.zeros(C, np.int_)))) for i in ids)
    envs, bags, tapes = [], [], [[] for _ in range(n_env)]
    for e in range(n_env):
        bag = Bag(len(ids), seed, e)
        bags.append(bag)

        def nxt(bag=bag, e=e):
            gid = ids[bag.draw()]
            tapes[e].append(gid)
            return gid
        env = ref["game"].Tetris(C, R, feature_directions=directions)   # draws from NumPy's global RNG: discarded
        env.tetrominos = [pieces[i] for i in range(9)]                  # indexable by global id
        env.tetromino_sampler = TapeSampler(env.tetrominos, nxt)
        env.reset()                                                     # first tape draw == construction draw
        envs.append(env)
    u = np.random.RandomState(seed).randint(0, 2 ** 31 - 1, size=(n_steps, n_env))
    rec = dict(piece=np.zeros((n_steps, n_env), np.int8), next_piece=np.zeros((n_steps, n_env), np.int8),
               reset_piece=np.full((n_steps, n_env), -1, np.int8),
               n_valid=np.zeros((n_steps, n_env), np.int16), n_all=np.zeros((n_steps, n_env), np.int16),
               action=np.zeros((n_steps, n_env), np.int16), reward=np.zeros((n_steps, n_env), np.int16),
               done=np.zeros((n_steps, n_env), bool), lines=np.zeros((n_steps, n_env), np.int8),
               obs2=np.zeros((n_steps, n_env, 8), np.int16),
               rows=np.zeros((n_steps, n_env, R + 4), np.uint16), heights=np.zeros((n_steps, n_env, C), np.int8),
               feat2=np.zeros((n_steps, n_env, a_max, 8), np.int16),
               valid=np.zeros((n_steps, n_env, a_max), bool))
    for t in range(n_steps):
        for e, env in enumerate(envs):
            rec["piece"][t, e] = piece_id(env.current_tetromino)
            feats, allf = env.get_after_states(include_terminal=True)
            nv, na = len(feats), len(allf)
            assert nv > 0
            rec["n_valid"][t, e], rec["n_all"][t, e] = nv, na
            und = allf * directions if directions is not None else allf   # directions are +-1: undo them
            rec["feat2"][t, e, :na] = feat2(und)
            term = np.array([c.terminal_state for c in env.current_tetromino.get_after_states(env.current_state)])
            rec["valid"][t, e, :na] = ~term
            if policy == "random":
                a = int(u[t, e] % nv)
            else:
                scores = np.array([env.fitness(s) for s in env.afterstates])
                assert scores.dtype == np.float32
                a = int(np.argmax(scores))
            obs, rew, done, lines = env.step(a)
            rec["action"][t, e], rec["reward"][t, e], rec["done"][t, e], rec["lines"][t, e] = a, rew, done, lines
            rec["next_piece"][t, e] = piece_id(env.current_tetromino)
            und = obs * directions if directions is not None else obs
            rec["obs2"][t, e] = feat2(und)
            rec["rows"][t, e] = rows_of(env.current_state.representation)
            rec["heights"][t, e] = env.current_state.lowest_free_rows
            if done:
                env.reset()
                rec["reset_piece"][t, e] = piece_id(env.current_tetromino)
    np.savez_compressed(os.path.join(HERE, "trace_%s.npz" % name), C=C, R=R, piece_set=piece_set, seed=seed,
                        policy=policy, a_max=a_max,
                        directions=np.zeros(0) if directions is None else directions, **rec)
    print("trace_%s.npz: %d envs x %d steps, %d game-overs, %d lines" % (
        name, n_env, n_steps, rec["done"].sum(), rec["lines"].sum()))


def gen_known_answer(ref):
    """corrected example_play.py flow, reference sampler, np.random.seed(0)."""
    np.random.seed(0)
    env = ref["game"].Tetris(10, 10, feature_directions=DIRECTIONS)
    env.reset()
    n = 100
    rec = dict(piece=np.zeros(n, np.int8), n_valid=np.zeros(n, np.int16), action=np.zeros(n, np.int16),
               reward=np.zeros(n, np.int16), lines=np.zeros(n, np.int8), done=np.zeros(n, bool),
               obs=np.zeros((n, 8), np.float64), next_piece=np.zeros(n, np.int8),
               reset_piece=np.full(n, -1, np.int8))
    hsh = hashlib.sha256()
    for t in range(n):
        rec["piece"][t] = piece_id(env.current_tetromino)
        feats, _ = env.get_after_states()
        i = int(np.argmax(feats.sum(axis=1)))
        obs, rew, done, lines = env.step(i)
        rec["n_valid"][t], rec["action"][t], rec["reward"][t] = len(feats), i, rew
        rec["lines"][t], rec["done"][t], rec["obs"][t] = lines, done, obs
        rec["next_piece"][t] = piece_id(env.current_tetromino)
        hsh.update(env.current_state.representation.astype(np.uint8).tobytes())
        hsh.update(np.asarray(obs, np.float64).tobytes())
        if done:
            env.reset()
            rec["reset_piece"][t] = piece_id(env.current_tetromino)
    digest = hsh.hexdigest()
    assert digest.startswith("253964af6c665fa0"), digest      # the value the survey recorded
    np.savez_compressed(os.path.join(HERE, "known_answer.npz"), sha256=digest,
                        final_rows=rows_of(env.current_state.representation),
                        final_heights=np.asarray(env.current_state.lowest_free_rows, np.int8),
                        first_piece=rec["piece"][0], **rec)
    print("known_answer.npz: sum reward", rec["reward"].sum(), "sha256", digest[:16])


def gen_fitness(ref):
    rng = np.random.default_rng(7)
    C, R = 10, 20
    pieces = make_pieces(ref, C)
    env = ref["game"].Tetris(C, R)
    env.tetrominos = [pieces[i] for i in SETS[1]]
    env.tetromino_sampler = TapeSampler(env.tetrominos, lambda: int(rng.integers(0, 7)))
    env.reset()
    rows, piece, start, count, fit, pol, f2 = [], [], [], [], [], [], []
    total = 0
    for t in range(60):
        feats, _ = env.get_after_states()
        children = env.current_tetromino.get_after_states(env.current_state)
        fv = np.array([env.fitness(c) for c in children])
        assert fv.dtype == np.float32
        rows.append(rows_of(env.current_state.representation)); piece.append(piece_id(env.current_tetromino))
        start.append(total); count.append(len(children)); total += len(children)
        fit.extend(fv); pol.extend(env.get_best_policy()); f2.extend(feat2(c.get_features()) for c in children)
        _, _, done, _ = env.step(int(rng.integers(0, len(feats))))
        if done:
            env.reset()
    np.savez_compressed(os.path.join(HERE, "fitness.npz"), C=C, R=R, rows=np.array(rows, np.uint16),
                        piece=np.array(piece, np.int8), start=np.array(start, np.int32),
                        count=np.array(count, np.int16), fitness=np.array(fit, np.float32),
                        best_policy=np.array(pol, np.float64), feat2=np.array(f2, np.int16))
    print("fitness.npz:", total, "afterstates")


BCTS = np.array([-24.04, -19.77, -13.08, -12.63, -10.49, -9.22, 6.6, -1.61], np.float32)


def greedy_policy_function(state, feats):
    """policy_function(state, action_features) for Tetris.single_rollout (game.py:139-140): first arg-max of the
    float32, left-to-right BCTS score of game.py:109-120 over the non-terminal afterstates' features."""
    f = np.asarray(feats).astype(np.float32)
    acc = f[:, 0] * BCTS[0]
    for i in range(1, 8):
        acc = acc + f[:, i] * BCTS[i]
    assert acc.dtype == np.float32
    return int(np.argmax(acc))


class ListSampler:
    """tetromino_sampler whose pieces come from a list that the harness swaps per rollout."""

    def __init__(self, pieces):
        self.pieces, self.tape, self.pos = pieces, [], 0

    def load(self, tape):
        self.tape, self.pos = list(tape), 0

    def next_tetromino(self):
        p = self.pieces[self.tape[self.pos]]
        self.pos += 1
        return p


def gen_rollouts(ref):
    """rollouts.npz: Tetris.single_rollout / perform_rollouts (game.py:129-160) of the reference itself.

    Part A (per config `c<k>_*`): for P parent states, every legal action and n forks, single_rollout() with a
    deterministic policy_function and the fork's own piece tape (injected through the replaceable sampler);
    get_after_states() is called on the parent before every rollout, as a caller must for `action` to index the
    parent's afterstates.  Recorded per enumeration slot: the sum of the forks' returns, legality, the tapes.
    Policies: 'greedy' (above) and 'random' -- the new framework's per-fork counter RNG restated here
    (stream 1 of (seed2, child id), counter = the fork's draw count), so the device can replay it.

    Part B (`s<k>_*`): perform_rollouts() called as shipped, ONE get_after_states() before it, pieces from one
    sequential tape: from the second rollout on `self.afterstates` is the list left behind by the previous rollout's
    last policy step (game.py:140 overwrites it, :147-148 do not restore it), so `action` indexes a stale list.  The
    compatibility class must reproduce these numbers to be a drop-in."""
    st = ref["state"]
    rng = np.random.default_rng(99)
    out = {}
    configs = [(10, 20, 1, "greedy", 4, 3, 1234), (10, 20, 1, "random", 4, 3, 77), (6, 12, 1, "greedy", 5, 2, 5),
               (6, 12, 1, "random", 3, 3, 6), (10, 10, 0, "greedy", 4, 2, 8), (8, 16, 1, "greedy", 3, 2, 9)]
    for k, (C, R, ps, policy, length, n_forks, seed2) in enumerate(configs):
        pieces = make_pieces(ref, C)
        ids = SETS[ps]
        env = ref["game"].Tetris(C, R)
        env.tetrominos = [pieces[i] for i in range(9)]
        sampler = ListSampler(env.tetrominos)
        env.tetromino_sampler = sampler
        # parents: states reached by random play, sampled late enough that some rollouts end the game
        parents = []
        sampler.load([ids[int(x)] for x in rng.integers(0, len(ids), 100000)])
        env.reset()
        want = 10 if C * R >= 200 else 14
        while len(parents) < want:
            feats, _ = env.get_after_states()
            _, _, done, _ = env.step(int(rng.integers(0, len(feats))))
            if done:
                env.reset()
            elif env.current_state.lowest_free_rows.max() >= R - 7 and rng.random() < 0.4:
                parents.append((env.current_state, env.current_tetromino))
        a_max = max(len(pieces[i].get_after_states(st.State(np.zeros((R + 4, C), np.int_), np.zeros(C, np.int_))))
                    for i in ids)
        P = len(parents)
        tape = np.zeros((P, a_max, n_forks, length), np.uint8)
        ret_sum = np.zeros((P, a_max), np.int32)
        valid = np.zeros((P, a_max), bool)
        n_ended = 0
        for p, (pstate, ppiece) in enumerate(parents):
            children = ppiece.get_after_states(pstate)
            legal = [s for s, ch in enumerate(children) if not ch.terminal_state]
            for a, s in enumerate(legal):
                valid[p, s] = True
                for f in range(n_forks):
                    tp = [ids[int(x)] for x in rng.integers(0, len(ids), length)]
                    tape[p, s, f] = tp
                    sampler.load(tp)
                    env.current_state, env.current_tetromino = pstate, ppiece
                    env.get_after_states()
                    if policy == "greedy":
                        fn = greedy_policy_function
                    else:
                        d = (p * a_max + s) * n_forks + f
                        calls = [0]

                        def fn(state, feats, d=d, calls=calls):
                            # parent draws = 1 (its reset draw); the action's draw makes 2; + one per policy step
                            r = rng32(seed2, d, 2 + calls[0], 1)
                            calls[0] += 1
                            return (r * len(feats)) >> 32
                    r = env.single_rollout(a, fn, length)
                    assert env.current_state is pstate and env.current_tetromino is ppiece
                    ret_sum[p, s] += r
                    n_ended += (r == -1)
        pre = "c%d_" % k
        out.update({pre + "C": C, pre + "R": R, pre + "piece_set": ps, pre + "policy": policy, pre + "length": length,
                    pre + "n_forks": n_forks, pre + "seed2": seed2,
                    pre + "rows": np.array([np.pad(rows_of(s.representation), (0, 32 - (R + 4))) for s, _ in parents], np.uint16),
                    pre + "piece": np.array([piece_id(t) for _, t in parents], np.uint8),
                    pre + "tape": tape, pre + "ret_sum": ret_sum, pre + "valid": valid})
        print("rollouts c%d: %dx%d set %d %s length %d forks %d: %d parents, %d legal actions, %d rollouts ended the game"
              % (k, C, R, ps, policy, length, n_forks, P, int(valid.sum()), n_ended))
    out["n_configs"] = len(configs)

    # ---- part B: perform_rollouts as shipped (stale self.afterstates from the second rollout on)
    stale = [(10, 20, 1, 4, 2), (6, 12, 1, 3, 3), (10, 10, 0, 4, 2)]
    for k, (C, R, ps, length, n) in enumerate(stale):
        pieces = make_pieces(ref, C)
        ids = SETS[ps]
        env = ref["game"].Tetris(C, R)
        env.tetrominos = [pieces[i] for i in range(9)]
        sampler = ListSampler(env.tetrominos)
        env.tetromino_sampler = sampler
        seq = [ids[int(x)] for x in rng.integers(0, len(ids), 4000)]
        sampler.load(seq)
        env.reset()
        for _ in range(14 if R >= 16 else 8):
            feats, _ = env.get_after_states()
            _, _, done, _ = env.step(int(rng.integers(0, len(feats))))
            assert not done
        start_pos = sampler.pos
        feats, _ = env.get_after_states()
        # perform_rollouts (game.py:150-160) call by call, so that a failure of the reference itself -- the stale list
        # can be shorter than the action index: IndexError at game.py:83 -- is recorded where it happens
        n_act = min(len(feats), 4)
        calls, err_at = [], -1
        for a in range(n_act):
            for i in range(n):
                try:
                    calls.append(env.single_rollout(a, greedy_policy_function, length))
                except IndexError:
                    err_at = len(calls)
                    break
            if err_at >= 0:
                break
        pre = "s%d_" % k
        out.update({pre + "C": C, pre + "R": R, pre + "piece_set": ps, pre + "length": length, pre + "n": n,
                    pre + "rows": rows_of(env.current_state.representation),
                    pre + "piece": piece_id(env.current_tetromino), pre + "n_actions": n_act,
                    pre + "tape": np.array(seq[start_pos:sampler.pos], np.uint8),
                    pre + "returns": np.array(calls, np.int32), pre + "error_at": err_at})
        print("rollouts s%d: single_rollout calls of perform_rollouts as shipped, %dx%d: returns %s, IndexError at call %d "
              "(%d pieces drawn)" % (k, C, R, calls, err_at, sampler.pos - start_pos))
    out["n_stale"] = len(stale)
    np.savez_compressed(os.path.join(HERE, "rollouts.npz"), **out)


def gen_learner(ref):
    """The learner math of the reference (utils.py:26-45) on seeded inputs: compute_action_probabilities,
    grad_of_log_action_probabilities and softmax, per env over its legal afterstates (slot order), for three temperatures.
    Features are float32 half-integers in the range the board features take; legal-slot masks are arbitrary subsets."""
    u = ref["utils"]
    rng = np.random.RandomState(20260)
    n, a_max = 96, 34
    feats = (rng.randint(-40, 400, size=(n, a_max, 8)) * 0.5).astype(np.float32)
    feats[:, :, 6] = rng.randint(0, 17, size=(n, a_max))                      # eroded cells: small integers
    valid = np.zeros(n, np.uint64)
    for e in range(n):
        k = 1 + rng.randint(a_max) if e else a_max                             # env 0: every slot legal
        for sl in rng.choice(a_max, size=k, replace=False):
            valid[e] |= np.uint64(1) << np.uint64(sl)
    weights = np.array([-2.4, -1.9, -1.3, -1.2, -1.0, -0.9, 0.6, -0.1])
    temps = np.array([1.0, 0.25, 7.5])
    bits = ((valid[:, None] >> np.arange(a_max, dtype=np.uint64)) & np.uint64(1)).astype(bool)
    actions = np.array([rng.choice(np.nonzero(b)[0]) for b in bits], np.int32)      # enumeration slot of the chosen action
    probs = np.zeros((len(temps), n, a_max))
    grads = np.zeros((len(temps), n, 8))
    soft = np.zeros((n, a_max))
    for e in range(n):
        fe = feats[e][bits[e]].astype(np.float64)
        k = int(np.nonzero(np.nonzero(bits[e])[0] == actions[e])[0][0])
        for ti, t in enumerate(temps):
            pr = u.compute_action_probabilities(fe, weights, t)
            probs[ti, e][bits[e]] = pr
            grads[ti, e] = u.grad_of_log_action_probabilities(fe, pr, k)
        soft[e][bits[e]] = u.softmax(fe.dot(weights))
    np.savez_compressed(os.path.join(HERE, "learner.npz"), feats=feats, valid=valid, weights=weights, temps=temps,
                        actions=actions, probs=probs, grads=grads, softmax=soft)
    print("learner.npz: %d envs x %d temperatures" % (n, len(temps)))


def main():
    ref = load_reference()
    only = sys.argv[1:]
    if only:                                     # e.g. `make_golden.py afterstates_dense rollouts`
        for name in only:
            globals()["gen_" + name](ref)
        return
    gen_known_answer(ref)
    gen_afterstates(ref)
    gen_afterstates_dense(ref)
    gen_rollouts(ref)
    gen_fitness(ref)
    gen_learner(ref)
    gen_trace(ref, "7p_10x20_random", 10, 20, 1, 32, 160, "random", None, 0x5EED)
    gen_trace(ref, "2p_10x10_random_dir", 10, 10, 0, 16, 120, "random", DIRECTIONS, 11)
    gen_trace(ref, "7p_6x12_random", 6, 12, 1, 16, 120, "random", None, 12)
    gen_trace(ref, "7p_10x20_greedy", 10, 20, 1, 8, 150, "greedy", None, 13)
    gen_trace(ref, "7p_6x12_greedy", 6, 12, 1, 16, 250, "greedy", None, 14)


if __name__ == "__main__":
    main()
