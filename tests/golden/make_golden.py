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


def main():
    ref = load_reference()
    gen_known_answer(ref)
    gen_afterstates(ref)
    gen_fitness(ref)
    gen_trace(ref, "7p_10x20_random", 10, 20, 1, 32, 160, "random", None, 0x5EED)
    gen_trace(ref, "2p_10x10_random_dir", 10, 10, 0, 16, 120, "random", DIRECTIONS, 11)
    gen_trace(ref, "7p_6x12_random", 6, 12, 1, 16, 120, "random", None, 12)
    gen_trace(ref, "7p_10x20_greedy", 10, 20, 1, 8, 150, "greedy", None, 13)
    gen_trace(ref, "7p_6x12_greedy", 6, 12, 1, 16, 250, "greedy", None, 14)


if __name__ == "__main__":
    main()
