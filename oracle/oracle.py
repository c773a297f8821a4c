"""ctypes wrapper over oracle/liboracle.so (the plain-C restatement of the
reference algorithm, oracle/tetris_oracle.c).  TEST INFRASTRUCTURE ONLY: the
product package ``tetris_b200`` must never import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

MAX_A = 64
STATS = ("placements", "episodes", "lines", "reward", "afterstates",
         "lines0", "lines1", "lines2", "lines3", "lines4",
         "max_ep_lines", "max_ep_steps", "sum_ep_steps", "sum_ep_lines", "reserved0", "reserved1")

PIECE_NAMES = ("Straight", "RCorner", "LCorner", "Square", "SnakeR", "SnakeL", "T", "ThreeL", "ThreeLine")
SETS = {0: (7, 8), 1: (0, 1, 2, 3, 4, 5, 6)}


def build(force=False):
    src = os.path.join(_HERE, "tetris_oracle.c")
    if force or not os.path.exists(_SO) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i32, i64, u64, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_uint32
        L.orc_num_afterstates.restype = C.c_int
        L.orc_num_afterstates.argtypes = [C.c_int, C.c_int]
        L.orc_afterstates.restype = C.c_int
        L.orc_afterstates.argtypes = [C.c_int, C.c_int, C.c_int] + [vp] * 9
        L.orc_board_features.restype = None
        L.orc_board_features.argtypes = [C.c_int, C.c_int] + [vp] * 6
        L.orc_reset_state_features.restype = None
        L.orc_reset_state_features.argtypes = [C.c_int, C.c_int, vp]
        L.orc_fitness.restype = C.c_float
        L.orc_fitness.argtypes = [vp, vp]
        L.orc_rng.restype = u32
        L.orc_rng.argtypes = [u64, u64, u32, u32]
        L.orc_batch_new.restype = vp
        L.orc_batch_new.argtypes = [C.c_int, C.c_int, C.c_int, i64, i64, u64]
        L.orc_batch_free.argtypes = [vp]
        for name in ("rep", "heights", "piece", "bag", "draws", "ep_steps", "ep_lines"):
            f = getattr(L, "orc_batch_" + name)
            f.restype = vp
            f.argtypes = [vp]
        L.orc_batch_reset.argtypes = [vp, vp]
        L.orc_batch_reset_masked.argtypes = [vp, vp, vp]
        L.orc_batch_afterstates.argtypes = [vp, C.c_int, vp, vp, vp, vp]
        L.orc_batch_step.restype = C.c_int
        L.orc_batch_step.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, vp]
        L.orc_batch_rollout_mt.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_int]
        L.orc_batch_rollout_values.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, u64, i64, vp, vp, vp]
        L.orc_stats_count.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_afterstates(piece, num_columns):
    return lib().orc_num_afterstates(piece, num_columns)


def a_max(num_columns, piece_set):
    return max(num_afterstates(p, num_columns) for p in SETS[piece_set])


def afterstates(num_columns, num_rows, piece, rep, heights=None):
    """All afterstates of `piece` on one board, in reference enumeration order.

    rep: (num_rows+4, num_columns) 0/1 array.  Returns a dict of arrays.
    """
    N = num_rows + 4
    rep = np.ascontiguousarray(rep, dtype=np.uint8).reshape(N, num_columns)
    if heights is None:
        heights = calc_lowest_free_rows(rep)
    h = np.ascontiguousarray(heights, dtype=np.int32)
    feats = np.zeros((MAX_A, 8), np.float32)
    term = np.zeros(MAX_A, np.uint8)
    ncl = np.zeros(MAX_A, np.int32)
    rep_out = np.zeros((MAX_A, N, num_columns), np.uint8)
    h_out = np.zeros((MAX_A, num_columns), np.int32)
    anchor = np.zeros((MAX_A, 2), np.int32)
    is_full = np.zeros((MAX_A, 4), np.uint8)
    n = lib().orc_afterstates(num_columns, num_rows, piece, _p(rep), _p(h), _p(feats), _p(term), _p(ncl),
                              _p(rep_out), _p(h_out), _p(anchor), _p(is_full))
    return dict(n=n, features=feats[:n], terminal=term[:n].astype(bool), n_cleared=ncl[:n],
                rep=rep_out[:n], heights=h_out[:n], anchor_col=anchor[:n, 0], anchor_row=anchor[:n, 1],
                is_full=is_full[:n].astype(bool))


def calc_lowest_free_rows(rep):
    rep = np.asarray(rep)
    N, Cc = rep.shape
    h = np.zeros(Cc, np.int32)
    for c in range(Cc):
        nz = np.nonzero(rep[:, c])[0]
        h[c] = nz[-1] + 1 if len(nz) else 0
    return h


def board_features(num_columns, num_rows, rep):
    """Features etc. of State(representation=rep) (state.py:5-38 defaults)."""
    N = num_rows + 4
    rep = np.ascontiguousarray(rep, dtype=np.uint8).reshape(N, num_columns)
    feat = np.zeros(8, np.float32)
    h = np.zeros(num_columns, np.int32)
    rep_out = np.zeros((N, num_columns), np.uint8)
    term = C.c_int32(0)
    ncl = C.c_int32(0)
    lib().orc_board_features(num_columns, num_rows, _p(rep), _p(feat), _p(h), _p(rep_out),
                             C.addressof(term), C.addressof(ncl))
    return dict(features=feat, heights=h, rep=rep_out, terminal=bool(term.value), n_cleared=ncl.value)


def reset_state_features(num_columns, num_rows):
    feat = np.zeros(8, np.float32)
    lib().orc_reset_state_features(num_columns, num_rows, _p(feat))
    return feat


def fitness(features, weights):
    f = np.ascontiguousarray(features, np.float32)
    w = np.ascontiguousarray(weights, np.float32)
    return np.float32(lib().orc_fitness(_p(f), _p(w)))


def rng(seed, env, ctr, stream):
    return lib().orc_rng(seed, env, ctr, stream)


BCTS_WEIGHTS = np.array([-24.04, -19.77, -13.08, -12.63, -10.49, -9.22, 6.6, -1.61], np.float32)


class Batch:
    """n independent envs stepped by the C oracle (plain arrays, one slot per env)."""

    def __init__(self, num_columns, num_rows, n_env, piece_set=1, seed=0, env_offset=0):
        self.C, self.R, self.N = num_columns, num_rows, num_rows + 4
        self.n, self.piece_set = int(n_env), piece_set
        self.a_max = a_max(num_columns, piece_set)
        self._b = lib().orc_batch_new(num_columns, num_rows, piece_set, self.n, env_offset, seed)
        L = lib()

        def view(fn, ctype, shape):
            ptr = C.cast(fn(self._b), C.POINTER(ctype))
            return np.ctypeslib.as_array(ptr, shape=shape)
        self.rep = view(L.orc_batch_rep, C.c_uint8, (self.n, self.N, self.C))
        self.heights = view(L.orc_batch_heights, C.c_int32, (self.n, self.C))
        self.piece = view(L.orc_batch_piece, C.c_int32, (self.n,))
        self.bag = view(L.orc_batch_bag, C.c_uint32, (self.n,))
        self.draws = view(L.orc_batch_draws, C.c_uint32, (self.n,))
        self.ep_steps = view(L.orc_batch_ep_steps, C.c_uint32, (self.n,))
        self.ep_lines = view(L.orc_batch_ep_lines, C.c_uint32, (self.n,))

    def __del__(self):
        try:
            lib().orc_batch_free(self._b)
        except Exception:
            pass

    def reset(self, tape=None):
        t = None if tape is None else np.ascontiguousarray(tape, np.uint8)
        lib().orc_batch_reset(self._b, _p(t))

    def reset_masked(self, mask, tape=None):
        m = np.ascontiguousarray(mask, np.uint8)
        t = None if tape is None else np.ascontiguousarray(tape, np.uint8)
        lib().orc_batch_reset_masked(self._b, _p(m), _p(t))

    def afterstates(self):
        """(feats[n, a_max, 8] by enumeration slot, valid mask u64[n], count[n], n_all[n])."""
        feats = np.zeros((self.n, self.a_max, 8), np.float32)
        valid = np.zeros(self.n, np.uint64)
        count = np.zeros(self.n, np.int32)
        n_all = np.zeros(self.n, np.int32)
        lib().orc_batch_afterstates(self._b, self.a_max, _p(feats), _p(valid), _p(count), _p(n_all))
        return feats, valid, count, n_all

    def step(self, actions, tape=None, auto_reset=False, action_is_slot=False):
        a = np.ascontiguousarray(actions, np.int32)
        t = None if tape is None else np.ascontiguousarray(tape, np.uint8)
        obs = np.zeros((self.n, 8), np.float32)
        reward = np.zeros(self.n, np.int32)
        done = np.zeros(self.n, np.uint8)
        lines = np.zeros(self.n, np.int32)
        rc = lib().orc_batch_step(self._b, _p(a), int(action_is_slot), _p(t), int(auto_reset),
                                  _p(obs), _p(reward), _p(done), _p(lines))
        if rc:
            raise IndexError("action out of range")
        return obs, reward, done.astype(bool), lines

    def rollout(self, T, policy, weights=None, threads=1):
        w = np.ascontiguousarray(BCTS_WEIGHTS if weights is None else weights, np.float32)
        stats = np.zeros(lib().orc_stats_count(), np.int64)
        lib().orc_batch_rollout_mt(self._b, int(T), int(policy), _p(w), _p(stats), int(threads))
        return stats

    def rollout_values(self, length, n_forks, policy, weights=None, seed2=0, child_offset=0, piece_tape=None):
        """(sum of fork returns int32[n, a_max], legal-slot mask u64[n]) -- game.py:129-160 for every env x action.
        piece_tape: optional uint8[n, a_max, n_forks, length], the pieces each fork draws (a recorded sampler)."""
        w = np.ascontiguousarray(BCTS_WEIGHTS if weights is None else weights, np.float32)
        ret = np.zeros((self.n, self.a_max), np.int32)
        valid = np.zeros(self.n, np.uint64)
        tape = None if piece_tape is None else np.ascontiguousarray(piece_tape, np.uint8)
        assert tape is None or tape.size == self.n * self.a_max * int(n_forks) * int(length)
        lib().orc_batch_rollout_values(self._b, self.a_max, int(n_forks), int(length), int(policy), _p(w),
                                       int(seed2), int(child_offset), _p(tape), _p(ret), _p(valid))
        return ret, valid

    def rows(self):
        """Row masks u16[n, N]: bit c of rows[e, r] = rep[e, r, c]."""
        w = (1 << np.arange(self.C)).astype(np.uint32)
        return (self.rep.astype(np.uint32) * w).sum(axis=2).astype(np.uint16)
