"""BatchedTetris: n lockstep envs on one GPU, tensors in / tensors out.

Semantics per env are exactly the reference's Tetris (game.py:8-127): `get_after_states` enumerates the
current piece's placements in the reference's action order, `step(a)` takes the a-th NON-terminal afterstate
(game.py:69,83), rewards are lines-1 (-100 more on game over), a finished env is reset by the caller
(`reset_masked`) or in place (`auto_reset=True`, what example_play.py:20-21 does).
All compute runs in the CUDA kernels of csrc/tb_kernels.cu; torch only owns memory and streams.
"""
import contextlib
import ctypes as C

import numpy as np
import torch

from . import BCTS_WEIGHTS, _lib

_PIECE_SET_NAMES = {"default": 0, "threes": 0, 0: 0, "tetrominoes": 1, "seven": 1, 1: 1}


def _lib_piece_names():
    from . import PIECE_NAMES
    return PIECE_NAMES


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class BatchedTetris:
    def __init__(self, num_columns, num_rows, n_env, piece_set=1, seed=0, env_offset=0, device=None,
                 feature_directions=None):
        L = _lib.lib()
        if not torch.cuda.is_available():
            raise RuntimeError("tetris_b200: no CUDA device; the batched environment has no CPU path")
        self.num_columns, self.num_rows = int(num_columns), int(num_rows)
        self.n_env, self.seed, self.env_offset = int(n_env), int(seed) & (2 ** 64 - 1), int(env_offset)
        self.piece_set = _PIECE_SET_NAMES[piece_set]
        _lib.ensure_shape(self.num_columns, self.num_rows)   # built in, or compiled on demand (game.py:21-31: any size)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.a_max = L.tb_a_max(self.num_columns, self.piece_set)
        self.n_stored_rows = self.num_rows + 4
        with self._on_device():
            self.state = torch.zeros(L.tb_state_bytes(self.num_columns, self.num_rows, self.n_env),
                                     dtype=torch.uint8, device=self.device)
            self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.stats = torch.zeros(len(_lib.STATS), dtype=torch.int64, device=self.device)
        self.set_directions(feature_directions)
        self.reset()

    # -- plumbing ---------------------------------------------------------------------------
    def _on_device(self):
        """Context that makes self.device current; free when it already is (the common single-GPU-per-process case)."""
        if torch.cuda.current_device() == self._dev_index:
            return contextlib.nullcontext()
        return torch.cuda.device(self.device)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _common(self):
        return (_ptr(self.state), self.num_columns, self.num_rows, self.n_env)

    def _dev_u8(self, x):
        if x is None:
            return None
        if isinstance(x, torch.Tensor):
            return x.to(device=self.device, dtype=torch.uint8).contiguous()
        return torch.as_tensor(np.ascontiguousarray(x, dtype=np.uint8), device=self.device)

    def _dev_pieces(self, x):
        """Piece ids for the device.  Host arrays are checked here (ids 0..8); ids in device tensors are checked by
        the kernels, where an id that names no piece makes the env inert."""
        if x is not None and not isinstance(x, torch.Tensor):
            a = np.asarray(x)
            if a.size and (a.min() < 0 or a.max() >= len(_lib_piece_names())):
                raise ValueError("piece ids must be in 0..%d" % (len(_lib_piece_names()) - 1))
        return self._dev_u8(x)

    def set_directions(self, feature_directions):
        """feature_directions (state.py:49-50): per-feature multipliers applied to get_after_states / step obs."""
        if feature_directions is None:
            self._dirs = None
        else:
            d = np.ascontiguousarray(feature_directions, dtype=np.float32)
            assert d.shape == (8,)
            self._dirs = d

    # -- reference surface, batched ---------------------------------------------------------
    def reset(self, tape=None):
        """Tetris.__init__ + reset for every env (game.py:21-63).  tape: uint8[n_env] first pieces (global ids)."""
        t = self._dev_pieces(tape)
        with self._on_device():
            _lib.check(_lib.lib().tb_reset(*self._common(), self.env_offset, self.seed, self.piece_set,
                                           _ptr(t), None, self._stream()))

    def reset_masked(self, mask, tape=None):
        """Tetris.reset() (game.py:53-63) on the envs where mask is set: board emptied, one more piece drawn."""
        m, t = self._dev_u8(mask), self._dev_pieces(tape)
        with self._on_device():
            _lib.check(_lib.lib().tb_reset(*self._common(), self.env_offset, self.seed, self.piece_set,
                                           _ptr(t), _ptr(m), self._stream()))

    def get_after_states(self, include_terminal=False, out=None, compact=False):
        """Tetris.get_after_states for every env (game.py:67-80).

        Returns (features float32[n_env, a_max, 8] by enumeration slot, valid int64[n_env] bit mask of
        non-terminal slots, count int32[n_env]).  Legal action k of env e is the k-th set bit of valid[e].
        Feature rows of terminal afterstates are written only with include_terminal=True (game.py:74-78);
        rows past the piece's slot count are never written.
        compact=True: features come as int16[n_env, a_max, 8] holding 2 x feature (every feature is a half-integer,
        so this is exact): half the bytes for a policy that reads them on the host; `features * 0.5` restores them.
        """
        if out is None:
            feats = torch.empty((self.n_env, self.a_max, 8), dtype=torch.int16 if compact else torch.float32,
                                device=self.device)
            valid = torch.empty(self.n_env, dtype=torch.int64, device=self.device)
            count = torch.empty(self.n_env, dtype=torch.int32, device=self.device)
        else:
            feats, valid, count = out
            if feats.dtype != (torch.int16 if compact else torch.float32):
                raise ValueError("out[0] must be %s" % ("int16 with compact=True" if compact else "float32"))
        d = None if self._dirs is None else self._dirs.ctypes.data_as(C.c_void_p)
        flags = (_lib.FLAG_INCLUDE_TERMINAL if include_terminal else 0) | (_lib.FLAG_FEATS_I16 if compact else 0)
        with self._on_device():
            _lib.check(_lib.lib().tb_afterstates(*self._common(), _ptr(feats), _ptr(valid), _ptr(count),
                                                 self.a_max, d, flags, self._stream()))
        return feats, valid, count

    def step(self, actions, tape=None, auto_reset=False, action_is_slot=False, check=True):
        """Tetris.step for every env (game.py:82-92).  Returns (obs f32[n,8], reward i32[n], done bool[n], lines i32[n]).

        check=True: like the reference, an out-of-range action raises IndexError BEFORE any env is stepped (a dry run
        of the kernel validates all actions first; one extra launch and a host synchronisation).  check=False is the
        asynchronous path (CUDA graphs, benchmarks): an env with an out-of-range action is left untouched and reports
        a zero observation / reward / lines and done = "no legal placement at all"."""
        if isinstance(actions, torch.Tensor):
            a = actions.to(device=self.device, dtype=torch.int32).contiguous()
        else:
            a = torch.as_tensor(np.ascontiguousarray(actions, dtype=np.int32), device=self.device)
        if a.shape != (self.n_env,):
            raise ValueError("actions must have shape (n_env,)")
        t = self._dev_pieces(tape)
        obs = torch.empty((self.n_env, 8), dtype=torch.float32, device=self.device)
        reward = torch.empty(self.n_env, dtype=torch.int32, device=self.device)
        done = torch.empty(self.n_env, dtype=torch.uint8, device=self.device)
        lines = torch.empty(self.n_env, dtype=torch.int32, device=self.device)
        flags = (_lib.FLAG_AUTO_RESET if auto_reset else 0) | (_lib.FLAG_ACTION_IS_SLOT if action_is_slot else 0)
        L = _lib.lib()
        with self._on_device():
            if check:
                self._status.zero_()
                _lib.check(L.tb_step(*self._common(), self.env_offset, self.seed, self.piece_set, _ptr(a), _ptr(t),
                                     None, None, None, None, _ptr(self._status), flags | _lib.FLAG_VALIDATE_ONLY,
                                     self._stream()))
                st = int(self._status.item())
                if st != 0:
                    raise IndexError("action out of range for env %d (game.py:83); no env was stepped" % (0x7FFFFFFF - st))
            _lib.check(L.tb_step(*self._common(), self.env_offset, self.seed, self.piece_set, _ptr(a), _ptr(t),
                                 _ptr(obs), _ptr(reward), _ptr(done), _ptr(lines), None, flags, self._stream()))
        if self._dirs is not None:
            obs = obs * torch.as_tensor(self._dirs, device=self.device)
        return obs, reward, done.bool(), lines

    def rollout(self, n_steps, policy="greedy", weights=None):
        """n_steps placements per env with an in-kernel policy and auto-reset; adds into self.stats (device)."""
        pol = {"random": _lib.POLICY_RANDOM, "greedy": _lib.POLICY_GREEDY, 0: 0, 1: 1}[policy]
        w = np.ascontiguousarray(BCTS_WEIGHTS if weights is None else weights, dtype=np.float32)
        assert w.shape == (8,)
        with self._on_device():
            _lib.check(_lib.lib().tb_rollout(*self._common(), self.env_offset, self.seed, self.piece_set, int(n_steps),
                                             pol, w.ctypes.data_as(C.c_void_p), _ptr(self.stats), self._stream()))
        return self.stats

    def rollout_values(self, length=5, n=5, policy="greedy", weights=None, seed=None, piece_tape=None):
        """Tetris.perform_rollouts (game.py:150-160) for every env and every legal action at once, on the device.

        Every (env, action) is forked n times: the action is applied, the next piece drawn from the fork's own
        stream, and the in-kernel policy followed for length - 1 more placements.  Returns (mean return float64
        [n_env, a_max] -- -1 for a fork that ended, else lines minus placements of the follow-up steps, 0 where the
        slot is not a legal action --, valid int64[n_env] bit mask of legal slots).  The envs themselves are not
        stepped.  Follow-up statistics are added to self.stats.
        piece_tape: optional uint8[n_env, a_max, n, length] -- the pieces each fork draws (the one after the action
        first) instead of its RNG stream: what a recorded sampler of the reference supplies.
        """
        pol = {"random": _lib.POLICY_RANDOM, "greedy": _lib.POLICY_GREEDY, 0: 0, 1: 1}[policy]
        w = np.ascontiguousarray(BCTS_WEIGHTS if weights is None else weights, dtype=np.float32)
        seed2 = (self.seed ^ 0xF02C) if seed is None else int(seed) & (2 ** 64 - 1)
        L = _lib.lib()
        n_child = self.n_env * self.a_max * int(n)
        tape = self._dev_pieces(piece_tape)
        if tape is not None and tape.numel() != n_child * int(length):
            raise ValueError("piece_tape must have shape (n_env, a_max, n, length)")
        with self._on_device():
            child = torch.empty(L.tb_state_bytes(self.num_columns, self.num_rows, n_child), dtype=torch.uint8,
                                device=self.device)
            ret = torch.empty((self.n_env, self.a_max), dtype=torch.int32, device=self.device)
            valid = torch.empty(self.n_env, dtype=torch.int64, device=self.device)
            _lib.check(L.tb_rollout_values(*self._common(), self.piece_set, _ptr(child), self.a_max, int(n), int(length),
                                           pol, w.ctypes.data_as(C.c_void_p), seed2,
                                           self.env_offset * self.a_max * int(n), _ptr(tape), _ptr(ret), _ptr(valid),
                                           _ptr(self.stats), self._stream()))
        return ret.double() / float(n), valid

    def action_probabilities(self, feats, valid, weights, temperature=1.0, actions=None):
        """utils.compute_action_probabilities (utils.py:26-31) for every env over its legal afterstates, and -- when
        `actions` (enumeration slots, int32[n_env]) is given -- utils.grad_of_log_action_probabilities
        (utils.py:35-38).  float64 on the device.  Returns probs float64[n_env, a_max] (0 for illegal slots), or
        (probs, grad float64[n_env, 8])."""
        w = np.ascontiguousarray(weights, dtype=np.float64)
        assert w.shape == (8,) and feats.shape == (self.n_env, self.a_max, 8) and feats.dtype == torch.float32
        probs = torch.empty((self.n_env, self.a_max), dtype=torch.float64, device=self.device)
        grad = a = None
        if actions is not None:
            a = actions.to(device=self.device, dtype=torch.int32).contiguous()
            grad = torch.empty((self.n_env, 8), dtype=torch.float64, device=self.device)
        with self._on_device():
            _lib.check(_lib.lib().tb_action_probabilities(self.n_env, self.a_max, _ptr(feats.contiguous()), _ptr(valid),
                                                          w.ctypes.data_as(C.c_void_p), float(temperature), _ptr(a),
                                                          _ptr(probs), _ptr(grad), self._stream()))
        return probs if actions is None else (probs, grad)

    def fitness(self, feats, weights=None):
        """Tetris.fitness (game.py:109-120) for every row of `feats` (float32 [..., 8], contiguous, on this device):
        float32 products summed left to right, no FMA.  Returns float32 [...]."""
        w = np.ascontiguousarray(BCTS_WEIGHTS if weights is None else weights, dtype=np.float32)
        f = feats.contiguous()
        out = torch.empty(f.shape[:-1], dtype=torch.float32, device=self.device)
        with self._on_device():
            _lib.check(_lib.lib().tb_fitness(out.numel(), _ptr(f), w.ctypes.data_as(C.c_void_p), _ptr(out), self._stream()))
        return out

    def capture_lockstep(self, policy, auto_reset=True, include_terminal=False):
        """Capture one lockstep iteration -- get_after_states -> policy -> step -- in a CUDA graph.

        policy(feats, valid, count) -> int32 actions [n_env] must be made of capturable torch ops on this device
        (no host synchronisation).  Returns (replay, outputs): replay() launches the whole iteration as one graph --
        for small batches the iteration is launch-bound, this removes the per-kernel launch cost -- and `outputs` is
        the dict of static tensors it fills (feats, valid, count, actions, obs, reward, done, lines)."""
        out = {}
        with self._on_device():
            feats = torch.empty((self.n_env, self.a_max, 8), dtype=torch.float32, device=self.device)
            valid = torch.empty(self.n_env, dtype=torch.int64, device=self.device)
            count = torch.empty(self.n_env, dtype=torch.int32, device=self.device)

            def iteration():
                self.get_after_states(include_terminal=include_terminal, out=(feats, valid, count))
                actions = policy(feats, valid, count).to(torch.int32)
                obs, reward, done, lines = self.step(actions, auto_reset=auto_reset, check=False)
                return actions, obs, reward, done, lines
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):                   # warm-up off the capture stream (allocator, lazy init)
                saved = self.state.clone()
                iteration()
                self.state.copy_(saved)
            torch.cuda.current_stream(self.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                actions, obs, reward, done, lines = iteration()
        out.update(feats=feats, valid=valid, count=count, actions=actions, obs=obs, reward=reward, done=done, lines=lines)
        return graph.replay, out

    def stats_dict(self, stats=None):
        s = (self.stats if stats is None else stats).cpu().tolist()
        return dict(zip(_lib.STATS, s))

    # -- state interchange --------------------------------------------------------------------
    def export_boards(self, first=0, count=None):
        """(rows uint16->int16 tensor [count, R+4] of row masks, heights uint8 [count, C], piece uint8 [count])."""
        count = self.n_env - first if count is None else count
        rows = torch.empty((count, self.n_stored_rows), dtype=torch.int16, device=self.device)
        heights = torch.empty((count, self.num_columns), dtype=torch.uint8, device=self.device)
        piece = torch.empty(count, dtype=torch.uint8, device=self.device)
        with self._on_device():
            _lib.check(_lib.lib().tb_export_boards(*self._common(), first, count, _ptr(rows), _ptr(heights), _ptr(piece),
                                                   self._stream()))
        return rows, heights, piece

    def import_boards(self, rows, piece=None, first=0):
        """Upload boards as row masks (uint16 per row, bit c = column c); heights are recomputed on the device.
        rows: numpy / torch array [count, num_rows + 4] (a CUDA int16 tensor is used in place, no host round trip)."""
        if isinstance(rows, torch.Tensor) and rows.is_cuda and rows.dtype == torch.int16:
            r = rows.contiguous()
        else:
            r = np.ascontiguousarray(rows.cpu().numpy() if isinstance(rows, torch.Tensor) else rows).astype(np.uint16)
            r = torch.as_tensor(r.view(np.int16), device=self.device)
        if r.dim() != 2 or r.shape[1] != self.n_stored_rows:
            raise ValueError("rows must have shape (count, num_rows + 4)")
        p = self._dev_u8(piece) if isinstance(piece, torch.Tensor) else self._dev_pieces(piece)
        with self._on_device():
            _lib.check(_lib.lib().tb_import_boards(*self._common(), first, r.shape[0], _ptr(r), _ptr(p), self._stream()))

    # numpy conveniences (host copies)
    def rows(self):
        return self.export_boards()[0].cpu().numpy().view(np.uint16)

    @property
    def heights(self):
        return self.export_boards()[1].cpu().numpy()

    @property
    def piece(self):
        return self.export_boards()[2].cpu().numpy()

    def representation(self, env=0):
        """(num_rows+4, num_columns) int64 0/1 array of one env, the reference's State.representation."""
        rows = self.export_boards(env, 1)[0].cpu().numpy().view(np.uint16)[0]
        return ((rows[:, None] >> np.arange(self.num_columns, dtype=np.uint16)) & 1).astype(np.int64)

    def board_string(self, env=0):
        """utils.print_board_to_string (utils.py:179-191) of one env of the batch: the board with its 4 buffer rows."""
        from .utils import print_board_to_string
        from types import SimpleNamespace
        return print_board_to_string(SimpleNamespace(representation=self.representation(env),
                                                     num_rows=self.n_stored_rows, num_columns=self.num_columns))

    def render(self, env=0):
        """Tetris.render (game.py:122-124) for one env of the batch: prints the board and the current piece."""
        from . import PIECE_NAMES
        print(self.board_string(env))
        p = int(self.export_boards(env, 1)[2].item())
        print(PIECE_NAMES[p] if p < len(PIECE_NAMES) else "(finished)")


class HostRollout:
    """Games that live in HOST memory, played on the device (the cycle of a caller that keeps its boards on the host).

    Every `play` copies the boards and pieces of all envs in from pinned host buffers (`h_rows` int16[n, R+4] row
    masks, `h_piece` uint8[n]), imports them (tb_import_boards), plays `n_steps` placements per env with the fused
    rollout kernel (tb_rollout), exports (tb_export_boards) and copies boards, heights and pieces back into the same
    host buffers, plus the episode statistics.  The env range is cut into `chunks` contiguous shards -- each a
    BatchedTetris with its global env_offset, on its own CUDA stream -- so the H2D copy of shard i+1 and the D2H copy
    of shard i-1 overlap the rollout kernel of shard i.  Results do not depend on `chunks`: every env's piece stream
    is keyed by its global env id (same argument as multi-GPU sharding).
    """

    def __init__(self, num_columns, num_rows, n_env, chunks=8, piece_set=1, seed=0, env_offset=0, device=None,
                 prioritized=True):
        from .distributed import shard_range
        self.n_env, self.chunks = int(n_env), max(1, min(int(chunks), int(n_env)))
        self.bounds = []
        self.envs = []
        for i in range(self.chunks):
            off, cnt = shard_range(self.n_env, i, self.chunks)
            self.bounds.append((off, off + cnt))
            self.envs.append(BatchedTetris(num_columns, num_rows, cnt, piece_set=piece_set, seed=seed,
                                           env_offset=int(env_offset) + off, device=device))
        e0 = self.envs[0]
        self.device = e0.device
        # earlier chunks at higher stream priority: their rollout CTAs are scheduled first, so chunk i finishes (and its
        # D2H copy starts) while chunk i+1 still computes, instead of all chunks' kernels sharing the SMs to the end
        self.streams = [torch.cuda.Stream(device=self.device, priority=-(self.chunks - 1 - i) if prioritized else 0)
                        for i in range(self.chunks)]
        self.h_rows = torch.zeros((self.n_env, e0.n_stored_rows), dtype=torch.int16).pin_memory()
        self.h_heights = torch.zeros((self.n_env, e0.num_columns), dtype=torch.uint8).pin_memory()
        self.h_piece = torch.zeros(self.n_env, dtype=torch.uint8).pin_memory()
        self.h_stats = torch.zeros((self.chunks, len(_lib.STATS)), dtype=torch.int64).pin_memory()
        self._d_rows = [torch.empty((b - a, e0.n_stored_rows), dtype=torch.int16, device=self.device) for a, b in self.bounds]
        self._d_piece = [torch.empty(b - a, dtype=torch.uint8, device=self.device) for a, b in self.bounds]
        self.h2d_bytes = self.h_rows.numel() * 2 + self.h_piece.numel() + 32          # + the 8 float32 weights
        self.d2h_bytes = (self.h_rows.numel() * 2 + self.h_heights.numel() + self.h_piece.numel()
                          + 8 * self.h_stats.numel())
        self.pull()                                         # host buffers <- the freshly reset games

    def pull(self):
        """Host buffers <- device state (boards, heights, pieces), synchronously."""
        for (a, b), env in zip(self.bounds, self.envs):
            r, hh, pp = env.export_boards()
            self.h_rows[a:b].copy_(r); self.h_heights[a:b].copy_(hh); self.h_piece[a:b].copy_(pp)

    def play(self, n_steps, policy="greedy", weights=None):
        """Host boards in -> n_steps placements per env -> host boards out.  Returns the combined cumulative episode
        statistics (int64[len(STATS)], host).  Synchronous: the host buffers are valid on return."""
        from .distributed import combine_stats
        cur = torch.cuda.current_stream(self.device)
        for i, (env, s) in enumerate(zip(self.envs, self.streams)):
            a, b = self.bounds[i]
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                self._d_rows[i].copy_(self.h_rows[a:b], non_blocking=True)
                self._d_piece[i].copy_(self.h_piece[a:b], non_blocking=True)
                env.import_boards(self._d_rows[i], piece=self._d_piece[i])
                env.rollout(n_steps, policy, weights)
                r, hh, pp = env.export_boards()
                self.h_rows[a:b].copy_(r, non_blocking=True)
                self.h_heights[a:b].copy_(hh, non_blocking=True)
                self.h_piece[a:b].copy_(pp, non_blocking=True)
                self.h_stats[i].copy_(env.stats, non_blocking=True)
        for s in self.streams:
            s.synchronize()
        return combine_stats(list(self.h_stats))
