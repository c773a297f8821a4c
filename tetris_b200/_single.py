"""Device plumbing for the single-env compatibility objects (state.State, tetromino.*, game.Tetris).

One scratch context per board shape: a one-env device state plus output buffers.  Every call here is a
kernel launch through the C ABI (tb_import_boards / tb_afterstates_export / tb_eval_states / tb_fitness);
the host only packs cells into row masks and unpacks the results.  No GPU -> RuntimeError, never a CPU path.
"""
import ctypes as C

import numpy as np

from . import _lib

A_STRIDE = 64        # >= TB_MAX_SLOTS (60: ThreeL on 16 columns)
_ctx = {}


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("tetris_b200: no CUDA device; the environment runs on the GPU only (no CPU fallback)")
    return torch


def pack_rows(representation):
    """(N, C) 0/1 cells -> uint16 row masks, bit c = column c."""
    rep = np.asarray(representation)
    if rep.ndim != 2:
        raise ValueError("representation must be a (num_rows + 4, num_columns) array")
    w = np.left_shift(1, np.arange(rep.shape[1]), dtype=np.int64)
    return ((rep != 0).astype(np.int64) * w).sum(axis=1).astype(np.uint16)


def unpack_rows(rows, num_columns):
    """uint16 row masks [..., N] -> int64 0/1 cells [..., N, C] (the reference's `representation` dtype)."""
    rows = np.asarray(rows, np.uint16)
    return ((rows[..., None] >> np.arange(num_columns, dtype=np.uint16)) & 1).astype(np.int64)


class _Ctx:
    def __init__(self, Cc, R):
        L = _lib.lib()
        torch = _torch()
        _lib.ensure_shape(Cc, R)                   # built in, or compiled on demand (any size: game.py:21-31)
        self.C, self.R, self.N = Cc, R, R + 4
        dev = torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        self.state = torch.zeros(L.tb_state_bytes(Cc, R, 1), dtype=torch.uint8, device=dev)
        self.feats = torch.empty((A_STRIDE, 8), dtype=torch.float32, device=dev)
        self.rows = torch.empty((A_STRIDE, self.N), dtype=torch.int16, device=dev)
        self.heights = torch.empty((A_STRIDE, Cc), dtype=torch.uint8, device=dev)
        self.info = torch.empty((A_STRIDE, 4), dtype=torch.int32, device=dev)
        self.score = torch.empty(A_STRIDE, dtype=torch.float32, device=dev)
        _lib.check(L.tb_reset(self._p(self.state), Cc, R, 1, 0, 0, 1, None, None, self._stream()))

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr())

    def _stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def enumerate(self, rows_u16, piece):
        """Every placement of `piece` on the board: (n, feats[n,8] f32, rows[n,N] u16, heights[n,C], info[n,4])."""
        import torch
        L = _lib.lib()
        n = L.tb_num_slots(piece, self.C)
        rin = torch.as_tensor(np.ascontiguousarray(rows_u16, np.uint16).view(np.int16).reshape(1, self.N), device=self.dev)
        pin = torch.as_tensor(np.array([piece], np.uint8), device=self.dev)
        st = self._stream()
        _lib.check(L.tb_import_boards(self._p(self.state), self.C, self.R, 1, 0, 1, self._p(rin), self._p(pin), st))
        _lib.check(L.tb_afterstates_export(self._p(self.state), self.C, self.R, 1, self._p(self.feats), self._p(self.rows),
                                           self._p(self.heights), self._p(self.info), A_STRIDE, st))
        return (n, self.feats[:n].cpu().numpy(), self.rows[:n].cpu().numpy().view(np.uint16),
                self.heights[:n].cpu().numpy(), self.info[:n].cpu().numpy())

    def eval_state(self, rows_u16, anchor_row, n_changed, ppcr_packed, bonus2):
        """State.__init__ on the device: clear, heights, terminal, features (tb_eval_states)."""
        import torch
        L = _lib.lib()
        rin = torch.as_tensor(np.ascontiguousarray(rows_u16, np.uint16).view(np.int16).reshape(1, self.N), device=self.dev)
        par = torch.as_tensor(np.array([[anchor_row, n_changed, ppcr_packed, bonus2]], np.int32), device=self.dev)
        _lib.check(L.tb_eval_states(self.C, self.R, 1, self._p(rin), self._p(par), self._p(self.rows), self._p(self.heights),
                                    self._p(self.info), self._p(self.feats), self._stream()))
        info = self.info[0].cpu().numpy()
        return (self.rows[0].cpu().numpy().view(np.uint16), self.heights[0].cpu().numpy(), int(info[0]), int(info[1]),
                bool(info[2]), self.feats[0].cpu().numpy())

    def fitness(self, feats, weights):
        """Tetris.fitness (game.py:109-120) of up to 36 feature rows, float32 on the device."""
        import torch
        f = np.ascontiguousarray(feats, np.float32).reshape(-1, 8)
        n = f.shape[0]
        assert 1 <= n <= A_STRIDE
        w = np.ascontiguousarray(weights, np.float32)
        self.feats[:n].copy_(torch.as_tensor(f))
        _lib.check(_lib.lib().tb_fitness(n, self._p(self.feats), w.ctypes.data_as(C.c_void_p), self._p(self.score),
                                         self._stream()))
        return self.score[:n].cpu().numpy()


def ctx(num_columns, num_rows):
    """The scratch context of a board shape (num_rows = legal rows, without the 4 buffer rows)."""
    key = (int(num_columns), int(num_rows))
    c = _ctx.get(key)
    if c is None:
        c = _ctx[key] = _Ctx(*key)
    return c


def slot_info(piece, num_columns, slot):
    out = (C.c_int32 * 17)()
    _lib.check(_lib.lib().tb_slot_info(piece, num_columns, slot, out))
    v = list(out)
    return dict(anchor_col=v[0], width=v[1], n_cells=v[2], n_changed=v[3], bonus2=v[4],
                ppcr=v[5:5 + v[3]], cells=[(v[9 + 2 * i], v[10 + 2 * i]) for i in range(v[2])])
