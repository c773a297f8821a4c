"""Pieces and samplers of the reference (tetromino.py:5-576) over the afterstate kernel.

Every piece class keeps the reference's constructor and `get_after_states(State) -> [State]`; the per-orientation
column loops, hard drop, painting and State construction of tetromino.py:41-576 are ONE launch of the afterstate
kernel (tb_afterstates_export) whose slots are already in the reference's enumeration (= action) order.  The
piece x rotation table itself lives in csrc/tb_core.cuh (TB_ORI_TABLE) and is readable from Python through
`piece_table()` below (tb_slot_info).

Global piece ids (include/tetris_b200.h): 0 Straight, 1 RCorner, 2 LCorner, 3 Square, 4 SnakeR, 5 SnakeL, 6 T
(the seven-piece order of game.py:41-47), 7 ThreeL, 8 ThreeLine (the default set, game.py:38-39).
"""
import numpy as np

from . import _single
from . import state

PIECE_NAMES = ("Straight", "RCorner", "LCorner", "Square", "SnakeR", "SnakeL", "T", "ThreeL", "ThreeLine")


def piece_table(piece_id, num_columns):
    """Rows of the piece x rotation table in enumeration order: one dict per afterstate slot with anchor_col,
    width, cells [(dx, dy)], n_changed, ppcr (pieces_per_changed_row) and bonus2 (2 * landing_height_bonus)."""
    from . import _lib
    n = _lib.lib().tb_num_slots(piece_id, num_columns)
    return [_single.slot_info(piece_id, num_columns, s) for s in range(n)]


class Tetromino:
    piece_id = None

    def __init__(self, feature_type, num_features, num_columns):
        self.feature_type = feature_type
        self.num_features = num_features
        self.num_columns = num_columns
        self._table = None

    def _slots(self):
        if self._table is None:
            self._table = piece_table(self.piece_id, self.num_columns)
        return self._table

    def get_after_states(self, current_state):
        """All rotation x column placements of this piece on `current_state`, as State objects, in the reference's
        enumeration order (terminal ones included, exactly like tetromino.py:*.get_after_states)."""
        if self.piece_id is None:
            raise NotImplementedError("Tetromino is abstract; use one of the piece classes")
        rep = np.asarray(current_state.representation)
        n_rows, n_cols = rep.shape
        if n_cols != self.num_columns:
            raise ValueError("state has %d columns, piece was built for %d" % (n_cols, self.num_columns))
        c = _single.ctx(n_cols, n_rows - 4)
        n, feats, rows, heights, info = c.enumerate(_single.pack_rows(rep), self.piece_id)
        reps = _single.unpack_rows(rows, n_cols)
        out = []
        for s, slot in enumerate(self._slots()):
            a, full_mask, terminal = int(info[s, 0]), int(info[s, 1]), bool(info[s, 2])
            is_full = np.array([(full_mask >> (a + k)) & 1 for k in range(slot["n_changed"])], dtype=bool)
            out.append(state.State._from_kernel(
                reps[s], heights[s], anchor_col=int(info[s, 3]), anchor_row=a,
                ppcr=np.array(slot["ppcr"]), bonus=slot["bonus2"] / 2 if slot["bonus2"] % 2 else slot["bonus2"] // 2,
                is_full=is_full, terminal=terminal, feats=feats[s], num_features=self.num_features,
                feature_type=self.feature_type))
        return out

    def __repr__(self):
        """ASCII picture of the first orientation (drawn from the table, top row first)."""
        cells = self._slots()[0]["cells"] if self.piece_id is not None else []
        if not cells:
            return object.__repr__(self)
        w = 1 + max(dx for dx, _ in cells)
        h = 1 + max(dy for _, dy in cells)
        lines = [" ".join("██" if (dx, dy) in cells else "  " for dx in range(w)).rstrip() for dy in range(h - 1, -1, -1)]
        return "\n" + "\n".join(lines)


class TetrominoSampler:
    """Shuffled-bag sampler on NumPy's global RNG, draw for draw like tetromino.py:12-22 (a permutation at
    construction and whenever the bag is empty), so `np.random.seed(s)` reproduces the reference's piece stream."""

    def __init__(self, tetrominos):
        self.tetrominos = tetrominos
        self.current_batch = np.random.permutation(len(self.tetrominos))

    def next_tetromino(self):
        if self.current_batch.size == 0:
            self.current_batch = np.random.permutation(len(self.tetrominos))
        head, self.current_batch = self.current_batch[0], self.current_batch[1:]
        return self.tetrominos[head]


class TetrominoSamplerRandom:
    """I.i.d. sampler.  (The reference's version, tetromino.py:25-30, returns a length-1 array instead of a
    piece and is unusable; this one returns the piece.)"""

    def __init__(self, tetrominos):
        self.tetrominos = tetrominos

    def next_tetromino(self):
        return self.tetrominos[np.random.randint(len(self.tetrominos))]


def _piece_class(name, piece_id, doc, extra=None):
    body = {"piece_id": piece_id, "__doc__": doc}
    body.update(extra or {})
    return type(name, (Tetromino,), body)


def _init_with_tet_ind(tet_ind):
    def __init__(self, feature_type, num_features, num_columns):
        Tetromino.__init__(self, feature_type, num_features, num_columns)
        self.tet_ind = tet_ind
    return {"__init__": __init__}


Straight = _piece_class("Straight", 0, "I piece: vertical then horizontal placements (tetromino.py:33-75)")
RCorner = _piece_class("RCorner", 1, "tetromino.py:417-495")
LCorner = _piece_class("LCorner", 2, "tetromino.py:498-576")
Square = _piece_class("Square", 3, "tetromino.py:78-104")
SnakeR = _piece_class("SnakeR", 4, "tetromino.py:107-154")
SnakeL = _piece_class("SnakeL", 5, "tetromino.py:285-331")
T = _piece_class("T", 6, "tetromino.py:334-414")
ThreeL = _piece_class("ThreeL", 7, "three-cell corner, tet_ind = 1 (tetromino.py:202-282)", _init_with_tet_ind(1))
ThreeLine = _piece_class("ThreeLine", 8, "three-cell line, tet_ind = 0 (tetromino.py:157-199)", _init_with_tet_ind(0))

PIECE_CLASSES = (Straight, RCorner, LCorner, Square, SnakeR, SnakeL, T, ThreeL, ThreeLine)


def standard_set(num_columns, feature_type='bcts', num_features=8):
    """The seven tetrominoes in the order of the commented-out list of game.py:41-47."""
    return [cls(feature_type, num_features, num_columns) for cls in PIECE_CLASSES[:7]]


def default_set(num_columns, feature_type='bcts', num_features=8):
    """The reference's active set [ThreeL, ThreeLine] (game.py:38-39)."""
    return [ThreeL(feature_type, num_features, num_columns), ThreeLine(feature_type, num_features, num_columns)]
