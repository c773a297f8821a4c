"""`State`: the reference's board-state object (state.py:4-107) over the CUDA kernels.

Same constructor, attributes and methods as the reference's State; the work its constructor and
`get_features` do in Python loops -- clearing the full rows among `changed_lines` (state.py:121-143), the
terminal test (state.py:111-117), column heights (state.py:162-172) and the eight BCTS features
(state.py:97-107,175-280) -- is one `tb_eval_states` launch here, or comes straight out of the afterstate
kernel when a Tetromino builds the State (`State._from_kernel`).
"""
import numpy as np

from . import _single

_DEFAULT_CHANGED = np.arange(1)
_DEFAULT_PPCR = np.array([0])


class State:
    def __init__(self, representation, lowest_free_rows=None, anchor_col=0, changed_lines=_DEFAULT_CHANGED,
                 pieces_per_changed_row=_DEFAULT_PPCR, landing_height_bonus=0.0, num_features=8,
                 feature_type='bcts'):
        rep = np.asarray(representation)
        changed = np.asarray(changed_lines).astype(np.int64).ravel()
        ppcr = np.asarray(pieces_per_changed_row).astype(np.int64).ravel()
        if not (1 <= len(changed) <= 4) or np.any(np.diff(changed) != 1):
            raise ValueError("changed_lines must be 1..4 consecutive rows (every reference piece passes "
                             "np.arange(anchor_row, anchor_row + k))")
        bonus2 = int(round(2 * float(landing_height_bonus)))
        if bonus2 != 2 * float(landing_height_bonus) or not 0 <= bonus2 <= 3:
            raise ValueError("landing_height_bonus must be one of 0, 0.5, 1, 1.5")
        packed = 0
        for k in range(min(len(ppcr), len(changed))):
            packed |= (int(ppcr[k]) & 15) << (4 * k)
        n_rows, n_cols = rep.shape
        if lowest_free_rows is not None:
            # The reference trusts the caller's heights (state.py:22-25) for hard drop, line clear and features; the
            # kernels derive them from the board.  Every reference call site passes heights that match the board, so a
            # mismatch is rejected loudly rather than silently ignored.
            given = np.asarray(lowest_free_rows).astype(np.int64).ravel()
            filled = rep != 0
            want = np.where(filled.any(axis=0), n_rows - np.argmax(filled[::-1], axis=0), 0)
            if given.shape != want.shape or np.any(given != want):
                raise ValueError("lowest_free_rows does not match the representation (expected %s); pass None to have "
                                 "it computed (state.py:162-172)" % want.tolist())
        c = _single.ctx(n_cols, n_rows - 4)
        rows, heights, n_cleared, full_mask, terminal, feats = c.eval_state(
            _single.pack_rows(rep), int(changed[0]), len(changed), packed, bonus2)
        self._fill(rep if n_cleared == 0 else _single.unpack_rows(rows, n_cols), heights, anchor_col, int(changed[0]),
                   pieces_per_changed_row, landing_height_bonus,
                   np.array([(full_mask >> int(r)) & 1 for r in changed], dtype=bool), n_cleared, terminal, feats,
                   num_features, feature_type)

    # -- construction from kernel outputs (tetromino.get_after_states) ----------------------------
    @classmethod
    def _from_kernel(cls, representation, heights, anchor_col, anchor_row, ppcr, bonus, is_full, terminal, feats,
                     num_features, feature_type):
        self = cls.__new__(cls)
        self._fill(representation, heights, anchor_col, anchor_row, ppcr, bonus, is_full, int(np.sum(is_full)),
                   terminal, feats, num_features, feature_type)
        return self

    def _fill(self, representation, heights, anchor_col, anchor_row, ppcr, bonus, is_full, n_cleared, terminal,
              feats, num_features, feature_type):
        self.representation = representation
        self.anchor_col = anchor_col
        self.pieces_per_changed_row = ppcr
        self.landing_height_bonus = bonus
        self.num_features = num_features
        self.feature_type = feature_type
        self.lowest_free_rows = np.asarray(heights).astype(np.int64)
        self.num_rows = representation.shape[0]          # includes the 4 buffer rows (state.py:27)
        self.num_columns = representation.shape[1]
        self.n_legal_rows = self.num_rows - 4
        self.n_cleared_lines = int(n_cleared)
        self.anchor_row = anchor_row
        self.cleared_rows_relative_to_anchor = is_full
        self.features = None                              # filled by get_features(), like the reference's lazy cache
        self._kernel_features = np.asarray(feats, np.float32)
        self.terminal_state = bool(terminal)
        self.reward = 0 if self.terminal_state else self.n_cleared_lines
        self.value_estimate = 0.0

    # -- reference surface ------------------------------------------------------------------------
    def __repr__(self):
        return self.print_board_to_string()

    def get_features(self, direct_by=None, order_by=None, standardize_by=None, addRBF=False):
        if self.features is None:
            self.calc_feature_values()
        if direct_by is None:
            return self.features
        return self.features * direct_by                  # float64 with -0.0 where a zero meets -1 (state.py:49-50)

    def calc_feature_values(self):
        if self.feature_type != 'bcts':
            raise ValueError("Only 'bcts' features implemented.")
        self.calc_bcts_features()

    def calc_bcts_features(self):
        f = np.zeros(self.num_features, dtype=np.float32)
        f[:8] = self._kernel_features
        self.features = f

    def clear_lines(self, changed_lines):
        """Re-run the constructor's line clear on this state (state.py:83-89)."""
        other = State(self.representation, changed_lines=changed_lines,
                      pieces_per_changed_row=np.zeros(len(changed_lines), dtype=np.int64))
        self.n_cleared_lines = other.n_cleared_lines
        self.representation, self.lowest_free_rows = other.representation, other.lowest_free_rows
        return other.cleared_rows_relative_to_anchor

    def _legal_rows_top_down(self):
        return self.representation[self.n_legal_rows - 1::-1] if self.n_legal_rows > 0 else self.representation[:0]

    def print_board(self):
        for row in self._legal_rows_top_down():
            print("| " + "".join("██ " if v else "   " for v in row) + "|")

    def print_board_to_string(self):
        return "\n" + "".join("|" + "".join("██" if v else "  " for v in row) + "|\n"
                              for row in self._legal_rows_top_down())


# ---------------------------------------------------------------------------------------------
# module-level helpers the reference exposes (state.py:110-172, 175-280); same names and results
# ---------------------------------------------------------------------------------------------
def check_terminal(representation, n_legal_rows):
    """Any cell in the first buffer row (state.py:111-117)."""
    return bool(np.any(np.asarray(representation)[n_legal_rows]))


def calc_lowest_free_rows(rep):
    """1 + index of the highest filled cell per column, 0 for an empty column (state.py:162-172); on the device."""
    rep = np.asarray(rep)
    c = _single.ctx(rep.shape[1], rep.shape[0] - 4)
    return c.eval_state(_single.pack_rows(rep), 0, 0, 0, 0)[1].astype(np.int64)


def clear_lines_jitted(changed_lines, representation, lowest_free_rows, num_columns):
    """state.py:121-143 -> (is_full, n_cleared_lines, representation, lowest_free_rows); on the device."""
    s = State(representation, changed_lines=changed_lines,
              pieces_per_changed_row=np.zeros(len(changed_lines), dtype=np.int64))
    return s.cleared_rows_relative_to_anchor, s.n_cleared_lines, s.representation, s.lowest_free_rows


def get_feature_values_jitted(lowest_free_rows, representation, num_rows, num_columns):
    """The six board features of state.py:175-280, in the reference's order
    [rows_with_holes, column_transitions, holes, cumulative_wells, row_transitions, hole_depth]; on the device."""
    rep = np.asarray(representation)
    f = _single.ctx(num_columns, num_rows).eval_state(_single.pack_rows(rep), 0, 0, 0, 0)[5]
    return [f[0], f[1], f[2], f[4], f[5], f[7]]


def minmaxavg_jitted(x):
    """Unused by the reference itself (state.py:147-158); kept importable.  Note its quirks: the mean leaves out
    x[0] from the sum, and a value that raises the maximum is never tested against the minimum."""
    x = list(x)
    lo = hi = x[0]
    total = 0
    for v in x[1:]:
        total += v
        if v > hi:
            hi = v
        elif v < lo:
            lo = v
    return lo, hi, total / len(x)
