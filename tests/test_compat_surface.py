"""CPU-only: the reference-facing Python surface (package `tetris` -> tetris_b200.game/state/tetromino/utils) has
the reference's names and signatures, its host-side helpers give the reference's numbers, the piece table read
through the C ABI matches SURVEY.md Appendix A, and the GPU-only objects fail loudly without a GPU."""
import inspect
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_package_alias_and_names():
    import tetris
    from tetris import state, tetromino, utils
    from tetris.game import Tetris
    import tetris_b200
    assert Tetris is tetris_b200.game.Tetris and state is tetris_b200.state
    for name in ("Tetromino", "TetrominoSampler", "TetrominoSamplerRandom", "Straight", "Square", "SnakeR", "SnakeL",
                 "T", "RCorner", "LCorner", "ThreeLine", "ThreeL"):
        assert inspect.isclass(getattr(tetromino, name)), name
    for name in ("State", "check_terminal", "clear_lines_jitted", "minmaxavg_jitted", "calc_lowest_free_rows",
                 "get_feature_values_jitted"):
        assert hasattr(state, name), name
    for name in ("Bunch", "one_hot_vector", "vert_one_hot", "compute_action_probabilities",
                 "grad_of_log_action_probabilities", "softmax", "plot_learning_curve", "plot_individual_agent",
                 "plot_analysis", "print_board_to_string"):
        assert hasattr(utils, name), name
    for name in ("reset", "get_after_states", "step", "is_game_over", "get_best_policy", "fitness", "render",
                 "get_state", "single_rollout", "perform_rollouts"):
        assert callable(getattr(Tetris, name)), name


def test_signatures_match_reference():
    from tetris import state, tetromino, utils
    from tetris.game import Tetris

    def params(f):
        return [(p.name, p.default if p.default is not inspect.Parameter.empty else "<req>")
                for p in inspect.signature(f).parameters.values()]
    assert params(Tetris.__init__) == [("self", "<req>"), ("num_columns", "<req>"), ("num_rows", "<req>"),
                                       ("feature_directions", None), ("feature_type", "bcts"), ("num_features", 8),
                                       ("tetromino_size", 4)]                       # game.py:21-23
    assert params(Tetris.get_after_states) == [("self", "<req>"), ("include_terminal", False)]
    assert params(Tetris.perform_rollouts)[1:] == [("actions", "<req>"), ("policy_function", "<req>"), ("length", 5), ("n", 5)]
    names = [p for p, _ in params(state.State.__init__)]
    assert names == ["self", "representation", "lowest_free_rows", "anchor_col", "changed_lines",
                     "pieces_per_changed_row", "landing_height_bonus", "num_features", "feature_type"]   # state.py:5-12
    assert [p for p, _ in params(state.State.get_features)] == ["self", "direct_by", "order_by", "standardize_by", "addRBF"]
    assert [p for p, _ in params(tetromino.Straight.__init__)] == ["self", "feature_type", "num_features", "num_columns"]
    assert [p for p, _ in params(utils.compute_action_probabilities)] == ["action_features", "weights", "temperature"]
    assert tetromino.ThreeLine("bcts", 8, 10).tet_ind == 0 and tetromino.ThreeL("bcts", 8, 10).tet_ind == 1


def test_reference_style_sys_path_use():
    """The reference's checkout is used with its directory on sys.path (`from game import Tetris`)."""
    d = os.path.join(ROOT, "tetris")
    sys.path.insert(0, d)
    try:
        import game
        import tetris_b200
        assert game.Tetris is tetris_b200.game.Tetris
    finally:
        sys.path.remove(d)


def test_utils_math():
    from tetris import utils
    rng = np.random.default_rng(0)
    f, w = rng.normal(size=(9, 8)), rng.normal(size=8)
    u = f.dot(w) / 0.7
    e = np.exp(u - u.max())
    p = utils.compute_action_probabilities(f, w, 0.7)
    assert np.array_equal(p, e / e.sum()) and abs(p.sum() - 1) < 1e-12
    assert np.array_equal(utils.softmax(u), e / e.sum())
    assert np.array_equal(utils.grad_of_log_action_probabilities(f, p, 3), f[3] - f.T.dot(p))
    assert utils.one_hot_vector(2, 5).tolist() == [0, 0, 1, 0, 0]
    assert utils.vert_one_hot(1, 3).shape == (3, 1) and utils.vert_one_hot(1, 3)[1, 0] == 1
    assert utils.Bunch({"a": 3}).a == 3

    class S:
        representation = np.array([[1, 0], [0, 0], [0, 1]])
        num_rows, num_columns = 3, 2
    assert utils.print_board_to_string(S) == "\n|  ██|\n|    |\n|██  |\n"       # top row first (utils.py:179-191)


def test_learner_math_vs_reference_fixture():
    """utils.compute_action_probabilities / grad_of_log_action_probabilities / softmax (utils.py:26-45) of the host mirror
    against vectors recorded from the reference's own functions (tests/golden/learner.npz, make_golden.py gen_learner).
    Same NumPy expressions in the same order: bit-identical."""
    from golden_util import load
    from tetris import utils
    g = load("learner")
    a_max = g["feats"].shape[1]
    bits = ((g["valid"][:, None] >> np.arange(a_max, dtype=np.uint64)) & np.uint64(1)).astype(bool)
    for e in range(len(bits)):
        fe = g["feats"][e][bits[e]].astype(np.float64)
        k = int(np.nonzero(np.nonzero(bits[e])[0] == g["actions"][e])[0][0])
        for ti, t in enumerate(g["temps"]):
            p = utils.compute_action_probabilities(fe, g["weights"], t)
            assert np.array_equal(p, g["probs"][ti, e][bits[e]]), (e, t)
            assert np.array_equal(utils.grad_of_log_action_probabilities(fe, p, k), g["grads"][ti, e]), (e, t)
        assert np.array_equal(utils.softmax(fe.dot(g["weights"])), g["softmax"][e][bits[e]])


def test_state_host_helpers():
    from tetris import state
    rep = np.zeros((14, 10), int)
    assert not state.check_terminal(rep, 10)
    rep[10, 3] = 1
    assert state.check_terminal(rep, 10)
    assert state.minmaxavg_jitted([3, 1, 4, 1, 5]) == (1, 5, (1 + 4 + 1 + 5) / 5)   # the sum leaves out x[0] (state.py:147-158)


def test_piece_table_matches_survey_appendix_a():
    from tetris import tetromino
    counts10 = {"Straight": 17, "RCorner": 34, "LCorner": 34, "Square": 9, "SnakeR": 17, "SnakeL": 17, "T": 34,
                "ThreeL": 36, "ThreeLine": 18}
    for pid, name in enumerate(tetromino.PIECE_NAMES):
        tab = tetromino.piece_table(pid, 10)
        assert len(tab) == counts10[name]
        for row in tab:
            assert row["n_cells"] == (3 if pid >= 7 else 4)
            assert sum(row["ppcr"]) <= row["n_cells"] and 1 <= row["n_changed"] <= 4
            assert row["anchor_col"] + row["width"] <= 10
    t = tetromino.piece_table(6, 10)                      # T: loop 1 interleaves O0/O1 per column (tetromino.py:347-378)
    assert [r["anchor_col"] for r in t[:4]] == [0, 0, 1, 1]
    assert t[0]["cells"] == [(0, 0), (1, 0), (1, 1), (2, 0)] and t[0]["ppcr"] == [3] and t[0]["bonus2"] == 1
    assert sorted(t[1]["cells"]) == [(0, 1), (1, 0), (1, 1), (2, 1)] and t[1]["ppcr"] == [1, 3]
    i = tetromino.piece_table(0, 10)                      # Straight: 10 vertical then 7 horizontal
    assert i[0]["width"] == 1 and i[0]["bonus2"] == 3 and i[10]["width"] == 4 and i[10]["ppcr"] == [4]
    assert repr(tetromino.Straight("bcts", 8, 10)) == "\n██\n██\n██\n██"
    assert repr(tetromino.T("bcts", 8, 10)) == "\n   ██\n██ ██ ██"


def test_sampler_follows_numpy_global_rng():
    """TetrominoSampler consumes np.random.permutation exactly like tetromino.py:12-22 (SURVEY 8c seed-0 order)."""
    from tetris import tetromino
    np.random.seed(0)
    s = tetromino.TetrominoSampler(list(range(7)))
    assert [s.next_tetromino() for _ in range(21)] == [6, 2, 1, 3, 0, 5, 4, 1, 0, 6, 3, 4, 2, 5, 6, 3, 5, 1, 2, 4, 0]


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from tetris.game import Tetris
    from tetris import state
    with pytest.raises(RuntimeError, match="no CUDA device"):
        Tetris(10, 10)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        state.State(np.zeros((14, 10), int))
    from tetris_b200 import BatchedTetris
    with pytest.raises(RuntimeError, match="no CUDA device"):
        BatchedTetris(10, 20, 8)
    from tetris_b200 import HostRollout
    with pytest.raises(RuntimeError, match="no CUDA device"):
        HostRollout(10, 20, 64, chunks=2)
