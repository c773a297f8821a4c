"""Side helpers with the names and results of the reference's utils.py.

Learner math (utils.py:6-45) is plain host NumPy on a handful of numbers -- it never touches env state.  The
plotting functions (utils.py:48-170) keep their signatures and output file names but import matplotlib only
when called (it is an optional dependency here).  `print_board_to_string` is the ASCII render `Tetris.render`
uses (utils.py:179-191), buffer rows included.
"""
import os

import numpy as np

FEATURE_NAMES = ('rows_with_holes', 'column_transitions', 'holes', 'landing_height', 'cumulative_wells',
                 'row_transitions', 'eroded', 'hole_depth')


class Bunch(object):
    """Attribute access to a dict's entries (utils.py:6-8)."""

    def __init__(self, adict):
        vars(self).update(adict)


def one_hot_vector(one_index, length):
    v = np.zeros(length)
    v[one_index] = 1.0
    return v


def vert_one_hot(one_index, length):
    return one_hot_vector(one_index, length).reshape(length, 1)


def softmax(U):
    """Numerically shifted softmax (utils.py:42-45)."""
    e = np.exp(U - np.max(U))
    return e / np.sum(e)


def compute_action_probabilities(action_features, weights, temperature):
    """Softmax policy over linear utilities features @ weights / temperature (utils.py:26-31)."""
    return softmax(action_features.dot(weights) / temperature)


def grad_of_log_action_probabilities(features, probabilities, action_index):
    """d/dw log pi(action): chosen features minus the probability-weighted mean features (utils.py:35-38)."""
    return features[action_index] - features.T.dot(probabilities)


# -- plots ---------------------------------------------------------------------------------------------
def _pyplot():
    try:
        import matplotlib
        matplotlib.use("Agg", force=False)
        import matplotlib.pyplot as plt
    except ImportError as exc:                                      # pragma: no cover - optional dependency
        raise ImportError("the plot_* helpers need matplotlib, which is not installed") from exc
    return plt


def _save_lines(plt, path, series, x=None):
    """One figure: every (label, y) in `series` as a line over x, legend, saved to `path`."""
    fig, ax = plt.subplots()
    for label, y in series:
        if x is None:
            ax.plot(y, label=label)
        else:
            ax.plot(x, y, label=label)
    ax.legend()
    fig.savefig(path)
    plt.close(fig)


def plot_learning_curve(plots_path, test_results, x_axis):
    """mean/median and max of test_results over axes (0, 2) -> mean_performance, max_performance."""
    plt = _pyplot()
    _save_lines(plt, os.path.join(plots_path, "mean_performance"),
                [("mean", np.mean(test_results, axis=(0, 2))), ("median", np.median(test_results, axis=(0, 2)))], x_axis)
    _save_lines(plt, os.path.join(plots_path, "max_performance"), [("max", np.max(test_results, axis=(0, 2)))], x_axis)


def _weight_paths(plt, plots_path, tested_weights, weights_storage, agent_ix, x_axis):
    names = FEATURE_NAMES
    _save_lines(plt, os.path.join(plots_path, "weight_paths_tested" + str(agent_ix)),
                [(names[i], tested_weights[:, i]) for i in range(tested_weights.shape[1])], x_axis)
    _save_lines(plt, os.path.join(plots_path, "weight_paths" + str(agent_ix)),
                [(names[i], weights_storage[:, i]) for i in range(weights_storage.shape[1])])


def plot_individual_agent(plots_path, tested_weights, test_results, weights_storage, agent_ix, x_axis):
    plt = _pyplot()
    _weight_paths(plt, plots_path, tested_weights, weights_storage, agent_ix, x_axis)
    _save_lines(plt, os.path.join(plots_path, "mean_performance" + str(agent_ix)),
                [("mean", np.mean(test_results, axis=1)), ("median", np.median(test_results, axis=1))], x_axis)


def plot_analysis(plots_path, tested_weights, test_results, weights_storage, agent_ix, x_axis):
    plt = _pyplot()
    _weight_paths(plt, plots_path, tested_weights, weights_storage, agent_ix, x_axis)
    step = np.diff(tested_weights, axis=0)
    _save_lines(plt, os.path.join(plots_path, "distances" + str(agent_ix)),
                [("l2 distance to previous", np.sqrt(np.sum(step ** 2, axis=1)))])
    rel = np.diff(tested_weights / np.abs(tested_weights[:, :1]), axis=0)
    _save_lines(plt, os.path.join(plots_path, "relative_distances" + str(agent_ix)),
                [("l2 RELATIVE distance to previous", np.sqrt(np.sum(rel ** 2, axis=1)))])
    _save_lines(plt, os.path.join(plots_path, "mean_performance" + str(agent_ix)),
                [("mean", np.mean(test_results, axis=1)), ("median", np.median(test_results, axis=1))], x_axis)
    _save_lines(plt, os.path.join(plots_path, "max_performance" + str(agent_ix)),
                [("max", np.max(test_results, axis=1))], x_axis)


# -- render --------------------------------------------------------------------------------------------
def print_board_to_string(state):
    """All stored rows (the 4 buffer rows included), top row first (utils.py:179-191)."""
    rows = np.asarray(state.representation)[::-1]
    return "\n" + "".join("|" + "".join("██" if v else "  " for v in row) + "|\n" for row in rows)
