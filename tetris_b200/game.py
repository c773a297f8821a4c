"""`Tetris`: the reference's single-game environment (game.py:8-160) on the GPU kernels.

Drop-in for a policy written against the reference: same constructor, attributes (`current_state`,
`current_tetromino`, `afterstates`, `tetrominos`, `tetromino_sampler`, `feature_directions`, ...), return types
and error behaviour (IndexError for an out-of-range action, AttributeError when `step` precedes
`get_after_states`, ValueError for a feature type other than 'bcts').  Enumeration, line clearing, terminal
tests and features are computed by the CUDA kernels through `tetromino.*.get_after_states` / `state.State`;
this file only holds the game loop.  For many games at once use `tetris_b200.BatchedTetris`.
"""
import numpy as np

from . import BCTS_WEIGHTS, _single
from . import state
from . import tetromino
from .utils import print_board_to_string


class Tetris:
    """
    Feature order (game.py:10-18):
    0 rows_with_holes, 1 column_transitions, 2 holes, 3 landing height, 4 cumulative_wells,
    5 row_transitions, 6 eroded pieces, 7 hole_depth
    """

    def __init__(self, num_columns, num_rows, feature_directions=None, feature_type='bcts', num_features=8,
                 tetromino_size=4):
        self.num_columns, self.num_rows = num_columns, num_rows
        self.feature_directions, self.feature_type, self.num_features = feature_directions, feature_type, num_features
        self.tetromino_size = tetromino_size             # only sizes the buffer rows above the board (game.py:56)

        # reward = lines cleared + timestep_reward (+ loss_reward when the game ends), game.py:34-35,86-90
        self.loss_reward, self.timestep_reward = -100, -1

        # active piece set of the reference (game.py:38-39); assign `tetrominos` and `tetromino_sampler` to change
        # it, e.g. tetromino.standard_set(num_columns) for the seven tetrominoes of game.py:41-47
        self.tetrominos = tetromino.default_set(num_columns, feature_type, num_features)
        self.tetromino_sampler = tetromino.TetrominoSampler(self.tetrominos)
        self.current_state, self.current_tetromino = self.reset()

    def reset(self):
        """Empty board, next piece from the (persisting) sampler (game.py:53-63)."""
        board = np.zeros((self.num_rows + self.tetromino_size, self.num_columns), dtype=np.int_)
        self.current_state = state.State(representation=board,
                                         lowest_free_rows=np.zeros(self.num_columns, dtype=np.int_),
                                         num_features=self.num_features, feature_type=self.feature_type)
        self.current_tetromino = self.tetromino_sampler.next_tetromino()
        return self.current_state, self.current_tetromino

    def _feature_matrix(self, states):
        m = np.zeros((len(states), self.num_features))
        for k, s in enumerate(states):
            m[k] = s.get_features(direct_by=self.feature_directions)
        return m

    def get_after_states(self, include_terminal=False):
        """(features of the non-terminal afterstates [n_valid, 8] float64, None) -- or, with include_terminal, the
        matrix over all afterstates as second element (game.py:67-80).  Caches `self.afterstates` for `step`."""
        children = self.current_tetromino.get_after_states(self.current_state)
        legal = [c for c in children if not c.terminal_state]
        self.afterstates = np.empty(len(legal), dtype=object)
        self.afterstates[:] = legal
        action_features = self._feature_matrix(legal)
        return action_features, (self._feature_matrix(children) if include_terminal else None)

    def step(self, action):
        """Take the action-th non-terminal afterstate of the last get_after_states() (game.py:82-92)."""
        chosen = self.afterstates[action]                 # IndexError / AttributeError exactly as in the reference
        self.current_state = chosen
        lines = chosen.n_cleared_lines
        self.current_tetromino = self.tetromino_sampler.next_tetromino()      # drawn before the game-over test
        done = self.is_game_over(chosen)
        reward = lines + self.timestep_reward + (self.loss_reward if done else 0)
        return self.get_state(), reward, done, lines

    def is_game_over(self, state):
        """True when the current piece has no non-terminal placement on `state` (game.py:94-100)."""
        return all(c.terminal_state for c in self.current_tetromino.get_after_states(state))

    def get_best_policy(self):
        """Uniform distribution over the fitness arg-max among ALL afterstates, terminal included (game.py:102-107)."""
        children = self.current_tetromino.get_after_states(self.current_state)
        feats = np.stack([c.get_features() for c in children])
        scores = _single.ctx(self.num_columns, self.num_rows).fitness(feats, BCTS_WEIGHTS)
        best = (scores == scores.max()).astype(float)
        return best / best.sum()

    def fitness(self, state):
        """BCTS linear score of one state: float32 products summed left to right (game.py:109-120), on the device."""
        rep = np.asarray(state.representation)
        return _single.ctx(rep.shape[1], rep.shape[0] - 4).fitness(state.get_features(), BCTS_WEIGHTS)[0]

    def render(self):
        print(print_board_to_string(self.current_state))
        print(self.current_tetromino)

    def get_state(self):
        return self.current_state.get_features(direct_by=self.feature_directions)

    # -- rollout helpers (game.py:129-160) ----------------------------------------------------------
    def single_rollout(self, action, policy_function, length):
        """Return of `length - 1` policy steps after taking `action`; -1 if the game ends on the way (game.py:129-148).

        Bug-for-bug with the reference, pinned by tests/golden/rollouts.npz: the env is restored to the saved
        (state, piece) afterwards, the sampler is not rewound, and the cached `afterstates` list is NOT restored --
        it is left as the last policy step's list, so `action` indexes the parent's afterstates only if the caller
        runs get_after_states() before every call (perform_rollouts does not: from its second rollout on it takes a
        placement of a stale state, and raises IndexError when that list is shorter).  For rollouts that mean what
        they say, and for all actions of many envs at once, use BatchedTetris.rollout_values."""
        saved = (self.current_state, self.current_tetromino)
        if self.is_game_over(saved[0]):
            return -1
        rollout_return = 0
        done = self.step(action)[2]
        if done:
            rollout_return = -1
        else:
            for _ in range(length - 1):
                choice = policy_function(self.current_state, self.get_after_states(include_terminal=True)[0])
                _, reward, done, _ = self.step(choice)
                rollout_return += reward
                if done:
                    rollout_return = -1
                    break
        self.current_state, self.current_tetromino = saved
        return rollout_return

    def perform_rollouts(self, actions, policy_function, length=5, n=5):
        """Mean return of n rollouts for every action index (game.py:150-160)."""
        returns = [np.mean([self.single_rollout(a, policy_function, length) for _ in range(n)])
                   for a in range(len(actions))]
        return list(actions), returns
