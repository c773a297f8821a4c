"""The reference's play flow (example_play.py:1-22): a 10x10 game, 100 placements, equal-weight directed arg-max.

The reference's script iterates the `(features, None)` tuple `get_after_states` returns and crashes in
`np.argmax`; this is the flow it intends -- unpack the tuple, pick the row with the largest directed sum.
Run it from anywhere: `python tetris/example_play.py [--quiet]`.
"""
import sys

import numpy as np

from game import Tetris


def main(n_placements=100, quiet=False, seed=None):
    if seed is not None:
        np.random.seed(seed)
    feature_directions = np.array([-1, -1, -1, -1, -1, -1, 1, -1])
    env = Tetris(10, 10, feature_directions=feature_directions)
    env.reset()
    total_reward = 0
    for _ in range(n_placements):
        after_state_features, _unused = env.get_after_states()
        i = int(np.argmax(after_state_features.sum(axis=1)))      # equal weights, directed
        observation, reward, done, lines = env.step(i)
        if not quiet:
            print(after_state_features[i])
            env.render()
        total_reward += reward
        if done:
            env.reset()
    print("placements %d  total reward %d" % (n_placements, total_reward))
    return total_reward


if __name__ == "__main__":
    main(quiet="--quiet" in sys.argv, seed=0 if "--seed0" in sys.argv else None)
