"""Constant-velocity predictor (reference: trajnetbaselines/classical/constant_velocity.py:4-20)."""
import numpy as np

from ..data import paths_to_xy


def predict(input_paths, predict_all=True, n_predict=12, obs_length=9):
    xy = paths_to_xy(input_paths)
    curr_position = xy[-1]
    curr_velocity = xy[-1] - xy[-2]
    output_rel_scenes = np.array([i * curr_velocity for i in range(1, n_predict + 1)])
    output_scenes = curr_position + output_rel_scenes
    return {0: (output_scenes[-n_predict:, 0], output_scenes[-n_predict:, 1:])}
