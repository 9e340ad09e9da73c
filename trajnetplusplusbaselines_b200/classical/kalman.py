"""Kalman predictor with the reference's `predict` signature (CPU: BASELINE configs[0]).

Mirrors trajnetbaselines/classical/kalman.py:6-73.  pykalman's em / smooth / sample are replaced
by tb2_kalman_predict (host C++ in csrc/kalman.cu, float64).  The reference averages 5 noisy
`kf.sample` draws from the unseeded global NumPy RNG; `n_samples=5` reproduces that (same RNG
source, noise drawn from the fitted Q, R), `n_samples=0` returns the expectation.
"""
import ctypes

import numpy as np

from .. import _lib

_A = np.array([[1, 1, 0, 0], [0, 1, 0, 0], [0, 0, 1, 1], [0, 0, 0, 1]], dtype=np.float64)
_C = np.array([[1, 0, 0, 0], [0, 0, 1, 0]], dtype=np.float64)


def predict_tracks(tracks, n_predict=12, n_samples=5, em_iterations=10):
    """tracks: list of [T_i, 2] arrays -> [n_tracks, n_predict, 2] float64."""
    lib = _lib.load()
    tracks = [np.ascontiguousarray(t, dtype=np.float64) for t in tracks]
    n = len(tracks)
    offs = np.zeros(n + 1, dtype=np.int64)
    offs[1:] = np.cumsum([len(t) for t in tracks])
    obs = np.ascontiguousarray(np.concatenate(tracks, axis=0)) if n else np.zeros((0, 2))
    pred = np.zeros((n, n_predict, 2), dtype=np.float64)
    q = np.zeros((n, 4, 4), dtype=np.float64)
    r = np.zeros((n, 2, 2), dtype=np.float64)
    last = np.zeros((n, 4), dtype=np.float64)
    _lib.check(lib.tb2_kalman_predict(obs.ctypes.data, offs.ctypes.data, n, n_predict, em_iterations,
                                      pred.ctypes.data, q.ctypes.data, r.ctypes.data, last.ctypes.data))
    if n_samples:
        for i in range(n):   # kalman.py:53-60: mean of n_samples x kf.sample(n_predict + 1)[observations][1:]
            acc = np.zeros((n_predict, 2))
            for _ in range(n_samples):
                x = last[i].copy()
                np.random.multivariate_normal(np.zeros(2), r[i])          # z_0 is drawn, then dropped
                for k in range(n_predict):
                    x = _A @ x + np.random.multivariate_normal(np.zeros(4), q[i])
                    acc[k] += _C @ x + np.random.multivariate_normal(np.zeros(2), r[i])
            pred[i] = acc / n_samples
    return pred


def predict(paths, predict_all=True, n_predict=12, obs_length=9, n_samples=5):
    neighbours_tracks = []
    primary = paths[0]
    start_frame = primary[obs_length - 1].frame
    if not predict_all:
        paths = paths[0:1]
    tracks, is_primary = [], []
    for i, path in enumerate(paths):
        past_path = [t for t in path if t.frame <= start_frame]
        past_frames = [t.frame for t in past_path]
        if start_frame not in past_frames or len(past_path) < 2:
            continue
        tracks.append(np.array([(r.x, r.y) for r in past_path], dtype=np.float64))
        is_primary.append(i == 0)
    pred = predict_tracks(tracks, n_predict=n_predict, n_samples=n_samples)
    primary_track = None
    for p, prim in zip(pred, is_primary):
        if prim:
            primary_track = p
        else:
            neighbours_tracks.append(p)
    if len(neighbours_tracks):
        neighbours_tracks = np.array(neighbours_tracks).transpose(1, 0, 2)
    return {0: (primary_track, neighbours_tracks)}
