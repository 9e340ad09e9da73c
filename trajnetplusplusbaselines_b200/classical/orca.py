"""ORCA predictor with the reference's `predict` signature, simulated on the GPU.

Mirrors trajnetbaselines/classical/orca.py:10-134.  rvo2.PyRVOSimulator + the per-agent
doStep / getAgentPosition / setAgentPrefVelocity loop (3 FFI calls per agent per step) is
replaced by tb2_orca_simulate (csrc/classical.cu): one persistent kernel, one CTA per scene,
float arithmetic like RVO2.
"""
import ctypes

import numpy as np
import torch

from .. import _lib
from ..engine import SceneLayout, _ptr, _stream
from .common import initial_states

MAX_SPEED_MULTIPLIER = 1.3   # applied inside the kernel (orca.py:8,36)


def simulate_batch(pos, vel, goals, speeds, batch_split, orca_params=(1.5, 1.5, 0.4), n_steps=97,
                   sample_every=8, fps=20, max_neighbors=10, end_range=0.05, device=None):
    """pos, vel [A, 2]; goals [A, 2]; speeds [A]; -> [n_steps // sample_every, A, 2] float32."""
    _lib.require_cuda()
    lib = _lib.load()
    device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
    pos_t = torch.as_tensor(np.asarray(pos), dtype=torch.float32).to(device).contiguous()
    vel_t = torch.as_tensor(np.asarray(vel), dtype=torch.float32).to(device).contiguous()
    goal_t = torch.as_tensor(np.asarray(goals), dtype=torch.float64).to(device).contiguous()
    speed_t = torch.as_tensor(np.asarray(speeds), dtype=torch.float64).to(device).contiguous()
    layout = SceneLayout(batch_split, device=device)
    if layout.num_tracks != pos_t.shape[0]:
        raise ValueError("batch_split[-1] != number of agents")
    p = _lib.OrcaParams()
    p.time_step = 1.0 / fps
    p.neighbor_dist = float(orca_params[0])
    p.max_neighbors = int(max_neighbors)
    p.time_horizon = float(orca_params[1])
    p.radius = float(orca_params[2])
    p.end_range = float(end_range)
    p.n_steps, p.sample_every = int(n_steps), int(sample_every)
    out = torch.empty((n_steps // sample_every, pos_t.shape[0], 2), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _lib.check(lib.tb2_orca_simulate(layout.handle, ctypes.byref(p), _ptr(pos_t), _ptr(vel_t),
                                         _ptr(goal_t), _ptr(speed_t), _ptr(out), _stream(device)))
    return out


def predict(input_paths, dest_dict=None, dest_type='interp', orca_params=[1.5, 1.5, 0.4],
            predict_all=True, n_predict=12, obs_length=9):
    pred_length = n_predict
    primary = input_paths[0]
    start_frame = primary[obs_length - 1].frame
    state, speeds = initial_states(input_paths, start_frame, pred_length, dest_dict, dest_type)
    fps = 20
    sampling_rate = int(fps / 2.5)
    n_steps = sampling_rate * pred_length + 1          # orca.py:99
    states = simulate_batch(state[:, 0:2], state[:, 2:4], state[:, 4:6], speeds, [0, len(state)],
                            orca_params, n_steps=n_steps, sample_every=sampling_rate, fps=fps)
    states = states.cpu().numpy().astype(np.float64)
    primary_track = states[:, 0, 0:2]
    neighbours_tracks = states[:, 1:, 0:2]
    if not predict_all:
        neighbours_tracks = []
    return {0: (primary_track, neighbours_tracks)}
