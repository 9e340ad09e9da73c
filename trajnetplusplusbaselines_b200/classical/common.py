"""Host-side scene preparation shared by the classical adapters (NumPy, O(scene)).

Mirrors the adapter code around the third-party simulators in the reference
(trajnetbaselines/classical/socialforce.py:15-72, orca.py:14-82): which pedestrians are
simulated, their initial velocity (stride-3 finite difference) and their destination (linear
extrapolation of the observed path).
"""
import numpy as np


def split_paths(paths, obs_length):
    primary = paths[0]
    start_frame = primary[obs_length - 1].frame
    return start_frame


def initial_states(input_paths, start_frame, pred_length, dest_dict=None, dest_type='interp'):
    """-> (state [K, 6] float64: x, y, vx, vy, dx, dy; speeds [K]) for the K pedestrians present at
    `start_frame`, in path order (socialforce.py:15-55 / orca.py:14-58)."""
    rows, speeds = [], []
    for path in input_paths:
        ped_id = path[0].pedestrian
        past_path = [t for t in path if t.frame <= start_frame]
        future_path = [t for t in path if t.frame > start_frame]
        past_frames = [t.frame for t in past_path]
        len_path = len(past_path)
        if start_frame not in past_frames:
            continue
        curr = past_path[-1]
        if len_path >= 4:
            stride, prev = 3, past_path[-4]
        else:
            stride, prev = len_path - 1, past_path[-len_path]
        if stride == 0:
            v_x = v_y = speed = 0.0
        else:
            diff = np.array([curr.x - prev.x, curr.y - prev.y])
            theta = np.arctan2(diff[1], diff[0])
            speed = np.linalg.norm(diff) / (stride * 0.4)
            v_x, v_y = speed * np.cos(theta), speed * np.sin(theta)
        if dest_type == 'true':
            if dest_dict is None:
                raise ValueError
            d_x, d_y = dest_dict[ped_id]
        elif dest_type == 'interp':
            if len_path == 1:
                d_x, d_y = curr.x, curr.y
            else:   # interp1d(..., fill_value='extrapolate') evaluated at len-1+pred_length
                p1, p0 = past_path[-1], past_path[-2]
                d_x = p1.x + (p1.x - p0.x) * pred_length
                d_y = p1.y + (p1.y - p0.y) * pred_length
        elif dest_type == 'vel':
            d_x, d_y = pred_length * v_x, pred_length * v_y
        elif dest_type == 'pred_end':
            d_x, d_y = future_path[-1].x, future_path[-1].y
        else:
            raise NotImplementedError
        rows.append([curr.x, curr.y, v_x, v_y, d_x, d_y])
        speeds.append(speed)
    return np.array(rows, dtype=np.float64).reshape(-1, 6), np.array(speeds, dtype=np.float64)
