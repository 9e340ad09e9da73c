"""Minimal stand-ins for the trajnetplusplustools types the predictor boundary touches.

The reference imports `trajnetplusplustools` (not vendored, absent from this image) for
`TrackRow` and `Reader.paths_to_xy` (lstm/lstm.py:289, classical/*.py).  The on-disk format is
visible in the reference's DATA_BLOCK/*.ndjson; these helpers implement just that.
"""
import json
from collections import namedtuple, defaultdict

import numpy as np

TrackRow = namedtuple('TrackRow', ['frame', 'pedestrian', 'x', 'y', 'prediction_number', 'scene_id'])
TrackRow.__new__.__defaults__ = (None, None, None, None, None, None)
SceneRow = namedtuple('SceneRow', ['scene', 'pedestrian', 'start', 'end', 'fps', 'tag'])
SceneRow.__new__.__defaults__ = (None, None, None, None, None, None)


def paths_to_xy(paths):
    """list of paths (primary first) -> xy [n_frames, n_peds, 2] float64, NaN where absent.

    Same contract as trajnetplusplustools.Reader.paths_to_xy: the frames are the SORTED set of the
    primary pedestrian's frames, and a pedestrian without a single row in those frames is dropped (it
    would be an all-NaN column that the writer later emits as NaN track rows).
    """
    if paths and paths[0] and not isinstance(paths[0][0], tuple):     # rows with attributes only: normalise to tuples
        paths = [[(r.frame, r.pedestrian, r.x, r.y) for r in path] for path in paths]
    frames = sorted({r[0] for r in paths[0]})              # TrackRow is a tuple: (frame, pedestrian, x, y, ...)
    frame_index = {f: i for i, f in enumerate(frames)}
    n_frames = len(frames)
    nan = float('nan')
    columns = []                                            # one flat [x0, y0, x1, y1, ...] list per kept pedestrian
    for path in paths:
        col = None
        for r in path:
            i = frame_index.get(r[0])
            if i is not None:
                if col is None:
                    col = [nan] * (2 * n_frames)
                col[2 * i] = r[2]
                col[2 * i + 1] = r[3]
        if col is not None:
            columns.append(col)
    if not columns:
        return np.full((n_frames, 0, 2), np.nan)
    xy = np.array(columns, dtype=np.float64).reshape(len(columns), n_frames, 2)
    return np.ascontiguousarray(xy.transpose(1, 0, 2))


def read_ndjson_scenes(filename):
    """Yield (scene_id, paths) with the primary pedestrian first (TrajNet++ ndjson)."""
    tracks_by_frame = defaultdict(list)
    scenes = []
    with open(filename) as f:
        lines = [line for line in f if line.strip()]
    # one C-level parse of the whole file instead of a json.loads call per line
    for d in json.loads('[' + ','.join(lines) + ']') if lines else ():
        if 'track' in d:
            t = d['track']
            row = TrackRow(t['f'], t['p'], t['x'], t['y'], t.get('prediction_number'), t.get('scene_id'))
            tracks_by_frame[row.frame].append(row)
        elif 'scene' in d:
            s = d['scene']
            scenes.append(SceneRow(s['id'], s['p'], s['s'], s['e'], s.get('fps'), s.get('tag')))
    for s in scenes:
        by_ped = defaultdict(list)
        for frame in range(s.start, s.end + 1):
            for r in tracks_by_frame.get(frame, ()):
                by_ped[r.pedestrian].append(r)
        if s.pedestrian not in by_ped:
            continue
        paths = [by_ped[s.pedestrian]] + [p for pid, p in by_ped.items() if pid != s.pedestrian]
        yield s.scene, paths


def preprocess_test(scene, obs_len):
    """Drop tracks that only appear after the observation period (evaluator/write_utils.py:31-39)."""
    obs_frames = [row.frame for row in scene[0]][:obs_len]
    last_obs_frame = obs_frames[-1]
    return [[row for row in ped if row.frame <= last_obs_frame]
            for ped in scene if ped[0].frame <= last_obs_frame]


def trajnet_line(row):
    """One ndjson line for a SceneRow / TrackRow (the trajnetplusplustools.writers.trajnet format as it
    appears in the reference's DATA_BLOCK files: coordinates rounded to 2 decimals)."""
    if isinstance(row, SceneRow):
        return json.dumps({'scene': {'id': row.scene, 'p': row.pedestrian, 's': row.start, 'e': row.end,
                                     'fps': row.fps, 'tag': row.tag}})
    x, y = round(float(row.x), 2), round(float(row.y), 2)
    if row.prediction_number is None:
        return json.dumps({'track': {'f': row.frame, 'p': row.pedestrian, 'x': x, 'y': y}})
    return json.dumps({'track': {'f': row.frame, 'p': row.pedestrian, 'x': x, 'y': y,
                                 'prediction_number': row.prediction_number, 'scene_id': row.scene_id}})


def write_predictions(pred_list, scenes, filename, obs_length=9, pred_length=12):
    """Append the predictions of a list of scenes to an ndjson file -- same records, in the same
    order, as evaluator/write_utils.py:42-81 (which goes through trajnetplusplustools.writers).

    pred_list : per scene {mode: [primary [pred_length, 2], neighbours [pred_length, K, 2] or []]}
    scenes    : per scene (anything, scene_id, paths) as the reference evaluator holds them
    """
    seq_length = obs_length + pred_length

    def track_line(frame, ped, x, y, mode, scene_id):
        # same text as trajnet_line(TrackRow(...)) without building the dict / calling json.dumps per row
        x, y = round(float(x), 2), round(float(y), 2)
        if (type(frame) is int and type(ped) is int and type(mode) is int and type(scene_id) is int
                and x - x == 0.0 and y - y == 0.0):          # finite
            return '{"track": {"f": %d, "p": %d, "x": %r, "y": %r, "prediction_number": %d, "scene_id": %d}}\n' % (
                frame, ped, x, y, mode, scene_id)
        return trajnet_line(TrackRow(frame, ped, x, y, mode, scene_id)) + '\n'

    with open(filename, "a") as out:
        for predictions, (_, scene_id, paths) in zip(pred_list, scenes):
            observed_path = paths[0]
            frame_diff = observed_path[1].frame - observed_path[0].frame
            first_frame = observed_path[obs_length - 1].frame + frame_diff
            ped_id = observed_path[0].pedestrian
            neigh_ids = [p[0].pedestrian for p in paths[1:]]
            out.write(trajnet_line(SceneRow(scene_id, ped_id, observed_path[0].frame,
                                            observed_path[0].frame + (seq_length - 1) * frame_diff, 2.5, 0)))
            out.write('\n')
            for m in range(len(predictions)):
                prediction, neigh_predictions = predictions[m]
                rows = [track_line(first_frame + i * frame_diff, ped_id, prediction[i][0], prediction[i][1], m, scene_id)
                        for i in range(len(prediction))]
                if len(neigh_predictions):
                    neigh = np.asarray(neigh_predictions).tolist()          # python floats once, not per element
                    for n in range(len(neigh[0])):
                        rows.extend(track_line(first_frame + j * frame_diff, neigh_ids[n], neigh[j][n][0], neigh[j][n][1],
                                               m, scene_id) for j in range(len(neigh)))
                out.write(''.join(rows))
