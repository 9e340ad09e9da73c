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

    Frames are those of the primary pedestrian, like trajnetplusplustools.Reader.paths_to_xy.
    """
    frames = [r.frame for r in paths[0]]
    frame_index = {f: i for i, f in enumerate(frames)}
    xy = np.full((len(frames), len(paths), 2), np.nan)
    for p, path in enumerate(paths):
        for r in path:
            i = frame_index.get(r.frame)
            if i is not None:
                xy[i, p] = (r.x, r.y)
    return xy


def read_ndjson_scenes(filename):
    """Yield (scene_id, paths) with the primary pedestrian first (TrajNet++ ndjson)."""
    tracks_by_frame = defaultdict(list)
    scenes = []
    with open(filename) as f:
        for line in f:
            if not line.strip():
                continue
            d = json.loads(line)
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
