"""ndjson boundary: reader / writer either side of the batched predictor."""
import json
import os

import numpy as np

from trajnetplusplusbaselines_b200.data import (SceneRow, TrackRow, paths_to_xy, preprocess_test,
                                                read_ndjson_scenes, trajnet_line, write_predictions)


def _scene(scene_id, n_peds, start, rng):
    paths = []
    for p in range(n_peds):
        x0 = rng.randn(2)
        paths.append([TrackRow(start + 10 * t, 100 * scene_id + p, float(x0[0] + 0.1 * t), float(x0[1] - 0.05 * t))
                      for t in range(21)])
    return paths


def test_writer_matches_data_block_line_format():
    # the two record kinds exactly as they appear in the reference's DATA_BLOCK/*.ndjson
    assert json.loads(trajnet_line(SceneRow(3, 7, 10, 210, 2.5, 0))) == \
        {"scene": {"id": 3, "p": 7, "s": 10, "e": 210, "fps": 2.5, "tag": 0}}
    assert json.loads(trajnet_line(TrackRow(10, 7, 1.23456, -2.5))) == {"track": {"f": 10, "p": 7, "x": 1.23, "y": -2.5}}
    assert json.loads(trajnet_line(TrackRow(10, 7, 1.0, 2.0, 0, 3)))["track"]["prediction_number"] == 0


def test_write_then_read_round_trip(tmp_path):
    rng = np.random.RandomState(0)
    scenes = [("file", sid, _scene(sid, n, 1000 * sid, rng)) for sid, n in ((1, 3), (2, 1), (5, 4))]
    preds = []
    for _, _, paths in scenes:
        K = len(paths) - 1
        prim = rng.randn(12, 2)
        neigh = rng.randn(12, K, 2) if K else []
        preds.append({0: [prim, neigh]})
    fn = os.path.join(tmp_path, "pred.ndjson")
    write_predictions(preds, scenes, fn, obs_length=9, pred_length=12)
    got = {sid: paths for sid, paths in read_ndjson_scenes(fn)}
    assert sorted(got) == [1, 2, 5]
    for (_, sid, paths), pred in zip(scenes, preds):
        back = got[sid]
        assert back[0][0].pedestrian == paths[0][0].pedestrian            # primary first
        assert len(back) == len(paths)
        assert [r.frame for r in back[0]] == [paths[0][8].frame + 10 * (k + 1) for k in range(12)]
        assert np.allclose([[r.x, r.y] for r in back[0]], np.round(pred[0][0], 2))
        by_id = {p[0].pedestrian: p for p in back[1:]}
        for n, path in enumerate(paths[1:]):
            assert np.allclose([[r.x, r.y] for r in by_id[path[0].pedestrian]], np.round(pred[0][1][:, n], 2))


def test_preprocess_test_drops_late_tracks():
    rng = np.random.RandomState(1)
    paths = _scene(1, 3, 0, rng)
    paths[2] = [r for r in paths[2] if r.frame >= 100]        # appears after the 9 observed frames
    out = preprocess_test(paths, 9)
    assert len(out) == 2 and all(r.frame <= 80 for p in out for r in p)
    assert paths_to_xy(out).shape == (9, 2, 2)


class _ConstantVelocity:
    """Stand-in with the reference's predictor call signature (CPU, no CUDA)."""

    def __call__(self, paths, scene_goal, n_predict=12, modes=1, predict_all=True, obs_length=9, start_length=0,
                 args=None):
        xy = paths_to_xy(paths)
        v = xy[obs_length - 1] - xy[obs_length - 2]
        pred = xy[obs_length - 1][None] + np.arange(1, n_predict + 1)[:, None, None] * v[None]
        return {0: [pred[:, 0], pred[:, 1:]]}


def test_evaluate_file_ndjson_in_ndjson_out(tmp_path):
    """load_test_scenes -> predict_scenes -> write_predictions on a file written in the DATA_BLOCK format."""
    from trajnetplusplusbaselines_b200.evaluator import evaluate_file, load_test_scenes
    rng = np.random.RandomState(2)
    infile, outfile = os.path.join(tmp_path, "in.ndjson"), os.path.join(tmp_path, "out.ndjson")
    truth = {}
    with open(infile, "w") as f:
        for sid, n in ((0, 3), (1, 2)):
            paths = _scene(sid, n, 5000 * sid, rng)
            late = [TrackRow(5000 * sid + 10 * t, 100 * sid + 50, 0.0, 0.1 * t) for t in range(12, 21)]   # enters after obs
            f.write(trajnet_line(SceneRow(sid, paths[0][0].pedestrian, paths[0][0].frame, paths[0][-1].frame, 2.5, 0)) + "\n")
            for p in paths + [late]:
                for r in p:
                    f.write(trajnet_line(r) + "\n")
            truth[sid] = paths
    scenes = load_test_scenes(infile, obs_length=9)
    assert [len(paths) for _, _, paths in scenes] == [3, 2]              # the late track is dropped
    assert all(len(p) == 9 for _, _, paths in scenes for p in paths)     # only the observed frames remain
    assert evaluate_file(_ConstantVelocity(), infile, outfile) == 2
    got = {sid: paths for sid, paths in read_ndjson_scenes(outfile)}
    for sid, paths in truth.items():
        xy = paths_to_xy(paths)
        v = xy[8, 0] - xy[7, 0]
        want = np.round(xy[8, 0][None] + np.arange(1, 13)[:, None] * v[None], 2)
        # the ndjson rows carry two decimals: compare in hundredths; a value within float rounding of x.xx5 may
        # round the other way, those entries are counted
        units = np.rint(np.abs(np.array([[r.x, r.y] for r in got[sid][0]]) - want) * 100).astype(int)
        assert units.max() <= 1 and int((units > 0).sum()) <= 1, units


def test_get_predictions_writes_one_file_per_dataset(tmp_path):
    """Directory layout of lstm/trajnet_evaluator.get_predictions: <root>/test/*.ndjson ->
    <root>/test_pred/<model>_modes<k>/<dataset>.ndjson; an existing model folder is skipped."""
    import types
    from trajnetplusplusbaselines_b200.evaluator import get_predictions
    rng = np.random.RandomState(4)
    test_dir = os.path.join(tmp_path, "test")
    os.makedirs(test_dir)
    for name, sids in (("a.ndjson", (0, 1)), ("b.ndjson", (2,))):
        with open(os.path.join(test_dir, name), "w") as f:
            for sid in sids:
                paths = _scene(sid, 3, 3000 * sid, rng)
                f.write(trajnet_line(SceneRow(sid, paths[0][0].pedestrian, paths[0][0].frame, paths[0][-1].frame, 2.5, 0)) + "\n")
                for p in paths:
                    for r in p:
                        f.write(trajnet_line(r) + "\n")
    args = types.SimpleNamespace(path=os.path.join(tmp_path, "test_pred") + os.sep, output=["models/cv.pkl"], modes=1,
                                 obs_length=9, pred_length=12, chunk=2, normalize_scene=False)
    os.makedirs(args.path)
    assert get_predictions(args, load_predictor=lambda fn: _ConstantVelocity()) == {"cv_modes1": 3}
    out_dir = os.path.join(args.path, "cv_modes1")
    assert sorted(os.listdir(out_dir)) == ["a.ndjson", "b.ndjson"]
    assert len(list(read_ndjson_scenes(os.path.join(out_dir, "a.ndjson")))) == 2
    assert get_predictions(args, load_predictor=lambda fn: _ConstantVelocity()) == {}      # skipped: already there


def test_fast_writer_is_byte_identical_to_the_line_writer(tmp_path):
    """write_predictions formats track rows directly; the text must equal trajnet_line row by row, also for
    negative zero, integral values, rounding ties and NaN neighbours."""
    rng = np.random.RandomState(7)
    scenes = [("f", 3, _scene(3, 4, 700, rng))]
    prim = rng.randn(12, 2) * 30
    prim[0] = (-0.0, 2.0)
    prim[1] = (0.125, -0.375)
    prim[2] = (1e-9, 123456.789)
    neigh = rng.randn(12, 3, 2)
    neigh[5:, 1] = np.nan
    fn = os.path.join(tmp_path, "p.ndjson")
    write_predictions([{0: [prim, neigh]}], scenes, fn)
    lines = open(fn).read().splitlines()
    first = scenes[0][2][0][8].frame + 10
    want = [trajnet_line(SceneRow(3, scenes[0][2][0][0].pedestrian, 700, 700 + 200, 2.5, 0))]
    want += [trajnet_line(TrackRow(first + 10 * i, scenes[0][2][0][0].pedestrian, prim[i, 0], prim[i, 1], 0, 3)) for i in range(12)]
    for n in range(3):
        want += [trajnet_line(TrackRow(first + 10 * j, scenes[0][2][n + 1][0].pedestrian, neigh[j, n, 0], neigh[j, n, 1], 0, 3))
                 for j in range(12)]
    assert lines == want


def test_paths_to_xy_contract():
    """trajnetplusplustools.Reader.paths_to_xy semantics: frames = sorted set of the primary's frames, pedestrians
    without a row in those frames are dropped, rows outside them ignored; rows may be plain objects with attributes."""
    import types
    from trajnetplusplusbaselines_b200.data import TrackRow, paths_to_xy
    primary = [TrackRow(30, 1, 3.0, 3.5), TrackRow(10, 1, 1.0, 1.5), TrackRow(20, 1, 2.0, 2.5)]     # unsorted on purpose
    other = [TrackRow(20, 2, 7.0, 7.5), TrackRow(40, 2, 9.0, 9.5)]                                  # frame 40 is outside
    ghost = [TrackRow(50, 3, 0.0, 0.0)]                                                            # never inside: dropped
    xy = paths_to_xy([primary, other, ghost])
    assert xy.shape == (3, 2, 2) and xy.dtype == np.float64 and xy.flags["C_CONTIGUOUS"]
    assert xy[:, 0].tolist() == [[1.0, 1.5], [2.0, 2.5], [3.0, 3.5]]
    assert np.isnan(xy[0, 1]).all() and xy[1, 1].tolist() == [7.0, 7.5] and np.isnan(xy[2, 1]).all()
    rows = [[types.SimpleNamespace(frame=r.frame, pedestrian=r.pedestrian, x=r.x, y=r.y) for r in path]
            for path in (primary, other, ghost)]
    assert np.array_equal(np.isnan(paths_to_xy(rows)), np.isnan(xy)) and np.nanmax(np.abs(paths_to_xy(rows) - xy)) == 0.0
