"""Batched evaluator path: what lstm/trajnet_evaluator.py:15-61 and
evaluator/write_utils.py do around the predictor -- read the test scenes of an ndjson file,
preprocess_test, predict, write_predictions -- with the per-scene joblib fan-out
(`Parallel(n_jobs=12)(delayed(predict_scene)...)`, trajnet_evaluator.py:61) replaced by chunks of
scenes going through ONE batched forward each (LSTMPredictor.predict_batch).
"""
import os

from .data import preprocess_test, read_ndjson_scenes, write_predictions


def load_test_scenes(filename, obs_length=9):
    """[(filename, scene_id, paths)] like evaluator/write_utils.load_test_datasets, already through
    preprocess_test (tracks that start after the observation period are dropped)."""
    name = os.path.basename(filename)
    return [(name, scene_id, preprocess_test(paths, obs_length)) for scene_id, paths in read_ndjson_scenes(filename)]


def predict_scenes(predictor, scenes, obs_length=9, pred_length=12, modes=1, chunk=1024, args=None):
    """Predictions for a list of (filename, scene_id, paths), in order.  A predictor with
    predict_batch (LSTMPredictor) gets `chunk` scenes per forward; any other predictor of the
    reference's call signature (S-GAN / VAE / classical) is called scene by scene."""
    out = []
    if hasattr(predictor, 'predict_batch') and modes == 1:
        for i in range(0, len(scenes), chunk):
            part = [paths for _, _, paths in scenes[i:i + chunk]]
            out.extend(predictor.predict_batch(part, n_predict=pred_length, obs_length=obs_length, args=args))
        return out
    import numpy as np
    for _, _, paths in scenes:
        out.append(predictor(paths, np.zeros((len(paths), 2)), n_predict=pred_length, obs_length=obs_length,
                             modes=modes, args=args))
    return out


def evaluate_file(predictor, infile, outfile, obs_length=9, pred_length=12, modes=1, chunk=1024, args=None):
    """ndjson in -> ndjson out (the records evaluator/write_utils.write_predictions appends).
    Returns the number of scenes written."""
    scenes = load_test_scenes(infile, obs_length)
    preds = predict_scenes(predictor, scenes, obs_length, pred_length, modes, chunk, args)
    if os.path.exists(outfile):
        os.remove(outfile)
    write_predictions(preds, scenes, outfile, obs_length=obs_length, pred_length=pred_length)
    return len(scenes)
