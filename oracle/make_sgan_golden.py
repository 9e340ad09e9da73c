"""Golden vectors for the S-GAN generator / discriminator, produced by the UNMODIFIED reference
(trajnetbaselines/sgan/sgan.py) imported from /root/reference in the build container, with its noise
source (get_noise) patched to return a fixed vector so that the run is reproducible.

    python -m oracle.make_sgan_golden        -> tests/golden/sgan_golden.npz

Per case: the generator outputs (free-running and teacher-forced) and the discriminator scores; inputs
and weights are regenerated from seeds (lstm_oracle.synthetic_scenes, sgan_oracle.sgan_weights).
"""
import os

import numpy as np

from . import lstm_oracle as O
from . import sgan_oracle as SO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# name, pool kind, scenes, peds, ragged, nan tracks, data seed, weight seed, no_noise
SGAN_CASES = [
    ("sgan_vanilla", "vanilla", 5, 6, True, True, 41, 1, False),
    ("sgan_directional", "directional", 4, 7, True, True, 42, 2, False),
    ("sgan_social_small", "social_small", 4, 5, False, True, 43, 3, False),
    ("sgan_vanilla_nonoise", "vanilla", 3, 4, False, False, 44, 4, True),
]
NOISE = np.array([0.3, -1.2, 0.8, 0.05, -0.4, 1.7, -0.9, 0.6], dtype=np.float32)


def main():
    import torch
    from oracle.ref_shim import import_reference
    import_reference()
    import trajnetbaselines.sgan.sgan as ref
    from trajnetbaselines.lstm.gridbased_pooling import GridBasedPooling
    ref.get_noise = lambda shape, noise_type, device: torch.from_numpy(NOISE.copy())
    out = {"noise": NOISE}
    for name, kind, B, N, ragged, nan_tracks, dseed, wseed, no_noise in SGAN_CASES:
        xy, bs = O.synthetic_scenes(B, N, seed=dseed, ragged=ragged, nan_tracks=nan_tracks)
        spec = O.MODEL_SPECS[kind]
        gen = ref.LSTMGenerator(pool=GridBasedPooling(**spec) if spec else None, no_noise=no_noise)
        dis = ref.LSTMDiscriminator(pool=GridBasedPooling(**spec) if spec else None)
        Wg, Wd = SO.sgan_weights(kind, wseed)
        for module, W in ((gen, Wg), (dis, Wd)):
            sd = module.state_dict()
            missing = [k for k in sd if k not in W and not k.startswith("goal_embedding.")]
            assert not missing, missing
            sd.update({k: torch.from_numpy(v.copy()) for k, v in W.items() if k in sd})
            module.load_state_dict(sd)
        scene, split = torch.from_numpy(xy), torch.from_numpy(bs)
        goals = torch.zeros(xy.shape[1], 2)
        with torch.no_grad():
            rel, pred = gen(scene[:9], goals, split, n_predict=12)
            rel_tf, pred_tf = gen(scene[:9].clone(), goals, split, scene[9:-1].clone())
            scores_real = dis(scene[:9], scene[9:21], goals, split)
            scores_fake = dis(scene[:9], pred[-12:], goals, split)
        out[name + "/rel"], out[name + "/pred"] = rel.numpy(), pred.numpy()
        out[name + "/rel_tf"], out[name + "/pred_tf"] = rel_tf.numpy(), pred_tf.numpy()
        out[name + "/scores_real"], out[name + "/scores_fake"] = scores_real.numpy(), scores_fake.numpy()
        print(name, "pred", pred.shape, "scores", scores_real.numpy().ravel()[:3])
    path = os.path.join(ROOT, "tests", "golden", "sgan_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
