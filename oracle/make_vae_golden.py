"""Golden vectors for the VAE forecaster at test time, produced by the UNMODIFIED reference
(trajnetbaselines/vae/vae.py) imported from /root/reference in the build container, with its latent
sampler (vae.utils.sample_multivariate_distribution) patched to return fixed samples.

    python -m oracle.make_vae_golden        -> tests/golden/vae_golden.npz
"""
import os

import numpy as np

from . import lstm_oracle as O
from . import sgan_oracle as SO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# name, pool kind, scenes, peds, ragged, nan tracks, data seed, weight seed, modes
VAE_CASES = [
    ("vae_vanilla", "vanilla", 5, 6, True, True, 51, 1, 2),
    ("vae_directional", "directional", 4, 7, False, True, 52, 2, 2),
    ("vae_social_small", "social_small", 3, 5, True, False, 53, 3, 1),
]


def fixed_z(name, modes, M, latent_dim=128):
    rng = np.random.RandomState(abs(hash(name)) % 1000 if False else sum(map(ord, name)))
    return (rng.standard_normal((modes, M, latent_dim)) * 1.6).astype(np.float32)


def main():
    import torch
    from oracle.ref_shim import import_reference
    import_reference()
    import trajnetbaselines.vae.vae as ref
    from trajnetbaselines.lstm.gridbased_pooling import GridBasedPooling
    out = {}
    for name, kind, B, N, ragged, nan_tracks, dseed, wseed, modes in VAE_CASES:
        xy, bs = O.synthetic_scenes(B, N, seed=dseed, ragged=ragged, nan_tracks=nan_tracks)
        spec = O.MODEL_SPECS[kind]
        model = ref.VAE(pool=GridBasedPooling(**spec) if spec else None, num_modes=modes)
        W = SO.vae_weights(kind, wseed)
        sd = model.state_dict()
        missing = [k for k in sd if k not in W and not k.startswith("goal_embedding.")]
        assert not missing, missing
        sd.update({k: torch.from_numpy(v.copy()) for k, v in W.items() if k in sd})
        model.load_state_dict(sd)
        model.eval()
        z = fixed_z(name, modes, xy.shape[1])
        calls = iter(range(modes))
        ref.sample_multivariate_distribution = lambda mean, var_log: torch.from_numpy(z[next(calls)].copy())
        scene, split = torch.from_numpy(xy), torch.from_numpy(bs)
        with torch.no_grad():
            rel_list, pred_list, _, _ = model(scene[:9], torch.zeros(xy.shape[1], 2), split, n_predict=12)
        for k in range(modes):
            out["%s/rel%d" % (name, k)] = rel_list[k].numpy()
            out["%s/pred%d" % (name, k)] = pred_list[k].numpy()
        print(name, [p.shape for p in pred_list])
    path = os.path.join(ROOT, "tests", "golden", "vae_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
