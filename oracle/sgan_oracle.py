"""CPU restatement of the S-GAN generator / discriminator forward -- TEST INFRASTRUCTURE (see
oracle/__init__.py).  Follows trajnetbaselines/sgan/sgan.py: LSTMGenerator.forward :301-394 (the
step is LSTM.step, restated in lstm_oracle.step), adding_noise :200-221, LSTMDiscriminator.forward
:524-581.  Pinned by tests/golden/sgan_golden.npz, which the unmodified reference produced with its
noise source patched to a fixed vector (oracle/make_sgan_golden.py).
"""
import numpy as np

from . import lstm_oracle as O

F32 = np.float32


def adding_noise(weights, h, c, noise):
    """h <- cat(ReLU(Linear(h)), z repeated for every track); c unchanged (sgan.py:200-221)."""
    w = weights["mlp_decoder_context.0.weight"].astype(F32)
    b = weights["mlp_decoder_context.0.bias"].astype(F32)
    new_h = np.maximum(h @ w.T + b, 0).astype(F32)
    z = np.tile(np.asarray(noise, dtype=F32)[None, :], (h.shape[0], 1))
    return np.concatenate([new_h, z], axis=1).astype(F32), c


def generator_forward(weights, pool_cfg, observed, batch_split, prediction_truth=None, n_predict=None, noise=None):
    """rel_pred_scene [S, M, 5], pred_scene [S, M, 2]; noise=None <-> no_noise=True."""
    between = None if noise is None else (lambda h, c: adding_noise(weights, h, c, noise))
    if prediction_truth is not None:
        # the generator chains (observed[-1:], prediction_truth[:-1]) (sgan.py:367-369): the last
        # teacher-forcing frame is never used (LSTM.forward chains the whole list, lstm.py:235-237)
        prediction_truth = prediction_truth[:-1]
    return O.forward(weights, pool_cfg, observed, batch_split, prediction_truth=prediction_truth,
                     n_predict=n_predict, between=between)


def discriminator_forward(weights, pool_cfg, observed, prediction, batch_split):
    """scores [B, 1] of the primaries: encoder-only LSTM over [observed; prediction], then the
    real_classifier MLP (an activation after every Linear, make_mlp sgan.py:34-45)."""
    seq = np.concatenate([np.asarray(observed, F32), np.asarray(prediction, F32)], axis=0)
    W = dict(weights)
    H = W["encoder.weight_hh"].shape[1]
    W.setdefault("hidden2normal.linear.weight", np.zeros((5, H), F32))     # the head is never read
    W.setdefault("hidden2normal.linear.bias", np.zeros((5,), F32))
    _, _, states = O.forward(W, pool_cfg, seq, batch_split, n_predict=1, return_states=True)
    h = states[-1][0][np.asarray(batch_split, dtype=np.int64)[:-1]]
    for k in (0, 2, 4):
        h = np.maximum(h @ W["real_classifier.%d.weight" % k].T + W["real_classifier.%d.bias" % k], 0).astype(F32)
    return h


def sgan_weights(kind, seed):
    """Seeded generator / discriminator weights in the reference's state_dict layout (numpy
    RandomState => identical on every machine; nothing but the seed is stored in the fixtures)."""
    import math
    gen = O.random_weights(kind, seed=seed)
    rng = np.random.RandomState(1000 + seed)
    H = gen["encoder.weight_hh"].shape[1]
    nd = 8
    k = 1.0 / math.sqrt(H)
    gen["mlp_decoder_context.0.weight"] = rng.uniform(-k, k, size=(H - nd, H)).astype(F32)
    gen["mlp_decoder_context.0.bias"] = rng.uniform(-k, k, size=(H - nd,)).astype(F32)
    base = O.random_weights(kind, seed=seed + 50)
    dis = {n: v for n, v in base.items() if not (n.startswith("decoder.") or n.startswith("hidden2normal."))}
    dims = [H, H // 2, H // 4, 1]
    for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):      # positive weights: the final ReLU stays open
        kk = 1.0 / math.sqrt(a)
        dis["real_classifier.%d.weight" % (2 * i)] = rng.uniform(0, kk, size=(b, a)).astype(F32)
        dis["real_classifier.%d.bias" % (2 * i)] = rng.uniform(0, kk, size=(b,)).astype(F32)
    return gen, dis


# ----------------------------------------------------------------------------------------
# VAE at test time (trajnetbaselines/vae/vae.py:188-315, add_noise :87-106)
# ----------------------------------------------------------------------------------------
def vae_forward(weights, pool_cfg, observed, batch_split, prediction_truth=None, n_predict=None, z=None):
    """One mode: h <- h * ReLU(fc z) between the observation encoder and the decoder."""
    W = dict(weights)
    for k in list(W):
        if k.startswith("obs_encoder."):
            W["encoder." + k[len("obs_encoder."):]] = W[k]

    def between(h, c):
        dec = np.maximum(np.asarray(z, F32) @ W["vae_decoder.fc.weight"].T + W["vae_decoder.fc.bias"], 0).astype(F32)
        return (h * dec).astype(F32), c
    return O.forward(W, pool_cfg, observed, batch_split, prediction_truth=prediction_truth, n_predict=n_predict,
                     between=between)


def vae_weights(kind, seed, latent_dim=128):
    import math
    base = O.random_weights(kind, seed=seed)
    rng = np.random.RandomState(2000 + seed)
    W = {}
    for k, v in base.items():
        W[("obs_encoder." + k[len("encoder."):]) if k.startswith("encoder.") else k] = v
    H = base["encoder.weight_hh"].shape[1]
    for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
        W["pred_encoder." + k] = rng.uniform(-0.08, 0.08, size=base["encoder." + k].shape).astype(F32)

    def lin(name, out_f, in_f, lo=None):
        kk = 1.0 / math.sqrt(in_f)
        W[name + ".weight"] = rng.uniform(-kk if lo is None else lo, kk, size=(out_f, in_f)).astype(F32)
        W[name + ".bias"] = rng.uniform(-kk if lo is None else lo, kk, size=(out_f,)).astype(F32)
    lin("vae_encoder_xy.fc_mu", latent_dim, 2 * H)
    lin("vae_encoder_xy.fc_var", latent_dim, 2 * H)
    lin("vae_encoder_x.fc_mu", latent_dim, H)
    lin("vae_encoder_x.fc_var", latent_dim, H)
    lin("vae_decoder.fc", H, latent_dim, lo=-0.02)          # mostly positive: the ReLU gate stays open
    return W
