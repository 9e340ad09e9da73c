"""VAE forecaster + predictor with the reference's API, inference side.

Mirrors trajnetbaselines/vae/vae.py: VAE :26-315, VAEEncoder :317-332, VAEDecoder :334-345,
VAEPredictor :347-398, utils.sample_multivariate_distribution (vae/utils.py:4-24).  The step is the
LSTM step of lstm/lstm.py (same kernels, `obs_encoder` in the encoder slot); at test time the latent
sample rescales the encoder state, h <- h * ReLU(fc z) (add_noise :87-106 -> tb2_vae_scale_hidden),
and each of the num_modes decodes from that state.  The observation encoder runs once.  Same
constructor arguments and state_dict keys.  Training (prediction encoder, KL term) is not built:
model.train() + forward raises.
"""
import ctypes

import numpy as np
import torch

from .. import _lib
from ..data import paths_to_xy
from ..engine import _ptr, _stream
from ..lstm.lstm import LSTM, center_scene, drop_distant, inverse_scene  # noqa: F401
from ..lstm.modules import Hidden2Normal, InputEmbedding


def sample_multivariate_distribution(mean, var_log):
    """vae/utils.py:4-24: one N(mean, diag(exp(var_log))) sample per track (numpy RNG, like the reference)."""
    mean = mean.detach().cpu().numpy()
    std = np.exp(0.5 * var_log.detach().cpu().numpy())
    return torch.from_numpy((mean + std * np.random.standard_normal(mean.shape)).astype(np.float32))


class VAEEncoder(torch.nn.Module):
    """vae.py:317-332 (parameters; used by the reference at training time only when desire=True)."""

    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.input_dim, self.output_dim = input_dim, output_dim
        self.fc_mu = torch.nn.Linear(self.input_dim, self.output_dim // 2)
        self.fc_var = torch.nn.Linear(self.input_dim, self.output_dim // 2)
        self.relu = torch.nn.ReLU()

    def forward(self, inputs):
        inputs = torch.reshape(torch.stack(list(inputs)) if isinstance(inputs, (list, tuple)) else inputs,
                               (-1, self.input_dim))
        return self.relu(self.fc_mu(inputs)), 0.01 + self.relu(self.fc_var(inputs))


class VAEDecoder(torch.nn.Module):
    """vae.py:334-345."""

    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.input_dim, self.output_dim = input_dim, output_dim
        self.fc = torch.nn.Linear(self.input_dim, self.output_dim)
        self.relu = torch.nn.ReLU()

    def forward(self, inputs):
        return self.relu(self.fc(torch.reshape(inputs, (-1, self.input_dim))))


class VAE(torch.nn.Module):
    """vae.py:26-315.  `fixed_z` (tensor [num_modes, M, latent_dim]) replaces the random draws when set."""

    def __init__(self, embedding_dim=64, hidden_dim=128, pool=None, pool_to_input=True, goal_dim=None,
                 goal_flag=False, num_modes=1, latent_dim=128):
        super().__init__()
        body = LSTM(embedding_dim, hidden_dim, pool, pool_to_input, goal_dim, goal_flag)
        self._body = [body]                    # engine owner; its modules are registered below under the reference's names
        self.hidden_dim = hidden_dim
        self.embedding_dim = embedding_dim
        self.pool = pool
        self.pool_to_input = pool_to_input
        self.input_embedding = body.input_embedding
        self.goal_flag = goal_flag
        self.goal_dim = body.goal_dim
        self.goal_embedding = body.goal_embedding
        self.obs_encoder = body.encoder
        in_dim = body.encoder.weight_ih.shape[1]
        self.pred_encoder = torch.nn.LSTMCell(in_dim, hidden_dim)
        self.decoder = body.decoder
        self.hidden2normal = body.hidden2normal
        self.latent_dim = latent_dim
        self.num_modes = num_modes
        self.desire = True
        self.vae_encoder_xy = VAEEncoder(2 * hidden_dim, 2 * latent_dim)
        self.vae_encoder_x = VAEEncoder(hidden_dim, 2 * latent_dim)
        self.vae_decoder = VAEDecoder(latent_dim, hidden_dim)
        self.fixed_z = None

    def forward(self, observed, goals, batch_split, prediction_truth=None, n_predict=None):
        """(rel_pred_scene list, pred_scene list, z_distr_xy, z_distr_x), vae.py:188-315 in eval mode."""
        assert ((prediction_truth is None) + (n_predict is None)) == 1
        if self.training or (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())):
            raise NotImplementedError("VAE training (prediction encoder, KL term) is not built; use "
                                      "model.eval() under torch.no_grad()")
        if not self.desire:
            raise NotImplementedError("desire=False (latent prior from vae_encoder_x) is not built")
        body = self._body[0]
        handle = body._engine()
        device = handle.device
        layout = body._layouts.get(batch_split.tolist() if torch.is_tensor(batch_split) else batch_split,
                                   device=body._device())
        M = layout.num_tracks
        obs = body._to_device(observed, device)
        obs_length = int(obs.shape[0])
        truth = None
        if prediction_truth is not None:
            if isinstance(prediction_truth, (list, tuple)):
                prediction_truth = torch.stack(list(prediction_truth))
            truth = body._to_device(prediction_truth, device)
            n_decode = int(truth.shape[0])
            if n_decode == 0:
                truth = None
        else:
            n_decode = int(n_predict) - 1
        S, S_enc = obs_length - 1 + n_decode, obs_length - 1
        f32 = dict(dtype=torch.float32, device=device)
        normals0, positions0 = torch.empty((S, M, 5), **f32), torch.empty((S, M, 2), **f32)
        h0, c0 = torch.empty((M, self.hidden_dim), **f32), torch.empty((M, self.hidden_dim), **f32)
        handle.forward_steps(layout, obs, truth, n_decode, 0, S_enc, normals0, positions0, h0, c0)
        lib = _lib.load()
        w = self.vae_decoder.fc.weight.detach().to(device=device, dtype=torch.float32).contiguous()
        b = self.vae_decoder.fc.bias.detach().to(device=device, dtype=torch.float32).contiguous()
        rel_list, pred_list = [], []
        for k in range(self.num_modes):
            if self.fixed_z is not None:
                z = torch.as_tensor(self.fixed_z[k], dtype=torch.float32)
            else:      # prior N(0, exp(1) I): z_mu_obs = 0, z_var_log_obs = 1 (vae.py:277-278)
                z = sample_multivariate_distribution(torch.zeros(M, self.latent_dim), torch.ones(M, self.latent_dim))
            z = z.to(device).contiguous()
            h, c = h0.clone(), c0.clone()
            normals, positions = normals0.clone(), positions0.clone()
            with torch.cuda.device(device):
                _lib.check(lib.tb2_vae_scale_hidden(_ptr(w), _ptr(b), _ptr(z), _ptr(h), M, int(self.hidden_dim),
                                                    int(self.latent_dim), _stream(device)))
            handle.forward_steps(layout, obs, truth, n_decode, S_enc, S, normals, positions, h, c)
            if observed.device != device:
                normals, positions = body._to_host(normals, positions)
            rel_list.append(normals)
            pred_list.append(positions)
        return rel_list, pred_list, None, None


class VAEPredictor(object):
    """vae.py:347-398."""

    def __init__(self, model):
        self.model = model

    def save(self, state, filename):
        with open(filename, 'wb') as f:
            torch.save(self, f)
        with open(filename + '.state', 'wb') as f:
            torch.save(state, f)

    @staticmethod
    def load(filename):
        with open(filename, 'rb') as f:
            return torch.load(f, weights_only=False)

    def __call__(self, paths, scene_goal, n_predict=12, modes=1, predict_all=True, obs_length=9, start_length=0,
                 args=None):
        self.model.eval()
        self.model.num_modes = modes
        with torch.no_grad():
            xy = paths_to_xy(paths)
            batch_split = [0, xy.shape[1]]
            normalize = bool(getattr(args, 'normalize_scene', False))
            if normalize:
                xy, rotation, center, scene_goal = center_scene(xy, obs_length, goals=np.asarray(scene_goal))
            xy = torch.Tensor(xy)
            scene_goal = torch.Tensor(np.asarray(scene_goal))
            batch_split = torch.Tensor(batch_split).long()
            multimodal_outputs = {}
            _, output_scenes_list, _, _ = self.model(xy[start_length:obs_length], scene_goal, batch_split,
                                                     n_predict=n_predict)
            for num_p, output_scenes in enumerate(output_scenes_list):
                output_scenes = output_scenes.cpu().numpy()
                if normalize:
                    output_scenes = inverse_scene(output_scenes, rotation, center)
                output_primary = output_scenes[-n_predict:, 0]
                output_neighs = output_scenes[-n_predict:, 1:]
                multimodal_outputs[num_p] = [output_primary, output_neighs if num_p == 0 else []]
        return multimodal_outputs
