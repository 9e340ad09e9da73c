"""Thin host-side owner of the C-ABI handles (model weights, scene layout, scratch).

PyTorch is plumbing here: it owns device memory and streams; every computation on the hot
path happens inside libtrajnet_b200.so.
"""
import ctypes
import weakref
from collections import OrderedDict

import torch

from . import _lib


# Parameter changes are detected through (data_ptr, _version) of every parameter.  Fused / foreach
# optimizers (torch.optim.Adam(fused=True)) update parameters without bumping `_version`, so every
# optimizer step additionally advances this epoch, which is part of the weight key.
_optimizer_epoch = [0]


def _on_optimizer_step(optimizer, args, kwargs):
    _optimizer_epoch[0] += 1


try:
    from torch.optim.optimizer import register_optimizer_step_post_hook
    register_optimizer_step_post_hook(_on_optimizer_step)
except Exception:      # very old torch: training forwards re-upload unconditionally (see LSTM._engine)
    pass


def weights_key(module):
    """Changes whenever a parameter of `module` may have changed."""
    return (_optimizer_epoch[0],) + tuple((p.data_ptr(), p._version) for p in module.parameters())


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _device_of(device):
    """Resolved CUDA device (the current one when `device` is None or an index-less 'cuda')."""
    device = torch.device('cuda' if device is None else device)
    if device.type != 'cuda':
        raise RuntimeError("scene layouts live in device memory: got device %s" % device)
    if device.index is None:
        device = torch.device('cuda', torch.cuda.current_device())
    return device


class SceneLayout:
    """tb2_layout wrapper: the `batch_split` partition of tracks into scenes.  The handle owns device
    buffers, allocated on `device` (default: the current CUDA device); a layout must only be used with
    models / tensors of that device."""

    def __init__(self, batch_split, pad_to_batch_max=True, device=None):
        _lib.require_cuda()
        lib = _lib.load()
        self.device = _device_of(device)
        offs = [int(v) for v in batch_split]
        self.offsets = offs
        arr = (ctypes.c_int64 * len(offs))(*offs)
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.tb2_layout_create(arr, len(offs) - 1, ctypes.byref(handle)))
        self.handle = handle
        self.num_scenes = len(offs) - 1
        self.num_tracks = offs[-1]
        self.max_scene = int(lib.tb2_layout_max_scene(handle))
        self.pad_to_batch_max = bool(pad_to_batch_max)
        if not pad_to_batch_max:     # evaluator semantics: every scene as if called on its own
            _lib.check(lib.tb2_layout_set_padding(handle, 0))
        self._finalizer = weakref.finalize(self, lib.tb2_layout_destroy, handle)


class LayoutCache:
    def __init__(self, capacity=8):
        self.capacity = capacity
        self._items = OrderedDict()

    def get(self, batch_split, pad_to_batch_max=True, device=None):
        device = _device_of(device)
        key = tuple(int(v) for v in batch_split) + (bool(pad_to_batch_max), device.index)
        item = self._items.get(key)
        if item is None:
            item = SceneLayout(key[:-2], pad_to_batch_max, device)
            self._items[key] = item
            if len(self._items) > self.capacity:
                self._items.popitem(last=False)
        else:
            self._items.move_to_end(key)
        return item


def plug_getstate(module):
    """`__getstate__` of the pooling modules: the per-process device handles of the stand-alone plug path (model
    handle, layouts, zero LSTM-cell weights, interaction-encoder state bookkeeping) are never pickled
    (LSTMPredictor.save pickles the whole model, lstm.py:270-277); they are rebuilt lazily after loading."""
    state = module.__dict__.copy()
    state.pop('_compiled_call_impl', None)         # like torch.nn.Module.__getstate__
    if '_handle' in state:
        state['_handle'] = None
    if '_layouts' in state:
        state['_layouts'] = LayoutCache()
    state.pop('_standalone_dummy', None)
    state.pop('_state_tracks', None)
    if '_reset_pending' in state:
        state['_reset_pending'] = True
    return state


class ModelHandle:
    """tb2_lstm wrapper: configuration + repacked weights on one device."""

    def __init__(self, config, device):
        _lib.require_cuda()
        lib = _lib.load()
        self.device = torch.device(device)
        self.config = config
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.tb2_lstm_create(ctypes.byref(config), ctypes.byref(handle)))
        self.handle = handle
        self._finalizer = weakref.finalize(self, lib.tb2_lstm_destroy, handle)
        self._weights_key = None
        self._workspace = None

    def weights_struct(self, named):
        """tb2_lstm_weights from a dict field name -> tensor (or list of 3 for the MLP)."""
        w = _lib.LstmWeights()
        keep = []
        for field, value in named.items():
            if isinstance(value, (list, tuple)):
                arr = getattr(w, field)
                for i, t in enumerate(value):
                    if t is not None:
                        t = self._prep(t)
                        keep.append(t)
                        arr[i] = t.data_ptr()
            elif value is not None:
                t = self._prep(value)
                keep.append(t)
                setattr(w, field, t.data_ptr())
        return w, keep

    def set_weights(self, named, key=None, force=False):
        """named: dict field name -> CUDA fp32 contiguous tensor (or list of 3 for the MLP)."""
        if not force and key is not None and key == self._weights_key:
            return
        lib = _lib.load()
        w = _lib.LstmWeights()
        keep = []
        for field, value in named.items():
            if isinstance(value, (list, tuple)):
                arr = getattr(w, field)
                for i, t in enumerate(value):
                    if t is not None:
                        t = self._prep(t)
                        keep.append(t)
                        arr[i] = t.data_ptr()
            elif value is not None:
                t = self._prep(value)
                keep.append(t)
                setattr(w, field, t.data_ptr())
        with torch.cuda.device(self.device):
            _lib.check(lib.tb2_lstm_set_weights(self.handle, ctypes.byref(w), _stream(self.device)))
        # the repack kernels read `keep` asynchronously on the current stream; record usage
        for t in keep:
            t.record_stream(torch.cuda.current_stream(self.device))
        self._weights_key = key

    def _prep(self, t):
        t = t.detach()
        if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(device=self.device, dtype=torch.float32).contiguous()
        return t

    def workspace(self, layout):
        lib = _lib.load()
        need = int(lib.tb2_lstm_workspace_bytes(self.handle, layout.handle))
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._workspace, need

    # -- compute entry points ---------------------------------------------------------------
    def grid_indices(self, layout, obs):
        lib = _lib.load()
        nm1 = max(layout.max_scene - 1, 0)
        cells = torch.empty((layout.num_tracks, nm1), dtype=torch.int32, device=self.device)
        flags = torch.empty((layout.num_tracks, nm1), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(lib.tb2_grid_indices(self.handle, layout.handle, _ptr(obs), _ptr(cells),
                                            _ptr(flags), _stream(self.device)))
        return cells, flags

    def pool_forward(self, layout, hidden, obs1, obs2, out_dim):
        lib = _lib.load()
        ws, need = self.workspace(layout)
        out = torch.empty((layout.num_tracks, out_dim), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(lib.tb2_pool_forward(self.handle, layout.handle, _ptr(hidden), _ptr(obs1),
                                            _ptr(obs2), _ptr(out), _ptr(ws), need, _stream(self.device)))
        return out

    def pool_state_reset(self, layout):
        """tb2_pool_state_reset: zero the interaction-encoder LSTM state of a stateful pool in this handle's workspace."""
        lib = _lib.load()
        ws, need = self.workspace(layout)
        with torch.cuda.device(self.device):
            _lib.check(lib.tb2_pool_state_reset(self.handle, layout.handle, _ptr(ws), need, _stream(self.device)))

    def step_forward(self, layout, phase, obs1, obs2, h, c):
        """One step; h, c updated in place.  Returns (normal [M,5], pos [M,2])."""
        lib = _lib.load()
        ws, need = self.workspace(layout)
        M = layout.num_tracks
        normal = torch.empty((M, 5), dtype=torch.float32, device=self.device)
        pos = torch.empty((M, 2), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(lib.tb2_lstm_step_forward(self.handle, layout.handle, phase, _ptr(obs1), _ptr(obs2),
                                                 _ptr(h), _ptr(c), _ptr(h), _ptr(c), _ptr(normal),
                                                 _ptr(pos), _ptr(ws), need, _stream(self.device)))
        return normal, pos

    def forward_steps(self, layout, observed, truth, n_decode, first_step, last_step, normals, positions, h, c):
        """Steps [first_step, last_step) of the time loop on caller-owned state (tb2_lstm_forward_steps)."""
        lib = _lib.load()
        ws, need = self.workspace(layout)
        with torch.cuda.device(self.device):
            _lib.check(lib.tb2_lstm_forward_steps(
                self.handle, layout.handle, _ptr(observed), int(observed.shape[0]), _ptr(truth),
                int(n_decode), int(first_step), int(last_step), _ptr(normals), _ptr(positions), _ptr(h), _ptr(c),
                _ptr(None), _ptr(ws), need, _stream(self.device)))

    def forward_sequence_host(self, layout, observed, truth, n_decode, normals, positions, h, c, normals_host,
                              positions_host, copy_stream):
        """tb2_lstm_forward_sequence_host: per-step device-to-host copies on `copy_stream`; the caller
        synchronises that stream before reading the pinned host tensors."""
        lib = _lib.load()
        ws, need = self.workspace(layout)
        with torch.cuda.device(self.device):
            _lib.check(lib.tb2_lstm_forward_sequence_host(
                self.handle, layout.handle, _ptr(observed), int(observed.shape[0]), _ptr(truth), int(n_decode),
                _ptr(normals), _ptr(positions), _ptr(h), _ptr(c), _ptr(ws), need, _ptr(normals_host),
                _ptr(positions_host), _stream(self.device), ctypes.c_void_p(copy_stream.cuda_stream)))

    def train_cache_bytes(self, layout, num_steps):
        return int(_lib.load().tb2_lstm_train_cache_bytes(self.handle, layout.handle, int(num_steps)))

    def forward_sequence_train(self, layout, observed, truth, n_decode, normals, positions, h, c, states, cache):
        """tb2_lstm_forward_sequence_train: per-step forward quantities of the social pooling stay in `cache`
        (uint8 device tensor of train_cache_bytes) for tb2_lstm_sequence_backward_cached."""
        lib = _lib.load()
        ws, need = self.workspace(layout)
        with torch.cuda.device(self.device):
            _lib.check(lib.tb2_lstm_forward_sequence_train(
                self.handle, layout.handle, _ptr(observed), int(observed.shape[0]), _ptr(truth), int(n_decode),
                _ptr(normals), _ptr(positions), _ptr(h), _ptr(c), _ptr(states), _ptr(cache), int(cache.numel()),
                _ptr(ws), need, _stream(self.device)))

    def forward_sequence(self, layout, observed, truth, n_decode, normals, positions, h, c, states=None):
        lib = _lib.load()
        ws, need = self.workspace(layout)
        with torch.cuda.device(self.device):
            _lib.check(lib.tb2_lstm_forward_sequence(
                self.handle, layout.handle, _ptr(observed), int(observed.shape[0]), _ptr(truth),
                int(n_decode), _ptr(normals), _ptr(positions), _ptr(h), _ptr(c), _ptr(states),
                _ptr(ws), need, _stream(self.device)))
