"""The C-ABI library loads and exports every symbol include/trajnet_b200.h declares (CPU only;
no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "trajnet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tb2_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    from trajnetplusplusbaselines_b200 import build, _lib
    build.build()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), "missing export: " + name
        assert name in _lib.PROTOTYPES, "ctypes prototype missing for " + name
    assert lib.tb2_version() >= 100


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from trajnetplusplusbaselines_b200.lstm import LSTM
    import numpy as np
    model = LSTM()
    obs = torch.zeros(9, 4, 2)
    with torch.no_grad(), pytest.raises(RuntimeError):
        model(obs, torch.zeros(4, 2), torch.tensor([0, 4]), n_predict=12)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "trajnetplusplusbaselines_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_state_dict_keys_match_reference_layout():
    from trajnetplusplusbaselines_b200.lstm import LSTM, GridBasedPooling
    pool = GridBasedPooling(type_='social', hidden_dim=128, cell_side=0.6, n=16, out_dim=256,
                            embedding_arch='two_layer', layer_dims=[1024], latent_dim=16)
    sd = LSTM(pool=pool).state_dict()
    expect = {  # SURVEY.md 8b/B2 (probe of the reference's own state_dict)
        'pool.hidden_dim_encoding.weight': (16, 128), 'pool.embedding.0.weight': (1024, 4096),
        'pool.embedding.2.weight': (256, 1024), 'input_embedding.input_embeddings.0.weight': (62, 2),
        'goal_embedding.input_embeddings.0.weight': (62, 2), 'encoder.weight_ih': (512, 320),
        'encoder.weight_hh': (512, 128), 'decoder.weight_ih': (512, 320),
        'hidden2normal.linear.weight': (5, 128),
    }
    for k, shape in expect.items():
        assert tuple(sd[k].shape) == shape, k
