"""Import shim for the UNMODIFIED reference at /root/reference (build container only).

TEST INFRASTRUCTURE -- never imported by the product package.  The reference's
third-party imports that are absent from this image are
replaced by empty stub modules so `trajnetbaselines.lstm` imports; the tensor-level
calls (LSTM.forward/step, GridBasedPooling.*, PredictionLoss) never touch the stubs.
/root/reference does not exist on the GPU box: only oracle/make_golden.py (run here,
outputs committed under tests/golden/) and tests marked `needs_reference` use this.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("TRAJNET_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "trajnetbaselines"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def import_reference():
    """Return the reference's `trajnetbaselines` package (stubs registered first)."""
    if not reference_available():
        raise ImportError("reference tree not present at %s" % REFERENCE_ROOT)
    _stub("trajnetplusplustools")
    _stub("trajnetplusplustools.show")
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    _stub("pykalman")
    _stub("socialforce")
    _stub("socialforce.potentials", PedPedPotential=object)
    _stub("socialforce.field_of_view", FieldOfView=object)
    _stub("rvo2")
    _stub("pysparkling")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import trajnetbaselines  # noqa: E402
    return trajnetbaselines
