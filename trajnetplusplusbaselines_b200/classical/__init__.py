"""Drop-in surface of trajnetbaselines.classical (reference: classical/__init__.py).

Like the reference, each submodule exposes `predict(...)`; the package-level name `predict` ends
up bound to the last import (constant_velocity), callers use the submodules.
"""
from . import socialforce, orca, kalman, constant_velocity
from .socialforce import predict
from .orca import predict
from .kalman import predict
from .constant_velocity import predict
