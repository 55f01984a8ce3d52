"""ode_b200 -- B200-native (sm_100a) implementation of ODE's per-step world update
(dSpaceCollide -> contact joints -> dWorldQuickStep -> dJointGroupEmpty) for batches of worlds.

The product is the CUDA shared library pair ode_b200/libode_b200_{single,double}.so behind the C-ABI of
include/ode_b200.h; this package is the thin ctypes mirror used by tests and bench.py.
There is no CPU fallback: loading fails loudly when the extension has not been built.
"""
import os
import numpy as np
from . import _binding
from ._binding import Scene, Batch, SceneLib, default_world_params  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path(precision="single"):
    return os.path.join(_HERE, "libode_b200_%s.so" % precision)


def load(precision="single"):
    """Open the CUDA library for one precision. Raises if it has not been built (no fallback)."""
    p = lib_path(precision)
    if not os.path.exists(p):
        raise RuntimeError("ode_b200: CUDA extension %s not built; run __graft_entry__.build()" % p)
    return SceneLib(p, "odeb_", np.float32 if precision == "single" else np.float64)
