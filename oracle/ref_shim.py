"""Import shim for the *real* reference (authoring container only).

TEST INFRASTRUCTURE.  `/root/reference` is pure Python and imports three
packages that are not installed here (`trimesh`, `matplotlib`, `tensorboardX`);
its env modules additionally import `gym` and `isaacgym` at module scope (no
arithmetic comes from them: the step-assembly functions pinned from
`envs/ig_parkour/mgdm_dm_util.py` and `envs/ig_char_env.py` are free TorchScript
functions over tensors), for which attribute-absorbing placeholders are injected.
This module injects minimal stand-ins and puts the reference on `sys.path`, so
that `oracle/make_golden.py` and the container-only tests can call the
reference's own functions.  `/root/reference` does not exist on the GPU box:
nothing that runs there may call `activate()`.

The only arithmetic the stubs provide is `trimesh.creation.icosphere
(subdivisions=0)`, i.e. the 12 vertices of a regular icosahedron scaled to the
requested radius (used by `util/geom_util.py:747`).  The vertex ORDER is
trimesh's (`trimesh/creation.py::icosahedron`): t = (1+sqrt5)/2,
  [-1, t, 0] [ 1, t, 0] [-1,-t, 0] [ 1,-t, 0]
  [ 0,-1, t] [ 0, 1, t] [ 0,-1,-t] [ 0, 1,-t]
  [ t, 0,-1] [ t, 0, 1] [-t, 0,-1] [-t, 0, 1]
each normalised to unit length, then multiplied by `radius`.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


def icosahedron_vertices(radius=1.0):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array(
        [[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0],
         [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
         [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v * radius


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "anim"))


def activate():
    """Make `import anim.motion_lib` etc. resolve to the reference."""
    if not available():
        raise RuntimeError("reference tree not present (expected only in the authoring container)")
    if "trimesh" not in sys.modules:
        tm = types.ModuleType("trimesh")
        cr = types.ModuleType("trimesh.creation")

        def icosphere(subdivisions=0, radius=1.0, **kw):
            assert subdivisions == 0, "stub only provides the base icosahedron"
            return types.SimpleNamespace(vertices=icosahedron_vertices(radius))

        cr.icosphere = icosphere
        tm.creation = cr
        sys.modules["trimesh"] = tm
        sys.modules["trimesh.creation"] = cr
    for name in ("matplotlib", "matplotlib.pyplot", "tensorboardX"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["tensorboardX"], "SummaryWriter"):
        sys.modules["tensorboardX"].SummaryWriter = object
    for name in ("gym", "gym.spaces", "isaacgym", "isaacgym.gymapi", "isaacgym.gymtorch", "isaacgym.gymutil",
                 "isaacgym.torch_utils"):
        if name not in sys.modules:
            sys.modules[name] = _placeholder_module(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


class _Absorb:
    """Accepts any construction / attribute / call; never computes anything."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Absorb()

    def __call__(self, *a, **k):
        return _Absorb()


def _placeholder_module(name):
    m = types.ModuleType(name)

    def _getattr(n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Absorb if n[0].isupper() else _Absorb()

    m.__getattr__ = _getattr
    m.__path__ = []
    return m
