"""Import the UNMODIFIED reference bilateral code from /root/reference (this container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Used by ``oracle/make_golden.py`` to mint
the committed fixtures in ``tests/golden/`` and by the CPU tests that are skipped when
``/root/reference`` is absent (it does not exist on the GPU box).

The reference imports three packages that are not installed here; none of them is touched
by the bilateral path, so they are replaced by empty stubs:

* ``tensorly``      - ``lib_bilagrid.py:48,53`` (``tl.set_backend`` at import time only)
* ``pytorch3d``     - ``models/modules.py:9``  (``knn_points``; VoxelDeformer only)
* ``nvdiffrast``    - ``models/modules.py:10`` (``dr.texture``; EnvLight only)
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("BDS_REFERENCE_ROOT", "/root/reference")
_PROJECT = os.path.join(REFERENCE_ROOT, "project")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_PROJECT, "bilateral", "lib_bilagrid.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _not_available(*_a, **_k):  # pragma: no cover
    raise RuntimeError("stubbed third-party function called; not on the bilateral path")


def load_reference():
    """Returns (lib_bilagrid module, models.modules module) of the reference."""
    if not reference_available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")
    _stub("tensorly", set_backend=lambda *_a, **_k: None)
    _stub("tensorly.decomposition", parafac=_not_available)
    p3d = _stub("pytorch3d")
    p3d.ops = _stub("pytorch3d.ops", knn_points=_not_available)
    p3d.transforms = _stub("pytorch3d.transforms", matrix_to_quaternion=_not_available)
    nvd = _stub("nvdiffrast")
    nvd.torch = _stub("nvdiffrast.torch", texture=_not_available)
    if _PROJECT not in sys.path:
        sys.path.insert(0, _PROJECT)
    import importlib

    lib = importlib.import_module("bilateral.lib_bilagrid")
    mods = importlib.import_module("models.modules")
    return lib, mods


def reference_apply_chain(rgb, affine_list):
    """Restates scene_graph.py:112-117 verbatim in meaning: sequential 3x4 apply."""
    x = rgb
    for aff in affine_list:
        aff = aff.reshape(rgb.shape[0], rgb.shape[1], 3, 4)
        x = (aff[..., :3, :3] @ x[..., None] + aff[..., :3, 3:])[..., 0]
    return x
