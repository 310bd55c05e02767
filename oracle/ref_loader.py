"""Import the UNMODIFIED reference code of the hot path.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Source tree, first that exists:
* ``/root/reference/project``        (the build container), or
* ``oracle/_ref/project``            (byte-for-byte copies made by ``oracle/build_ref.py``; git-ignored, travels to
                                      the GPU box with the snapshot - ``/root/reference`` does not exist there).

Used by ``oracle/make_golden.py`` to mint the committed fixtures in ``tests/golden/``, by the tests that compare
with the live reference (skipped when neither tree exists), and by ``bench.py``'s CPU legs (``cpu_baseline`` /
``--impl reference``), which time the reference's own bilateral module on the host cores.

Packages the reference imports that are not installed here are replaced by the stand-ins of ``oracle/ref_stubs.py``.
"""
import importlib
import os
import sys

from . import ref_stubs

REFERENCE_ROOT = os.environ.get("BDS_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = (os.path.join(REFERENCE_ROOT, "project"), os.path.join(_HERE, "_ref", "project"))


def project_root():
    for p in _CANDIDATES:
        if os.path.isfile(os.path.join(p, "bilateral", "lib_bilagrid.py")):
            return p
    return None


def reference_available() -> bool:
    return project_root() is not None


def reference_kind() -> str:
    """"reference" = /root/reference itself, "_ref" = the copies under oracle/_ref."""
    p = project_root()
    if p is None:
        return "absent"
    return "reference" if p == _CANDIDATES[0] else "_ref"


def _prepare():
    root = project_root()
    if root is None:
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT} nor oracle/_ref (run python -m oracle.build_ref)")
    ref_stubs.install()
    if root not in sys.path:
        sys.path.insert(0, root)
    return root


def load_reference():
    """Returns (lib_bilagrid module, models.modules module) of the reference."""
    _prepare()
    lib = importlib.import_module("bilateral.lib_bilagrid")
    mods = importlib.import_module("models.modules")
    return lib, mods


def load_reference_trainers(gsplat_pkg_dir=None):
    """Returns the reference's ``models.trainers.scene_graph`` module (``MultiTrainer``; ``BasicTrainer`` is its
    base).  ``models/gaussians/basics.py:12-15`` imports ``gsplat``: ``gsplat_pkg_dir`` is put first on ``sys.path``
    for it (``<repo>/shim`` = the sm_100a renderer, the import seam; ``<repo>/oracle/gsplat_cpu`` = the CPU oracle)."""
    _prepare()
    if gsplat_pkg_dir is not None and gsplat_pkg_dir not in sys.path:
        sys.path.insert(0, gsplat_pkg_dir)
    return importlib.import_module("models.trainers.scene_graph")


def load_reference_eval(gsplat_pkg_dir=None):
    """Returns the reference's ``models.video_utils`` module (``render_images`` / ``render``, the eval harness of
    ``tools/eval.py`` and ``tools/train.py``)."""
    load_reference_trainers(gsplat_pkg_dir)
    return importlib.import_module("models.video_utils")


def reference_apply_chain(rgb, affine_list):
    """Restates scene_graph.py:112-117 verbatim in meaning: sequential 3x4 apply."""
    x = rgb
    for aff in affine_list:
        aff = aff.reshape(rgb.shape[0], rgb.shape[1], 3, 4)
        x = (aff[..., :3, :3] @ x[..., None] + aff[..., :3, 3:])[..., 0]
    return x
