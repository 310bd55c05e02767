"""Recipe for ``oracle/_ref/``: the reference's OWN Python files for the hot path, laid out so that they can be
imported where ``/root/reference`` does not exist (the GPU box).  TEST INFRASTRUCTURE ONLY.

    python -m oracle.build_ref            # in the build container; also run by __graft_entry__.build()

``oracle/_ref/`` is git-ignored (reference sources never enter the history) but NOT gpurun-ignored, so it travels
to the GPU box with the snapshot, like the built ``.so``.  Nothing is edited: files are copied byte for byte from
where they lie under ``/root/reference/project`` and their SHA-256 is recorded in ``oracle/_ref/MANIFEST.json``.
The third-party packages those files import but never use on this path (omegaconf, kornia, pytorch3d, ...) are
NOT copied from anywhere: ``oracle/ref_stubs.py`` (our code) stands in for them.

What is copied and why:
* ``bilateral/lib_bilagrid.py``, ``models/modules.py``       the reference's bilateral half of the path (A6-A9)
* ``models/trainers/*.py``                                   ``BasicTrainer.compute_losses`` / ``MultiTrainer.forward``
                                                              - the callers the drop-in trainer subclasses
* ``models/gaussians/*.py``, ``models/losses.py``            ``VanillaGaussians`` (activations + SH call site, A1/A2,
                                                              densification bookkeeping N3), the loss functions
* ``utils/misc.py``, ``utils/geometry.py``                   ``import_str`` (the plug-in loader), rotation helpers
"""
import hashlib
import json
import os
import shutil
import subprocess

REFERENCE_ROOT = os.environ.get("BDS_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

FILES = [
    "bilateral/__init__.py",
    "bilateral/lib_bilagrid.py",
    "models/__init__.py",
    "models/modules.py",
    "models/losses.py",
    "models/trainers/__init__.py",
    "models/trainers/base.py",
    "models/trainers/scene_graph.py",
    "models/trainers/single.py",
    "models/gaussians/__init__.py",
    "models/gaussians/basics.py",
    "models/gaussians/vanilla.py",
    "models/gaussians/deformgs.py",
    "models/gaussians/pvg.py",
    "models/gaussians/scaffold.py",
    "utils/__init__.py",
    "utils/misc.py",
    "utils/geometry.py",
    "models/video_utils.py",      # the eval harness (render_images / render), tests/test_trainer_reference.py
    "utils/visualization.py",
]


def reference_present() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "project", "bilateral", "lib_bilagrid.py"))


def built() -> bool:
    return os.path.isfile(os.path.join(OUT, "MANIFEST.json"))


def build(verbose: bool = False) -> bool:
    """Copies the files; returns False (and leaves an existing ``_ref`` alone) when the reference is absent."""
    if not reference_present():
        return False
    src_root = os.path.join(REFERENCE_ROOT, "project")
    dst_root = os.path.join(OUT, "project")
    if os.path.isdir(dst_root):
        shutil.rmtree(dst_root)
    manifest = {"reference_root": REFERENCE_ROOT, "files": {}}
    try:
        manifest["reference_commit"] = subprocess.run(["git", "-C", REFERENCE_ROOT, "rev-parse", "HEAD"],
                                                      capture_output=True, text=True, timeout=10).stdout.strip() or None
    except Exception:
        manifest["reference_commit"] = None
    for rel in FILES:
        src = os.path.join(src_root, rel)
        dst = os.path.join(dst_root, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.isfile(src):
            if rel.endswith("__init__.py"):   # namespace-style directory in the reference: keep it importable
                open(dst, "w").close()
                manifest["files"][rel] = None
                continue
            raise FileNotFoundError(src)
        shutil.copyfile(src, dst)
        manifest["files"][rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
        if verbose:
            print("copied", rel)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    return True


if __name__ == "__main__":
    ok = build(verbose=True)
    print("oracle/_ref built" if ok else f"reference not found under {REFERENCE_ROOT}: nothing done")
