"""CPU: the gsplat import seam.  models/gaussians/basics.py:12-15 of the reference imports exactly these four
names from gsplat; with <repo>/shim first on sys.path they resolve to the sm_100a host mirror."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROBE = r"""
import sys
sys.path[:0] = [{shim!r}, {root!r}]
from gsplat.rendering import rasterization
from gsplat.cuda_legacy._wrapper import num_sh_bases
from gsplat.cuda_legacy._torch_impl import quat_to_rotmat
from gsplat.cuda._wrapper import spherical_harmonics
import bilateral_driving_b200.render as R
assert rasterization is R.rasterization and spherical_harmonics is R.spherical_harmonics
assert num_sh_bases(3) == 16 and num_sh_bases(0) == 1
import torch
q = torch.tensor([[2.0, 0.0, 0.0, 0.0], [0.5, 0.5, 0.5, 0.5]])
m = quat_to_rotmat(q)
assert torch.allclose(m[0], torch.eye(3)) and torch.allclose(m[1] @ m[1].T, torch.eye(3), atol=1e-6)
print("ok")
"""


def test_gsplat_shim_exposes_the_reference_import_paths():
    code = PROBE.format(shim=os.path.join(ROOT, "shim"), root=ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
