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


def test_grid_slots_of_cameras_inside_the_band_are_required():
    """Mode 2 runs the bilateral chain for every camera that owns a tile row of the band: a None slot there would read
    an uninitialised workspace (round-1 advisor finding) and is refused; None stays legal outside the band."""
    import pytest
    import torch

    sys.path.insert(0, ROOT)
    from bilateral_driving_b200.render import RenderCfg, check_grid_slots

    sizes = ((2, 2, 1), (4, 4, 2), (8, 8, 4))
    slot = [torch.zeros(12, 1, 2, 2), torch.zeros(12, 2, 4, 4), torch.zeros(12, 4, 8, 8)]
    full = RenderCfg(width=64, height=48, mode=2, bil_sizes=sizes)                    # 3 tile rows per camera
    check_grid_slots(full, 2, slot + slot)
    with pytest.raises(ValueError, match="camera 1"):
        check_grid_slots(full, 2, slot + [None] * 3)
    with pytest.raises(ValueError, match="level"):
        check_grid_slots(full, 2, slot + [slot[0], None, slot[2]])
    band = RenderCfg(width=64, height=48, mode=2, bil_sizes=sizes, row_begin=0, row_end=3)   # camera 0 only
    check_grid_slots(band, 2, slot + [None] * 3)
    with pytest.raises(ValueError, match="camera 0"):
        check_grid_slots(band, 2, [None] * 3 + slot)
    check_grid_slots(RenderCfg(width=64, height=48, mode=1), 2, [])                   # glue only: nothing to check
    with pytest.raises(ValueError, match="expected"):
        check_grid_slots(full, 2, slot)
