"""Import-seam shim: put ``<repo>/shim`` (and ``<repo>``) ahead of site-packages on PYTHONPATH and the
reference's ``models/gaussians/basics.py:12-15`` imports

    from gsplat.rendering import rasterization
    from gsplat.cuda_legacy._wrapper import num_sh_bases
    from gsplat.cuda_legacy._torch_impl import quat_to_rotmat
    from gsplat.cuda._wrapper import spherical_harmonics

resolve to the sm_100a kernels of ``bilateral_driving_b200`` with zero edits to the reference."""
__version__ = "1.3.0+bds_b200"

from bilateral_driving_b200.render import rasterization, spherical_harmonics  # noqa: F401,E402
