"""Drop-in Gaussian model for the reference's YAML (``model.Background.type=bilateral_driving_b200.gaussians.
FusedVanillaGaussians``): the reference's own ``VanillaGaussians`` (``models/gaussians/vanilla.py``) with the two
per-step pieces that sit next to the hot path moved onto the sm_100a kernels:

* ``after_train`` (``vanilla.py:163-191``): the running densification statistics ``xys_grad_norm / vis_counts /
  max_2Dsize`` are updated by ONE launch (``bds_densify_stats``) from the taps the renderer emits
  (``info["radii"]``, ``info["means2d"].absgrad``) instead of a dozen boolean-mask gathers / scatters;
* everything else - parameters, activations, ``get_gaussians``, split / duplicate / cull, optimiser surgery,
  checkpoints - is inherited unchanged.

``densify_stats_update`` is the functional form.  No CPU fallback exists.
"""
import ctypes as C

import torch

from ._lib import BdsError, check, lib, ptr, stream_ptr


def densify_stats_update(radii, xys_grad, last_size, xys_grad_norm, vis_counts, max_2dsize, first: bool,
                         scale_xy=(1.0, 1.0)):
    """In-place update of the three running statistics (all ``[n]`` fp32 CUDA tensors) from ``radii [n]`` and
    ``xys_grad [n,2]`` of one step; semantics of ``VanillaGaussians.after_train`` with ``filter_mask`` all True."""
    for t in (radii, xys_grad, xys_grad_norm, vis_counts, max_2dsize):
        if not t.is_cuda:
            raise BdsError("bds operators run on CUDA tensors only (no CPU fallback exists)")
    n = radii.numel()
    if xys_grad.shape != (n, 2) or any(t.shape != (n,) for t in (xys_grad_norm, vis_counts, max_2dsize)):
        raise ValueError("densify_stats_update: shapes must be radii [n], xys_grad [n,2], statistics [n]")
    r = radii.reshape(-1).to(torch.int32).contiguous()
    g = xys_grad.contiguous().float()
    for t in (xys_grad_norm, vis_counts, max_2dsize):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("densify_stats_update: statistics must be contiguous fp32 (updated in place)")
    with torch.cuda.device(radii.device):
        check(lib.bds_densify_stats(C.c_int64(n), ptr(r), ptr(g), C.c_float(scale_xy[0]), C.c_float(scale_xy[1]),
                                    C.c_float(float(last_size)), C.c_int(int(bool(first))), ptr(xys_grad_norm),
                                    ptr(vis_counts), ptr(max_2dsize), stream_ptr()), "bds_densify_stats")


def _build():
    try:
        from models.gaussians.vanilla import VanillaGaussians  # the reference's own class
    except Exception as exc:  # pragma: no cover - needs the reference tree and its dependencies
        raise ImportError("FusedVanillaGaussians needs the reference tree on PYTHONPATH "
                          "(export PYTHONPATH=<reference>/project, as scripts/train.sh:21 does)") from exc

    class FusedVanillaGaussians(VanillaGaussians):
        """See module docstring."""

        def after_train(self, radii, xys_grad, last_size):
            # vanilla.py:163-191.  get_gaussians (vanilla.py:378-379) sets filter_mask to all-True, so the
            # reference's full_mask[filter_mask] = visible_mask is the visibility mask itself.
            if not radii.is_cuda or not bool(getattr(self, "filter_mask", torch.ones(1, dtype=torch.bool)).all()):
                return super().after_train(radii, xys_grad, last_size)
            with torch.no_grad():
                n = self.num_points
                first = self.xys_grad_norm is None
                if first:
                    self.xys_grad_norm = torch.empty(n, device=radii.device, dtype=torch.float32)
                    self.vis_counts = torch.empty(n, device=radii.device, dtype=torch.float32)
                if self.max_2Dsize is None:
                    self.max_2Dsize = torch.zeros(n, device=radii.device, dtype=torch.float32)
                densify_stats_update(radii.reshape(-1), xys_grad.reshape(n, 2), float(last_size), self.xys_grad_norm,
                                     self.vis_counts, self.max_2Dsize, first)

    return FusedVanillaGaussians


def __getattr__(name):  # PEP 562: resolve the class on first use (import_str does getattr on the module)
    if name == "FusedVanillaGaussians":
        cls = _build()
        globals()[name] = cls
        return cls
    raise AttributeError(name)
