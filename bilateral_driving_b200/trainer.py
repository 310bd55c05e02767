"""Drop-in trainer for the reference's ``tools/train.py`` (selected with the CLI dotlist
``trainer.type=bilateral_driving_b200.trainer.FusedMultiTrainer``).

``FusedMultiTrainer`` subclasses the reference's own ``MultiTrainer``
(``models/trainers/scene_graph.py``) and replaces only what sits on the hot path:

* ``affine_transformation`` (``scene_graph.py:86-120``): for ``MultiScaleBilateralAffineTransform`` /
  ``BilateralAffineTransform`` it calls the module's fused ``transform`` (slice + sequential apply in one
  op, no 100 MB-per-level affine fields) instead of ``forward`` + the Python apply loop;
* ``render_gaussians`` (``base.py:385-432``) hands back a ``render_fn`` whose masked re-renders
  (``scene_graph.py:296-313``) reuse the sorted tile lists of the first render (``render.rasterize_masked``);
* ``forward`` additionally emits ``outputs["original_rgb"]``, which ``compute_losses`` reads at
  ``base.py:628-631`` but ``MultiTrainer.forward`` never sets (a bug of the published ms-bilateral configs:
  the first training step raises ``KeyError``).

The render itself is replaced through the ``gsplat`` import seam (``shim/gsplat``), not here, so everything else
of the reference trainer (losses, optimiser, densification, checkpoints) runs unchanged.

The reference tree must be importable (``PYTHONPATH=<reference>/project``); this module imports it lazily so
that the rest of the package does not depend on it.
"""
from typing import Dict

import torch


def _reference_multi_trainer():
    try:
        from models.trainers.scene_graph import MultiTrainer  # the reference's own class
    except Exception as exc:  # pragma: no cover - needs the reference tree and its dependencies
        raise ImportError(
            "FusedMultiTrainer needs the reference tree on PYTHONPATH (export PYTHONPATH=<reference>/project, as "
            "scripts/train.sh:21 does)") from exc
    return MultiTrainer


_FUSED_AFFINE_TYPES = (
    "bilateral_driving_b200.bilateral.MultiScaleBilateralAffineTransform",
    "bilateral_driving_b200.bilateral.BilateralAffineTransform",
)


def _build():
    MultiTrainer = _reference_multi_trainer()

    class FusedMultiTrainer(MultiTrainer):
        """See module docstring."""

        guidance_factor = [4, 4, 2]  # the reference's default (modules.py:505); set None for full-res guidance

        def affine_transformation(self, rgb_blended: torch.Tensor, image_infos: Dict[str, torch.Tensor]):
            if "Affine" in self.models and self.model_config.Affine.type in _FUSED_AFFINE_TYPES:
                self._original_rgb = rgb_blended
                affine = self.models["Affine"]
                if hasattr(affine, "grid_size"):  # multi-scale module
                    return affine.transform(rgb_blended, image_infos, guidance_factor=self.guidance_factor)
                return affine.transform(rgb_blended, image_infos)
            return super().affine_transformation(rgb_blended, image_infos)

        def render_gaussians(self, gs, cam, **kwargs):
            """``base.py:385-432``: same results, but the ``render_fn(opacity_mask)`` closure handed back to
            ``MultiTrainer.forward`` (per-class / dynamic-only renders, ``scene_graph.py:296-313``) composites
            again over the sorted tile lists of the first render instead of re-running the whole rasterization."""
            results, render_fn = super().render_gaussians(gs, cam, **kwargs)
            info = self.info

            def cached_render_fn(opaticy_mask=None, return_info=False):
                reusable = (opaticy_mask is not None and not return_info and not torch.is_grad_enabled()
                            and isinstance(info, dict) and "_bds_cache" in info
                            and opaticy_mask.dtype in (torch.bool, torch.uint8))
                if not reusable:
                    return render_fn(opaticy_mask, return_info)
                from .render import rasterize_masked
                renders, alphas = rasterize_masked(info, opaticy_mask)
                renders, alphas = renders[0], alphas[0].squeeze(-1)
                rendered_rgb, rendered_depth = torch.split(renders, [3, 1], dim=-1)
                return torch.clamp(rendered_rgb, max=1.0), rendered_depth, alphas[..., None]

            return results, cached_render_fn

        def forward(self, image_infos, camera_infos, novel_view: bool = False):
            self._original_rgb = None
            outputs = super().forward(image_infos, camera_infos, novel_view)
            if "original_rgb" not in outputs:
                pre = self._original_rgb
                if pre is None:  # affine not fused (other Affine types): rebuild as scene_graph.py:293 does
                    pre = outputs["rgb_gaussians"] + outputs["rgb_sky"] * (1.0 - outputs["opacity"])
                outputs["original_rgb"] = pre
            return outputs

    return FusedMultiTrainer


def __getattr__(name):  # PEP 562: resolve the class on first use (import_str does getattr on the module)
    if name == "FusedMultiTrainer":
        cls = _build()
        globals()[name] = cls
        return cls
    raise AttributeError(name)
