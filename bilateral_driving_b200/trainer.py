"""Drop-in trainer for the reference's ``tools/train.py`` (selected with the CLI dotlist
``trainer.type=bilateral_driving_b200.trainer.FusedMultiTrainer``).

``FusedMultiTrainer`` subclasses the reference's own ``MultiTrainer``
(``models/trainers/scene_graph.py``) and replaces only what sits on the hot path:

* ``forward`` (``scene_graph.py:231-315``): render -> clamp -> sky composite -> bilateral affine runs as ONE call of
  ``render.render_fused`` - the fused composite kernel with the reference glue in its epilogue (mode 1) followed by
  the low-resolution-guidance bilateral kernels for the reference's default ``guidance_factor=[4,4,2]``, or the
  whole chain inside the composite kernel (mode 2) for ``guidance_factor=None`` / the single-grid
  ``BilateralAffineTransform``.  Same output dictionary as the reference, plus ``outputs["original_rgb"]``, which
  ``compute_losses`` reads at ``base.py:628-631`` but ``MultiTrainer.forward`` never sets (a bug of the published
  ms-bilateral configs: the first training step raises ``KeyError``).
* ``render_fn(opacity_mask)`` (per-class / dynamic-only renders, ``scene_graph.py:296-313``) composites again over
  the sorted tile lists of the first render (``render.rasterize_masked``).
* ``compute_losses`` (``base.py:518-659``): the reference dispatches the affine regularisers on the LITERAL strings
  ``'models.modules.*'`` (``base.py:587-635``); with ``model.Affine.type`` pointing at this package no branch
  would match and the TV + cycle losses would silently vanish.  The override presents the reference's own type
  string while the unchanged ``BasicTrainer.compute_losses`` runs, so ``loss_dict["affine_loss"] =
  affine.w * tv_loss() + affine.w1 * inverse_loss(gt, original_rgb)`` is there exactly as in the reference.
  With ``w1 == 0`` (``configs/omnire_ms_bilateral.yaml:34``) the cycle term - a per-pixel 4x4 ``torch.inverse`` in
  the reference, multiplied by zero - is not evaluated.

Everything else of the reference trainer (other losses, optimiser, densification, checkpoints, viewer) runs
unchanged; any ``Affine`` type this package does not implement falls back to the reference's own ``forward``.

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


# this package's Affine types -> the literal the reference's compute_losses dispatches on (base.py:587-635)
_REFERENCE_AFFINE_TYPE = {
    "bilateral_driving_b200.bilateral.MultiScaleBilateralAffineTransform": "models.modules.MultiScaleBilateralAffineTransform",
    "bilateral_driving_b200.bilateral.BilateralAffineTransform": "models.modules.BilateralAffineTransform",
}
_FUSED_AFFINE_TYPES = tuple(_REFERENCE_AFFINE_TYPE)


def _build():
    MultiTrainer = _reference_multi_trainer()

    class FusedMultiTrainer(MultiTrainer):
        """See module docstring."""

        guidance_factor = [4, 4, 2]  # the reference's default (modules.py:505); set None for full-res guidance
        fused_render = True          # False: keep the reference's forward, fuse only the affine (round-1 behaviour)

        # ---- which path -------------------------------------------------------------------------------------
        def _affine_type(self):
            if "Affine" not in self.models:
                return None
            return self.model_config.Affine.type

        def _fusable(self) -> bool:
            t = self._affine_type()
            return bool(self.fused_render and (t is None or t in _FUSED_AFFINE_TYPES) and "Sky" in self.models
                        and self.render_cfg.batch_size == 1)

        # ---- scene_graph.py:86-120 (only reached on the non-fused path) ---------------------------------------
        def affine_transformation(self, rgb_blended: torch.Tensor, image_infos: Dict[str, torch.Tensor]):
            if self._affine_type() in _FUSED_AFFINE_TYPES:
                self._original_rgb = rgb_blended
                affine = self.models["Affine"]
                if hasattr(affine, "grid_size"):  # multi-scale module
                    return affine.transform(rgb_blended, image_infos, guidance_factor=self.guidance_factor)
                return affine.transform(rgb_blended, image_infos)
            return super().affine_transformation(rgb_blended, image_infos)

        # ---- base.py:385-432 (non-fused path and the viewer) --------------------------------------------------
        def render_gaussians(self, gs, cam, **kwargs):
            """Same results as the reference; the ``render_fn(opacity_mask)`` closure handed back composites again
            over the sorted tile lists of the first render instead of re-running the whole rasterization."""
            results, render_fn = super().render_gaussians(gs, cam, **kwargs)
            return results, self._cached_render_fn(self.info, render_fn)

        def _cached_render_fn(self, info, full_render_fn):
            def cached_render_fn(opaticy_mask=None, return_info=False):
                reusable = (opaticy_mask is not None and not return_info and not torch.is_grad_enabled()
                            and isinstance(info, dict) and info.get("_bds_cache") is not None
                            and opaticy_mask.dtype in (torch.bool, torch.uint8))
                if not reusable:
                    return full_render_fn(opaticy_mask, return_info)
                from . import render as R
                renders, alphas = R.rasterize_masked(info, opaticy_mask)
                renders, alphas = renders[0], alphas[0].squeeze(-1)
                rendered_rgb, rendered_depth = torch.split(renders, [3, 1], dim=-1)
                return torch.clamp(rendered_rgb, max=1.0), rendered_depth, alphas[..., None]

            return cached_render_fn

        # ---- the fused hot path ---------------------------------------------------------------------------------
        def render_fused_path(self, gs, cam, rgb_sky, image_infos, **kwargs):
            """render_gaussians (base.py:385-432) + clamp (base.py:417) + sky composite (scene_graph.py:287-294) +
            affine_transformation (scene_graph.py:86-120) in one ``render_fused`` call.  Returns (results,
            render_fn) like ``render_gaussians``; ``results`` additionally holds the post-affine ``rgb``."""
            from . import render as R
            W, H = R._as_int(cam.W), R._as_int(cam.H)
            affine = self.models.get("Affine") if self._affine_type() in _FUSED_AFFINE_TYPES else None
            slots = sizes = gf = None
            if affine is not None:
                slots = affine._slots(image_infos)
                sizes = affine.level_sizes()
                gf = self.guidance_factor if hasattr(affine, "grid_size") else None
            params = dict(_means=gs.means, _quats=gs.quats, _scales=gs.scales, _opacities=gs.opacities.squeeze(-1),
                          _rgbs=gs.rgbs)
            viewmats = torch.linalg.inv(cam.camtoworlds)[None, ...]     # base.py:399 (CamPose gradient flows here)
            out = R.render_fused(
                params, viewmats, cam.Ks[None, ...], W, H, sky=rgb_sky,
                grid_slots=None if slots is None else [slots], bil_sizes=() if sizes is None else sizes,
                near_plane=kwargs.get("near_plane", 0.01), far_plane=kwargs.get("far_plane", 1e10),
                radius_clip=kwargs.get("radius_clip", 0.0), absgrad=bool(self.render_cfg.absgrad), activated=True,
                dense_info=True, antialiased=bool(self.render_cfg.antialiased), guidance_factor=gf)
            holder = out["info"]
            self.info = {"means2d": out["means2d"], "radii": out["radii"], "width": W, "height": H, "n_cameras": 1,
                         "tile_size": 16, "n_isect": holder.get("n_isect"), "n_visible": holder.get("n_visible"),
                         "depths": holder.get("depths"), "conics": holder.get("conics"),
                         "_bds_cache": holder.get("cache")}
            if self.training:
                self.info["means2d"].retain_grad()                      # base.py:430
            results = {"rgb_gaussians": out["rgb_gaussians"], "depth": out["depth"], "opacity": out["opacity"],
                       "rgb": out["rgb"]}

            def full_render_fn(opaticy_mask=None, return_info=False):   # base.py:392-419, for soft masks / with grad
                from models.gaussians.basics import rasterization
                renders, alphas, info = rasterization(
                    means=gs.means, quats=gs.quats, scales=gs.scales,
                    opacities=gs.opacities.squeeze() * opaticy_mask if opaticy_mask is not None else gs.opacities.squeeze(),
                    colors=gs.rgbs, viewmats=viewmats, Ks=cam.Ks[None, ...], width=cam.W, height=cam.H,
                    packed=self.render_cfg.packed, absgrad=self.render_cfg.absgrad,
                    sparse_grad=self.render_cfg.sparse_grad,
                    rasterize_mode="antialiased" if self.render_cfg.antialiased else "classic",
                    render_mode="RGB+ED", **kwargs)
                renders, alphas = renders[0], alphas[0].squeeze(-1)
                rendered_rgb, rendered_depth = torch.split(renders, [3, 1], dim=-1)
                res = (torch.clamp(rendered_rgb, max=1.0), rendered_depth, alphas[..., None])
                return res + (info,) if return_info else res

            if affine is not None and hasattr(affine, "remember_for_inverse_loss"):
                self._cycle_args = (affine, slots, gf)
            return results, self._cached_render_fn(self.info, full_render_fn)

        def forward(self, image_infos, camera_infos, novel_view: bool = False):
            self._original_rgb = None
            self._cycle_args = None
            if not self._fusable():
                outputs = super().forward(image_infos, camera_infos, novel_view)
                if "original_rgb" not in outputs:
                    pre = self._original_rgb
                    if pre is None:  # affine not fused (other Affine types): rebuild as scene_graph.py:293 does
                        pre = outputs["rgb_gaussians"] + outputs["rgb_sky"] * (1.0 - outputs["opacity"])
                    outputs["original_rgb"] = pre
                return outputs

            # scene_graph.py:247-273, unchanged in meaning
            normed_time = image_infos["normed_time"].flatten()[0]
            self.cur_frame = torch.argmin(torch.abs(self.normalized_timestamps - normed_time))
            for model in self.models.values():
                if hasattr(model, "in_test_set"):
                    model.in_test_set = self.in_test_set
            for class_name in self.gaussian_classes.keys():
                model = self.models[class_name]
                if hasattr(model, "set_cur_frame"):
                    model.set_cur_frame(self.cur_frame)
            image_ids = image_infos["img_idx"].flatten()[0]
            processed_cam = self.process_camera(camera_infos=camera_infos, image_ids=image_ids, novel_view=novel_view)
            gs = self.collect_gaussians(cam=processed_cam, image_ids=image_ids)

            # the sky colour only depends on image_infos (scene_graph.py:287-289): evaluated first, composited in
            # the kernel's epilogue
            rgb_sky = self.models["Sky"](image_infos)
            outputs, render_fn = self.render_fused_path(
                gs, processed_cam, rgb_sky, image_infos, near_plane=self.render_cfg.near_plane,
                far_plane=self.render_cfg.far_plane, radius_clip=self.render_cfg.get("radius_clip", 0.))
            outputs["rgb_sky"] = rgb_sky
            outputs["rgb_sky_blend"] = rgb_sky * (1.0 - outputs["opacity"])
            # pre-affine image (scene_graph.py:293); read by compute_losses at base.py:630
            outputs["original_rgb"] = outputs["rgb_gaussians"] + outputs["rgb_sky_blend"]
            if self._cycle_args is not None:
                affine, slots, gf = self._cycle_args
                affine.remember_for_inverse_loss(outputs["original_rgb"], slots, gf)

            # scene_graph.py:296-313
            if not self.training and self.render_each_class:
                with torch.no_grad():
                    for class_name in self.gaussian_classes.keys():
                        gaussian_mask = self.pts_labels == self.gaussian_classes[class_name]
                        sep_rgb, sep_depth, sep_opacity = render_fn(gaussian_mask)
                        outputs[class_name + "_rgb"] = sep_rgb
                        outputs[class_name + "_opacity"] = sep_opacity
                        outputs[class_name + "_depth"] = sep_depth
            if not self.training or self.render_dynamic_mask:
                with torch.no_grad():
                    gaussian_mask = self.pts_labels != self.gaussian_classes["Background"]
                    sep_rgb, sep_depth, sep_opacity = render_fn(gaussian_mask)
                    outputs["Dynamic_rgb"] = sep_rgb
                    outputs["Dynamic_opacity"] = sep_opacity
                    outputs["Dynamic_depth"] = sep_depth
            return outputs

        # ---- base.py:518-659 ------------------------------------------------------------------------------------
        def compute_losses(self, outputs, image_infos, cam_infos):
            t = self._affine_type()
            ref_t = _REFERENCE_AFFINE_TYPE.get(t)
            if ref_t is None:
                return super().compute_losses(outputs, image_infos, cam_infos)
            affine = self.models["Affine"]
            affine_reg = self.losses_dict.get("affine", None)
            skip_cycle = (affine_reg is not None and hasattr(affine, "inverse_loss")
                          and float(affine_reg.get("w1", 0.0)) == 0.0)
            if skip_cycle:
                # w1 * inverse_loss with w1 == 0: same loss value and same (zero) gradients without the per-pixel inverse
                zero = outputs["rgb"].new_zeros(())
                affine.inverse_loss = lambda gt, render: zero
            self.model_config.Affine.type = ref_t    # the literal base.py:587-635 compares with
            try:
                return super().compute_losses(outputs, image_infos, cam_infos)
            finally:
                self.model_config.Affine.type = t
                if skip_cycle:
                    del affine.inverse_loss          # back to the class's method

    return FusedMultiTrainer


def __getattr__(name):  # PEP 562: resolve the class on first use (import_str does getattr on the module)
    if name == "FusedMultiTrainer":
        cls = _build()
        globals()[name] = cls
        return cls
    raise AttributeError(name)
