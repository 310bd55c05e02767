"""CPU: host logic of the drop-in trainer (bilateral_driving_b200/trainer.py) against a stub of the reference's
MultiTrainer (models/trainers/scene_graph.py) - the real one needs omegaconf / kornia / pytorch3d / datasets."""
import sys
import types

import pytest
import torch


@pytest.fixture()
def fused_trainer_cls(monkeypatch):
    calls = {"render_fn": [], "masked": []}

    class MultiTrainer:  # the slice of the reference class the subclass touches
        training = False

        def __init__(self):
            self.models = {}
            self.model_config = types.SimpleNamespace(Affine=types.SimpleNamespace(type="x"))
            self.info = {"_bds_cache": object()}

        def render_gaussians(self, gs, cam, **kwargs):
            def render_fn(opaticy_mask=None, return_info=False):
                calls["render_fn"].append((opaticy_mask, return_info))
                return "rgb", "depth", "opacity"

            return {"rgb_gaussians": 1}, render_fn

        def forward(self, image_infos, camera_infos, novel_view=False):
            return {"rgb_gaussians": torch.ones(2, 2, 3), "rgb_sky": torch.full((2, 2, 3), 0.5),
                    "opacity": torch.full((2, 2, 1), 0.25)}

        def affine_transformation(self, rgb, image_infos):
            return rgb

    pkg = types.ModuleType("models")
    sub = types.ModuleType("models.trainers")
    mod = types.ModuleType("models.trainers.scene_graph")
    mod.MultiTrainer = MultiTrainer
    monkeypatch.setitem(sys.modules, "models", pkg)
    monkeypatch.setitem(sys.modules, "models.trainers", sub)
    monkeypatch.setitem(sys.modules, "models.trainers.scene_graph", mod)
    import bilateral_driving_b200.trainer as T

    monkeypatch.delitem(T.__dict__, "FusedMultiTrainer", raising=False)
    import bilateral_driving_b200.render as R

    def fake_masked(info, mask):
        calls["masked"].append(mask)
        H, W = 4, 6
        return torch.full((1, H, W, 4), 2.0), torch.full((1, H, W, 1), 0.5)

    monkeypatch.setattr(R, "rasterize_masked", fake_masked)
    cls = T.FusedMultiTrainer
    yield cls, calls
    T.__dict__.pop("FusedMultiTrainer", None)


def test_masked_render_fn_reuses_lists_only_when_it_is_exact(fused_trainer_cls):
    cls, calls = fused_trainer_cls
    tr = cls()
    results, render_fn = tr.render_gaussians(None, None)
    assert results == {"rgb_gaussians": 1}
    mask = torch.tensor([True, False, True])
    with torch.no_grad():
        rgb, depth, opacity = render_fn(mask)           # bool mask, no grad, no info wanted -> cached lists
    assert len(calls["masked"]) == 1 and not calls["render_fn"]
    assert rgb.shape == (4, 6, 3) and float(rgb.max()) == 1.0     # clamp(max=1) as base.py:417
    assert depth.shape == (4, 6, 1) and opacity.shape == (4, 6, 1)
    render_fn(mask)                                      # grad enabled -> the reference's own render_fn
    render_fn(None)                                      # no mask
    with torch.no_grad():
        render_fn(mask.float())                          # soft mask: not a keep flag
        render_fn(mask, return_info=True)
    assert len(calls["masked"]) == 1 and len(calls["render_fn"]) == 4


def test_forward_adds_original_rgb(fused_trainer_cls):
    cls, _ = fused_trainer_cls
    out = cls().forward({}, {})
    # scene_graph.py:293: rgb_gaussians + rgb_sky * (1 - opacity)
    assert torch.allclose(out["original_rgb"], torch.ones(2, 2, 3) + 0.5 * 0.75)
