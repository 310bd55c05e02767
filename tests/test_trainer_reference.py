"""The drop-in trainer against the reference's UNMODIFIED trainer code.

Reference arm: ``models.trainers.scene_graph.MultiTrainer`` + ``models.modules.MultiScaleBilateralAffineTransform``
imported from the reference tree (``/root/reference/project`` or its byte-for-byte copies under ``oracle/_ref``),
running on the CPU with the oracle as ``gsplat``.  Checked arm: ``FusedMultiTrainer`` + this package's Affine module
with the same parameters.

* CPU test (``-m "not gpu"``): the kernel-backed entry points are replaced by their oracle restatements, so what is
  checked is the trainer's host logic - the output dictionary, ``loss_dict`` (``affine_loss`` present and equal to the
  reference's), every parameter gradient, the densification taps, the per-class eval renders.
* GPU test (``-m gpu``): the same comparison with the real sm_100a kernels on the fused arm.
"""
import pytest
import torch

import trainer_harness as TH
from oracle.ref_loader import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="needs the reference tree or oracle/_ref")


def _reference_arm(w1, single=False):
    sg = TH.load_scene_graph()
    cfg = TH.make_cfg(TH.REF_SINGLE if single else TH.REF_MS, w1=w1, single=single)
    tr = TH.build_trainer(sg.MultiTrainer, cfg, torch.device("cpu"))
    TH.init_scene(tr, "cpu")
    return tr


def _fused_arm(ref_trainer, device, w1, single=False, guidance_factor="default", fused_gaussians=False):
    TH.load_scene_graph()
    import bilateral_driving_b200.trainer as T

    bg = "bilateral_driving_b200.gaussians.FusedVanillaGaussians" if fused_gaussians else "models.gaussians.VanillaGaussians"
    cfg = TH.make_cfg(TH.OUR_SINGLE if single else TH.OUR_MS, w1=w1, single=single, background_type=bg)
    cfg.trainer.type = "bilateral_driving_b200.trainer.FusedMultiTrainer"
    tr = TH.build_trainer(T.FusedMultiTrainer, cfg, torch.device(device))
    for m in tr.models.values():
        m.to(device)
    TH.copy_models(ref_trainer, tr, device)
    if guidance_factor != "default":
        tr.guidance_factor = guidance_factor
    return tr


def _step(tr, image_infos, cam_infos, fix_reference_bug=False):
    tr.set_train()
    tr.preprocess_per_train_step(tr.step)
    for p in TH.all_params(tr).values():
        p.grad = None
    outputs = tr(image_infos, cam_infos)
    tr.update_visibility_filter()
    if fix_reference_bug:
        # the published MultiTrainer.forward never sets outputs["original_rgb"] although compute_losses reads it
        # (base.py:628-631): the first step of every ms-bilateral config raises KeyError
        with pytest.raises(KeyError):
            tr.compute_losses(outputs, image_infos, cam_infos)
        outputs["original_rgb"] = outputs["rgb_gaussians"] + outputs["rgb_sky"] * (1.0 - outputs["opacity"])  # base.py:496
    loss_dict = tr.compute_losses(outputs, image_infos, cam_infos)
    sum(loss_dict.values()).backward()
    return outputs, loss_dict


def _patch_cpu(monkeypatch):
    import bilateral_driving_b200.bilateral as BL
    import bilateral_driving_b200.render as R

    monkeypatch.setattr(R, "render_fused", TH.oracle_render_fused)
    monkeypatch.setattr(BL, "multiscale_bilateral", TH.oracle_multiscale_bilateral)
    monkeypatch.setattr(BL, "total_variation_loss", TH.oracle_tv)
    monkeypatch.setattr(BL, "total_variation_loss_levels", TH.oracle_tv_levels)


def _compare(ref, fused, o_ref, o_fu, l_ref, l_fu, tol_img, tol_loss, tol_grad, keep=None):
    dev = o_fu["rgb"].device
    k = None if keep is None else keep.to(dev)[..., None]
    for key in ("rgb", "rgb_gaussians", "opacity", "rgb_sky", "rgb_sky_blend", "original_rgb"):
        assert key in o_fu, key
        d = (o_fu[key] - o_ref[key].to(dev)).abs()
        assert float((d if k is None else d * k).max()) < tol_img, key
    d = (o_fu["depth"] - o_ref["depth"].to(dev)).abs() / o_ref["depth"].to(dev).abs().clamp(min=1.0)
    assert float((d if k is None else d * k).max()) < 2 * tol_img
    assert set(l_fu) == set(l_ref), (sorted(l_fu), sorted(l_ref))
    assert "affine_loss" in l_fu
    for key in l_ref:
        a, b = float(l_fu[key]), float(l_ref[key])
        assert abs(a - b) <= tol_loss * max(1.0, abs(b)), (key, a, b)
    pr, pf = TH.all_params(ref), TH.all_params(fused)
    assert set(pr) == set(pf)
    for name in pr:
        gr, gf = pr[name].grad, pf[name].grad
        assert (gr is None) == (gf is None), name
        if gr is None:
            continue
        scale = float(gr.abs().max().clamp(min=1e-12))
        assert float((gf.cpu() - gr).abs().max()) / scale < tol_grad, name
    # densification taps (base.py:279-297): values, not just shapes
    ir, iff = ref.info, fused.info
    assert int((iff["radii"].cpu() != ir["radii"]).sum()) <= 2
    for attr in ("grad", "absgrad"):
        a, b = getattr(iff["means2d"], attr), getattr(ir["means2d"], attr)
        assert a is not None and b is not None and a.shape == b.shape, attr
        assert float((a.cpu() - b).abs().max()) / float(b.abs().max().clamp(min=1e-12)) < tol_grad, attr
    assert int(iff["width"]) == int(ir["width"]) and int(iff["height"]) == int(ir["height"])


@pytest.mark.parametrize("w1", [0.0, 0.5])
def test_fused_trainer_host_logic_against_reference_trainer(monkeypatch, w1):
    """CPU: FusedMultiTrainer.forward / compute_losses against the reference's own BasicTrainer.compute_losses /
    MultiTrainer.forward on the same parameters (default guidance_factor=[4,4,2])."""
    _patch_cpu(monkeypatch)
    ref = _reference_arm(w1)
    fused = _fused_arm(ref, "cpu", w1)
    image_infos, cam_infos = TH.make_batch("cpu")
    o_ref, l_ref = _step(ref, image_infos, cam_infos, fix_reference_bug=True)
    o_fu, l_fu = _step(fused, image_infos, cam_infos)
    # the reference's expression, spelled out: affine.w * tv_loss() + affine.w1 * inverse_loss(gt, original_rgb)
    aff = ref.models["Affine"]
    mask = (1.0 - image_infos["egocar_masks"]).float()[..., None]
    expect = 0.01 * aff.tv_loss() + w1 * aff.inverse_loss(image_infos["pixels"] * mask, o_ref["original_rgb"] * mask)
    assert abs(float(l_fu["affine_loss"]) - float(expect)) < 1e-6 * max(1.0, abs(float(expect)))
    assert float(l_fu["affine_loss"]) != 0.0
    _compare(ref, fused, o_ref, o_fu, l_ref, l_fu, tol_img=1e-6, tol_loss=1e-6, tol_grad=1e-4)
    # the type string is back in place after compute_losses
    assert fused.model_config.Affine.type == TH.OUR_MS
    # eval: per-class and dynamic-only renders (scene_graph.py:296-313)
    ref.set_eval(); fused.set_eval()
    with torch.no_grad():
        e_ref, e_fu = ref(image_infos, cam_infos), fused(image_infos, cam_infos)
    assert set(e_ref) | {"original_rgb"} == set(e_fu)
    for key in ("Background_rgb", "Background_opacity", "Background_depth", "Dynamic_rgb", "Dynamic_opacity"):
        assert float((e_fu[key] - e_ref[key]).abs().max()) < 1e-6, key


def test_single_grid_affine_type_also_gets_its_tv_loss(monkeypatch):
    """base.py:589-593: the BilateralAffineTransform branch (TV only)."""
    _patch_cpu(monkeypatch)
    ref = _reference_arm(0.0, single=True)
    fused = _fused_arm(ref, "cpu", 0.0, single=True)
    image_infos, cam_infos = TH.make_batch("cpu")
    o_ref, l_ref = _step(ref, image_infos, cam_infos)
    o_fu, l_fu = _step(fused, image_infos, cam_infos)
    o_ref["original_rgb"] = o_fu["original_rgb"]   # not produced by the reference on this branch (and not read)
    _compare(ref, fused, o_ref, o_fu, l_ref, l_fu, tol_img=1e-6, tol_loss=1e-6, tol_grad=1e-4)


def test_chain_inverse_closed_form_matches_torch_inverse():
    """modules.py:474-492 composes 4x4 homogeneous matrices and calls torch.inverse per pixel."""
    from bilateral_driving_b200.bilateral import affine_to_homogeneous_batch, chain_inverse_apply

    g = torch.Generator().manual_seed(0)
    H, W = 9, 7
    fields = [torch.eye(3, 4).expand(1, H, W, 3, 4) + 0.2 * torch.randn(1, H, W, 3, 4, generator=g) for _ in range(3)]
    gt = torch.rand(H, W, 3, generator=g)
    mat = torch.eye(4).view(1, 1, 1, 4, 4).repeat(1, H, W, 1, 1)
    for arr in fields:
        mat = affine_to_homogeneous_batch(arr) @ mat
    inv = torch.inverse(mat.view(-1, 4, 4)).view(1, H, W, 4, 4)[:, :, :, :3, :].reshape(H, W, 3, 4)
    want = (inv[..., :3, :3] @ gt[..., None] + inv[..., :3, 3:])[..., 0]
    got = chain_inverse_apply([f.reshape(H, W, 3, 4) for f in fields], gt)
    assert float((got - want).abs().max()) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("guidance", ["default", None])
def test_fused_trainer_on_gpu_against_reference_trainer(guidance):
    """GPU: the drop-in trainer with the real kernels (composite mode 1 + low-res bilateral kernels for the default
    guidance_factor=[4,4,2]; composite mode 2 for None) against the reference's trainer on the CPU oracle."""
    ref = _reference_arm(0.5)
    if guidance is None:
        # the reference trainer cannot be asked for guidance_factor=None (scene_graph.py:113 does not pass it):
        # give the reference module that default for this comparison
        aff = ref.models["Affine"]
        orig = aff.forward
        aff.forward = lambda rgb, infos, guidance_factor=None: orig(rgb, infos, guidance_factor=None)
    fused = _fused_arm(ref, "cuda", 0.5, guidance_factor=guidance, fused_gaussians=True)
    image_infos, cam_infos = TH.make_batch("cpu")
    g_infos, g_cam = TH.make_batch("cuda")
    o_ref, l_ref = _step(ref, image_infos, cam_infos, fix_reference_bug=True)
    o_fu, l_fu = _step(fused, g_infos, g_cam)
    keep = ~ref.info["ambiguous"][0]
    _compare(ref, fused, o_ref, o_fu, l_ref, l_fu, tol_img=2e-5, tol_loss=2e-4, tol_grad=2e-3, keep=keep)
    assert "_bds_cache" in fused.info and fused.info["_bds_cache"] is not None
    # densification bookkeeping (base.py:279-297 -> vanilla.py:151-191): the reference's after_train on the CPU against
    # FusedVanillaGaussians.after_train (one bds_densify_stats launch), first step (statistics None) and a second one
    for tr in (ref, fused):
        tr.initialize_optimizer()
    for rep in range(2):
        ref.postprocess_per_train_step(ref.step)
        fused.postprocess_per_train_step(fused.step)
        a, b = ref.models["Background"], fused.models["Background"]
        assert type(b).__name__ == "FusedVanillaGaussians"
        for name in ("xys_grad_norm", "vis_counts", "max_2Dsize"):
            want, got = getattr(a, name), getattr(b, name).cpu()
            assert want.shape == got.shape, name
            assert float((got - want).abs().max()) <= 2e-3 * float(want.abs().max().clamp(min=1e-12)), (name, rep)
        assert float(b.vis_counts.max()) == rep + 1.0
    # eval: masked re-renders reuse the sorted lists of the fused render
    from bilateral_driving_b200 import _lib

    ref.set_eval(); fused.set_eval()
    with torch.no_grad():
        e_ref = ref(image_infos, cam_infos)
        n0 = _lib.lib.bds_launch_count()
        e_fu = fused(g_infos, g_cam)
        torch.cuda.synchronize()
        launches = _lib.lib.bds_launch_count() - n0
    keep_g = keep.cuda()[..., None]
    for key in ("rgb", "Background_rgb", "Background_opacity", "Dynamic_rgb", "Dynamic_opacity"):
        assert float(((e_fu[key] - e_ref[key].cuda()).abs() * keep_g).max()) < 2e-5, key
    assert launches < 40   # one pipeline + two masked composites, not three rasterizations


class _EvalSplit:
    """Stands in for datasets.base.SplitWrapper (split_wrapper.py:22-27): hands out (image_infos, cam_infos) per index."""
    split = "test"

    def __init__(self, n=2):
        self.items = []
        for i in range(n):
            image_infos, cam_infos = TH.make_batch("cpu", img_idx=i, seed=20 + i)
            image_infos.pop("lidar_depth_map")          # the lidar overlay needs cv2 / matplotlib colour maps
            g = torch.Generator().manual_seed(40 + i)
            H, W = image_infos["pixels"].shape[:2]
            for k, frac in (("dynamic_masks", 0.2), ("human_masks", 0.05), ("vehicle_masks", 0.1)):
                image_infos[k] = (torch.rand(H, W, generator=g) < frac).float()
            cam_infos["cam_name"] = f"CAM_{i}"
            cam_infos["cam_id"] = torch.full((H, W), i, dtype=torch.long)
            self.items.append((image_infos, cam_infos))

    def __len__(self):
        return len(self.items)

    def get_image(self, idx, camera_downscale):
        assert camera_downscale == 1          # eval renders at full resolution (base.py:142-146)
        image_infos, cam_infos = self.items[idx]
        return dict(image_infos), dict(cam_infos)


def test_eval_harness_runs_the_drop_in_trainer(monkeypatch):
    """the reference's own eval harness (models/video_utils.py:47-620, render_images -> render, what
    tools/eval.py and the end of tools/train.py call), unmodified, over the reference trainer and over the drop-in
    trainer with the same parameters: same metrics, same per-frame arrays, same per-class renders.  skimage / lpips /
    imageio are stand-ins (oracle/ref_stubs.py) applied to both arms; `.cuda()` is the identity on this CPU run."""
    import numpy as np

    from oracle.ref_loader import load_reference_eval
    import os

    _patch_cpu(monkeypatch)
    ref = _reference_arm(0.5)
    fused = _fused_arm(ref, "cpu", 0.5)
    VU = load_reference_eval(os.path.join(TH.ROOT, "oracle", "gsplat_seam"))
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    data = _EvalSplit(2)
    r_ref = VU.render_images(ref, data, compute_metrics=True, compute_error_map=True)
    r_fu = VU.render_images(fused, data, compute_metrics=True, compute_error_map=True)
    assert not ref.training and not fused.training
    assert set(r_ref) == set(r_fu)
    for k in ("psnr", "ssim", "lpips", "occupied_psnr", "occupied_ssim", "masked_psnr", "masked_ssim", "human_psnr",
              "human_ssim", "vehicle_psnr", "vehicle_ssim"):
        assert r_ref[k] != -1 and np.isfinite(r_ref[k]), k
        assert abs(r_fu[k] - r_ref[k]) <= 1e-5 * max(1.0, abs(r_ref[k])), (k, r_fu[k], r_ref[k])
    for k in ("rgbs", "depths", "opacities", "gt_rgbs", "rgb_error_maps", "rgb_sky_blend", "rgb_sky", "Background_rgbs",
              "Background_depths", "Background_opacities"):
        assert k in r_fu and len(r_fu[k]) == len(r_ref[k]) == 2, k
        for a, b in zip(r_fu[k], r_ref[k]):
            assert a.shape == b.shape and float(np.abs(a - b).max()) <= 1e-5 * max(1.0, float(np.abs(b).max())), k
    assert r_fu["cam_names"] == r_ref["cam_names"] == ["CAM_0", "CAM_1"]


def _train_iteration(tr, step, image_infos, cam_infos, fix_reference_bug=False):
    """tools/train.py:252-284, one iteration of the reference's training loop."""
    tr.set_train()
    tr.preprocess_per_train_step(step=step)
    tr.optimizer_zero_grad()
    outputs = tr(image_infos, cam_infos)
    tr.update_visibility_filter()
    if fix_reference_bug:   # see _step
        outputs["original_rgb"] = outputs["rgb_gaussians"] + outputs["rgb_sky"] * (1.0 - outputs["opacity"])
    loss_dict = tr.compute_losses(outputs=outputs, image_infos=image_infos, cam_infos=cam_infos)
    for k, v in loss_dict.items():
        assert torch.isfinite(v).all(), (k, step)
    tr.backward(loss_dict)                      # backward + optimizer step + lr schedule (base.py:502-516)
    tr.postprocess_per_train_step(step=step)    # densification statistics; split / duplicate / cull every 100 steps
    return {k: float(v.detach()) for k, v in loss_dict.items()}


def test_training_loop_with_densification_matches_reference(monkeypatch):
    """CPU: four iterations of the reference's loop (tools/train.py:252-284) over both trainers - optimizer built by
    the reference's initialize_optimizer from each arm's param groups (Affine#grid{i} learning rates, base.py:182-208),
    steps 3198..3201 so that step 3200 runs the reference's split / duplicate / cull surgery on the statistics the
    taps of the drop-in fed it.  Same losses every step, same Gaussian count after the surgery, same parameters."""
    _patch_cpu(monkeypatch)
    ref = _reference_arm(0.5)
    fused = _fused_arm(ref, "cpu", 0.5)
    for tr in (ref, fused):
        tr.initialize_optimizer()
    names = lambda tr: sorted(g["name"] for g in tr.optimizer.param_groups)  # noqa: E731
    assert names(ref) == names(fused) and any(n.startswith("Affine#grid") for n in names(fused))
    lrs = lambda tr: {g["name"]: g["lr"] for g in tr.optimizer.param_groups}  # noqa: E731
    n0 = ref.models["Background"].num_points
    counts = []
    for i, step in enumerate(range(3198, 3202)):
        image_infos, cam_infos = TH.make_batch("cpu", img_idx=i % TH.N_IMAGES, seed=60 + i)
        torch.manual_seed(1000 + step)
        l_ref = _train_iteration(ref, step, image_infos, cam_infos, fix_reference_bug=True)
        torch.manual_seed(1000 + step)
        l_fu = _train_iteration(fused, step, image_infos, cam_infos)
        assert set(l_ref) == set(l_fu) and "affine_loss" in l_fu
        for k in l_ref:
            assert abs(l_fu[k] - l_ref[k]) <= 1e-5 * max(1.0, abs(l_ref[k])), (step, k, l_fu[k], l_ref[k])
        assert lrs(ref) == lrs(fused)
        a, b = ref.models["Background"], fused.models["Background"]
        assert a.num_points == b.num_points, step
        counts.append(a.num_points)
    assert counts[1] == n0 and counts[2] != n0, counts      # the surgery ran at step 3200 and changed the scene
    pr, pf = TH.all_params(ref), TH.all_params(fused)
    assert set(pr) == set(pf)
    for name in pr:
        assert pr[name].shape == pf[name].shape, name
        scale = float(pr[name].abs().max().clamp(min=1e-12))
        assert float((pf[name] - pr[name]).abs().max()) / scale < 1e-4, name


def test_checkpoints_cross_load_between_reference_and_drop_in_trainer(monkeypatch, tmp_path):
    """a checkpoint written by the reference trainer (base.py:739-752) resumes in the drop-in trainer
    (base.py:727-737, strict) and the other way round, and the resumed trainer renders what the writer renders."""
    import os

    _patch_cpu(monkeypatch)
    ref = _reference_arm(0.5)
    fused = _fused_arm(ref, "cpu", 0.5)
    image_infos, cam_infos = TH.make_batch("cpu")
    # move both arms off their common initialisation, differently
    g = torch.Generator().manual_seed(8)
    for tr, amp in ((ref, 0.01), (fused, 0.02)):
        for p in TH.all_params(tr).values():
            p.data += amp * torch.randn(p.shape, generator=g)
    (tmp_path / "ref").mkdir(); (tmp_path / "fused").mkdir()
    ref.save_checkpoint(str(tmp_path / "ref"), is_final=True)
    fused.save_checkpoint(str(tmp_path / "fused"), is_final=True)
    ck_ref = torch.load(os.path.join(tmp_path, "ref", "checkpoint_final.pth"))
    ck_fu = torch.load(os.path.join(tmp_path, "fused", "checkpoint_final.pth"))
    assert set(ck_ref) == set(ck_fu)
    for name in ck_ref["models"]:                    # same keys, same shapes: the checkpoint FORMAT is the reference's
        a, b = ck_ref["models"][name], ck_fu["models"][name]
        assert sorted(a) == sorted(b), name
        assert all(a[k].shape == b[k].shape for k in a if torch.is_tensor(a[k])), name

    def eval_render(tr):
        tr.set_eval()
        with torch.no_grad():
            return tr(image_infos, cam_infos)

    want_ref, want_fu = eval_render(ref), eval_render(fused)
    ref2 = _reference_arm(0.5)
    fused2 = _fused_arm(ref2, "cpu", 0.5)
    ref2.resume_from_checkpoint(os.path.join(tmp_path, "fused", "checkpoint_final.pth"))   # drop-in -> reference
    fused2.resume_from_checkpoint(os.path.join(tmp_path, "ref", "checkpoint_final.pth"))   # reference -> drop-in
    assert ref2.step == fused.step and fused2.step == ref.step
    got_ref2, got_fu2 = eval_render(ref2), eval_render(fused2)
    for key in ("rgb", "depth", "opacity", "Background_rgb"):
        assert float((got_ref2[key] - want_fu[key]).abs().max()) < 1e-6, key
        assert float((got_fu2[key] - want_ref[key]).abs().max()) < 1e-6, key
