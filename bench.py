#!/usr/bin/env python
"""bench.py - fwd+bwd Mpix/s of the fused render + bilateral hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm

A step = one pass of the hot path over one batch of synthetic input:
  SH + project -> bin + sort -> fused composite + bilateral forward -> photometric loss (+ TV) ->
  full backward to _means/_scales/_quats/_features_dc/_features_rest/_opacities and the 3 grids
  (-> one all-reduce of the flat gradient when N > 1).
Workload at N=1: BASELINE.json configs[2] - 2 M synthetic Gaussians, 6-camera nuScenes-shaped rig,
1920x1080, 3-scale grids (8,8,4)/(16,16,8)/(32,32,16), full-resolution guidance.  For N > 1 the
408 tile rows of the rig are split into N contiguous bands (strong scaling, fixed total work).

Prints ONE JSON line (rank 0).  ``value`` = whole-job Mpix/s with inputs resident in HBM;
``e2e`` = the same step with the GT images copied from pinned host memory and the loss read back
inside the timed region.  ``roofline`` is for the dominant kernel named by BASELINE.json (the fused
composite + bilateral forward), timed live with CUDA events on the launching stream.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "fwd+bwd Mpix/s at 2M Gaussians, 6x1080p; achieved HBM GB/s vs B200 peak"
UNIT = "Mpix/s"
LAMBDA_D, LAMBDA_A, TV_W = 0.01, 0.05, 1.0


def ncu_profile_numbers(kernel_key):
    """DRAM traffic and executed warp-instructions of ONE launch of the roofline kernel on the default workload, from
    the committed parser output of the round's `ncu --set full` capture (scripts/ncu_traffic.py ->
    profiles/ncu_traffic.json).  None when the file or the kernel is missing."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        doc = json.load(open(path))
        for name, k in doc["kernels"].items():
            if kernel_key in name:
                return {"traffic": k["dram_bytes_read"] + k["dram_bytes_write"], "warp_instructions": k["warp_instructions"],
                        "source": f"profiles/ncu_traffic.json <- {doc.get('source', '')}"}
    except Exception:
        pass
    return None


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-gauss", type=int, default=2_000_000)
    ap.add_argument("--cams", type=int, default=6)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="print the per-phase CUDA-event times of rank 0 to stderr")
    ap.add_argument("--grad-exchange", default="compact", choices=["compact", "splats", "dense"],
                    help="N > 1: compact (default) = ONE all-reduce of the Gaussian gradients with the SH part as a colour "
                         "cotangent per (camera, Gaussian) - 11 + 3 C floats per Gaussian instead of 59 - expanded "
                         "afterwards; dense = ONE all-reduce of the flat 236 B x N parameter gradient (NCCL runs it as "
                         "NVLS in-switch reduction, 1.2 ms at 8 GPUs); splats = all-gather the per-splat gradient records "
                         "and run the projection backward over all of them on every rank (less data, but the uneven "
                         "all-gather measured slower: 3.89 vs 3.04 ms/step at 8 GPUs, profiles/r02_bench_n8_*.json)")
    ap.add_argument("--bands", default="balanced", choices=["balanced", "equal"],
                    help="N > 1: balanced = band boundaries chosen so that every rank gets about the same work (records of "
                         "an untimed probe render + a per-tile constant); equal = the same number of tile rows per rank")
    ap.add_argument("--guidance", default="full", choices=["full", "lowres"],
                    help="full: guidance_factor=None fused in the composite kernel (headline); lowres: the reference's "
                         "default [4,4,2] = composite mode 1 + stand-alone low-res bilateral kernels")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_threads() -> int:
    """Threads for the CPU legs: every host core, stated in ``cores``."""
    return max(1, os.cpu_count() or 1)


def _reference_bilateral_step(H, W, grids_cpu, rgb_in, G, guidance_factor, kind_box):
    """One fwd+bwd of the reference's pure-PyTorch bilateral path on the host cores: the reference's OWN
    ``models.modules.MultiScaleBilateralAffineTransform`` + the apply loop of scene_graph.py:112-117, imported from
    oracle/_ref (byte-for-byte copies, oracle/build_ref.py) or /root/reference; the oracle port only if neither
    exists.  Returns a closure that runs one timed iteration."""
    import torch

    from oracle import ref_loader

    sizes = [[g.shape[3], g.shape[2], g.shape[1]] for g in grids_cpu]   # (grid_X, grid_Y, grid_W)
    if ref_loader.reference_available():
        _, mods = ref_loader.load_reference()
        kind_box["kind"] = "reference"
        kind_box["source"] = ref_loader.reference_kind()
        m = mods.MultiScaleBilateralAffineTransform("Affine", n=1, grid=sizes, device="cpu")
        for i, g in enumerate(grids_cpu):
            getattr(m, f"bil_grids{i}").grids.data.copy_(g[None])
        info = {"img_idx": torch.zeros(H, W, dtype=torch.long)}

        def run():
            rgb = rgb_in.clone().requires_grad_(True)
            for p_ in m.parameters():
                p_.grad = None
            out = ref_loader.reference_apply_chain(rgb, m(rgb, info, guidance_factor=guidance_factor))
            (out * G).sum().backward()
        return run
    from oracle import bilateral_ref as B

    kind_box["kind"] = "port"
    kind_box["source"] = "oracle/bilateral_ref.py (oracle/_ref absent)"

    def run_port():
        rgb = rgb_in.clone().requires_grad_(True)
        grids = [x.clone().requires_grad_(True) for x in grids_cpu]
        (B.multiscale_forward(grids, rgb, guidance_factor) * G).sum().backward()
    return run_port


def cpu_reference_timings(H, W, grids_cpu, warmup=1, repeats=5, repeats_fullres=3):
    """Times the CPU reference for BOTH guidance modes: ``guidance_factor=None`` (the
    semantics of the fused kernel = the benchmark workload) and the reference's default ``[4,4,2]``.  Returns a dict
    with per-mode lists of seconds, the thread count and what was timed."""
    import torch

    torch.set_num_threads(cpu_threads())
    g = torch.Generator(); g.manual_seed(17)
    rgb_in = torch.rand(H, W, 3, generator=g)
    G = torch.randn(H, W, 3, generator=g)
    res = {"threads": torch.get_num_threads(), "torch": torch.__version__}
    for tag, gf, reps in (("none", None, repeats_fullres), ("f442", [4, 4, 2], repeats)):
        box = {}
        run = _reference_bilateral_step(H, W, grids_cpu, rgb_in, G, gf, box)
        times = []
        for it in range(warmup + reps):
            t0 = time.perf_counter()
            run()
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
        res[tag] = times
        res.update(box)
    return res


def _cpu_model_name():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown CPU"


def _cpu_baseline_block(H, W, t):
    best_none, best_442 = min(t["none"]), min(t["f442"])
    v = H * W / best_none / 1e6
    return {
        "value": v, "unit": UNIT, "cores": t["threads"], "kind": t["kind"], "source": t["source"],
        "guidance_442_value": H * W / best_442 / 1e6,
        "sample": (f"the reference's own MultiScaleBilateralAffineTransform + apply (bilateral half of the path; the "
                   f"rasteriser half is gsplat CUDA, no CPU implementation exists) fwd+bwd of sum(out*G) on ONE {W}x{H} "
                   f"image: value = guidance_factor=None (the workload's semantics), best of {len(t['none'])} after 1 "
                   f"warm-up = {best_none:.2f} s; guidance_442_value = reference default [4,4,2], best of "
                   f"{len(t['f442'])} = {best_442:.2f} s; {_cpu_model_name()}, {t['threads']} threads, torch {t['torch']}")}


def run_reference(args):
    """Reference arm: the reference has NO CPU (or any own) implementation of the rasteriser half (it is gsplat CUDA,
    a pip dependency); its own code on this path is the pure-PyTorch bilateral module, which is what runs here on the
    host cores, imported unmodified from oracle/_ref (or /root/reference).  Each step = fwd+bwd of one 1920x1080
    camera image with guidance_factor=None (the semantics of the GPU arm's workload)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from bilateral_driving_b200 import synthetic as S

    H, W = args.height, args.width
    grids = [x[0] for x in S.make_grids(1)]
    torch.set_num_threads(cpu_threads())
    g = torch.Generator(); g.manual_seed(17)
    rgb_in = torch.rand(H, W, 3, generator=g)
    Gm = torch.randn(H, W, 3, generator=g)
    box = {}
    gf = [4, 4, 2] if args.guidance == "lowres" else None
    run = _reference_bilateral_step(H, W, grids, rgb_in, Gm, gf, box)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        run()
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    val = H * W * len(times) / total / 1e6
    sample = (f"bilateral half only (the reference's rasteriser is gsplat CUDA, no CPU path exists): the reference's own "
              f"MultiScaleBilateralAffineTransform(guidance_factor={gf})+apply, fwd+bwd, one {W}x{H} image per step, "
              f"{_cpu_model_name()}, {torch.get_num_threads()} threads, torch {torch.__version__}")
    _emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # same workload name as the GPU arm; what a step of THIS arm covers is the bounded sample below
        "config": {"workload": _workload_name(args.n_gauss, args.cams, W, H, args.guidance, args.gpus), "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": box["kind"],
                         "source": box["source"], "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def _workload_name(n_gauss, cams, W, H, guidance, world=1):
    """BASELINE.json config the run corresponds to."""
    idx = {(500_000, 1): 1, (2_000_000, 6): 2 if world == 1 else 3, (8_000_000, 6): 4}.get((n_gauss, cams))
    tag = f"configs[{idx}]" if idx is not None and (W, H) == (1920, 1080) else "custom"
    return (f"{tag}: {n_gauss} synthetic Gaussians, {cams}-cam nuScenes-shaped rig {W}x{H}, 3-scale grids 8/16/32 "
            f"{'full-res guidance (fused)' if guidance == 'full' else 'guidance_factor=[4,4,2] (two-phase)'}, SH degree 3")


_JSON_OUT = None


def _claim_stdout():
    """stdout carries exactly ONE JSON line: everything else that writes to fd 1 (NCCL's version banner, library
    chatter) is sent to stderr; the JSON line goes to the saved descriptor."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _claim_stdout()
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torch.distributed.run)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from bilateral_driving_b200 import _lib, render, synthetic as S
    from bilateral_driving_b200.bilateral import total_variation_loss_levels
    from bilateral_driving_b200.dist import (allreduce_grads, balanced_bands, band_for_rank, band_pixel_rows,
                                             cameras_in_band)

    N, Cn, W, H = args.n_gauss, args.cams, args.width, args.height
    sizes = S.GRID_SIZES_BASELINE
    p_cpu = S.make_gaussians(N)
    vm, Ks = S.make_rig(Cn, W, H)
    grids_cpu = S.make_grids(Cn, sizes)
    rb, re = band_for_rank(rank, world, Cn, H)
    band_note = "equal tile-row counts"
    if world > 1 and args.bands == "balanced" and args.guidance == "full":
        # untimed probe: records per tile row of an equal split -> band boundaries of about equal work.  (A trainer
        # gets these counts for free from the previous step.)  Cost model from the N=1 breakdown: a tile costs about
        # as much as 146 records (per-pixel epilogues and the bilateral chain against the per-record walk).
        tw_, th_ = (W + 15) // 16, (H + 15) // 16
        with torch.no_grad():
            probe = render.render_fused({k: v.to(dev) for k, v in p_cpu.items()}, vm.to(dev), Ks.to(dev), W, H, sky=None,
                                        grid_slots=None, sh_degree=3, near_plane=0.1, row_begin=rb, row_end=re,
                                        dense_info=False)
        offs = probe["info"]["tile_offsets"].long()
        per_row = (offs[1:] - offs[:-1]).view(re - rb, tw_).sum(1).float() + 146.0 * tw_
        weights = torch.zeros(Cn * th_, device=dev)
        weights[rb:re] = per_row
        dist.all_reduce(weights)
        # + about 1 M records' worth per camera a band touches: projection and emission run once per camera of the band
        rb, re = balanced_bands(weights.tolist(), world, rows_per_camera=th_, camera_cost=1.0e6)[rank]
        band_note = "tile-row bands of equal estimated work (records + 146 per tile + 1e6 per camera touched)"
        del probe
        torch.cuda.empty_cache()
    r0, r1 = band_pixel_rows(rb, re, Cn, H)
    cams = cameras_in_band(rb, re, H)
    rows = r1 - r0
    # band slices of the per-pixel images (generated per camera to bound host memory)
    sky_rows, gt_rows = [], []
    for c in cams:
        g = torch.Generator(); g.manual_seed(17 + 1000 * c)
        sky_c = torch.rand(H, W, 3, generator=g)
        gt_c = torch.rand(H, W, 3, generator=g)
        lo, hi = max(r0, c * H) - c * H, min(r1, (c + 1) * H) - c * H
        sky_rows.append(sky_c[lo:hi]); gt_rows.append(gt_c[lo:hi])
    sky = torch.cat(sky_rows).to(dev)
    gt_host = torch.cat(gt_rows).pin_memory()
    gt_dev = gt_host.to(dev)
    params = {k: v.to(dev).requires_grad_(True) for k, v in p_cpu.items()}
    grids = [g.to(dev).requires_grad_(True) for g in grids_cpu]
    vm_d, Ks_d = vm.to(dev), Ks.to(dev)
    vm_host, Ks_host = vm.pin_memory(), Ks.pin_memory()
    leaves = list(params.values()) + grids
    total_px = Cn * H * W
    info_box = {}

    def step(gt, vmx, Ksx, gt_ready=None):
        for t in leaves:
            t.grad = None
        per_cam = [g.unbind(0) for g in grids]  # one autograd node per level (backward = one stack), not C selects
        slots = [[u[c] for u in per_cam] if c in cams else None for c in range(Cn)]
        out = render.render_fused(params, vmx, Ksx, W, H, sky=sky, grid_slots=slots, bil_sizes=sizes, sh_degree=3,
                                  near_plane=0.1, row_begin=rb, row_end=re, absgrad=True, dense_info=False,
                                  guidance_factor=(4, 4, 2) if args.guidance == "lowres" else None,
                                  exchange_group=True if (world > 1 and args.grad_exchange != "dense") else None,
                                  exchange_mode=args.grad_exchange if args.grad_exchange != "dense" else "splats")
        if gt_ready is not None:  # the GT image copy ran on a side stream, overlapped with the render
            torch.cuda.current_stream().wait_event(gt_ready)
        loss = render.photometric_loss(out["rgb"], gt, out["depth"], out["opacity"], LAMBDA_D, LAMBDA_A, count=total_px,
                                       unit_cotangent=True)
        # TV over all image slots (all levels in one launch).  It does not depend on the band: with the compact
        # exchange (grid gradients reduced inside the render backward) every rank adds it itself, otherwise rank 0
        # alone does and the all-reduce below spreads it
        tv_everywhere = world > 1 and args.grad_exchange == "compact"
        if rank == 0 or tv_everywhere:
            tv = total_variation_loss_levels(grids, [TV_W * 0.5 * (sx * sy * sl) ** 0.5 for sx, sy, sl in sizes])
            loss = loss + tv
        loss.backward()
        with render._timed("allreduce"):
            if out["info"].get("grids_are_global"):    # everything was reduced inside the render backward
                pass
            elif out["info"].get("grads_are_global"):  # the Gaussian gradients are already the job's: grids only
                allreduce_grads([t.grad for t in grids])
            else:
                allreduce_grads([t.grad for t in leaves], flat=out["info"].get("grad_flat"))
        info_box.update(n_isect=out["info"]["n_isect"], n_visible=out["info"]["n_visible"])
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(max(args.warmup, 3)):
        step(gt_dev, vm_d, Ks_d)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    render.KERNEL_EVENTS = {}
    launches0 = _lib.lib.bds_launch_count()
    profile_range = os.environ.get("BDS_PROFILE_RANGE") == "1"  # ncu --profile-from-start off: the timed steps only
    if profile_range:
        torch.cuda.profiler.start()
    ms = timed(lambda: step(gt_dev, vm_d, Ks_d), args.steps)
    if profile_range:
        torch.cuda.profiler.stop()
    launches = (_lib.lib.bds_launch_count() - launches0) // args.steps
    ev = render.KERNEL_EVENTS
    render.KERNEL_EVENTS = None
    t_fwd = sum(a.elapsed_time(b) for a, b in ev.get("composite_fwd", [])) / max(len(ev.get("composite_fwd", [])), 1)
    t_bwd = sum(a.elapsed_time(b) for a, b in ev.get("composite_bwd", [])) / max(len(ev.get("composite_bwd", [])), 1)

    if args.breakdown and rank == 0:
        parts = {k: sum(a.elapsed_time(b) for a, b in v) / args.steps for k, v in ev.items()}
        parts["other (loss, TV, fills, gaps)"] = ms / args.steps - sum(parts.values())
        sys.stderr.write("phase ms/step: " + json.dumps({k: round(v, 3) for k, v in parts.items()}) + "\n")

    if os.environ.get("BDS_TIMELINE") and rank == 0:
        # profiling aid: kernel timeline of three steps (torch.profiler / CUPTI) -> gaps between kernels
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(3):
                step(gt_dev, vm_d, Ks_d)
            torch.cuda.synchronize()
        prof.export_chrome_trace(os.environ["BDS_TIMELINE"])

    # end to end through the public API with HOST buffers: GT image + cameras copied from pinned
    # memory every step, loss read back
    copy_stream = torch.cuda.Stream()

    def e2e_step():
        # cameras first (the render needs them), then the GT images on a side stream so that the 149 MB
        # PCIe copy overlaps projection / sort / composite; the loss kernel waits for it
        v = vm_host.to(dev, non_blocking=True)
        k = Ks_host.to(dev, non_blocking=True)
        copy_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(copy_stream):
            gt = gt_host.to(dev, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(copy_stream)
        gt.record_stream(torch.cuda.current_stream())
        loss = step(gt, v, k, gt_ready=ready)
        return float(loss.detach())  # device -> host read of the step's result

    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    default_workload = (world == 1 and (N, Cn, W, H) == (2_000_000, 6, 1920, 1080))
    I, Nv = info_box["n_isect"], info_box["n_visible"]
    V = 12 * sum(gx * gy * gl for gx, gy, gl in sizes)
    band_px = rows * W
    tw = (W + 15) // 16
    n_tiles = (re - rb) * tw
    # B_fwd = 44 I + 48 Px + 4 tiles + 4 C V   (rank-0 band; I = records actually read)
    b_fwd = 44 * I + 48 * band_px + 4 * n_tiles + 4 * len(cams) * V
    achieved = b_fwd / (t_fwd * 1e-3) / 1e9 if t_fwd > 0 else 0.0
    value = total_px * args.steps / (ms * 1e-3) / 1e6
    e2e_v = total_px * args.steps / (ms_e2e * 1e-3) / 1e6
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": _workload_name(N, Cn, W, H, args.guidance, world),
                   "gsplat": "not installed on the B200 box (profiles/r02_gsplat_probe.txt): no GPU reference comparator",
                   "parallelism": f"tile-row bands x{world}" + ("" if world == 1 else f" ({band_note}), gradient exchange: {args.grad_exchange}"),
                   "n_isect_rank0": I, "n_visible_rank0": Nv,
                   "l2": "inputs larger than L2 (472 MB of parameters + images per step)",
                   "composite_fwd_ms": t_fwd, "composite_bwd_ms": t_bwd},
        "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": gt_host.numel() * 4 + vm.numel() * 4 + Ks.numel() * 4,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    kernel_key = "composite_fwd_kernel<2>" if args.guidance == "full" else "composite_fwd_kernel<1>"
    prof = ncu_profile_numbers(kernel_key) if default_workload else None
    line["roofline"] = {
        "bound": "hbm",
        "kernel": ("composite_fwd_kernel<2> (fused composite + glue + bilateral)" if args.guidance == "full"
                   else "composite_fwd_kernel<1> (composite + glue)"),
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
        "algorithmic_bytes": b_fwd,
        "traffic": prof["traffic"] if prof else None, "traffic_source": prof["source"] if prof else None,
    }
    if prof and t_fwd > 0:
        # SURVEY 8d caveat: the kernel is bound by instruction issue, not by HBM - its second roofline is the issue
        # ceiling: one warp-instruction per clock per SM sub-partition (148 SMs x 4) at the SM clock sampled in this run
        mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
        issue_peak = 148 * 4 * mhz * 1e6
        issue_ach = prof["warp_instructions"] / (t_fwd * 1e-3)
        line["roofline"]["secondary"] = {
            "bound": "issue", "achieved": issue_ach / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-instr/s",
            "frac": issue_ach / issue_peak, "warp_instructions_per_launch": prof["warp_instructions"],
            "source": prof["source"] + "; peak = 148 SMs x 4 schedulers x sampled SM clock"}
    if world == 1 and not args.no_cpu_baseline:
        # CPU baseline on the box's host cores: bounded sample = ONE camera image, both guidance modes
        t = cpu_reference_timings(H, W, [g[0] for g in grids_cpu])
        line["cpu_baseline"] = _cpu_baseline_block(H, W, t)
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
