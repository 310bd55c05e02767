"""CPU: the product's host/device math headers (projection + VJP, SH, exact ellipse-vs-rectangle
bound, resize taps, trilinear stencil), compiled with g++ (tests/hostcheck), against the oracle."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest
import torch

from bilateral_driving_b200 import synthetic as S
from oracle import bilateral_ref as B
from oracle import raster_ref as R
from oracle import sh_ref

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hc(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hc") / "libhc.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-I/usr/local/cuda/include", "-o", out,
                    os.path.join(HERE, "hostcheck", "hostcheck.cpp")], check=True)
    lib = C.CDLL(out)
    lib.hc_min_sigma_rect.restype = C.c_float
    lib.hc_min_sigma_rect.argtypes = [C.c_float] * 9
    lib.hc_lattice_coord.restype = C.c_float
    return lib


def fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _scene(n=400, W=96, H=64):
    p = S.make_gaussians(n, extent=8.0, scale_mean=0.15)
    p["_means"][:, 2] *= 0.4
    vm, Ks = S.make_rig(2, W, H)
    a = S.activate(p)
    return a, vm, Ks, W, H


def test_projection_forward_and_vjp(hc):
    a, vm, Ks, W, H = _scene()
    n = a["means"].shape[0]
    for c in range(2):
        means = a["means"].double().requires_grad_(True)
        quats = (a["quats"] * 1.7).double().requires_grad_(True)  # un-normalised on purpose
        scales = a["scales"].double().requires_grad_(True)
        view = vm[c].double().requires_grad_(True)
        pr = R.project(means, quats, scales, view, Ks[c].double(), W, H, near_plane=0.1)
        out = np.zeros((n, 8), np.float32)
        args = [np.ascontiguousarray(t.detach().float().numpy()) for t in (means, quats, scales, view, Ks[c])]
        hc.hc_project(n, *map(fp, args), W, H, C.c_float(0.3), C.c_float(0.1), C.c_float(1e10), C.c_float(0.0), fp(out))
        vis = pr["radii"] > 0
        assert int(vis.sum()) > 50
        # radius ceil() may legitimately flip when 3*sqrt(lambda) sits on an integer: allow ambiguous ones
        ok = (torch.from_numpy(out[:, 6]).long() == pr["radii"]) | pr["ambiguous"] | \
             ((pr["radii"] == 0) & (torch.from_numpy(out[:, 6]) == 0))
        assert bool(ok.all())
        v = vis.numpy() & (out[:, 6] > 0)
        assert np.abs(out[v, 0:2] - pr["means2d"].detach().numpy()[v]).max() < 2e-3
        assert np.abs(out[v, 2] - pr["depths"].detach().numpy()[v]).max() < 1e-5
        con = pr["conics"].detach().numpy()[v]
        assert (np.abs(out[v, 3:6] - con) / np.maximum(np.abs(con), 1e-3)).max() < 2e-3
        assert np.abs(out[v, 7] - pr["compensations"].detach().numpy()[v]).max() < 1e-4
        # VJP with random cotangents on the visible ones
        gen = torch.Generator(); gen.manual_seed(c)
        cot = torch.randn(n, 7, generator=gen, dtype=torch.float64) * vis[:, None]
        loss = (pr["means2d"] * cot[:, 0:2]).sum() + (pr["depths"] * cot[:, 2]).sum() + \
               (pr["conics"] * cot[:, 3:6]).sum() + (pr["compensations"] * cot[:, 6]).sum()
        loss.backward()
        vme, vq, vs = (np.zeros((n, k), np.float32) for k in (3, 4, 3))
        vview = np.zeros(12, np.float32)
        cot32 = np.ascontiguousarray(cot.float().numpy())
        hc.hc_project_vjp(n, *map(fp, args), W, H, C.c_float(0.3), fp(cot32), fp(vme), fp(vq), fp(vs), fp(vview))
        for ours, ref in ((vme, means.grad), (vq, quats.grad), (vs, scales.grad)):
            ref = ref.numpy() * vis.numpy()[:, None]
            ours = ours * vis.numpy()[:, None]
            scale = np.abs(ref).max(axis=1, keepdims=True) + 1e-6
            assert (np.abs(ours - ref) / scale).max() < 5e-3, (np.abs(ours - ref) / scale).max()
        gv = view.grad.numpy()
        ref_view = np.concatenate([gv[:3, :3].reshape(-1), gv[:3, 3]])
        assert np.abs(vview - ref_view).max() < 2e-3 * np.abs(ref_view).max()


def test_projection_culls_radius_clip_and_far_plane(hc):
    """The viewer path renders with radius_clip=4.0 (base.py:811-826) and the trainers pass a far plane: the device
    projection math (compiled for the host) must cull exactly the Gaussians the oracle culls (radius <= radius_clip,
    depth outside [near, far])."""
    a, vm, Ks, W, H = _scene(n=600)
    n = a["means"].shape[0]
    a["scales"] = a["scales"].clone()
    a["scales"][::3] *= 0.03                                          # every third Gaussian projects to 2-4 pixels
    args = [np.ascontiguousarray(t.float().numpy()) for t in (a["means"], a["quats"], a["scales"], vm[0], Ks[0])]
    base = R.project(a["means"].double(), a["quats"].double(), a["scales"].double(), vm[0].double(), Ks[0].double(), W, H,
                     near_plane=0.1)
    far = float(base["depths"][base["radii"] > 0].median())          # a far plane that cuts the scene in two
    for clip, far_plane in ((4.0, 1e10), (0.0, far), (4.0, far)):
        pr = R.project(a["means"].double(), a["quats"].double(), a["scales"].double(), vm[0].double(), Ks[0].double(),
                       W, H, near_plane=0.1, far_plane=far_plane, radius_clip=clip)
        out = np.zeros((n, 8), np.float32)
        hc.hc_project(n, *map(fp, args), W, H, C.c_float(0.3), C.c_float(0.1), C.c_float(far_plane), C.c_float(clip),
                      fp(out))
        ours = torch.from_numpy(out[:, 6]).long()
        near_far = (base["depths"] - far_plane).abs() < 1e-4 * far_plane          # depth within rounding of the plane
        ok = (ours == pr["radii"]) | pr["ambiguous"] | base["ambiguous"] | near_far
        assert bool(ok.all()), (clip, far_plane, int((~ok).sum()))
        culled = int(((base["radii"] > 0) & (pr["radii"] == 0)).sum())
        assert culled > 10, (clip, far_plane, culled)                             # the option does cull something here


def test_projection_clamped_fov_branch(hc):
    # Gaussians far off-axis exercise the 1.3*tan(fov) clamp in J (appendix A.4)
    W, H = 64, 48
    vm, Ks = S.make_rig(1, W, H)
    c2w = torch.linalg.inv(vm[0])
    pts_cam = torch.tensor([[9.0, 0.2, 2.0], [-7.0, 3.0, 1.5], [0.5, -6.0, 1.2], [0.1, 0.1, 3.0]])
    means = (pts_cam @ c2w[:3, :3].T + c2w[:3, 3]).double().requires_grad_(True)
    quats = torch.tensor([[1.0, 0.2, -0.1, 0.3]] * 4).double().requires_grad_(True)
    scales = torch.tensor([[0.3, 0.1, 0.2]] * 4).double().requires_grad_(True)
    view = vm[0].double()
    pr = R.project(means, quats, scales, view, Ks[0].double(), W, H, near_plane=0.1)
    cot = torch.randn(4, 7, dtype=torch.float64)
    cot[:, 6] = 0
    ((pr["means2d"] * cot[:, 0:2]).sum() + (pr["depths"] * cot[:, 2]).sum() + (pr["conics"] * cot[:, 3:6]).sum()).backward()
    args = [np.ascontiguousarray(t.detach().float().numpy()) for t in (means, quats, scales, view, Ks[0])]
    vme, vq, vs = (np.zeros((4, k), np.float32) for k in (3, 4, 3))
    vview = np.zeros(12, np.float32)
    hc.hc_project_vjp(4, *map(fp, args), W, H, C.c_float(0.3), fp(np.ascontiguousarray(cot.float().numpy())),
                      fp(vme), fp(vq), fp(vs), fp(vview))
    for ours, ref in ((vme, means.grad), (vq, quats.grad), (vs, scales.grad)):
        ref = ref.numpy()
        assert (np.abs(ours - ref) / (np.abs(ref).max(axis=1, keepdims=True) + 1e-6)).max() < 5e-3


def test_min_sigma_rect_is_exact_box_minimum(hc):
    rng = np.random.default_rng(0)
    for _ in range(300):
        # random SPD conic, random centre, random rectangle
        a, c = rng.uniform(0.01, 2.0, 2)
        b = rng.uniform(-0.95, 0.95) * math.sqrt(a * c)
        gx, gy = rng.uniform(-30, 30, 2)
        x0, y0 = rng.uniform(-20, 10, 2)
        w, h = rng.uniform(0.0, 15.0, 2)
        s = hc.hc_min_sigma_rect(gx, gy, a, b, c, x0, x0 + w, y0, y0 + h)
        xs = np.linspace(x0, x0 + w, 301)
        ys = np.linspace(y0, y0 + h, 301)
        X, Y = np.meshgrid(xs, ys)
        dx, dy = gx - X, gy - Y
        brute = (a * dx * dx + b * dx * dy + c * dy * dy).min()
        assert s <= brute * (1 + 1e-4) + 1e-4           # never above the true minimum (conservative cull)
        assert s >= brute * (1 - 2e-3) - 2e-3           # and tight


def test_sh_basis_and_gradient(hc):
    gen = torch.Generator(); gen.manual_seed(3)
    for _ in range(20):
        d = torch.randn(3, generator=gen, dtype=torch.float64)
        d = (d / d.norm()).requires_grad_(True)
        for deg in range(4):
            b = np.zeros(16, np.float32); gx = np.zeros(16, np.float32); gy = np.zeros(16, np.float32); gz = np.zeros(16, np.float32)
            hc.hc_sh_basis(deg, C.c_float(float(d[0])), C.c_float(float(d[1])), C.c_float(float(d[2])), fp(b), fp(gx), fp(gy), fp(gz))
            nb = (deg + 1) ** 2
            # oracle basis WITHOUT re-normalisation: evaluate the polynomial at d directly
            x, y, z = d.unbind()
            ref = sh_ref.sh_bases(deg, d)  # normalises; d is unit so identical value
            assert np.abs(b[:nb] - ref.detach().numpy()).max() < 1e-6
    # gradient of the polynomial: finite differences of the product's own basis
    d0 = np.array([0.3, -0.5, 0.81], np.float64); d0 /= np.linalg.norm(d0)
    b0 = np.zeros(16, np.float32); gx = np.zeros(16, np.float32); gy = np.zeros(16, np.float32); gz = np.zeros(16, np.float32)
    hc.hc_sh_basis(3, C.c_float(d0[0]), C.c_float(d0[1]), C.c_float(d0[2]), fp(b0), fp(gx), fp(gy), fp(gz))
    eps = 1e-3
    for ax, g in enumerate((gx, gy, gz)):
        dp, dm = d0.copy(), d0.copy()
        dp[ax] += eps; dm[ax] -= eps
        bp = np.zeros(16, np.float32); bm = np.zeros(16, np.float32); t = np.zeros(16, np.float32)
        hc.hc_sh_basis(3, C.c_float(dp[0]), C.c_float(dp[1]), C.c_float(dp[2]), fp(bp), fp(t), fp(t), fp(t))
        hc.hc_sh_basis(3, C.c_float(dm[0]), C.c_float(dm[1]), C.c_float(dm[2]), fp(bm), fp(t), fp(t), fp(t))
        assert np.abs((bp - bm) / (2 * eps) - g).max() < 2e-3


def test_resize_taps_and_trilinear_stencil(hc):
    for (n_in, n_out) in ((270, 1080), (1080, 270), (7, 3), (3, 7), (5, 5)):
        i0r, i1r, tr = B.lin_src(n_out, n_in)
        for d in range(n_out):
            i0, i1, t = C.c_int(), C.c_int(), C.c_float()
            hc.hc_lin_src(d, n_in, n_out, C.byref(i0), C.byref(i1), C.byref(t))
            assert (i0.value, i1.value) == (int(i0r[d]), int(i1r[d])) and abs(t.value - float(tr[d])) < 1e-5
    assert abs(hc.hc_lattice_coord(5, 11, 8) - 0.5 * 7) < 1e-5
    g = torch.randn(12, 4, 5, 6, dtype=torch.float64)
    for (fx, fy, fz) in ((2.3, 1.7, 0.4), (-1.0, 9.0, 3.0), (5.0, 4.0, 2.999), (0.0, 0.0, -0.5)):
        nodes = (C.c_int * 8)(); w = (C.c_float * 8)()
        inside = hc.hc_tri(C.c_float(fx), C.c_float(fy), C.c_float(fz), 4, 5, 6, nodes, w)
        flat = g.permute(0, 2, 3, 1).reshape(12, -1)  # product repack order [GY][GX][L]
        ours = sum(flat[:, nodes[k]] * w[k] for k in range(8))
        ref = B.trilerp(g, torch.tensor(fx).double(), torch.tensor(fy).double(), torch.tensor(fz).double())
        assert (ours - ref).abs().max() < 1e-5
        assert inside == int(0 < fz < 3)


def test_value_repack_index_maps(hc):
    """The quad-major value repack [GY][GX][3][L][4]: the repack kernels' element -> parameter index map is a
    bijection onto [12][L][GY][GX], the slice's (node, slab, affine row) offset addresses the same element, and
    the hoisted lattice coordinate equals the reference form bit for bit."""
    hc.hc_value_param_index.restype = C.c_long
    hc.hc_unit_lin01.restype = C.c_float
    for (L, GY, GX) in ((4, 8, 8), (8, 16, 16), (3, 5, 6), (1, 2, 2)):
        n = 12 * L * GY * GX
        seen = np.zeros(n, dtype=np.int32)
        for i in range(n):
            seen[hc.hc_value_param_index(i, L, GY, GX)] += 1
        assert (seen == 1).all()
        for (x, y, z, ch) in ((0, 0, 0, 0), (GX - 1, GY - 1, L - 1, 11), (GX // 2, GY // 3, L // 2, 6), (1, 0, 0, 5)):
            node = hc.hc_node(x, y, z, L, GX)
            off = hc.hc_value_offset(node, z, ch // 4, L) + ch % 4
            assert hc.hc_value_param_index(off, L, GY, GX) == ((ch * L + z) * GY + y) * GX + x
    for (j, n_, g) in ((0, 1920, 32), (5, 11, 8), (1919, 1920, 8), (540, 1080, 16), (7, 8, 2)):
        assert hc.hc_unit_lin01(j, n_, g) == hc.hc_lattice_coord(j, n_, g)


def test_candidate_rect_contains_every_hit_tile(hc):
    """The ellipse-bbox pruning of the tile enumeration never drops a tile the exact test accepts."""
    rng = np.random.default_rng(1)
    hc.hc_candidate_rect_check.argtypes = [C.c_float] * 7 + [C.c_int, C.c_int, C.POINTER(C.c_int)]
    tot = np.zeros(3, np.int64)
    for _ in range(3000):
        W, H = 1920, 1080
        sx, sy = np.exp(rng.uniform(np.log(0.6), np.log(120.0), 2))  # pixel sigmas
        rho = rng.uniform(-0.95, 0.95)
        cov = np.array([[sx * sx, rho * sx * sy], [rho * sx * sy, sy * sy]])
        con = np.linalg.inv(cov)
        lam = 0.5 * (cov[0, 0] + cov[1, 1]) + np.sqrt(max(0.01, (0.5 * (cov[0, 0] + cov[1, 1])) ** 2 - np.linalg.det(cov)))
        radius = float(np.ceil(3 * np.sqrt(lam)))
        op = float(np.clip(rng.uniform(0.0, 1.0) ** 2, 1.0 / 255 + 1e-4, 1.0))
        LOG2E = 1.4426950408889634
        qa, qb, qc = 0.5 * LOG2E * con[0, 0], LOG2E * con[0, 1], 0.5 * LOG2E * con[1, 1]
        mx, my = rng.uniform(-100, W + 100), rng.uniform(-100, H + 100)
        stats = (C.c_int * 3)()
        bad = hc.hc_candidate_rect_check(mx, my, radius, qa, qb, qc, float(np.log2(255 * op)), W, H, stats)
        assert bad == 0
        tot += np.array(list(stats))
    assert tot[2] > 0 and tot[1] < tot[0]  # it does prune, and hits exist
