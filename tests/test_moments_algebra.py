"""CPU check of the algebra behind the deferred per-record reduction of the composite backward
(bilateral_driving_b200/csrc/composite.cu: flush_batch; projection.cu: project_bwd_kernel).

The walk stores w = [alpha unclamped] * araw * v_alpha per (pixel, record); the gradient record accumulates
pixel-LOCAL moments of w over each warp rectangle, re-centred on the splat mean, and project_bwd turns the
moments into v_mean2d / v_conic / v_opacity / absgrad.  Here the same steps in fp64 numpy against the textbook
per-pixel sums (gsplat rasterize_to_pixels_bwd as restated in oracle/raster_ref.py)."""
import numpy as np

LN2 = np.log(2.0)
LOG2E = 1.0 / LN2


def test_moment_pipeline_equals_per_pixel_sums():
    rng = np.random.default_rng(0)
    for _ in range(20):
        # one splat: mean, conic (a, b, c) positive definite, opacity
        mean = rng.uniform(0, 64, 2)
        L = rng.normal(size=(2, 2)) * 0.2
        conic = L @ L.T + 0.02 * np.eye(2)
        a, b, c = conic[0, 0], conic[0, 1], conic[1, 1]
        qa, qb, qc = 0.5 * LOG2E * a, LOG2E * b, 0.5 * LOG2E * c       # record fields (a', b', c')
        # several 8x4 warp rectangles, each with its own per-pixel w (any real numbers: the algebra is linear)
        tot = dict(vx=0.0, vy=0.0, va=0.0, vb=0.0, vc=0.0, absx=0.0, absy=0.0, m0=0.0)
        rec = np.zeros(12)
        for _r in range(5):
            x0, y0 = rng.integers(0, 8) * 8, rng.integers(0, 16) * 4
            u, v = np.meshgrid(np.arange(8), np.arange(4))
            w = rng.normal(size=(4, 8)) * (rng.random((4, 8)) < 0.8)
            # textbook per-pixel terms: sigma' = a' dx^2 + b' dx dy + c' dy^2, v_sigma' = -ln2 * w
            dx = mean[0] - (x0 + u + 0.5)
            dy = mean[1] - (y0 + v + 0.5)
            vs = -LN2 * w
            gx, gy = 2 * qa * dx + qb * dy, qb * dx + 2 * qc * dy
            tot["vx"] += (vs * gx).sum(); tot["vy"] += (vs * gy).sum()
            tot["va"] += (vs * dx * dx).sum() * 0.5 * LOG2E            # d sigma'/d a = log2e/2 dx^2
            tot["vb"] += (vs * dx * dy).sum() * LOG2E
            tot["vc"] += (vs * dy * dy).sum() * 0.5 * LOG2E
            tot["absx"] += np.abs(vs * gx).sum(); tot["absy"] += np.abs(vs * gy).sum()
            tot["m0"] += w.sum()
            # flush_batch: pixel-local moments, re-centred on the mean relative to pixel (0, 0) of the rectangle
            X, Y = mean[0] - (x0 + 0.5), mean[1] - (y0 + 0.5)
            m0, mu, mv = w.sum(), (w * u).sum(), (w * v).sum()
            muu, muv, mvv = (w * u * u).sum(), (w * u * v).sum(), (w * v * v).sum()
            mx, my = X * m0 - mu, Y * m0 - mv
            mxx = X * (mx - mu) + muu
            mxy = X * my - Y * mu + muv
            myy = Y * (my - mv) + mvv
            A2, B, C2 = 2 * qa, qb, 2 * qc
            ax = np.abs(w * (A2 * (X - u) + B * (Y - v))).sum()
            ay = np.abs(w * (B * (X - u) + C2 * (Y - v))).sum()
            rec += np.array([mx, my, mxx, mxy, myy, m0, 0, 0, 0, 0, ax, ay])
        # project_bwd_kernel: moments -> gradients
        vmx = -LN2 * (2 * qa * rec[0] + qb * rec[1])
        vmy = -LN2 * (qb * rec[0] + 2 * qc * rec[1])
        va, vb, vc = -0.5 * rec[2], -rec[3], -0.5 * rec[4]
        np.testing.assert_allclose([vmx, vmy], [tot["vx"], tot["vy"]], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose([va, vb, vc], [tot["va"], tot["vb"], tot["vc"]], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose([LN2 * rec[10], LN2 * rec[11]], [tot["absx"], tot["absy"]], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(rec[5], tot["m0"], rtol=1e-12, atol=1e-12)


def test_running_scalar_equals_four_colour_buffers():
    """v_alpha_i = sum_c (c_i T_i - buf_c / (1 - alpha_i)) v_C + T_final v_A / (1 - alpha_i) with one scalar."""
    rng = np.random.default_rng(1)
    n = 40
    alpha = rng.uniform(0.01, 0.9, n)
    col = rng.random((n, 4))
    vC = rng.normal(size=4)
    vA = rng.normal()
    T = np.concatenate([[1.0], np.cumprod(1 - alpha)])       # T[i] = transmittance in front of record i
    Tfin = T[-1]
    # textbook, back to front
    buf = np.zeros(4)
    ref = np.zeros(n)
    for i in range(n - 1, -1, -1):
        ra = 1.0 / (1.0 - alpha[i])
        ref[i] = ((col[i] * T[i] - buf * ra) * vC).sum() + Tfin * ra * vA
        buf += col[i] * alpha[i] * T[i]
    # one running scalar (composite.cu): bufdot starts at -T_final * v_A
    bufdot = -Tfin * vA
    Tcur = Tfin
    got = np.zeros(n)
    for i in range(n - 1, -1, -1):
        ra = 1.0 / (1.0 - alpha[i])
        Tcur *= ra
        cdot = (col[i] * vC).sum()
        got[i] = Tcur * cdot - ra * bufdot
        bufdot += alpha[i] * Tcur * cdot
    np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-12)
