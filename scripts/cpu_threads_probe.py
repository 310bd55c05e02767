"""Probe: CPU bilateral baseline (oracle port) speed vs torch thread count on this host."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bilateral_driving_b200 import synthetic as S
from oracle import bilateral_ref as B

H, W = 1080, 1920
g = torch.Generator(); g.manual_seed(17)
rgb_in = torch.rand(H, W, 3, generator=g)
grids = [x[0] for x in S.make_grids(1)]
G = torch.randn(H, W, 3, generator=g)
print("cpu_count", os.cpu_count())
for nt in (8, 16, 32, 64, os.cpu_count()):
    torch.set_num_threads(nt)
    best = 1e9
    for it in range(2):
        rgb = rgb_in.clone().requires_grad_(True)
        gr = [x.clone().requires_grad_(True) for x in grids]
        t0 = time.perf_counter()
        (B.multiscale_forward(gr, rgb, None) * G).sum().backward()
        best = min(best, time.perf_counter() - t0)
    print(nt, "threads:", round(best, 2), "s", round(H * W / best / 1e6, 3), "Mpix/s", flush=True)
