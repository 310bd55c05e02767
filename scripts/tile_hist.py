"""Profiling aid: histogram of per-tile list lengths of the bench scene (one render_fused forward)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bilateral_driving_b200 import render, synthetic as S

if __name__ == "__main__":
    N, Cn, W, H = 2_000_000, 6, 1920, 1080
    dev = "cuda"
    params = {k: v.to(dev) for k, v in S.make_gaussians(N).items()}
    vm, Ks = S.make_rig(Cn, W, H)
    with torch.no_grad():
        out = render.render_fused(params, vm.to(dev), Ks.to(dev), W, H, sky=None, grid_slots=None, bil_sizes=(), sh_degree=3,
                                  near_plane=0.1, dense_info=False)
    offs = out["info"]["tile_offsets"].long()
    n = (offs[1:] - offs[:-1]).cpu()
    edges = [0, 1, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 1 << 30]
    hist = {f"<{edges[i + 1]}": int(((n >= edges[i]) & (n < edges[i + 1])).sum()) for i in range(len(edges) - 1)}
    rec = {f"<{edges[i + 1]}": int(n[(n >= edges[i]) & (n < edges[i + 1])].sum()) for i in range(len(edges) - 1)}
    print(json.dumps(dict(tiles=int(n.numel()), records=int(n.sum()), max=int(n.max()), mean=float(n.float().mean()),
                          tiles_by_length=hist, records_by_length=rec)))
