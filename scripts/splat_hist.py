"""Profiling aid: distribution of tiles touched per visible splat of the bench scene (load balance of the
warp-cooperative tile enumeration in projection / emission)."""
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
    t = out["info"]["tiles_touched"].reshape(-1).long()
    t = t[t > 0]
    edges = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 1 << 30]
    res = dict(visible=int(t.numel()), records=int(t.sum()), max=int(t.max()))
    res["splats_by_tiles"] = {f"<{edges[i + 1]}": int(((t >= edges[i]) & (t < edges[i + 1])).sum()) for i in range(len(edges) - 1)}
    res["records_by_tiles"] = {f"<{edges[i + 1]}": int(t[(t >= edges[i]) & (t < edges[i + 1])].sum()) for i in range(len(edges) - 1)}
    # per-warp totals in slot order approximated by Gaussian order: 32 consecutive (cam, gaussian) pairs
    tt = out["info"]["tiles_touched"].reshape(-1).long()
    pad = (-tt.numel()) % 32
    w = torch.cat([tt, tt.new_zeros(pad)]).view(-1, 32).sum(1)
    res["per_warp_of_32_pairs"] = dict(mean=float(w.float().mean()), p99=float(w.float().quantile(0.99)), max=int(w.max()))
    print(json.dumps(res))
