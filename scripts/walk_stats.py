"""Profiling aid: sub-warp hit statistics of the composite walk.  `make -C bilateral_driving_b200/csrc stats`
builds libbds_b200_stats.so (-DBDS_STATS); this script selects it through BDS_LIB."""
import ctypes as C, json, sys, os
os.environ.setdefault("BDS_LIB", "libbds_b200_stats.so")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bilateral_driving_b200 import _lib
import bench  # noqa: F401  (reuses the bench scene through its CLI below)

if __name__ == "__main__":
    lib = _lib.lib
    out = (C.c_ulonglong * 16)()
    lib.bds_debug_stats(out, 1)
    sys.argv = ["bench.py", "--steps", "1", "--warmup", "3", "--no-cpu-baseline"]
    bench.main()
    lib.bds_debug_stats(out, 0)
    v = [int(x) for x in out]
    names = ["batches", "hits_8x4", "max2_batch", "max2_chunk", "max4_batch", "max4_chunk", "sum2", "sum4", "valid_lanes", "union2",
             "bwd_survivors", "bwd_evals", "bwd_valid_pairs", "bwd_valid_4x4_halves", "bwd_evals_le2", "bwd_evals_le8"]
    d = dict(zip(names, v))
    d["calls"] = 4 + 1 + 1  # warm-up(3 -> max(3)) + timed + e2e ... informational only
    print(json.dumps(d))
