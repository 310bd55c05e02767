"""Summarise a chrome trace written by bench.py (BDS_TIMELINE): kernels of the second step, idle gaps."""
import gzip, json, sys


def main(path, min_us=20.0):
    op = gzip.open if path.endswith(".gz") else open
    d = json.load(op(path))
    ev = [e for e in d["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
    ev.sort(key=lambda e: e["ts"])
    starts = [i for i, e in enumerate(ev) if "project_fwd_kernel" in e["name"]]
    a, b = starts[1], starts[2]
    t0, prev_end, gaps, busy = ev[a]["ts"], None, 0.0, 0.0
    for e in ev[a:b]:
        gap = (e["ts"] - prev_end) if prev_end is not None else 0.0
        gaps += max(gap, 0.0)
        busy += e["dur"]
        if e["dur"] > min_us or gap > min_us:
            print(f"{e['ts'] - t0:9.1f} +{e['dur']:8.1f} gap {gap:7.1f}  {e['name'][:72]}")
        prev_end = max(prev_end or 0, e["ts"] + e["dur"])
    print(f"step span {ev[b]['ts'] - t0:.1f} us, kernels {busy:.1f} us, idle {gaps:.1f} us, {b - a} launches")


if __name__ == "__main__":
    main(sys.argv[1])
