"""Extract what bench.py's roofline block takes from a profile: per kernel of an .ncu-rep (an `ncu --set full` capture
of `bench.py --steps 1 --warmup 1 --no-cpu-baseline`, default workload, N=1) the DRAM traffic and the executed
warp-instructions of ONE launch.  Writes profiles/ncu_traffic.json (committed; bench.py reads it - no constants pasted
into bench.py).

    python scripts/ncu_traffic.py gpurun_out/r02_full.ncu-rep profiles/r02_ncu_full_key_metrics.txt
"""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "second": 1e3, "nsecond": 1e-6}


def main(rep, source_note):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    res = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        short = name.split("(")[0].replace("void ", "").replace("bds::", "").strip()

        def val(metric, table):
            i = ix[metric]
            return float(r[i].replace(",", "")) * table.get(units[i], 1.0)
        res[short] = {
            "dram_bytes_read": val("dram__bytes_read.sum", UNIT),
            "dram_bytes_write": val("dram__bytes_write.sum", UNIT),
            "warp_instructions": float(r[ix["smsp__inst_executed.sum"]].replace(",", "")),
            "duration_ms_under_ncu": val("gpu__time_duration.sum", TIME),
            "registers_per_thread": int(float(r[ix["launch__registers_per_thread"]])),
        }
    doc = {"workload": "configs[2]: 2000000 synthetic Gaussians, 6 x 1920x1080, N=1 (bench.py defaults)",
           "source": source_note, "report": os.path.basename(rep), "kernels": res}
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    json.dump(doc, open(path, "w"), indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
