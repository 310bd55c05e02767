"""Per-source-line instruction counts of one kernel: joins the SASS page of an .ncu-rep (per-instruction
`Instructions Executed`, in address order) with `nvdisasm -g` line info of the same cubin (needs ncu, cuobjdump,
nvdisasm on PATH; no GPU).

    python scripts/ncu_source_lines.py gpurun_out/r01_full.ncu-rep composite_bwd composite [top_n]

argv[2] = regex for the kernel name in the report, argv[3] = substring of the cubin (source file stem).
"""
import collections, csv, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "bilateral_driving_b200", "libbds_b200.so")


def sass_lines(cubin_stem, kernel_re):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if cubin_stem in f and f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    sections, cur, where = {}, None, ("?", 0)
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            cur = m.group(1)
            sections[cur] = []
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            chain = [(os.path.basename(m.group(1)), int(m.group(2)))]
            for f, l in re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3)):
                chain.append((os.path.basename(f), int(l)))
            where = chain
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur:
            sections[cur].append((int(m.group(1), 16), m.group(2).strip(), where))
    name = [k for k in sections if re.search(kernel_re, k)]
    assert name, f"no section matches {kernel_re}: {list(sections)[:8]}"
    # prefer the largest match (template instances)
    name.sort(key=lambda k: -len(sections[k]))
    return name[0], sections[name[0]]


def main():
    rep, kre, stem = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    sec, sass = sass_lines(stem, kre)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # several launches may match: blocks start with a "Kernel Name" row; take the last block whose length fits
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    blk = [b for b in blocks if len(b["rows"]) == len(sass)]
    assert blk, f"no report block with {len(sass)} instructions ({[(b['name'], len(b['rows'])) for b in blocks]})"
    b = blk[-1]
    hi = b["hdr"].index("Instructions Executed")
    si = b["hdr"].index("# Samples")
    per_line = collections.Counter()
    per_line_samples = collections.Counter()
    per_outer = collections.Counter()
    total = 0
    for (off, ins, chain), r in zip(sass, b["rows"]):
        n = int(r[hi] or 0)
        total += n
        per_line[chain[0]] += n
        per_line_samples[chain[0]] += int(r[si] or 0)
        per_outer[chain[-1]] += n
    print(f"kernel {b['name']}  ({sec}): {len(sass)} SASS instructions, {total} warp-instructions executed")
    print("warp-instructions  share   stall samples   source line")
    for (f, l), n in per_line.most_common(top):
        print(f"{n:14d} {100.0 * n / total:5.1f}%  samples {per_line_samples[(f, l)]:7d}  {f}:{l}")


if __name__ == "__main__":
    main()
