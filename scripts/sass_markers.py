"""SASS evidence per kernel of libbds_b200.so (no GPU needed): counts of the mnemonics that prove the B200 mapping -
UBLKCP (1-D TMA bulk copy), SYNCS (mbarrier), FFMA2 / FMUL2 / FADD2 (packed fp32x2), REDG.*F32x4 (128-bit vector
reductions), ATOMS (shared atomics), MATCH / REDUX, MUFU - next to the ptxas register / spill table.

    python scripts/sass_markers.py > profiles/r02_sass_markers.txt
"""
import collections, glob, os, re, subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "bilateral_driving_b200", "libbds_b200.so")
MARKS = ["UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "FMUL2", "FADD2", "RED", "ATOMS", "ATOMG", "MATCH", "REDUX", "MUFU",
         "LDS", "STS", "LDG", "STG", "BAR", "UTCHMMA", "HMMA", "LDL", "STL"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    cur, counts, total = None, collections.OrderedDict(), {}
    for ln in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            counts[cur] = collections.Counter()
            total[cur] = 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Za-z0-9_.]+)", ln)
        if m and cur:
            op = m.group(1)
            total[cur] += 1
            for k in MARKS:
                if op.startswith(k):
                    key = k
                    if k == "RED" and "F32x4" in op:
                        key = "RED.F32x4"
                    elif k == "RED" and not op.startswith("REDUX"):
                        key = "RED(scalar)"
                    elif k == "RED":
                        continue
                    counts[cur][key] += 1
                    break
    print("# SASS mnemonic counts per kernel (cuobjdump -sass libbds_b200.so); static instruction counts, not executed")
    for k, c in counts.items():
        if total[k] < 40:
            continue
        print(f"{k}  [{total[k]} SASS instructions]")
        print("    " + "  ".join(f"{m}={n}" for m, n in sorted(c.items())))
    print("\n# ptxas -v: registers / spills / shared memory per kernel")
    for log in sorted(glob.glob(os.path.join(ROOT, "bilateral_driving_b200", "csrc", "*.ptxas.log"))):
        if "variant" in log:
            continue
        name = None
        for ln in open(log):
            m = re.search(r"Compiling entry function '(\S+)'", ln)
            if m:
                name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
            if m:
                spill = f"stack {m.group(1)} B, spill stores {m.group(2)} B, spill loads {m.group(3)} B"
            m = re.search(r"Used (\d+) registers.*?(?:, (\d+) bytes smem)?", ln)
            if m and name:
                smem = re.search(r"(\d+) bytes smem", ln)
                print(f"{name:70s} {m.group(1):>3s} regs, {spill}, static smem {smem.group(1) if smem else 0} B")
                name = None


if __name__ == "__main__":
    main()
