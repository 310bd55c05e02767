"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, sys

def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"][:64]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        tot.setdefault(name, [0, 0.0])
        tot[name][0] += 1
        tot[name][1] += v
    s = sum(v for _, v in tot.values())
    for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{v:10.1f} us {100 * v / s:5.1f}% x{n:3d}  {k}")
    print(f"total {s:.1f} us over {sum(n for n, _ in tot.values())} launches")

if __name__ == "__main__":
    main(sys.argv[1])
