"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals.
Usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for name, (c, t) in agg.items():
        print(f"| `{name}` | {c} | {t:.3f} | {100 * t / tot:.1f}% |")
    print(f"| **all** | {sum(a[0] for a in agg.values())} | {tot:.3f} | 100% |")


if __name__ == "__main__":
    main(sys.argv[1])
