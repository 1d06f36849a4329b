"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares.
Usage: python tools/launch_shares.py launches.csv [top_n]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        n = row["Kernel Name"]
        m = re.search(r"(\w+)(<[^(]*>)?\(", n)
        n = (m.group(1) + (m.group(2) or "")) if m else n
        t = float(row["Metric Value"]) / 1e3
        agg[n[:90]][0] += 1
        agg[n[:90]][1] += t
        tot += t
    nl = sum(v[0] for v in agg.values())
    print(f"# {path}: {nl} launches, {tot / 1e3:.3f} ms of kernel time (ncu: cold cache, serialised)")
    print(f"{'us':>12s} {'share':>7s} {'n':>6s} {'avg us':>9s}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1]:12.1f} {100 * v[1] / tot:6.1f}% {v[0]:6d} {v[1] / v[0]:9.1f}  {k}")


if __name__ == "__main__":
    main()
