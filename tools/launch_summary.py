"""Summarise an ncu launch list (gpu__time_duration.sum CSV) for ONE train step (between two projection kernels).

    python tools/launch_summary.py launches.csv [marker-substring] > profiles/rN_launches_summary.txt
"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    marker = sys.argv[2] if len(sys.argv) > 2 else "project_forward_kernel"
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v, unit = float(row["Metric Value"].replace(",", "")), row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v if unit.startswith("us") else v * 1e3)
        rows.append((row["Kernel Name"][:100], v))
    idx = [i for i, (k, _) in enumerate(rows) if marker in k]
    if len(idx) < 2:
        print("marker kernel not found twice:", marker, "launches:", len(rows))
        return
    a, b = idx[0], idx[1]
    step = rows[a:b]
    tot = sum(v for _, v in step)
    agg, cnt = collections.defaultdict(float), collections.Counter()
    for k, v in step:
        agg[k] += v
        cnt[k] += 1
    print(f"# one train step = {b - a} launches, {tot:.1f} us of kernel time (ncu gpu__time_duration.sum, --clock-control none;")
    print("# cold-cache, serialised: compare SHARES, not absolutes)")
    print(f"{'us':>9} {'share':>6} {'n':>3}  kernel")
    for k, v in sorted(agg.items(), key=lambda x: -x[1]):
        print(f"{v:9.1f} {100 * v / tot:5.1f}% {cnt[k]:3d}  {k}")
    ours = sum(v for k, v in agg.items() if "b200::" in k or "cub::" in k)
    print(f"# libb200splat kernels (b200:: + cub::): {ours:.1f} us = {100 * ours / tot:.1f}% ; PyTorch glue + optimizer: {tot - ours:.1f} us")


if __name__ == "__main__":
    main()
