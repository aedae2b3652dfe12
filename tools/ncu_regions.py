"""Group the SASS of an ncu report into regions of equal execution count (poor man's hot-spot table).

    python tools/ncu_regions.py report.ncu-rep [min_share_percent]
"""
import csv
import io
import subprocess
import sys
from collections import Counter


def main():
    rep = sys.argv[1]
    thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[1]
    ia, isrc, ith, ism = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Avg. Threads Executed"), hdr.index("# Samples")
    data = [r for r in rows[2:] if len(r) > ia and r[ia].isdigit()]
    # the page repeats the kernel once per (launch x view); keep the first copy
    first = data[0][isrc]
    for j in range(1, len(data)):
        if data[j][isrc] == first and data[j][ia] == data[0][ia] and j > 50:
            data = data[:j]
            break
    tot = sum(int(r[ia]) for r in data)
    print(f"# {rep}: {len(data)} SASS instructions, {tot} warp-instructions executed")
    groups, cur = [], None
    for i, r in enumerate(data):
        c = int(r[ia])
        op = r[isrc].split()[1] if r[isrc].strip().startswith("@") else r[isrc].split()[0]
        if cur and abs(c - cur["c"]) <= 0.03 * max(cur["c"], 1):
            cur["n"] += 1; cur["inst"] += c; cur["ops"].append(op); cur["end"] = i; cur["thr"] += float(r[ith]); cur["smp"] += int(r[ism])
        else:
            cur = dict(c=c, n=1, inst=c, ops=[op], start=i, end=i, thr=float(r[ith]), smp=int(r[ism]))
            groups.append(cur)
    for g in groups:
        if 100.0 * g["inst"] / tot >= thr:
            oc = ", ".join(f"{k}x{v}" for k, v in Counter(o.split(".")[0] for o in g["ops"]).most_common(7))
            print(f"[{g['start']:4d}-{g['end']:4d}] n={g['n']:3d} exec/instr={g['c']:>9d} share={100 * g['inst'] / tot:5.1f}% "
                  f"lanes={g['thr'] / g['n']:4.1f}  {oc}")


if __name__ == "__main__":
    main()
