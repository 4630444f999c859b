#!/usr/bin/env python
"""Group an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel and print the markdown table kept under profiles/.

    python tools/launch_list.py gpurun_out/launches.csv [skip_first_n [count]] > profiles/rNN_launches_<what>.md
"""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    limit = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    n = 0
    for r in rows[1:]:
        if len(r) <= col["Metric Value"] or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        n += 1
        if n <= skip or n > skip + limit:
            continue
        unit = r[col["Metric Unit"]]
        v = float(r[col["Metric Value"]].replace(",", ""))
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        key = (r[col["Kernel Name"]][:90], r[col["Grid Size"]], r[col["Block Size"]])
        a = agg.setdefault(key, [0.0, 0])
        a[0] += us
        a[1] += 1
    total = sum(a[0] for a in agg.values())
    count = sum(a[1] for a in agg.values())
    print("| share | total us | launches | avg us | kernel | grid | block |")
    print("|---|---|---|---|---|---|---|")
    for (name, grid, block), (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"| {100 * us / total:.1f}% | {us:.1f} | {c} | {us / c:.1f} | `{name}` | {grid} | {block} |")
    print(f"\nTotal {total:.0f} us over {count} launches.")


if __name__ == "__main__":
    main()
