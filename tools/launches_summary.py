"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name.

    python tools/launches_summary.py gpurun_out/launches.csv [first_id last_id] > profiles/rNN_launches.md

Per-launch times under ncu are cold-cache and serialised: read the SHARES, not the absolutes."""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("void ", "").replace("at::native::", "").replace("(anonymous namespace)::", "")
    return name[:110]


def main():
    path = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        i = int(r["ID"])
        if lo <= i <= hi:
            rows.append((i, short(r["Kernel Name"]), float(r["Metric Value"]) / 1e3, r["Grid Size"], r["Block Size"]))
    agg = OrderedDict()
    for _, n, us, g, b in rows:
        a = agg.setdefault(n, [0, 0.0, g, b])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    OURS = ("vqa::", "sb::", "gm::")          # ncu drops the outer namespace of templated kernels
    ours = sum(a[1] for n, a in agg.items() if n.startswith(OURS))
    print(f"launches {len(rows)} (ids {rows[0][0]}..{rows[-1][0]}), total device time {tot:.1f} us; "
          f"kernels of libvqa_sm100.so {ours:.1f} us = {100 * ours / tot:.1f} %\n")
    print("| kernel | launches | total us | share | avg us | grid | block |")
    print("|---|---:|---:|---:|---:|---|---|")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f} % | {a[1] / a[0]:.1f} | {a[2]} | {a[3]} |")


if __name__ == "__main__":
    main()
