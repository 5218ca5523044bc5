"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, mean, share)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in r:
        if len(row) <= vi:
            continue
        name = re.sub(r"\(.*", "", row[ki])
        name = re.sub(r"^.*::", "", name)
        t = float(row[vi].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row[ui].strip(), 1e-3)
        agg[name][0] += 1
        agg[name][1] += t * scale
    tot = sum(v[1] for v in agg.values())
    print("%-44s %7s %12s %10s %7s" % ("kernel", "count", "total_ms", "mean_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %7d %12.3f %10.1f %7.3f" % (k[:44], v[0], v[1] / 1e3, v[1] / v[0], v[1] / tot))
    print("%-44s %7d %12.3f" % ("TOTAL", sum(v[0] for v in agg.values()), tot / 1e3))


if __name__ == "__main__":
    main(sys.argv[1])
