"""Summarise an ncu launch list (gpu__time_duration.sum CSV) per kernel: count, total, average, share."""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void gb::", "").replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms total (cold-cache, serialised: compare shares)")
    print(f"{'kernel':58s} {'n':>5s} {'total ms':>10s} {'avg us':>9s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:58s} {v[0]:5d} {v[1] / 1e6:10.3f} {v[1] / v[0] / 1e3:9.1f} {100 * v[1] / tot:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
