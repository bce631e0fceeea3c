"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the sequence."""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    ki, vi, ui, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("Grid Size")
    out = []
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        out.append((r[ki], v, r[gi]))
    return out


if __name__ == "__main__":
    data = load(sys.argv[1])
    seq = len(sys.argv) > 2
    tot = sum(d[1] for d in data)
    print(len(data), "launches; total %.1f us" % tot)
    agg = collections.OrderedDict()
    for n, v, g in data:
        n = n[:70]
        agg.setdefault(n, [0, 0.0])
        agg[n][0] += 1
        agg[n][1] += v
    for n, (c, v) in sorted(agg.items(), key=lambda t: -t[1][1]):
        print("%10.1f us %5.1f%%  x%3d  %s" % (v, 100 * v / tot, c, n))
    if seq:
        for n, v, g in data:
            print("%9.1f %18s %s" % (v, g, n[:70]))
