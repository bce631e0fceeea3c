"""Top stall sites of one kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass [--launch-skip N --launch-count 1]`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
S, N, I = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data = [r for r in rows[hi + 1:] if len(r) > N and r[N].isdigit()]
tot = sum(int(r[N]) for r in data)
inst = sum(int(r[I]) for r in data if r[I].isdigit())
print(rows[0][1] if len(rows[0]) > 1 else "", "| samples", tot, "| warp instructions", inst)
top = sorted(enumerate(data), key=lambda t: -int(t[1][N]))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for idx, r in sorted(top):
    st = sorted(((int(r[c]), h[c][6:]) for c in stall_cols if r[c].isdigit() and int(r[c]) > 0), reverse=True)[:3]
    print("%5d %6.2f%% %9s  %-70s %s" % (idx, 100.0 * int(r[N]) / tot, r[I], r[S].strip()[:70], " ".join("%s=%d" % (b, a) for a, b in st)))
