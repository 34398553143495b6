#!/usr/bin/env python
"""Top stalled SASS instructions from `ncu -i X.ncu-rep --page source --csv` output (stdin or file)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        s = int(r[ci["# Samples"]])
    except ValueError:
        continue
    top = sorted(((int(r[ci[k]] or 0), k) for k in stalls), reverse=True)[:2]
    data.append((s, r[ci["Source"]][:90], int(r[ci["Instructions Executed"]] or 0), top,
                 r[ci["L1 Conflicts Shared N-Way"]], r[ci["L2 Theoretical Sectors Global Excessive"]]))
tot = sum(d[0] for d in data) or 1
print("total samples", tot, "instructions", len(data))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for s, src, ex, top, conf, exc in sorted(data, key=lambda d: -d[0])[:n]:
    print("%6d %5.1f%% ex=%9d %-26s %s" % (s, 100.0 * s / tot, ex, ",".join("%s:%d" % (k[6:], v) for v, k in top if v), src))
