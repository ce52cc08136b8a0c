#!/usr/bin/env python
"""Per-region view of `ncu --page source --csv` output: where the executed warp
instructions and stall samples of one kernel sit (SASS order, `chunk` at a time).
    ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --launch-count 1 > k.csv
    python scripts/ncu_hot.py k.csv [chunk]
"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 60
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
end = next((i for i in range(h + 1, len(rows)) if rows[i] and rows[i][0] in ("Kernel Name", "Address")),
           len(rows))
hdr, data = rows[h], [r for r in rows[h + 1:end] if len(r) == len(rows[h])]
ci = {n: i for i, n in enumerate(hdr)}
I, S, T = ci["Instructions Executed"], ci["# Samples"], ci["Avg. Threads Executed"]
ti = sum(int(r[I]) for r in data)
ts = sum(int(r[S]) for r in data)
print("kernel:", rows[0][1][:90] if rows[0] else "?")
print("total warp insts %d, samples %d, SASS lines %d" % (ti, ts, len(data)))
for i in range(0, len(data), chunk):
    c = data[i:i + chunk]
    n = sum(int(r[I]) for r in c)
    s = sum(int(r[S]) for r in c)
    ops = Counter(r[ci["Source"]].split()[0].split(".")[0] if not r[ci["Source"]].strip().startswith("@")
                  else r[ci["Source"]].split()[1].split(".")[0] for r in c)
    thr = sum(float(r[T]) * int(r[I]) for r in c) / max(n, 1)
    print("%5d  insts %5.1f%%  samples %5.1f%%  thr %4.1f  %s" % (
        i, 100.0 * n / ti, 100.0 * s / max(ts, 1), thr,
        " ".join("%s:%d" % kv for kv in ops.most_common(5))))
