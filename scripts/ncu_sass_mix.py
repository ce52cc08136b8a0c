#!/usr/bin/env python
"""Instruction mix + stall samples of one kernel from `ncu -i X.ncu-rep --page source --csv
--kernel-name regex:NAME` output (read on stdin or from a file)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
iS, iE, iSm = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
st = [(c, h.index(c)) for c in h if c.startswith("stall_") and "Not Issued" not in c]
ops, samp, stalls = collections.Counter(), collections.Counter(), collections.Counter()
tot = totS = n_static = 0
for r in rows[hi + 1:]:
    if len(r) < len(h) or r[0] == "Address":
        continue
    parts = r[iS].split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    f = op.split(".")
    op = f[0] + ("." + f[-1] if f[0] in ("LDS", "STS", "LDG", "STG") and len(f) > 1 else "")
    e, s = int(r[iE]), int(r[iSm])
    ops[op] += e
    samp[op] += s
    tot += e
    totS += s
    n_static += 1
    for n, i in st:
        stalls[n] += int(r[i])
print("warp instructions", tot, " samples", totS, " static SASS", n_static)
for op, e in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 24):
    print("%-10s exec %9d (%4.1f%%)  samples %6d (%4.1f%%)" % (op, e, 100 * e / tot, samp[op], 100 * samp[op] / max(totS, 1)))
print(", ".join("%s %.1f%%" % (n, 100 * v / max(totS, 1)) for n, v in stalls.most_common(8)))
