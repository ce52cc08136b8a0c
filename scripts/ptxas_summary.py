#!/usr/bin/env python
"""Registers, static shared memory, stack and spills of every kernel, from the `-Xptxas -v` logs
the Makefile leaves under pytorch-unsup-pc_b200/build/ (no GPU needed):
    python scripts/ptxas_summary.py [substring ...]  > profiles/rNN_ptxas_summary.txt"""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?Function properties for \S+\n"
                 r"\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                 r".*?Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes cumulative stack size)?"
                 r"(?:, (\d+) bytes smem)?", re.S)


def main():
    want = sys.argv[1:]
    rows = []
    for f in sorted(glob.glob(os.path.join(ROOT, "pytorch-unsup-pc_b200", "build", "*.ptxas.log"))):
        for m in PAT.finditer(open(f).read()):
            rows.append((os.path.basename(f).split(".")[0], m.group(1), int(m.group(5)), int(m.group(8) or 0),
                         int(m.group(2)), int(m.group(3)), int(m.group(4))))
    names = subprocess.run(["c++filt"] + [r[1] for r in rows], capture_output=True, text=True).stdout.split("\n")
    print("%-18s %-86s %5s %8s %6s %7s" % ("source", "kernel", "regs", "smem B", "stack", "spill B"))
    for name, r in zip(names, rows):
        short = re.sub(r"\(.*", "", name).replace("void ", "")
        if want and not any(w in short for w in want):
            continue
        print("%-18s %-86s %5d %8d %6d %7d" % (r[0], short[:86], r[2], r[3], r[4], r[5] + r[6]))
    print("# %d kernels, %d with spills" % (len(rows), sum(1 for r in rows if r[5] or r[6])))


if __name__ == "__main__":
    main()
