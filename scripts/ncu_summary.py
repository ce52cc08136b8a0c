#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small CSV for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<what>.csv

One row per profiled launch: duration, DRAM bytes read/written, DRAM and issue
utilisation, registers, occupancy and the top warp-stall reasons.  Runs on the
CPU box (ncu -i needs no GPU).
"""
import csv
import subprocess
import sys

KEEP = [
    ("Kernel Name", "kernel"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem_B"),
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "fmaheavy_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True,
                         stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    stall_cols = [h for h in head if h.startswith(STALL) and h.endswith("_per_issue_active.ratio")]
    out = csv.writer(sys.stdout)
    names = [n for k, n in KEEP if k in col]
    out.writerow(names + ["top_stalls(warps per issue)"])
    for r in data:
        vals = []
        for k, n in KEEP:
            if k not in col:
                continue
            v, u = r[col[k]], units[col[k]]
            if n == "kernel":
                v = v.split("(")[0][:60]
            elif u and n in ("time", "dram_read", "dram_write"):
                v = "%s %s" % (v, u)
            vals.append(v)
        st = []
        for h in stall_cols:
            try:
                st.append((float(r[col[h]]), h[len(STALL):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
        st.sort(reverse=True)
        vals.append(" ".join("%s=%.2f" % (n, v) for v, n in st[:4]))
        out.writerow(vals)


if __name__ == "__main__":
    main()
