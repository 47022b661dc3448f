#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one line per kernel launch with the metrics DESIGN.md quotes."""
import csv, sys
COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst")]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def val(r, k):
    if k not in ix: return float("nan")
    try: v = float(r[ix[k]].replace(",", ""))
    except ValueError: return float("nan")
    u = units[ix[k]]
    if k.startswith("gpu__time"):
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
    if k.startswith("dram__bytes"):
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
    return v
print("%-28s" % "kernel" + "".join("%9s" % c[1] for c in COLS))
for r in rows[2:]:
    if len(r) < len(hdr): continue
    print("%-28s" % r[ix["Kernel Name"]][:27] + "".join("%9.1f" % val(r, c[0]) for c in COLS))
