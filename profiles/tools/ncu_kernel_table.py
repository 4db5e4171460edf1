"""Condense `ncu --page raw --csv` output (one row per profiled launch, one column per metric) into a per-kernel table:
duration, DRAM bytes, DRAM / tensor-pipe / issue utilisation, occupancy, registers.  Usage:
    ncu -i prof.ncu-rep --page raw --csv > raw.csv ; python profiles/tools/ncu_kernel_table.py raw.csv
"""
import csv
import re
import sys
from collections import OrderedDict

WANT = OrderedDict([
    ("gpu__time_duration.sum", "us"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_%"),
    ("lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("sm__inst_executed.avg.per_cycle_active", "ipc"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
])


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


rows = list(csv.reader(open(sys.argv[1], newline="")))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, units = rows[hdr_i], rows[hdr_i + 1]
col = {h: i for i, h in enumerate(hdr)}
name_i = col["Kernel Name"]
print("%-58s %8s %10s %10s %7s %8s %6s %6s %6s %5s %6s %5s %6s %6s" % (
    "kernel", "us", "dram_rd_MB", "dram_wr_MB", "dram_%", "tensor_%", "sm_%", "smem_%", "l2_%", "ipc", "occ_%", "regs", "grid",
    "block"))
for r in rows[hdr_i + 2:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[name_i]).replace("void ", "").replace("mm::", "")[:58]
    v = {}
    for m, short in WANT.items():
        if m in col:
            x, u = num(r[col[m]]), units[col[m]]
            if x is None:
                continue
            if short == "us":
                x = x / 1e3 if u in ("ns", "nsecond") else (x * 1e3 if u in ("ms", "msecond") else x)
            if short.endswith("_MB"):
                x = x / 1e6 if u in ("byte", "bytes") else (x / 1e3 if u.startswith("K") else (x if u.startswith("M") else x * 1e3))
            v[short] = x
    f = lambda k, fmt: (fmt % v[k]) if k in v else "-"
    print("%-58s %8s %10s %10s %7s %8s %6s %6s %6s %5s %6s %5s %6s %6s" % (
        name, f("us", "%.1f"), f("dram_rd_MB", "%.1f"), f("dram_wr_MB", "%.1f"), f("dram_%", "%.1f"), f("tensor_%", "%.1f"),
        f("sm_%", "%.1f"), f("smem_%", "%.1f"), f("l2_%", "%.1f"), f("ipc", "%.2f"), f("occ_%", "%.1f"), f("regs", "%d"),
        f("grid", "%d"), f("block", "%d")))
