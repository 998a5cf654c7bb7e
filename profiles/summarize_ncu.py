#!/usr/bin/env python
"""Key figures of one kernel from `ncu -i X.ncu-rep --page raw --csv` (stdin or file): duration, launch shape, pipe
utilisation, issue activity, DRAM traffic, stall mix.   python profiles/summarize_ncu.py raw.csv [kernel-substring]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
hdr = rows[0]
unit = rows[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
]
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    rec = dict(zip(hdr, r))
    name = rec.get("Kernel Name", "")
    if want and want not in name:
        continue
    print("kernel:", name[:120])
    for k in KEYS:
        if k in rec:
            print(f"  {k:75s} {rec[k]:>16s} {unit[hdr.index(k)]}")
    stalls = [(k, float(rec[k])) for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and rec[k]]
    print("  stalls per issue-active cycle:", ", ".join(f"{k.split('stalled_')[1].split('_per_')[0]} {v:.2f}" for k, v in sorted(stalls, key=lambda kv: -kv[1])[:8]))
