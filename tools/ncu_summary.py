#!/usr/bin/env python3
"""Key metrics of every kernel in an `ncu --set full` report, as JSON (for profiles/).

  ncu -i rep.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv "note" > profiles/xyz_summary.json
"""
import csv, json, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
KEEP = [r"^Kernel Name$", r"^gpu__time_duration\.sum$", r"^launch__grid_size$", r"^launch__block_size$", r"^launch__registers_per_thread$",
        r"^launch__occupancy_limit_(registers|shared_mem|warps)$", r"^launch__waves_per_multiprocessor$", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
        r"^smsp__warps_active\.avg\.per_cycle_active$", r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$", r"^smsp__inst_executed\.sum$",
        r"^sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_active$", r"^sm__inst_executed_pipe_tensor_subpipe_dmma\.avg\.pct_of_peak_sustained_active$",
        r"^sm__inst_executed_pipe_lsu\.avg\.pct_of_peak_sustained_active$", r"^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum(\.pct_of_peak_sustained_elapsed)?$",
        r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$", r"^dram__bytes_(read|write)\.sum$", r"^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$",
        r"^sm__icc_request_hit_rate\.pct$", r"^smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$", r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$"]
out = {"note": sys.argv[2] if len(sys.argv) > 2 else "", "kernels": []}
for r in rows[2:]:
    d = {}
    for i, h in enumerate(hdr):
        if any(re.search(k, h) for k in KEEP):
            v = r[i]
            try:
                v = float(v.replace(",", ""))
            except ValueError:
                pass
            if isinstance(v, float) and h.startswith("smsp__average_warps_issue_stalled") and v < 0.05:
                continue
            d[h + (f" [{units[i]}]" if units[i] else "")] = v
    out["kernels"].append(d)
json.dump(out, sys.stdout, indent=1)
print()
