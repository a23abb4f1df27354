#!/usr/bin/env python
"""Condenses an .ncu-rep (ncu --set full) into the small JSON kept under profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/NAME_ncu_full_summary.json"""
import csv, io, json, subprocess, sys
KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'smsp__inst_executed.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'sm__inst_executed_pipe_fp64.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum']
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0]}
    for k in KEEP:
        if k in hdr:
            d[k] = f"{r[hdr.index(k)]} {units[hdr.index(k)]}".strip()
    stalls = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                stalls.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    d["stalls_per_issue"] = {n: v for v, n in sorted(stalls, reverse=True)[:8]}
    out.append(d)
json.dump(out, open(sys.argv[2], "w"), indent=1)
for d in out:
    print(d["kernel"], d.get("gpu__time_duration.sum"), "fp64 pipe", d.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
          "regs", d.get("launch__registers_per_thread"), "dram", d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum"))
    print("   ", d["stalls_per_issue"])
