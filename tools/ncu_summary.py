#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU): the metrics B200_PROFILING.md names + stall reasons.
usage: tools/ncu_summary.py rep.ncu-rep [--source N]  (top-N source lines by stall samples)"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_blocks",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fp64.sum",
        "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "smsp__inst_executed_op_global_ld.sum",
        "sm__icc_requests_lookup_miss.sum", "sm__icc_requests.sum"]
for i, h in enumerate(hdr):
    if h in keys or ("issue_stalled" in h and "per_issue_active" in h):
        vals = [r[i] for r in data]
        if "issue_stalled" in h:
            try:
                if max(float(v) for v in vals) < 0.05: continue
            except ValueError: pass
        print(f"{h} [{units[i]}] {vals}")
if "--source" in sys.argv:
    n = int(sys.argv[sys.argv.index("--source") + 1])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    open("/tmp/ncu_source.csv", "w").write(src)
    print("source page written to /tmp/ncu_source.csv", len(src))
