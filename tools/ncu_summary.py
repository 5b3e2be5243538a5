#!/usr/bin/env python
"""profiles/*_ncu_summary.txt from an .ncu-rep (run where ncu is installed; no GPU needed):
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep "header line" > profiles/name.txt"""
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.max.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.per_second",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__cluster_max_active", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpc__cycles_elapsed.max",
        "gpc__cycles_elapsed.max.per_second", "sm__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
rep = sys.argv[1]
for line in sys.argv[2:]:
    print(line)
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
for h, u, v in zip(hdr, units, vals):
    if h in KEEP:
        print(h, "[" + u + "]", v)
