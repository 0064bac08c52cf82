"""Print the handful of ncu metrics we track from a .ncu-rep (development aid).
usage: python tools/ncu_summary.py report.ncu-rep [kernel-index]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2 + idx]
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
        "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]
for h, u, v in zip(hdr, units, vals):
    if h in WANT or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        try:
            if float(v) == 0.0:
                continue
        except ValueError:
            pass
        print(f"{h:90s} {u:12s} {v}")
