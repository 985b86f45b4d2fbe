"""Reads an `ncu --set full` report of scripts/ncu_c3.py (the tensor-core neighbour search on C3) and writes
profiles/<tag>_ncu_c3_tensor.csv: one line per profiled launch with the tensor-pipe / instruction-pipe / memory
metrics DESIGN.md quotes.
usage: python scripts/ncu_c3_summary.py gpurun_out/<name>.ncu-rep <tag> ["comment"]
"""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], sys.argv[2]
comment = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.max.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
idx = [hdr.index(w) for w in want if w in hdr]
out = os.path.join(ROOT, "profiles", f"{tag}_ncu_c3_tensor.csv")
with open(out, "w", newline="") as fp:
    wr = csv.writer(fp)
    wr.writerow([f"# ncu --set full --clock-control none, N=200000 x D=64 (C3), scripts/ncu_c3.py. {comment}"])
    wr.writerow([hdr[i] for i in idx])
    wr.writerow([units[i] for i in idx])
    for r in rows[2:]:
        wr.writerow([r[i] for i in idx])
print(out)
