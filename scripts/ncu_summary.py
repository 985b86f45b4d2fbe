"""Reads an `ncu --set full` report (gpurun_out/prof.ncu-rep) and writes
  profiles/<tag>_ncu_full.csv      one line per profiled launch with the metrics DESIGN.md quotes
  profiles/ncu_traffic.json        kernel -> DRAM bytes per launch (bench.py's `roofline.traffic`)
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep r1d
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
idx = [hdr.index(w) for w in want if w in hdr]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out_csv = os.path.join(ROOT, "profiles", f"{tag}_ncu_full.csv")
traffic = {}
with open(out_csv, "w", newline="") as fp:
    wr = csv.writer(fp)
    wr.writerow([hdr[i] for i in idx])
    wr.writerow([units[i] for i in idx])
    for r in rows[2:]:
        wr.writerow([r[i] for i in idx])
        name = r[hdr.index("Kernel Name")].split("(")[0].split("<")[0].replace("void ", "").strip()
        rd = float(r[hdr.index("dram__bytes_read.sum")]) * SCALE[units[hdr.index("dram__bytes_read.sum")]]
        wrb = float(r[hdr.index("dram__bytes_write.sum")]) * SCALE[units[hdr.index("dram__bytes_write.sum")]]
        t = traffic.setdefault(name, {"dram_bytes": 0.0, "launches": 0, "time_ms": 0.0})
        t["dram_bytes"] += rd + wrb
        t["launches"] += 1
        t["time_ms"] += float(r[hdr.index("gpu__time_duration.sum")])
for t in traffic.values():   # per launch
    t["dram_bytes"] = t["dram_bytes"] / t["launches"]
    t["time_ms"] = t["time_ms"] / t["launches"]
with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as fp:
    json.dump({"source": f"profiles/{tag}_ncu_full.csv (ncu --set full --clock-control none, scripts/ncu_c2.py: one C2 fit + one "
               "count-mode pass)", "kernels": traffic}, fp, indent=1)
print(out_csv, {k: round(v["dram_bytes"] / 1e6, 1) for k, v in traffic.items()})
