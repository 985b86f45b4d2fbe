"""Profiling target for ncu: warm fits of C2 (MD-like 3-D, N = 1e6), then ONE fit and ONE count-mode
pass inside a cudaProfilerStart/Stop window (run with `ncu --profile-from-start off ...`).

    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:'grid_rows|compact_rows|comp_|cn_count' -o gpurun_out/prof python scripts/ncu_c2.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from commonnnclustering_b200 import _lib, datasets

n = int(os.environ.get("N", 1_000_000))
count_mode = os.environ.get("COUNT", "1") != "0"
p = datasets.C2_PARAMS
x = datasets.md_like_3d(n)
r2, c = p["radius_cutoff"] ** 2, p["cnn_cutoff"]
ctx = _lib.Context(0, torch_allocator=True)
ctx.set_stream(torch.cuda.current_stream())
xd = torch.from_numpy(x).cuda()
labd = torch.empty(n, dtype=torch.int64, device="cuda")
for _ in range(2):
    ctx.fit_points(xd, r2, c + 1, c + 2, labd, member_cutoff=p["member_cutoff"])
_, nnz, _ = ctx.neighbours_info()
counts = torch.empty(nnz, dtype=torch.int32, device="cuda")
if count_mode:
    ctx.cn_count(c + 1, counts)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ctx.fit_points(xd, r2, c + 1, c + 2, labd, member_cutoff=p["member_cutoff"])
if count_mode:
    ctx.cn_count(c + 1, counts)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one fit", "+ one cn_count" if count_mode else "", "nnz", nnz, "clusters", int(labd.max()))
