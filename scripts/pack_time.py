"""pack_rows_kernel timing on C2 (the packed rows are rebuilt after every neighbour search)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from commonnnclustering_b200 import _lib, datasets
n = 1000000
xd = torch.from_numpy(datasets.md_like_3d(n)).cuda()
ctx = _lib.Context(0)
ctx.neighbours_from_points(xd, 0.09 ** 2)
cnt = torch.empty(ctx.neighbours_info()[1], dtype=torch.int32, device="cuda")
ctx.set_timing(True)
for it in range(6):
    ctx.neighbours_from_points(xd, 0.09 ** 2)
    ctx.timings(reset=True)
    ctx.cn_count(21, cnt)
    t, _ = ctx.timings(reset=True)
    print("pack ms", round(t["ingest"][0], 3), "count ms", round(t["cn_count"][0], 3), flush=True)
