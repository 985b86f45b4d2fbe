"""Stage times of one sweep cell (components from counts + numbering + size sort) on C2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from commonnnclustering_b200 import _lib, datasets
n = 1000000
xd = torch.from_numpy(datasets.md_like_3d(n)).cuda()
ctx = _lib.Context(0)
nnz = ctx.neighbours_from_points(xd, 0.09 ** 2)
counts = torch.empty(nnz, dtype=torch.int32, device="cuda")
ctx.cn_count(3, counts)
lab = torch.empty(n, dtype=torch.int64, device="cuda")
ctx.set_timing(True)
for c in (2, 5, 10, 20, 30, 50, 20, 20):
    ctx.timings(reset=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.components_begin(c + 1, c + 2)
    ctx.components_from_counts(counts)
    ctx.components_labels(1, lab, fresh=True)
    ctx.sort_by_size(lab, 10, None)
    s = ctx.label_summary()
    torch.cuda.synchronize(); wall = time.perf_counter() - t0
    t, _ = ctx.timings(reset=True)
    print("c", c, "wall ms", round(wall * 1e3, 3), {k: round(v[0], 3) for k, v in t.items() if v[1]}, s, flush=True)
