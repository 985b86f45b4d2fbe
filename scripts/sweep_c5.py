"""C5 (BASELINE configs[4]): 8x8 (radius_cutoff, cnn_cutoff) sweep on the C2 data (N = 1e6, 3-D), then isolate
the clusters of the softest cell, recluster every child with >= 1000 points and reel (SURVEY.md 8d "C5").
Prints one JSON line: sweep time (shared radius search + count mode per radius, components per cell) next
to 64 independent fits on the same GPU, with the labels of every cell compared between the two."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from commonnnclustering_b200 import _lib, cluster, datasets, sweep

n = int(os.environ.get("N", 1_000_000))
RS = [0.045, 0.050, 0.056, 0.063, 0.071, 0.079, 0.089, 0.100]
CS = [2, 5, 10, 15, 20, 30, 50, 80]
MC = 10
x = datasets.md_like_3d(n)
ctx = _lib.Context(0, torch_allocator=True)
xd = torch.from_numpy(x).cuda()
sweep.fit_sweep(xd, RS[:2], CS[:2], MC, ctx=ctx, return_labels=False)   # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
labels, summary = sweep.fit_sweep(xd, RS, CS, MC, ctx=ctx)
torch.cuda.synchronize()
t_sweep = time.perf_counter() - t0
t0 = time.perf_counter()
_, _ = sweep.fit_sweep(xd, RS, CS, MC, ctx=ctx, return_labels=False)
torch.cuda.synchronize()
t_sweep_nolabels = time.perf_counter() - t0
t0 = time.perf_counter()
_, _ = sweep.fit_sweep(xd, RS, CS, MC, ctx=ctx, labels_out=labels)     # caller's (page-locked) label array reused
torch.cuda.synchronize()
t_sweep_prealloc = time.perf_counter() - t0

# the same 64 cells as independent fits (radius search and early-exit checks repeated per cell)
single = torch.empty(n, dtype=torch.int64, device="cuda")
equal = 0
t_fits = 0.0
for ir, r in enumerate(RS):
    for ic, c in enumerate(CS):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.fit_points(xd, r * r, c + 1, c + 2, single, member_cutoff=MC)
        torch.cuda.synchronize()
        t_fits += time.perf_counter() - t0
        equal += int(np.array_equal(single.cpu().numpy(), labels[ir, ic]))

# hierarchy: isolate the softest cell, recluster the big children, reel
cl = cluster.Clustering(x)
cl.fitter._ctx = ctx
t0 = time.perf_counter()
cl.labels = labels[len(RS) - 1, 0].copy()
cl.isolate()
n_children = n_reclustered = 0
for label, child in cl.root._children.items():
    n_children += 1
    if child._input_data.n_points >= 1000:
        cl.fit(0.063, 20, member_cutoff=MC, v=False, bundle=child)
        n_reclustered += 1
cl.reel()
t_hier = time.perf_counter() - t0
print(json.dumps({
    "workload": f"C5: 8x8 sweep on MD-like 3-D N={n}, member_cutoff={MC}; isolate (r=0.100, c=2), recluster children >= 1000 "
                "points at (0.063, 20), reel",
    "cells": len(RS) * len(CS), "sweep_s": t_sweep, "sweep_s_without_label_copies": t_sweep_nolabels,
    "sweep_s_labels_into_preallocated_array": t_sweep_prealloc,
    "independent_fits_s": t_fits, "cells_equal_to_independent_fits": equal,
    "point_fits_per_s_sweep": len(RS) * len(CS) * n / t_sweep, "point_fits_per_s_independent": len(RS) * len(CS) * n / t_fits,
    "hierarchy_s": t_hier, "children": n_children, "children_reclustered": n_reclustered,
    "clusters_after_reel": int(cl.labels.max()),
    "n_clusters_grid": [[summary[i * len(CS) + j].n_clusters for j in range(len(CS))] for i in range(len(RS))],
}), flush=True)
