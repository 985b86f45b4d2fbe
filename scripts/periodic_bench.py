"""Periodic reduced Euclidean metric on the C2 generator in a periodic box (bounds 12 x 12 x 12, the generator's extent):
cell list on the image-augmented set vs the all-pairs f64 kernel -- identical CSR -- and the time of a whole fit.

    python scripts/periodic_bench.py [N] > gpurun_out/periodic.json"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from commonnnclustering_b200 import _lib, datasets

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n_check = min(n, 300_000)
p = datasets.C2_PARAMS
r, c, mc = p["radius_cutoff"], p["cnn_cutoff"], p["member_cutoff"]
bounds = np.array([12.0, 12.0, 12.0])
x = datasets.md_like_3d(n)                       # coordinates in [-1, 11): negative ones included
ctx = _lib.Context(0)
ctx.set_periodic_bounds(bounds)

# identical neighbour sets from both methods (the all-pairs kernel is pinned to the oracle in tests/)
xs = np.ascontiguousarray(x[:n_check])
rs = r * (n / n_check) ** (1.0 / 3.0)            # same mean degree on the subset
t0 = time.perf_counter()
ctx.neighbours_from_points(xs, rs * rs, metric="euclidean_periodic_r", method="tiled")
t_tiled = time.perf_counter() - t0
rp_t, ci_t = ctx.neighbours_export()
t0 = time.perf_counter()
ctx.neighbours_from_points(xs, rs * rs, metric="euclidean_periodic_r", method="grid")
t_grid = time.perf_counter() - t0
st = ctx.stats()
rp_g, ci_g = ctx.neighbours_export()
same = bool(np.array_equal(rp_t, rp_g) and np.array_equal(ci_t, ci_g))
assert same, "cell list and all-pairs kernel disagree"
assert st["distance_evals"] < n_check * n_check // 8, "the cell list did not serve the request"
ctx.neighbours_from_points(xs, rs * rs, metric="euclidean_r", method="grid")
nnz_open = ctx.neighbours_info()[1]

# whole fit at full size, device-resident coordinates
xd = torch.from_numpy(x).cuda()
lab = torch.empty(n, dtype=torch.int64, device="cuda")


def fit(metric):
    return ctx.fit_points(xd, r * r, c + 1, c + 2, lab, metric=metric, member_cutoff=mc)


out = {}
for metric in ("euclidean_periodic_r", "euclidean_r"):
    for _ in range(3):
        fit(metric)
    ctx.set_timing(True)
    ctx.timings(reset=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    reps = 10
    ev[0].record()
    for _ in range(reps):
        k = fit(metric)
    ev[1].record()
    torch.cuda.synchronize()
    t, _ = ctx.timings(reset=True)
    ctx.set_timing(False)
    out[metric] = {"fit_ms": ev[0].elapsed_time(ev[1]) / reps, "n_clusters": int(k), "nnz": ctx.neighbours_info()[1],
                   "stage_ms": {s: v[0] / reps for s, v in t.items() if v[1]}}
print(json.dumps({
    "workload": f"C2 generator N={n} in a periodic box 12^3, r={r}, c={c}, member_cutoff={mc}",
    "check": {"n": n_check, "r": rs, "csr_equal_all_pairs_kernel": same, "nnz": int(rp_g[-1]), "nnz_without_periods": int(nnz_open),
              "all_pairs_s": t_tiled, "cell_list_s": t_grid, "filter_candidates": st["distance_evals"]},
    "periodic": out["euclidean_periodic_r"], "open_boundaries": out["euclidean_r"],
}))
