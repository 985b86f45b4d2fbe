"""Wall-clock parts of one C5 step (bench.py --workload c5): the 64-cell sweep, the softest cell's fit, isolate, the
child fits, reel."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from commonnnclustering_b200 import _lib, cluster, datasets, sweep

n = 1_000_000
x = datasets.md_like_3d(n)
xd = torch.from_numpy(x).cuda()
ctx = _lib.Context(0, torch_allocator=True)
ctx.set_stream(torch.cuda.current_stream())
radii, cnn = datasets.C5_RADII, datasets.C5_CNN
hcl = cluster.Clustering(x); hcl.fitter._ctx = ctx
soft = torch.empty(n, dtype=torch.int64, device="cuda")
def sync(): torch.cuda.synchronize()
for rep in range(4):
    t = {}
    sync(); t0 = time.perf_counter()
    sweep.fit_sweep(xd, radii, cnn, 10, ctx=ctx, return_labels=False); sync(); t["sweep64"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    ctx.fit_points(xd, radii[-1] ** 2, cnn[0] + 1, cnn[0] + 2, soft, member_cutoff=10); sync(); t["soft_fit"] = time.perf_counter() - t0
    t0 = time.perf_counter(); lab = soft.cpu().numpy(); t["d2h"] = time.perf_counter() - t0
    t0 = time.perf_counter(); hcl.labels = lab; t["set_labels"] = time.perf_counter() - t0
    t0 = time.perf_counter(); hcl.isolate(); t["isolate"] = time.perf_counter() - t0
    t0 = time.perf_counter(); k = 0
    for child in hcl.root._children.values():
        if child._input_data.n_points >= 1000:
            hcl.fit(0.063, 20, member_cutoff=10, v=False, bundle=child); k += 1
    t["child_fits"] = time.perf_counter() - t0
    t0 = time.perf_counter(); hcl.reel(); t["reel"] = time.perf_counter() - t0
    print({k_: round(1e3 * v, 2) for k_, v in t.items()}, "sum", round(1e3 * sum(t.values()), 1), "children fitted", k, flush=True)
# where the sweep's time goes
pr = cProfile.Profile(); pr.enable()
sweep.fit_sweep(xd, radii, cnn, 10, ctx=ctx, return_labels=False); sync()
pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(12)
pr = cProfile.Profile(); pr.enable()
for child in hcl.root._children.values():
    if child._input_data.n_points >= 1000:
        hcl.fit(0.063, 20, member_cutoff=10, v=False, bundle=child)
pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(10)
