"""Where the host time of the C5 hierarchy step goes (isolate, child fits, reel): python scripts/hier_profile.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from commonnnclustering_b200 import _lib, cluster, datasets
x = datasets.md_like_3d(1_000_000)
ctx = _lib.Context(0, torch_allocator=True)
cl = cluster.Clustering(x); cl.fitter._ctx = ctx
cl.fit(0.1, 2, member_cutoff=10, v=False)
soft = cl.labels.copy()
for rep in range(3):
    t = [time.perf_counter()]
    cl.labels = soft.copy(); t.append(time.perf_counter())
    cl.isolate(); t.append(time.perf_counter())
    nfit = 0
    for child in cl.root._children.values():
        if child._input_data.n_points >= 1000:
            cl.fit(0.063, 20, member_cutoff=10, v=False, bundle=child); nfit += 1
    t.append(time.perf_counter())
    cl.reel(); t.append(time.perf_counter())
    print("labels %.1f ms, isolate %.1f ms (%d children), %d child fits %.1f ms, reel %.1f ms" % (
        1e3 * (t[1] - t[0]), 1e3 * (t[2] - t[1]), len(cl.root._children), nfit, 1e3 * (t[3] - t[2]), 1e3 * (t[4] - t[3])))
    import cProfile, pstats
    if rep == 2:
        cl.labels = soft.copy()
        pr = cProfile.Profile(); pr.enable(); cl.isolate(); pr.disable()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(8)
