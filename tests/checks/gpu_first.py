"""First contact with the GPU: parity on small sets, then C2 timing by stage."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from commonnnclustering_b200 import _lib, datasets
from oracle import port

ctx = _lib.Context(0)

def gpu_fit(x, r, c, member_cutoff=None, metric="euclidean_r", method="auto"):
    lab = np.empty(x.shape[0], dtype=np.int64)
    rc = r * r if metric == "euclidean_r" else r
    ctx.fit_points(x, rc, c + 1, c + 2, lab, metric=metric, method=method, member_cutoff=member_cutoff)
    return lab

ok = True
x = datasets.README_POINTS
print("README", gpu_fit(x, 1.5, 1), flush=True)
for name, x in [("blobs3000", datasets.blobs_2d(3000)), ("md20000", datasets.md_like_3d(20000))]:
    for r, c in [(0.3, 10), (0.1, 20), (0.05, 5), (0.19, 37)] if x.shape[1] == 2 else [(0.4, 5), (0.25, 10), (0.6, 20)]:
        want = port.fit_points(x, r, c)
        for method in ("grid", "tiled"):
            got = gpu_fit(x, r, c, method=method)
            same = np.array_equal(want, got)
            ok &= same
            print(name, r, c, method, "OK" if same else "MISMATCH", int(want.max()), int(got.max()), ctx.stats(), flush=True)
        rp, ci = port.neighbours(x, r * r)
        ctx.neighbours_from_points(x, r * r, method="grid")
        rp2, ci2 = ctx.neighbours_export()
        print("  csr equal:", np.array_equal(rp, rp2) and np.array_equal(ci, ci2), flush=True)
x5 = np.random.default_rng(1).standard_normal((2000, 5))
want = port.fit_points(x5, 1.0, 5); got = gpu_fit(x5, 1.0, 5)
print("5d", np.array_equal(want, got)); ok &= np.array_equal(want, got)

import torch
n = int(os.environ.get("N", 1000000))
x = datasets.md_like_3d(n)
r, c = 0.09, 20
lab = np.empty(n, dtype=np.int64)
for it in range(3):
    t0 = time.perf_counter()
    ctx.fit_points(x, r * r, c + 1, c + 2, lab, member_cutoff=10)
    t1 = time.perf_counter()
    print("C2 e2e pageable", t1 - t0, "clusters", lab.max(), (lab == 0).mean(), ctx.neighbours_info(), ctx.stats(), flush=True)
xd = torch.from_numpy(x).cuda()
labd = torch.empty(n, dtype=torch.int64, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
for it in range(3):
    torch.cuda.synchronize()
    ev[0].record()
    ctx.neighbours_from_points(xd, r * r)
    ev[1].record()
    ctx.components_begin(c + 1, c + 2)
    ctx.components_edges(0, n)
    ev[2].record()
    labd.zero_()
    ctx.components_labels(1, labd)
    ctx.sort_by_size(labd, 10)
    ev[3].record()
    torch.cuda.synchronize()
    print("C2 stages ms: search %.3f comp %.3f label %.3f" % (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])), flush=True)
np.save("gpurun_out/c2_labels.npy", lab)
if n <= 1000000:
    t0 = time.perf_counter()
    want = port.fit_points(x, r, c, member_cutoff=10)
    print("oracle port time", time.perf_counter() - t0, "equal", np.array_equal(want, lab), flush=True)
    ok &= np.array_equal(want, lab)
print("ALL OK" if ok else "FAILURES")
