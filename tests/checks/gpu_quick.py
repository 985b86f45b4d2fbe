"""Quick GPU check: parity on a few sets + C2 stage timings (for kernel iteration).
SWEEP_VAR=<CNNB200_* environment variable> VALUES=a,b,c runs the C2 fit once per value (fresh context each)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from commonnnclustering_b200 import _lib, datasets
from oracle import port

ctx = _lib.Context(0)
ok = True
for name, x, r, c in [("blobs3000", datasets.blobs_2d(3000), 0.1, 20), ("md20000", datasets.md_like_3d(20000), 0.25, 10),
                      ("c1", datasets.blobs_2d(10000), 0.19, 37), ("uni1d", np.random.default_rng(0).random((3000, 1)), 0.01, 5)]:
    want = port.fit_points(x, r, c)
    got = np.empty(x.shape[0], dtype=np.int64)
    ctx.fit_points(x, r * r, c + 1, c + 2, got)
    same = np.array_equal(want, got)
    ok &= same
    rp, ci = port.neighbours(x, r * r)
    ctx.neighbours_from_points(x, r * r)
    rp2, ci2 = ctx.neighbours_export()
    same2 = np.array_equal(rp, rp2) and np.array_equal(ci, ci2)
    ok &= same2
    print(name, "labels", same, "csr", same2, flush=True)
n = int(os.environ.get("N", 1000000))
x = datasets.md_like_3d(n)
r, c = 0.09, 20
xd = torch.from_numpy(x).cuda()
labd = torch.empty(n, dtype=torch.int64, device="cuda")
ref = np.load("gpurun_out/c2_labels.npy") if os.path.exists("gpurun_out/c2_labels.npy") and n == 1000000 else None
VAR = os.environ.get("SWEEP_VAR", "CNNB200_FILTER")
for value in os.environ.get("VALUES", "1").split(","):
    os.environ[VAR] = value
    ctx = _lib.Context(0)
    ctx.set_timing(True)
    for it in range(4):
        ctx.fit_points(xd, r * r, c + 1, c + 2, labd, member_cutoff=10)
        t, launches = ctx.timings()
        if it:
            print(VAR, value, {k: round(v[0], 3) for k, v in t.items() if v[1]}, "total", round(sum(v[0] for v in t.values()), 3), ctx.stats(), flush=True)
    if ref is None:
        ref = labd.cpu().numpy()
    else:
        same = np.array_equal(ref, labd.cpu().numpy()); ok &= same
        print("labels equal to the previous variant / stored run:", same, flush=True)
_, nnz, _ = ctx.neighbours_info()
cnt = torch.empty(nnz, dtype=torch.int32, device="cuda")
for it in range(3):
    _, ne = ctx.cn_count(c + 1, cnt)
    t, _l = ctx.timings()
    print("cn_count ms", round(t["cn_count"][0], 3), "pack ms", round(t["ingest"][0], 3), "edges", ne, flush=True)
lab = labd.cpu().numpy()
print("ALL OK" if ok else "FAILURES")
