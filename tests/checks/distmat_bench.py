"""Distance-matrix input ("distances" recipe): the thresholding kernel against the HBM roofline, the whole fit from a
device-resident and from a host matrix, and the oracle's thresholding + BFS on the host beside it.

    python tests/checks/distmat_bench.py [N] > gpurun_out/distmat.json

The matrix holds the Euclidean distances of the C2 generator's first N points (N = 40000: 12.8 GB, larger than L2)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from commonnnclustering_b200 import _lib, datasets
from oracle import port

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
r, c, mc = 0.25, 10, 10
x = datasets.md_like_3d(n)
xd = torch.from_numpy(x).cuda()
dm = torch.cdist(xd, xd)                      # f64 [n, n] on the device (input preparation, not timed)
dm = torch.minimum(dm, dm.T).contiguous()     # exactly symmetric
torch.cuda.synchronize()
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6554.2) \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "..", "MEASURED_PEAKS.json")) else 6554.2

ctx = _lib.Context(0)
lab = torch.zeros(n, dtype=torch.int64, device="cuda")


def fit_device():
    ctx.neighbours_from_distances(dm, r)
    lab.zero_()
    ctx.fit(c + 1, c + 2, 1, lab)
    ctx.sort_by_size(lab, member_cutoff=mc)


for _ in range(3):
    fit_device()
ctx.set_timing(True)
ctx.timings(reset=True)
reps = 10
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(reps):
    fit_device()
ev[1].record()
torch.cuda.synchronize()
t, launches = ctx.timings(reset=True)
ctx.set_timing(False)
ms_fit = ev[0].elapsed_time(ev[1]) / reps
_, nnz, max_deg = ctx.neighbours_info()
ms_mask = t["search_count"][0] / t["search_count"][1]
ms_fill = t["search_fill"][0] / t["search_fill"][1]
bytes_mask = 8 * n * n + 4 * n * ((n + 31) // 32) + 4 * n
labels_dev = lab.cpu().numpy()

# from a host matrix (pageable numpy array, as a user holds it): bound by the host link
dm_h = dm.cpu().numpy()
lab_h = np.zeros(n, dtype=np.int64)
ts = []
for _ in range(3):
    t0 = time.perf_counter()
    ctx.neighbours_from_distances(dm_h, r)
    lab_h[:] = 0
    ctx.fit(c + 1, c + 2, 1, lab_h)
    ctx.sort_by_size(lab_h, member_cutoff=mc)
    ts.append(time.perf_counter() - t0)
assert np.array_equal(lab_h, labels_dev)
pinned = _lib.host_register(dm_h)              # what InputDataComponents.pin() does for Clustering.fit
tp = []
for _ in range(3):
    t0 = time.perf_counter()
    ctx.neighbours_from_distances(dm_h, r)
    lab_h[:] = 0
    ctx.fit(c + 1, c + 2, 1, lab_h)
    ctx.sort_by_size(lab_h, member_cutoff=mc)
    tp.append(time.perf_counter() - t0)
assert np.array_equal(lab_h, labels_dev)

# CPU: the oracle's thresholding (numpy) + BFS (C restatement), one core
t0 = time.perf_counter()
rp, ci = port.neighbours_from_distances(dm_h, r)
t_thr = time.perf_counter() - t0
t0 = time.perf_counter()
want = port.fit_csr(rp, ci, c, member_cutoff=mc)
t_bfs = time.perf_counter() - t0
assert np.array_equal(want, labels_dev), "labels differ from the oracle"
g_rp, g_ci = ctx.neighbours_export()
assert np.array_equal(g_rp, rp) and np.array_equal(g_ci, ci), "neighbour sets differ from the oracle"

print(json.dumps({
    "workload": f"distance matrix N={n} f64 ({8 * n * n / 1e9:.1f} GB), r={r}, c={c}, member_cutoff={mc}; nnz={nnz}, max degree {max_deg}",
    "fit_ms_device_matrix": ms_fit, "points_per_s_device_matrix": n / (ms_fit * 1e-3),
    "fit_ms_host_matrix": 1e3 * min(ts), "host_matrix_gb_per_s": 8 * n * n / min(ts) / 1e9,
    "fit_ms_host_matrix_page_locked": 1e3 * min(tp), "page_locked": bool(pinned),
    "host_matrix_page_locked_gb_per_s": 8 * n * n / min(tp) / 1e9,
    "roofline": {"kernel": "dist_mask_kernel", "bound": "hbm", "algorithmic_bytes": bytes_mask, "ms_per_launch": ms_mask,
                 "achieved": bytes_mask / (ms_mask * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                 "frac": bytes_mask / (ms_mask * 1e-3) / 1e9 / peak,
                 "note": "8*N^2 matrix in + N^2/8 mask out + 4*N degrees"},
    "dist_fill_kernel_ms": ms_fill,
    "stage_ms": {k: v[0] / reps for k, v in t.items() if v[1]},
    "cpu_baseline": {"kind": "port", "cores": 1, "threshold_s": t_thr, "bfs_and_sort_s": t_bfs,
                     "points_per_s": n / (t_thr + t_bfs)},
    "labels_equal_oracle": True, "n_clusters": int(labels_dev.max()),
}))
