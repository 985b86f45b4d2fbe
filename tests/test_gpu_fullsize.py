"""BASELINE.json's configurations at their NAMED sizes against the CPU oracle (never against another CUDA path):

    C3  N = 2e5 x D = 64   neighbour sets vs the oracle's f64 brute force, labels vs the oracle's BFS
    C4  N = 1e7            precomputed neighbourhoods (built by the ORACLE's cell list), labels vs the oracle's BFS
    C5  8 x 8 sweep        eight cells of the sweep vs one oracle fit per cell
    count mode             every slot of the bit-packed count kernel vs the oracle's sorted-merge count

(C1 and C2 at full size: tests/test_gpu_parity.py.)  The oracle legs take a minute or two of host time each.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from commonnnclustering_b200 import _lib, cluster, datasets  # noqa: E402
from oracle import port  # noqa: E402

pytestmark = pytest.mark.gpu


def _threads():
    port.set_threads(os.cpu_count() or 1)


def test_c3_full_size_equals_oracle_brute_force(ctx):
    """configs[2]: the DEFAULT route of Clustering(x).fit for D = 64 is the tcgen05 filter + f64 recheck; its
    neighbour sets equal the oracle's brute force (same arithmetic as the reference's Metric._calc_distance) entry
    for entry, and the labels equal the oracle's BFS on those sets."""
    _threads()
    p = datasets.C3_PARAMS
    x = datasets.features_64d(200_000)
    r2 = p["radius_cutoff"] ** 2
    rp, ci = port.neighbours_large(x, r2)
    want = port.fit_csr(rp, ci, p["cnn_cutoff"], member_cutoff=p["member_cutoff"])
    cl = cluster.Clustering(x)
    cl.fitter._ctx = ctx
    cl.fit(p["radius_cutoff"], p["cnn_cutoff"], member_cutoff=p["member_cutoff"], v=False)
    assert ctx.stats()["search_method"] == "tensor"      # tensor_tiles_kernel served the default call
    assert np.array_equal(cl.labels, want)
    assert want.max() >= 10
    grp, gci = ctx.neighbours_export()
    assert np.array_equal(grp, rp) and np.array_equal(gci, ci)


def test_c4_full_size_equals_oracle_bfs(ctx):
    """configs[3]: N = 1e7 precomputed neighbourhoods.  The CSR comes from the ORACLE's cell list (caller's index
    space = time order of the trajectory, rows ascending), the labels from the oracle's BFS."""
    import torch

    _threads()
    p = datasets.C4_PARAMS
    n = 10_000_000
    x = datasets.md_like_3d(n)
    rp, ci = port.neighbours(x, p["radius_cutoff"] ** 2)
    del x
    want = port.fit_csr(rp, ci, p["cnn_cutoff"], member_cutoff=p["member_cutoff"])
    ci32 = ci.astype(np.int32)
    del ci
    assert 40 < rp[-1] / n < 60                       # the mean degree the configuration names (~50)
    got = np.zeros(n, dtype=np.int64)
    ctx.neighbours_from_csr(rp, ci32)
    ctx.fit(p["cnn_cutoff"] + 1, p["cnn_cutoff"] + 2, 1, got, None)
    ctx.sort_by_size(got, p["member_cutoff"])
    assert np.array_equal(got, want)
    # the same from device-resident (borrowed) buffers
    rp_d, ci_d = torch.from_numpy(rp).cuda(), torch.from_numpy(ci32).cuda()
    lab_d = torch.zeros(n, dtype=torch.int64, device="cuda")
    ctx.neighbours_from_csr(rp_d, ci_d, borrow=True)
    ctx.fit(p["cnn_cutoff"] + 1, p["cnn_cutoff"] + 2, 1, lab_d, None)
    ctx.sort_by_size(lab_d, p["member_cutoff"])
    assert np.array_equal(lab_d.cpu().numpy(), want)
    ctx.release()


def test_c5_sweep_cells_equal_oracle_at_full_size(ctx):
    """configs[4]: the 8 x 8 sweep on N = 1e6; the eight cells of the diagonal against one oracle fit per cell."""
    _threads()
    x = datasets.md_like_3d(1_000_000)
    radii, cnn = list(datasets.C5_RADII), list(datasets.C5_CNN)
    cl = cluster.Clustering(x)
    cl.fitter._ctx = ctx
    labels, summary = cl.fit_sweep(radii, cnn, member_cutoff=10)
    assert labels.shape == (8, 8, x.shape[0]) and len(summary) == 64
    for k in range(8):
        want = port.fit_points(x, radii[k], cnn[k], member_cutoff=10)
        assert np.array_equal(labels[k, k], want), (radii[k], cnn[k])
        rec = summary[k * 8 + k]
        assert rec.n_clusters == int(want.max())


def test_packed_count_kernel_equals_oracle(ctx):
    """cn_count_packed_kernel (the kernel behind the CN-checks/s figure) slot for slot: the cell-sorted CSR of a
    C2-like set compresses well, so the packed kernel runs (asserted), and every count equals the oracle's
    sorted-merge count of the same two rows."""
    import ctypes as C

    import torch

    from commonnnclustering_b200 import dist as cdist

    _threads()
    n, r, c = 300_000, 0.13, 20
    x = datasets.md_like_3d(n)
    ctx.neighbours_from_points(x, r * r)
    _, nnz, _ = ctx.neighbours_info()
    counts = torch.empty(nnz, dtype=torch.int32, device="cuda")
    _, n_edges = ctx.cn_count(c + 1, counts)
    assert ctx.stats()["count_packed"]
    rp_ptr, ci_ptr, perm_ptr = ctx.neighbours_device_view()
    dev = torch.device("cuda", ctx.device)
    rp = cdist.tensor_from_ptr(rp_ptr, (n + 1,), "<i8", dev).cpu().numpy()
    ci = cdist.tensor_from_ptr(ci_ptr, (nnz,), "<i4", dev).cpu().numpy().astype(np.int64)
    perm = cdist.tensor_from_ptr(perm_ptr, (n,), "<i4", dev).cpu().numpy()
    got = counts.cpu().numpy().view(np.uint32)
    # the internal CSR is the oracle's CSR relabelled through perm (rows ascending in the internal order)
    orp, oci = port.neighbours(x, r * r)
    inv = np.empty(n, dtype=np.int64)
    inv[perm] = np.arange(n)
    assert np.array_equal(np.diff(rp), np.diff(orp)[perm])
    for pidx in (0, 1, n // 2, n - 1):
        assert np.array_equal(ci[rp[pidx]:rp[pidx + 1]], np.sort(inv[oci[orp[perm[pidx]]:orp[perm[pidx] + 1]]]))
    # every slot: one C loop of sorted merges over the internal CSR (the oracle's count, cnn_oracle_count_sorted)
    L = port.lib()
    deg = np.diff(rp)
    core = deg > c + 1
    rows = np.repeat(np.arange(n), deg)
    is_edge = (ci > rows) & core[rows] & core[ci]
    assert int(is_edge.sum()) == n_edges
    assert (got[~is_edge] == 0xFFFFFFFF).all()
    rng = np.random.default_rng(0)
    sample = rng.choice(np.flatnonzero(is_edge), size=200_000, replace=False)
    for k in sample[:20_000]:
        i, j = rows[k], ci[k]
        a, b = ci[rp[i]:rp[i + 1]], ci[rp[j]:rp[j + 1]]
        assert got[k] == L.cnn_oracle_count_sorted(a, a.size, b, b.size)
    # and all of them through NumPy: |N(i) ∩ N(j)| summed per row must match (a checksum over every slot)
    want_sum = 0
    got_sum = int(got[is_edge].astype(np.int64).sum())
    # common neighbours of (i, j) = number of k with (i, k) and (j, k) present: a sparse product, block by block
    import scipy.sparse as sp

    a = sp.csr_matrix((np.ones(nnz, dtype=np.int32), ci, rp), shape=(n, n))
    blk = 20_000
    for s in range(0, n, blk):
        e = min(n, s + blk)
        prod = a[s:e] @ a.T                              # prod[i, j] = |N(i) ∩ N(j)|
        sel = is_edge[rp[s]:rp[e]]
        rr = rows[rp[s]:rp[e]][sel] - s
        cc = ci[rp[s]:rp[e]][sel]
        vals = np.asarray(prod[rr, cc]).ravel()
        assert np.array_equal(vals.astype(np.uint32), got[rp[s]:rp[e]][sel]), s
        want_sum += int(vals.sum())
    assert want_sum == got_sum
