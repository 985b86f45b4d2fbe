"""Row-sharded (multi-GPU) path.

CPU part (gloo, world_size 2): the sharding scheme and the collective plumbing of
``commonnnclustering_b200.dist`` -- degree exchange, variable-size CSR slice exchange, parent
all-gather + min-hooking merge -- reproduce the oracle's labels when each rank only settles the
edges of its own row block.  The per-rank compute is stood in for by NumPy here (no GPU); the GPU
part runs the real thing under torchrun when two devices are visible."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find(parent, x):
    while parent[x] != x:
        parent[x] = parent[parent[x]]
        x = parent[x]
    return x


def _union_min(parent, a, b):
    ra, rb = _find(parent, a), _find(parent, b)
    if ra != rb:
        parent[max(ra, rb)] = min(ra, rb)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        from commonnnclustering_b200 import datasets
        from commonnnclustering_b200 import dist as cdist
        from oracle import port as oracle

        x = datasets.md_like_3d(900)
        r, c = 0.45, 4
        rp, ci = oracle.neighbours(x, r * r)
        n = x.shape[0]
        bounds = np.array([0, 416, n], dtype=np.int64)       # two unequal row blocks
        b, e = cdist.block_of(bounds, rank)
        # exchange 1a: degrees (each rank knows only its block)
        deg = torch.zeros(n + 1, dtype=torch.int32)
        deg[b:e] = torch.from_numpy(np.diff(rp)[b:e].astype(np.int32))
        row_ptr = cdist.exchange_degrees(deg)
        assert np.array_equal(row_ptr.numpy(), rp)
        # exchange 1b: col_idx slices of different sizes
        sb = rp[bounds]
        col = torch.full((int(rp[-1]),), -1, dtype=torch.int32)
        col[sb[rank]:sb[rank + 1]] = torch.from_numpy(ci[sb[rank]:sb[rank + 1]].astype(np.int32))
        cdist.exchange_slices(col, sb)
        assert np.array_equal(col.numpy(), ci)
        # exchange 1, equal-chunk form (the precomputed-CSR path): flat chunks that ignore row boundaries, gathered in place
        for arr in (rp, ci.astype(np.int32)):
            chunk = cdist.chunk_of(arr.shape[0], world)
            full = torch.full((world * chunk,), -9, dtype=torch.from_numpy(arr).dtype)
            a, b2 = rank * chunk, min((rank + 1) * chunk, arr.shape[0])
            full[a:b2] = torch.from_numpy(arr[a:b2])
            cdist.allgather_chunks(full, chunk)
            assert np.array_equal(full[:arr.shape[0]].numpy(), arr)
        # per-rank components on the own row block only (NumPy stand-in for K5/K6)
        core = np.diff(rp) > c + 1
        parent = np.where(core, np.arange(n), -1).astype(np.int32)
        for i in range(b, e):
            if not core[i]:
                continue
            ni = ci[rp[i]:rp[i + 1]]
            for j in ni[ni > i]:
                if core[j] and oracle.check(ni, ci[rp[j]:rp[j + 1]], c + 2):
                    _union_min(parent, i, int(j))
        # exchange 2: parents, then the merge every rank performs
        allp = cdist.gather_parents(torch.from_numpy(parent)).numpy()
        assert allp.shape == (world, n)
        for g in range(world):
            for i in np.flatnonzero(allp[g] >= 0):
                _union_min(parent, int(i), int(allp[g][i]))
        roots = np.array([_find(parent, i) if core[i] else -1 for i in range(n)])
        # numbering by ascending smallest member (= the root, hooking goes to the minimum)
        uniq = np.unique(roots[roots >= 0])
        labels = np.zeros(n, dtype=np.int64)
        labels[roots >= 0] = np.searchsorted(uniq, roots[roots >= 0]) + 1
        want = oracle.fit_points(x, r, c, sort=False)
        assert np.array_equal(labels, want)
        # the windowed form of exchange 2 (dist.plan_windows; cnnb200_fit_sharded_ranges / _end_windows on the device): a
        # rank that built only its block and the rows its lists refer to knows nothing outside that range
        xs = x[np.argsort(x[:, 0], kind="stable")]             # index locality, as after the cell sort
        rp, ci = oracle.neighbours(xs, r * r)
        core = np.diff(rp) > c + 1
        touched = np.unique(np.concatenate([np.arange(b, e), ci[rp[b]:rp[e]]]))
        lo, hi = int(touched.min()), int(touched.max()) + 1
        parent = np.full(n, -1, dtype=np.int32)
        parent[lo:hi] = np.where(core[lo:hi], np.arange(lo, hi), -1)
        for i in range(b, e):
            if not core[i]:
                continue
            ni = ci[rp[i]:rp[i + 1]]
            for j in ni[ni > i]:
                if core[j] and oracle.check(ni, ci[rp[j]:rp[j + 1]], c + 2):
                    _union_min(parent, i, int(j))
        mine = torch.tensor([lo, hi], dtype=torch.int64)
        ranges = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(ranges, mine)
        ranges = torch.stack(ranges).numpy()
        length, starts = cdist.plan_windows(ranges, n)
        assert length < n and np.all(starts <= ranges[:, 0]) and np.all(starts + length >= ranges[:, 1])
        wins = [torch.zeros(length, dtype=torch.int32) for _ in range(world)]
        dist.all_gather(wins, torch.from_numpy(parent[starts[rank]:starts[rank] + length].copy()))
        for g in range(world):                                  # rows another rank knows to be core become singleton trees
            w = wins[g].numpy()
            idx = starts[g] + np.flatnonzero(w >= 0)
            fresh = idx[parent[idx] < 0]
            parent[fresh] = fresh
        for g in range(world):
            w = wins[g].numpy()
            for k in np.flatnonzero(w >= 0):
                _union_min(parent, int(starts[g] + k), int(w[k]))
        roots = np.array([_find(parent, i) if parent[i] >= 0 else -1 for i in range(n)])
        uniq = np.unique(roots[roots >= 0])
        labels = np.zeros(n, dtype=np.int64)
        labels[roots >= 0] = np.searchsorted(uniq, roots[roots >= 0]) + 1
        assert np.array_equal(labels, oracle.fit_points(xs, r, c, sort=False))
        out[rank] = True
    finally:
        dist.destroy_process_group()


def test_sharded_plumbing_gloo_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out.get(0) and out.get(1)


@pytest.mark.gpu
def test_sharded_fit_two_gpus_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = os.path.join(ROOT, "tests", "checks", "dist_check.py")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29777", script],
                         capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "DIST OK" in res.stdout
