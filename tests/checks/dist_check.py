"""torchrun target: row-sharded fit on all visible GPUs must equal the single-GPU fit and the oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import torch.distributed as dist

from commonnnclustering_b200 import _lib, datasets
from commonnnclustering_b200 import dist as cdist

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
ctx = _lib.Context(local, torch_allocator=True)
ctx.set_stream(torch.cuda.current_stream())
sf = cdist.ShardedFit(ctx)
ok = True
for n, r, c, mc in [(20000, 0.25, 10, None), (300000, 0.13, 20, 10)]:
    x = datasets.md_like_3d(n)
    lab = np.empty(n, dtype=np.int64)
    sf.fit_points(x, r * r, c + 1, c + 2, lab, member_cutoff=mc)
    # the windowed exchange of the parent arrays (cnnb200_fit_sharded_ranges / _end_windows)
    lab_w = np.empty(n, dtype=np.int64)
    sf.windows = True
    sf.fit_points(x, r * r, c + 1, c + 2, lab_w, member_cutoff=mc)
    sf.windows = False
    assert np.array_equal(lab_w, lab), "windowed exchange differs"
    # device-resident coordinates and labels; labels only on rank 0 (the other ranks take part in the merge only)
    xd = torch.from_numpy(x).cuda()
    lab_d = torch.full((n,), -7, dtype=torch.int64, device="cuda") if rank == 0 else None
    sf.fit_points(xd, r * r, c + 1, c + 2, lab_d, member_cutoff=mc)
    # unsorted numbering (FitterExtBFS's own labels)
    lab_u = np.empty(n, dtype=np.int64)
    sf.fit_points(x, r * r, c + 1, c + 2, lab_u, sort_by_size=False)
    single_u = np.empty(n, dtype=np.int64)
    c0 = _lib.Context(local)
    c0.neighbours_from_points(x, r * r)
    c0.fit(c + 1, c + 2, 1, single_u, None)
    single = np.empty(n, dtype=np.int64)
    _lib.Context(local).fit_points(x, r * r, c + 1, c + 2, single, member_cutoff=mc)
    same = np.array_equal(lab, single) and np.array_equal(lab_u, single_u)
    if rank == 0:
        same &= np.array_equal(lab_d.cpu().numpy(), single)
    lab_x = np.empty(n, dtype=np.int64)
    sf.fit_points(x, r * r, c + 1, c + 2, lab_x, member_cutoff=mc, exchange=True)   # CSR all-gather layout
    same &= np.array_equal(lab_x, single)
    if n <= 20000:
        from oracle import port
        same &= np.array_equal(lab, port.fit_points(x, r, c, member_cutoff=mc))
    # precomputed-CSR route: every rank uploads only its rows
    rp, ci = _lib.Context(local), None
    c1 = _lib.Context(local)
    c1.neighbours_from_points(x, r * r)
    rp, ci = c1.neighbours_export()
    bounds = sf.neighbours_from_csr_host(rp, ci)
    lab2 = np.empty(n, dtype=np.int64)
    sf.fit(bounds, c + 1, c + 2, lab2, member_cutoff=mc)
    same &= np.array_equal(lab2, single)
    print(f"rank {rank}/{world} n={n}: sharded == single: {same}; clusters {int(lab.max())}", flush=True)
    ok &= bool(same)
t = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DIST OK" if t.item() else "DIST FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if t.item() else 1)
