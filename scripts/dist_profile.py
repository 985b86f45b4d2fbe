"""torchrun target: wall-clock breakdown of ShardedFit.fit_points (synchronised between phases)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from commonnnclustering_b200 import _lib, datasets
from commonnnclustering_b200 import dist as cdist

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
ctx = _lib.Context(local, torch_allocator=True)
ctx.set_stream(torch.cuda.current_stream())
sf = cdist.ShardedFit(ctx)
n = 1_000_000
x = torch.from_numpy(datasets.md_like_3d(n)).cuda()
lab = torch.empty(n, dtype=torch.int64, device="cuda")
r2, c = 0.09 ** 2, 20
for _ in range(3):
    sf.fit_points(x, r2, c + 1, c + 2, lab, member_cutoff=10)
marks = {}
def tick(name, t0):
    torch.cuda.synchronize()
    marks[name] = marks.get(name, 0.0) + time.perf_counter() - t0
    return time.perf_counter()
reps = 10
for _ in range(reps):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    ctx.grid_begin(x, r2, "euclidean_r"); t = tick("grid_begin", t)
    bounds = ctx.grid_balance(world); b, e = cdist.block_of(bounds, rank); t = tick("balance", t)
    lo, hi = ctx.grid_halo(b, e); t = tick("halo", t)
    deg = torch.zeros(n + 1, dtype=torch.int32, device="cuda"); t = tick("zeros", t)
    ctx.grid_degrees(lo, hi, deg); t = tick("degrees(search)", t)
    row_ptr = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    torch.cumsum(deg[:-1], 0, dtype=torch.int64, out=row_ptr[1:])
    host = torch.stack([row_ptr[-1], deg.max().to(torch.int64)]).cpu().numpy(); t = tick("scan+readback", t)
    nnz, max_deg = int(host[0]), int(host[1])
    col_idx = torch.empty(max(nnz, 1), dtype=torch.int32, device="cuda")[:nnz]
    ctx.grid_fill(lo, hi, row_ptr, col_idx); ctx.neighbours_adopt(row_ptr, col_idx, max_deg, grid_perm=True); t = tick("fill", t)
    ctx.components_begin(c + 1, c + 2); ctx.components_edges(b, e); t = tick("components", t)
    parents = cdist.tensor_from_ptr(ctx.components_parents_ptr(), (n,), "<i4", torch.device("cuda", local)); t = tick("tensor_from_ptr", t)
    gathered = cdist.gather_parents(parents); t = tick("all_gather", t)
    ctx.components_merge(gathered, world); t = tick("merge", t)
    lab.zero_(); ncl = ctx.components_labels(1, lab, fresh=True); t = tick("labels", t)
    ctx.sort_by_size(lab, 10, None); t = tick("sort_by_size", t)
if rank == 0:
    print({k: round(1e3 * v / reps, 3) for k, v in marks.items()}, "sum ms", round(1e3 * sum(marks.values()) / reps, 3))
dist.destroy_process_group()
