"""torchrun target: per-rank wall-clock of the three phases of ShardedFit.fit_points (begin / all-gather / end),
synchronised between phases, plus the library's stage times -- shows load imbalance between the ranks."""
import os, sys, time, json
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
    sf.fit_points(x, r2, c + 1, c + 2, lab if rank == 0 else None, member_cutoff=10)
reps = 10
marks = {"begin": 0.0, "gather": 0.0, "end": 0.0, "whole": 0.0}
ctx.set_timing(True); ctx.timings()
for _ in range(reps):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    bounds, pptr = ctx.fit_sharded_begin(x, r2, c + 1, c + 2, rank, world, "euclidean_r")
    torch.cuda.synchronize(); t1 = time.perf_counter()
    parents = cdist.tensor_from_ptr(pptr, (n,), "<i4", torch.device("cuda", local))
    gathered = sf._gather_buffer(n)
    dist.all_gather_into_tensor(gathered, parents)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    ctx.fit_sharded_end(gathered, world, lab if rank == 0 else None, True, 1, 10, None)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    marks["begin"] += t1 - t0; marks["gather"] += t2 - t1; marks["end"] += t3 - t2
t, _ = ctx.timings()
for _ in range(reps):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    sf.fit_points(x, r2, c + 1, c + 2, lab if rank == 0 else None, member_cutoff=10)
    torch.cuda.synchronize(); marks["whole"] += time.perf_counter() - t0
out = {"rank": rank, "bounds": [int(b) for b in bounds], "ms": {k: round(1e3 * v / reps, 3) for k, v in marks.items()},
       "stage_ms": {k: round(v[0] / reps, 3) for k, v in t.items() if v[1]}}
for r in range(world):
    dist.barrier()
    if r == rank:
        print(json.dumps(out), flush=True)
dist.destroy_process_group()
