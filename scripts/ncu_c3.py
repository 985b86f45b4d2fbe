"""Profiling target: tensor-core neighbour search on the C3 generator (N x 64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from commonnnclustering_b200 import _lib, datasets
n = int(os.environ.get("N", 50_000))
x = datasets.features_64d(n)
ctx = _lib.Context(0)
xd = torch.from_numpy(x).cuda()
ctx.neighbours_from_points(xd, 0.36, method="tensor")
torch.cuda.synchronize()
torch.cuda.profiler.start()
nnz = ctx.neighbours_from_points(xd, 0.36, method="tensor")
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("nnz", nnz, ctx.stats())
