import os, sys
sys.path.insert(0, os.getcwd())
import torch
from commonnnclustering_b200 import _lib, datasets
x = datasets.features_64d(200000); xd = torch.from_numpy(x).cuda()
ctx = _lib.Context(0); ctx.set_timing(True)
for it in range(3):
    ctx.timings()
    try:
        ctx.neighbours_from_points(xd, 0.36, method="tensor")
    except Exception as e:
        print("err", e)
    t, _ = ctx.timings()
print(os.environ.get("CNNB200_TENSOR_DBG"), {k: round(v[0], 3) for k, v in t.items() if v[1]})
