"""Empirical head-room of the tensor-core filter's error bound: the band eps is scaled down by
CNNB200_TENSOR_EPS_SCALE and the neighbour sets are compared with the exact f64 sets.  With the derived bound
(scale 1) the sets must be equal; the smallest scale that still gives equal sets says how far the real error of the
fp16-split MMA chain stays below the bound on this data.
    python scripts/tensor_eps_margin.py [n]        # child processes, one per scale
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(n):
    import numpy as np
    import torch

    from commonnnclustering_b200 import _lib, datasets

    out = {}
    ctx = _lib.Context(0)
    for name, x, r in (("c3_features_64d", datasets.features_64d(n), 0.6),
                       ("gauss_33d_offset50", np.random.default_rng(1).standard_normal((min(n, 20000), 33)) + 50.0, 7.0)):
        xd = torch.from_numpy(np.ascontiguousarray(x)).cuda()
        ctx.neighbours_from_points(xd, r * r, method="tiled")
        want = ctx.neighbours_export()
        ctx.neighbours_from_points(xd, r * r, method="tensor")
        got = ctx.neighbours_export()
        same = bool(np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]))
        diff = 0
        if not same:
            a = set(zip(np.repeat(np.arange(len(want[0]) - 1), np.diff(want[0])).tolist(), want[1].tolist()))
            b = set(zip(np.repeat(np.arange(len(got[0]) - 1), np.diff(got[0])).tolist(), got[1].tolist()))
            diff = len(a ^ b)
        out[name] = {"equal": same, "entries_differing": diff, "rechecks": ctx.stats()["rechecks"], "nnz": int(want[0][-1])}
    print(json.dumps(out))


if __name__ == "__main__":
    if os.environ.get("_EPS_CHILD"):
        child(int(sys.argv[1]))
    else:
        n = sys.argv[1] if len(sys.argv) > 1 else "50000"
        res = {}
        for scale in ("1", "0.3", "0.1", "0.03", "0.01", "0.003", "0.001", "0"):
            env = dict(os.environ, _EPS_CHILD="1", CNNB200_TENSOR_EPS_SCALE=scale)
            p = subprocess.run([sys.executable, os.path.abspath(__file__), n], env=env, capture_output=True, text=True)
            line = [l for l in p.stdout.splitlines() if l.startswith("{")]
            res[scale] = json.loads(line[-1]) if line else {"error": p.stderr[-300:]}
            print(scale, res[scale], flush=True)
        json.dump({"n": int(n), "by_eps_scale": res}, open(os.path.join(ROOT, "gpurun_out", "tensor_eps_margin.json"), "w"), indent=1)
