"""Per-stage device times of one workload's fit (CUDA events inside the library), for A/B runs of tuning knobs:
    python scripts/stage_times.py [c1|c2] [KEY=VALUE ...]     # each KEY=VALUE group (comma separated) is one variant
e.g. python scripts/stage_times.py c2 CNNB200_MASK_MINB=4 CNNB200_MASK_MINB=3
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(name):
    import numpy as np
    import torch

    from commonnnclustering_b200 import _lib, datasets

    if name == "c1":
        x, p = datasets.blobs_2d(10_000), datasets.C1_PARAMS
    elif name == "c3":
        x, p = datasets.features_64d(int(os.environ.get("N", 200_000))), datasets.C3_PARAMS
    elif name == "c4":
        x, p = datasets.md_like_3d(int(os.environ.get("N", 10_000_000))), datasets.C4_PARAMS
    else:
        x, p = datasets.md_like_3d(int(os.environ.get("N", 1_000_000))), datasets.C2_PARAMS
    n = x.shape[0]
    ctx = _lib.Context(0, torch_allocator=True)
    ctx.set_stream(torch.cuda.current_stream())
    xd = torch.from_numpy(x).cuda()
    lab = torch.empty(n, dtype=torch.int64, device="cuda")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    r2, c = p["radius_cutoff"] ** 2, p["cnn_cutoff"]
    if name == "c4":     # precomputed neighbourhoods in the caller's order, device resident
        ctx.neighbours_from_points(xd, r2)
        rp_d, ci_d = ctx.neighbours_export_device()
        ctx.release()

        class _Fit:
            def fit_points(self, *_a, **_k):
                ctx.neighbours_from_csr(rp_d, ci_d, borrow=True)
                ctx.fit_sorted(c + 1, c + 2, lab, member_cutoff=p["member_cutoff"])
        fitter = _Fit()
    else:
        fitter = ctx
    for _ in range(3):
        fitter.fit_points(xd, r2, c + 1, c + 2, lab, member_cutoff=p["member_cutoff"])
    torch.cuda.synchronize()
    ctx.set_timing(True)
    ctx.timings(reset=True)
    steps = 10
    evs = []
    for i in range(steps):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fitter.fit_points(xd, r2, c + 1, c + 2, lab, member_cutoff=p["member_cutoff"])
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    t, launches = ctx.timings(reset=True)
    total = sum(a.elapsed_time(b) for a, b in evs) / steps
    out = {k: round(v[0] / steps, 4) for k, v in t.items() if v[1]}
    out["total_ms"] = round(total, 4)
    out["launches"] = launches // steps
    out["stats"] = ctx.stats()
    import hashlib
    out["labels_sha1"] = hashlib.sha1(lab.cpu().numpy().tobytes()).hexdigest()[:12]
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2])
        sys.exit(0)
    name = sys.argv[1] if len(sys.argv) > 1 and "=" not in sys.argv[1] else "c2"
    variants = [a for a in sys.argv[1:] if "=" in a] or [""]
    for v in variants:
        env = dict(os.environ)
        for kv in filter(None, v.split(",")):
            k, val = kv.split("=", 1)
            env[k] = val
        r = subprocess.run([sys.executable, __file__, "--child", name], env=env, capture_output=True, text=True)
        print(f"[{v or 'default'}] {r.stdout.strip() or r.stderr[-2000:]}", flush=True)
