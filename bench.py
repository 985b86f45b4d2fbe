#!/usr/bin/env python
"""Benchmark of the fit hot path (BASELINE.json metric: fit() points/s and CN-checks/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c5]   # this repository's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W] [--workload ...]        # the reference's CPU path

Workloads (`config.workload`) are BASELINE.json's five configurations:
    c1  2-D blobs N = 1e4, r = 0.3, c = 10                       (configs[0], the reference's own CPU-runnable case)
    c2  MD-like 3-D trajectory N = 1e6, r = 0.09, c = 20          (configs[1]; the DEFAULT: the metric is quoted on it)
    c3  features N = 2e5 x D = 64, r = 0.6, c = 20                (configs[2]; tcgen05 filter + f64 recheck)
    c4  precomputed neighbourhoods N = 1e7, c = 10                (configs[3]; rows sharded, NCCL CSR all-gather)
    c5  8 x 8 (r, c) sweep on the C2 data                          (configs[4]; one step = all 64 cells)
One step = one complete pass of the hot path: neighbour search (or CSR ingest) -> common-neighbour checks +
union-find -> cluster numbering -> sort_by_size, labels int64[N] out.

Prints ONE JSON line (rank 0).  `value` = points/s with the inputs already resident in HBM (CUDA events, L2 flushed
between steps); `e2e.value` = the same through the C-ABI on pinned HOST buffers, H2D of the inputs and D2H of the
labels inside the timed region (wall clock) -- SURVEY.md 8d's definition of the metric.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fit_points_per_s"
UNIT = "points/s"


def workload(name, n_override=None):
    from commonnnclustering_b200 import datasets

    if name == "c2":
        n = n_override or 1_000_000
        p = datasets.C2_PARAMS
        return dict(key="c2", name=f"C2: MD-like 3-D trajectory N={n}, r={p['radius_cutoff']}, c={p['cnn_cutoff']}, "
                                   f"member_cutoff={p['member_cutoff']} (BASELINE configs[1])",
                    make=lambda: datasets.md_like_3d(n), n=n, d=3, r=p["radius_cutoff"], c=p["cnn_cutoff"],
                    member_cutoff=p["member_cutoff"])
    if name == "c1":
        p = datasets.C1_PARAMS
        return dict(key="c1", name="C1: 2-D blobs N=10000, r=0.3, c=10 (BASELINE configs[0])",
                    make=lambda: datasets.blobs_2d(10_000), n=10_000, d=2, r=p["radius_cutoff"], c=p["cnn_cutoff"],
                    member_cutoff=None)
    if name == "c3":
        n = n_override or 200_000
        p = datasets.C3_PARAMS
        return dict(key="c3", name=f"C3: features N={n} x D=64, r={p['radius_cutoff']}, c={p['cnn_cutoff']}, "
                                   f"member_cutoff={p['member_cutoff']} (BASELINE configs[2])",
                    make=lambda: datasets.features_64d(n), n=n, d=64, r=p["radius_cutoff"], c=p["cnn_cutoff"],
                    member_cutoff=p["member_cutoff"])
    if name == "c4":
        n = n_override or 10_000_000
        p = datasets.C4_PARAMS
        r = p["radius_cutoff"] * (10_000_000 / n) ** (1.0 / 3.0)   # reduced N (debugging): same mean degree
        return dict(key="c4", name=f"C4: precomputed neighbourhoods (sorted CSR, int32 indices) N={n}, mean degree ~50, "
                                   f"c={p['cnn_cutoff']}, member_cutoff={p['member_cutoff']} (BASELINE configs[3])",
                    make=lambda: datasets.md_like_3d(n), n=n, d=3, r=r, c=p["cnn_cutoff"], member_cutoff=p["member_cutoff"])
    if name == "c5":
        n = n_override or 1_000_000
        return dict(key="c5", name=f"C5: 8x8 (radius_cutoff, cnn_cutoff) sweep on MD-like 3-D N={n}, member_cutoff=10; "
                                   "one step = all 64 cells + isolate / child reclustering / reel of the softest cell (BASELINE configs[4])",
                    make=lambda: datasets.md_like_3d(n), n=n, d=3, radii=list(datasets.C5_RADII), cnn=list(datasets.C5_CNN),
                    r=0.063, c=20, member_cutoff=10)
    raise SystemExit(f"unknown workload {name}")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fp:
            d = json.load(fp)
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d.get("bf16_tflops", 1640.0)),
                    tf_sustained=float(d.get("bf16_tflops_sustained", 1343.0)), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1640.0, tf_sustained=1343.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def sha(labels):
    return hashlib.sha1(np.ascontiguousarray(labels, dtype=np.int64).tobytes()).hexdigest()[:16]


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own fit on the box's host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_setup(w, threads):
    """Neighbourhoods for the reference's `sorted_neighbourhoods` recipe (the only CPU route the reference's authors
    use beyond ~1e4 points; its brute-force default would need ~14 h on C2).  C4 / C5 run a bounded sample (see
    `sample`): C4 the same generator at N = 1e6 with the radius scaled to the same mean degree, C5 one cell of the
    grid per step."""
    from oracle import port, ref

    port.set_threads(threads)
    key = w["key"]
    n_sample = w["n"]
    r = w["r"]
    if key == "c4" and w["n"] > 1_000_000:
        n_sample = 1_000_000
        r = w["r"] * (w["n"] / n_sample) ** (1.0 / 3.0)
        from commonnnclustering_b200 import datasets
        x = datasets.md_like_3d(n_sample)
    else:
        x = w["make"]()
    t0 = time.perf_counter()
    if x.shape[1] <= 3:
        rp, ci = port.neighbours(x, r * r)
    else:
        rp, ci = port.neighbours_large(x, r * r)
    t_nb = time.perf_counter() - t0
    kind = "reference" if ref.available() else "port"
    state = dict(kind=kind, rp=rp, ci=ci, t_neighbours=t_nb, n=n_sample, r=r)
    if kind == "reference":
        mod = ref.module()  # puts oracle/_ref on sys.path
        from cnnclustering import _types as rt

        padded, counts = port.csr_to_padded(rp, ci)
        state["clustering"] = mod.Clustering(rt.InputDataExtNeighbourhoodsMemoryview(padded, counts),
                                             recipe="sorted_neighbourhoods")
    return state


def cpu_reference_step(w, state):
    """One fit of the reference on the prepared neighbourhoods; returns seconds of `fitter.fit`
    (the reference's Record.execution_time, cluster.pyx:332-337) and the labels."""
    if state["kind"] == "reference":
        cl = state["clustering"]
        cl.fit(state["r"], w["c"], member_cutoff=w["member_cutoff"], v=False, record=True, record_time=True, info=False)
        return cl.summary[-1].execution_time, np.asarray(cl.labels)
    from oracle import port

    t0 = time.perf_counter()
    lab = port.fit_csr(state["rp"], state["ci"], w["c"], member_cutoff=w["member_cutoff"])
    return time.perf_counter() - t0, lab


def reference_sample_text(w, state, threads):
    n = state["n"]
    what = {"c4": f"bounded sample: the same generator at N={n} with the radius scaled to the same mean degree "
                  f"(the padded int64 matrix of the full N=1e7 input does not fit a step of seconds)",
            "c5": f"bounded sample: ONE cell of the 8x8 grid per step (r={w['r']}, c={w['c']}) on the full N={n}; the "
                  "reference runs every cell as an independent fit, so points/s per cell = point-fits/s of the sweep"}.get(
        w["key"], f"full workload, N={n}")
    return (f"{what}; each step = the reference's FitterExtBFS.fit + sort_by_size on PRECOMPUTED sorted neighbourhoods "
            f"(sorted_neighbourhoods recipe); the neighbour search is NOT in the reference's step (done once outside by the "
            f"oracle's C search with {threads} threads in {state['t_neighbours']:.1f} s; the reference's own routes are brute "
            f"force ~5e-8*N^2 s or sklearn KDTree + a Python padding loop)")


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    w = workload(args.workload, args.n)
    threads = os.cpu_count() or 1
    state = cpu_reference_setup(w, threads)
    n = state["n"]
    times = []
    for i in range(args.warmup + args.steps):
        t, _ = cpu_reference_step(w, state)
        if i >= args.warmup:
            times.append(t)
    total = float(np.sum(times))
    value = n * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64 distances / int64 indices", "data": "synthetic",
        "config": {"workload": w["name"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": state["kind"],
                         "sample": reference_sample_text(w, state, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# this repository's arm
# ------------------------------------------------------------------------------------------------
def ncu_traffic(world):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the hand-written kernels, from the
    committed `ncu --set full` capture of the C2 fit on ONE GPU (profiles/ncu_traffic.json, written by
    scripts/ncu_summary.py); {} when there is none or when the kernels run on row blocks (N > 1)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if world > 1 or not os.path.exists(path):
        return {}
    with open(path) as fp:
        return json.load(fp).get("kernels", {})


def stage_rooflines(w, timings, n, d, nnz, stats, peak, world, words):
    """Algorithmic bytes per launch (SURVEY.md 8d / DESIGN.md 4) over the measured mean duration of the stage."""
    traffic = ncu_traffic(world) if w["key"] == "c2" else {}

    def entry(stage, kernels, nbytes, note, bound_by):
        ms, calls = timings.get(stage, (0, 0))
        if calls == 0 or ms <= 0:
            return None
        dur = ms / calls
        ach = nbytes / (dur * 1e-3) / 1e9
        tr = [traffic.get(k, {}).get("dram_bytes") for k in kernels]
        return {"kernel": " + ".join(kernels), "stage": stage, "bound": "hbm", "achieved": ach, "peak": peak["hbm"],
                "unit": "GB/s", "frac": ach / peak["hbm"], "traffic": sum(tr) if all(t is not None for t in tr) else None,
                "ms_per_launch": dur, "algorithmic_bytes": nbytes, "limited_by": bound_by, "note": note}

    cand, unsettled = stats["candidate_edges"], stats["candidate_edges"] - stats["skipped"]
    out = []
    if stats.get("search_method") == "grid":
        out += [
            entry("search_count", ["grid_plan_kernel", "grid_mask_kernel"], 8 * d * n + 16 * n + 4 * n + 4 * words,
                  "coords in (f64 for the exact recheck + float4 cell-local copy) + 4*N degrees + 4 B per (query, 32-candidate "
                  "chunk) hit mask out", "issue slots: 6.5 instructions per candidate and warp (DESIGN.md 4)"),
            entry("search_fill", ["grid_expand_kernel"], 4 * words + 8 * (n + 1) + 4 * nnz,
                  "hit masks in + row offsets + 4*nnz column indices out (coalesced)", "issue slots (bit extraction)"),
        ]
    elif stats.get("search_method") == "tensor":
        ms, calls = timings.get("search_count", (0, 0))
        if calls:
            dur = ms / calls
            useful = float(d) * n * (n - 1)
            kpad = ((3 * d + 63) // 64) * 64
            visited = float(stats["distance_evals"])       # pairs in the 128x128 tiles the kernel put through the MMA
            issued = 2.0 * visited * kpad                  # hi/lo split: K = 3*D padded to 64
            tf = issued / (dur * 1e-3) / 1e12
            out.append({"kernel": "tensor_tiles_kernel", "stage": "search_count", "bound": "tensor",
                        "achieved": tf, "peak": peak["tf_sustained"], "unit": "TFLOP/s",
                        "frac": tf / peak["tf_sustained"], "traffic": None, "ms_per_launch": dur,
                        "issued_flops": issued, "issued_flops_unsplit": issued * d / kpad,
                        "tiles_visited_frac": visited / (float(n) * n / 2.0),
                        "whole_problem_flops_survey_8d": useful,
                        "whole_problem_tflops_equivalent": useful / (dur * 1e-3) / 1e12,
                        "limited_by": "tensor pipe (an MMA-only run of the same tile walk takes 0.42 ms on C3) + epilogue",
                        "note": "achieved = flops the tensor pipe really executed (visited tiles x K = 3*D of the fp16 hi/lo "
                                "split) over the kernel time; tile pairs whose bounding balls are further apart than the "
                                "radius are never visited (data dependent) -- whole_problem_* relates SURVEY 8d's "
                                "D*N*(N-1) to the same time and is NOT a hardware rate"})
    if "ingest" in timings and timings["ingest"][1] and w["key"] == "c4":
        out.append(entry("ingest", ["csr ingest (validate + adopt)"], 4 * nnz + 8 * (n + 1),
                         "row offsets + column indices read once for validation (borrowed device buffers)", "HBM"))
    out += [
        entry("comp_sample", ["comp_sample_kernel", "comp_far_emit_kernel", "comp_worklist_kernel"], 8 * (n + 1) + 4 * n + 4 * n,
              "row_ptr + parents + first-above positions (accepted edges are linked in place); rows read by the executed "
              "CN checks are not counted", "memory latency (dependent row look-ups, union-find chains)"),
        entry("comp_edges", ["comp_scan_kernel", "comp_worklist_kernel"],
              2 * nnz + 8 * (n + 1) + 8 * n + 4 * cand + 16 * unsettled,
              "upper half of every core row (about 2*nnz bytes) + row_ptr + parents + first-above positions + one 4-byte "
              "parent look-up per candidate edge + 8 B written and read per work-list entry; rows re-read by the "
              "executed CN checks are not counted", "memory latency (scan) / issue slots (work list)"),
    ]
    return [e for e in out if e]


def count_mode_bytes(ctx, n_member_cutoff, rows=None):
    """SURVEY 8d's algorithmic bytes of count mode for the CSR the context holds (computed on the device); `rows`
    restricts the sum to the candidate edges of a row block."""
    import torch

    from commonnnclustering_b200 import dist as cdist

    n, nnz, _ = ctx.neighbours_info()
    rp_ptr, ci_ptr, _ = ctx.neighbours_device_view()
    dev = torch.device("cuda", ctx.device)
    rp = cdist.tensor_from_ptr(rp_ptr, (n + 1,), "<i8", dev)
    ci = cdist.tensor_from_ptr(ci_ptr, (nnz,), "<i4", dev)
    deg = rp[1:] - rp[:-1]
    core = deg > n_member_cutoff
    total = 0
    edges = 0
    step = 1 << 17  # rows per slab keeps the temporaries small
    r0, r1 = (0, n) if rows is None else rows
    for b in range(r0, r1, step):
        e = min(r1, b + step)
        s0, s1 = int(rp[b]), int(rp[e])
        rws = torch.repeat_interleave(torch.arange(b, e, device=dev), deg[b:e])
        cols = ci[s0:s1].long()
        m = (cols > rws) & core[rws] & core[cols]
        total += int((deg[rws][m] + deg[cols][m]).sum())
        edges += int(m.sum())
    return 4 * total + 8 * edges, edges


def run_b200(args):
    import torch
    import torch.distributed as dist

    from commonnnclustering_b200 import _lib, cluster, sweep
    from commonnnclustering_b200 import dist as cdist

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = workload(args.workload, args.n)
    key = w["key"]
    x = w["make"]()
    n, d = x.shape
    r2 = w["r"] ** 2
    nmc, sim = w["c"] + 1, w["c"] + 2
    peak = peaks()
    steps = args.steps
    warmup = max(args.warmup, 3)

    ctx = _lib.Context(local, torch_allocator=True)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream)
    # c1 / c2 (and c4) shard their rows over the ranks; c3 / c5 do not shard: N independent replicas
    sharded = world > 1 and key in ("c1", "c2", "c4")
    replicas = world > 1 and not sharded
    runner = cdist.ShardedFit(ctx, dist.group.WORLD) if sharded else None
    want_labels = rank == 0 or replicas          # sharded runs: only the caller's rank takes the labels

    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    extra = {}
    h2d = d2h = 0

    # ---- the workload's device-resident and host-resident steps ---------------------------------------------
    if key in ("c1", "c2", "c3"):
        xd = torch.from_numpy(x).cuda()
        labd = torch.empty(n, dtype=torch.int64, device="cuda") if want_labels else None
        xh = torch.empty((n, d), dtype=torch.float64).pin_memory()
        xh.copy_(torch.from_numpy(x))
        lh = torch.empty(n, dtype=torch.int64).pin_memory() if want_labels else None
        xh_np, lh_np = xh.numpy(), (lh.numpy() if want_labels else None)
        h2d, d2h = int(x.nbytes), int(n * 8)

        def step_device():
            if runner is None:
                ctx.fit_points(xd, r2, nmc, sim, labd, member_cutoff=w["member_cutoff"])
            else:
                runner.fit_points(xd, r2, nmc, sim, labd, member_cutoff=w["member_cutoff"])

        def step_host():
            if runner is None:
                ctx.fit_points(xh_np, r2, nmc, sim, lh_np, member_cutoff=w["member_cutoff"])
            else:
                runner.fit_points(xh_np, r2, nmc, sim, lh_np, member_cutoff=w["member_cutoff"])

        e2e_call = ("cnnb200_fit_points (C-ABI, pinned host buffers)" if runner is None else
                    "dist.ShardedFit.fit_points over the C-ABI (cnnb200_fit_sharded_begin / _end), pinned host buffers: every "
                    "rank uploads 1/G of the coordinates, NCCL all-gather over NVLink, rank 0 reads the labels back")
        units = n
    elif key == "c4":
        # the "precomputed" input: neighbourhoods of the C2 generator at N = 1e7 in the caller's index space (the
        # time order of the trajectory), rows ascending, int32 indices / int64 offsets -- produced once, outside
        # every timed region, then treated as caller data
        ctx.neighbours_from_points(torch.from_numpy(x).cuda(), r2)
        rp_d, ci_d = ctx.neighbours_export_device()
        ctx.release()
        nnz = int(rp_d[-1])
        rp_h = torch.empty(n + 1, dtype=torch.int64).pin_memory()
        ci_h = torch.empty(nnz, dtype=torch.int32).pin_memory()
        rp_h.copy_(rp_d); ci_h.copy_(ci_d)
        rp_np, ci_np = rp_h.numpy(), ci_h.numpy()
        labd = torch.empty(n, dtype=torch.int64, device="cuda") if want_labels else None
        lh = torch.empty(n, dtype=torch.int64).pin_memory() if want_labels else None
        lh_np = lh.numpy() if want_labels else None
        h2d, d2h = int(rp_np.nbytes + ci_np.nbytes), int(n * 8)
        extra["nnz"] = nnz
        if runner is None:
            def step_device():
                ctx.neighbours_from_csr(rp_d, ci_d, borrow=True)
                ctx.fit_sorted(nmc, sim, labd, member_cutoff=w["member_cutoff"])

            def step_host():
                ctx.neighbours_from_csr(rp_np, ci_np)
                ctx.fit_sorted(nmc, sim, lh_np, member_cutoff=w["member_cutoff"])
            e2e_call = "cnnb200_neighbours_from_csr + cnnb200_fit_sorted (C-ABI, page-locked host CSR)"
        else:
            del rp_d, ci_d
            torch.cuda.empty_cache()
            blk = runner.upload_csr_block(rp_np, ci_np)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            xtimes = []

            def step_device():       # this rank's rows resident in HBM; exchange over NVLink inside the step
                ev[0].record(stream)
                bounds = runner.exchange_csr(blk)
                ev[1].record(stream)
                runner.fit(bounds, nmc, sim, labd, member_cutoff=w["member_cutoff"])
                xtimes.append((ev[0], ev[1]))
                ev[0], ev[1] = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

            def step_host():
                bounds = runner.exchange_csr(runner.upload_csr_block(rp_np, ci_np))
                runner.fit(bounds, nmc, sim, lh_np, member_cutoff=w["member_cutoff"])
            e2e_call = ("dist.ShardedFit: every rank uploads 1/G of the CSR rows from page-locked host arrays, NCCL broadcast of "
                        "the slices (CSR all-gather), row-sharded components, parent all-gather + merge, rank 0 reads the labels")
        units = n
    else:  # c5
        xd = torch.from_numpy(x).cuda()
        radii, cnn = w["radii"], w["cnn"]
        cells = len(radii) * len(cnn)
        lab_all = torch.empty((len(radii), len(cnn), n), dtype=torch.int64).pin_memory().numpy()
        h2d, d2h = int(x.nbytes), int(cells * n * 8)

        hcl = cluster.Clustering(x)
        hcl.fitter._ctx = ctx
        soft = torch.empty(n, dtype=torch.int64, device="cuda")

        def hierarchy(labels_softest):
            """isolate the clusters of the softest cell (r = 0.100, c = 2), recluster every child with >= 1000 points
            at (0.063, 20), reel."""
            hcl.labels = labels_softest
            hcl.isolate()
            for child in hcl.root._children.values():
                if child._input_data.n_points >= 1000:
                    hcl.fit(0.063, 20, member_cutoff=w["member_cutoff"], v=False, bundle=child)
            hcl.reel()
            return hcl.labels

        def step_device():
            sweep.fit_sweep(xd, radii, cnn, w["member_cutoff"], ctx=ctx, return_labels=False)
            # the softest cell once more for its labels (the sweep above keeps none): one fit, one 8 MB copy
            ctx.fit_points(xd, radii[-1] ** 2, cnn[0] + 1, cnn[0] + 2, soft, member_cutoff=w["member_cutoff"])
            extra["clusters_after_reel"] = int(hierarchy(soft.cpu().numpy()).max())

        def step_host():
            sweep.fit_sweep(x, radii, cnn, w["member_cutoff"], ctx=ctx, labels_out=lab_all)
            hierarchy(lab_all[len(radii) - 1, 0].copy())
        e2e_call = ("sweep.fit_sweep (one radius search + one count-mode pass per radius, components per cell over the C-ABI), "
                    "coordinates from the host, all 64 label arrays copied into a page-locked host array; then "
                    "Clustering.isolate on the softest cell, Clustering.fit on every child with >= 1000 points, Clustering.reel")
        units = cells * n
        steps = min(steps, 5)
        labd = None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup):
        step_device()
    barrier()

    # ---- timed region 1: inputs resident in HBM, CUDA events on the launching stream -------------
    ctx.set_timing(True)
    ctx.timings(reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    if key == "c4" and runner is not None:
        xtimes.clear()
    barrier()
    for i in range(steps):
        flush.fill_(i & 0xff)  # L2 flush between timed iterations (outside the event pair)
        starts[i].record(stream)
        step_device()
        ends[i].record(stream)
    barrier()
    dev_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    timings, launches = ctx.timings(reset=True)
    ctx.set_timing(False)
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    stats = ctx.stats()
    _, nnz, max_deg = ctx.neighbours_info()
    lab_dev = labd.cpu().numpy() if labd is not None else None
    if key == "c4" and runner is not None and xtimes:
        x_ms = float(np.mean([a.elapsed_time(b) for a, b in xtimes]))
        recv = (extra["nnz"] * 4 + (n + 1) * 8) * (world - 1) / world
        extra["csr_all_gather"] = {"ms": x_ms, "bytes_received_per_rank": recv, "bus_gbs": recv / (x_ms * 1e-3) / 1e9,
                                   "note": "NCCL broadcasts of the row_ptr / col_idx slices from their owners (variable sizes)"}

    # ---- timed region 2: end to end through the C-ABI on pinned HOST buffers ----------------------
    for _ in range(3 if key != "c5" else 1):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        step_host()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    if key != "c5" and want_labels:
        assert np.array_equal(lh_np, lab_dev), "host-buffer and device-buffer paths disagree"

    # ---- count mode (SimilarityChecker._get for every candidate edge): the intersection kernel on its own -------
    count_mode = None
    if key in ("c1", "c2") and not replicas:
        step_device()
        rows = None if runner is None else cdist.block_of(runner.last["bounds"], rank)
        nb_count, n_edges = count_mode_bytes(ctx, nmc, rows)
        _, nnz_local, _ = ctx.neighbours_info()
        counts = torch.empty(max(nnz_local, 1), dtype=torch.int32, device="cuda")
        ctx.set_timing(True)
        ctx.timings(reset=True)
        for _ in range(2):
            ctx.cn_count(nmc, counts, rows=rows)
        t_warm, _ = ctx.timings(reset=True)
        pack_ms = t_warm["ingest"][0] if t_warm.get("ingest", (0, 0))[1] else None   # pack_rows_kernel: once per CSR
        barrier()
        reps = 5
        for i in range(reps):
            flush.fill_(i)
            ctx.cn_count(nmc, counts, rows=rows)
        t_count, _ = ctx.timings(reset=True)
        ctx.set_timing(False)
        agg = torch.tensor([float(nb_count), float(n_edges)], dtype=torch.float64, device="cuda")
        tmax = torch.tensor([t_count["cn_count"][0] / reps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(agg)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        count_mode = (float(tmax.item()), float(agg[0].item()), float(agg[1].item()), pack_ms, ctx.stats()["count_packed"])
        del counts

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the user-facing call: Clustering.fit (host wrapper included), single GPU only -----------
    api = None
    if world == 1 and key in ("c1", "c2", "c3"):
        cl = cluster.Clustering(xh_np)
        cl.fitter._ctx = ctx
        ts = []
        for i in range(3 + min(steps, 5)):
            t0 = time.perf_counter()
            cl.fit(w["r"], w["c"], member_cutoff=w["member_cutoff"], v=False)
            ts.append(time.perf_counter() - t0)
        api = {"call": "Clustering(x).fit(r, c, member_cutoff)", "value": n / float(np.mean(ts[3:])), "unit": UNIT,
               "fitter_points_per_s": n / cl.summary[-1].execution_time}
        assert np.array_equal(cl.labels, lab_dev)

    words = stats["distance_evals"] // 32
    roofs = stage_rooflines(w, timings, n, d, nnz, stats, peak, world, words) if key != "c5" else []
    if count_mode:
        ms, nb, edges, pack_ms, packed = count_mode
        kname = "cn_count_packed_kernel" if packed else "cn_count_kernel"
        tr = ncu_traffic(world).get(kname, {}).get("dram_bytes") if key == "c2" else None
        roofs.append({
            "kernel": kname, "stage": "cn_count (count mode: every candidate edge, no early exit; not part of the timed fit "
            "step)" + (f"; rows sharded over {world} GPUs, slowest rank's time" if world > 1 else ""),
            "bound": "hbm", "achieved": (tr / (ms * 1e-3) / 1e9) if tr else None, "peak": peak["hbm"], "unit": "GB/s",
            "frac": (tr / (ms * 1e-3) / 1e9 / peak["hbm"]) if tr else None, "traffic": tr, "ms_per_launch": ms,
            "cn_checks_per_s": edges / (ms * 1e-3), "pack_rows_ms_once_per_csr": pack_ms,
            "no_reuse_model_bytes": nb, "no_reuse_model_gbs": nb / (ms * 1e-3) / 1e9,
            "limited_by": "issue slots (one AND+POPC per streamed 32-index word; the rows come from L1/L2)",
            "note": "`achieved` / `frac` = MEASURED DRAM traffic of the committed ncu capture over this run's duration (1 GPU "
                    "only); the headline of this kernel is cn_checks_per_s.  no_reuse_model_* is SURVEY 8d's algorithmic model "
                    "(every row re-read per edge), a work measure, not a bandwidth"})
    fit_roofs = [e for e in roofs if not e["stage"].startswith("cn_count")]
    top = max(fit_roofs, key=lambda e: e["ms_per_launch"]) if fit_roofs else None
    stage_ms = {k: v[0] / steps for k, v in timings.items() if v[1]}
    t_cn = (timings["comp_sample"][0] + timings["comp_edges"][0]) / steps * 1e-3
    line = {
        "metric": METRIC, "value": units * steps / (dev_ms * 1e-3) * (world if replicas else 1), "unit": UNIT, "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": dev_ms / steps, "higher_is_better": True,
        "scaling": "weak" if replicas else "strong", "vs_baseline": None,
        "value_is": "inputs resident in HBM, CUDA events (bench contract); e2e.value is SURVEY 8d's metric: H2D of the inputs, "
                    "all kernels and collectives, D2H of the labels and the final synchronisation inside the timed region",
        "dtype": {"c3": "fp16 hi/lo split tcgen05 filter, f32 accumulation + exact f64 recheck of borderline pairs / int32 indices "
                        "/ int64 labels"}.get(key, "f64 (fp32 filter + exact f64 recheck of borderline pairs) / int32 indices / "
                                                   "int64 labels") if key != "c4" else "int32 indices / int64 offsets and labels",
        "data": "synthetic",
        "config": {"workload": w["name"], "l2_flush_between_steps": True,
                   "parallelism": (f"row-sharded x{world}" if sharded else f"{world} independent replicas (the path does not "
                                   "shard)" if replicas else "1 GPU"),
                   "n_points": n, "n_dim": d, "nnz": nnz, "max_degree": max_deg},
        "e2e": {"value": units * steps / e2e_s * (world if replicas else 1), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / steps, "call": e2e_call},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": ({k: top[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic")} | {
            "kernel": top["kernel"], "limited_by": top["limited_by"], "peak_source": peak["source"]}) if top else None,
        "kernels": roofs,
        "stage_ms_per_step": stage_ms,
        "unstaged_ms_per_step": dev_ms / steps - sum(stage_ms.values()),
        "cn_checks": {"candidate_edges": stats["candidate_edges"], "checks_executed": stats["checks"],
                      "accepted": stats["accepted"], "skipped_already_connected": stats["skipped"],
                      "edges_settled_per_s": stats["candidate_edges"] / t_cn if t_cn > 0 else None,
                      "checks_executed_per_s": stats["checks"] / t_cn if t_cn > 0 else None},
        "search": {"method": stats["search_method"], "distance_evals": stats["distance_evals"], "f64_rechecks": stats["rechecks"]},
    }
    if lab_dev is not None:
        line["n_clusters"] = int(lab_dev.max())
        line["labels_sha1"] = sha(lab_dev)
    if key == "c5":
        line["labels_sha1"] = sha(lab_all)
        line["cells"] = len(w["radii"]) * len(w["cnn"])
    line.update(extra)
    if api:
        line["api"] = api

    # ---- CPU baseline: the reference's own fit on this box's host cores, checked against the GPU labels ---------
    if not args.no_cpu_baseline and (world == 1 or key in ("c1", "c2")):
        state = cpu_reference_setup(w, os.cpu_count() or 1)
        t_fit, lab_cpu = cpu_reference_step(w, state)
        full = state["n"] == n and key != "c5"
        line["cpu_baseline"] = {
            "value": state["n"] / t_fit, "unit": UNIT, "cores": 1, "kind": state["kind"],
            "sample": reference_sample_text(w, state, os.cpu_count() or 1) + f"; ONE fit, {t_fit:.2f} s",
            "labels_equal_gpu": bool(np.array_equal(lab_cpu, lab_dev)) if full else None,
        }
        if key == "c5":   # the sampled cell against the sweep's labels for that cell
            ir, ic = w["radii"].index(w["r"]), w["cnn"].index(w["c"])
            line["cpu_baseline"]["labels_equal_gpu"] = bool(np.array_equal(lab_cpu, lab_all[ir, ic]))
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--n", type=int, default=None, help="override the number of points (debugging only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
