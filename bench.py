#!/usr/bin/env python
"""Benchmark of the fit hot path (BASELINE.json metric: fit() points/s and CN-checks/s).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repository's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path

Workload (`config.workload`): BASELINE.json configs[1] -- synthetic MD-like 3-D trajectory,
N = 1,000,000 frames, radius_cutoff 0.09, cnn_cutoff 20, member_cutoff 10 (SURVEY.md 8d "C2").
One step = one complete fit: radius-neighbour search (cell list) -> common-neighbour check +
union-find -> cluster numbering -> sort_by_size, labels int64[N] out.

Prints ONE JSON line (rank 0).  `value` = points/s with the coordinates already resident in HBM
(CUDA events); `e2e` = the same through the C-ABI call `cnnb200_fit_points` on pinned HOST
buffers, H2D of the coordinates and D2H of the labels inside the timed region (wall clock).
"""
import argparse
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
        return dict(name=f"C2: MD-like 3-D trajectory N={n}, r={p['radius_cutoff']}, c={p['cnn_cutoff']}, "
                         f"member_cutoff={p['member_cutoff']} (BASELINE configs[1])",
                    x=datasets.md_like_3d(n), r=p["radius_cutoff"], c=p["cnn_cutoff"], member_cutoff=p["member_cutoff"])
    if name == "c1":
        p = datasets.C1_PARAMS
        return dict(name="C1: 2-D blobs N=10000, r=0.3, c=10 (BASELINE configs[0])", x=datasets.blobs_2d(10_000),
                    r=p["radius_cutoff"], c=p["cnn_cutoff"], member_cutoff=None)
    raise SystemExit(f"unknown workload {name}")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fp:
            d = json.load(fp)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline
# ------------------------------------------------------------------------------------------------
def cpu_reference_setup(w, threads):
    """Neighbourhoods for the reference's `sorted_neighbourhoods` recipe (the only CPU route the
    reference's authors use beyond ~1e4 points; its brute-force default would need ~14 h here)."""
    from oracle import port, ref

    port.set_threads(threads)
    t0 = time.perf_counter()
    rp, ci = port.neighbours(w["x"], w["r"] ** 2)
    t_nb = time.perf_counter() - t0
    kind = "reference" if ref.available() else "port"
    state = dict(kind=kind, rp=rp, ci=ci, t_neighbours=t_nb)
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
        cl.fit(w["r"], w["c"], member_cutoff=w["member_cutoff"], v=False, record=True, record_time=True, info=False)
        return cl.summary[-1].execution_time, np.asarray(cl.labels)
    from oracle import port

    t0 = time.perf_counter()
    lab = port.fit_csr(state["rp"], state["ci"], w["c"], member_cutoff=w["member_cutoff"])
    return time.perf_counter() - t0, lab


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    w = workload(args.workload, args.n)
    n = w["x"].shape[0]
    threads = os.cpu_count() or 1
    state = cpu_reference_setup(w, threads)
    times = []
    for i in range(args.warmup + args.steps):
        t, _ = cpu_reference_step(w, state)
        if i >= args.warmup:
            times.append(t)
    total = float(np.sum(times))
    value = n * len(times) / total
    sample = (f"full workload, N={n}; each step = the reference's FitterExtBFS.fit + sort_by_size on PRECOMPUTED sorted "
              f"neighbourhoods (sorted_neighbourhoods recipe); the neighbour search is NOT in the reference's step "
              f"(done once outside by the oracle's C cell list with {threads} threads in {state['t_neighbours']:.1f} s; "
              f"the reference's own routes are brute force ~5e-8*N^2 s or sklearn KDTree ~23 s + a ~30 s Python padding loop)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64 distances / int64 indices", "data": "synthetic",
        "config": {"workload": w["name"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": state["kind"], "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# this repository's arm
# ------------------------------------------------------------------------------------------------
def ncu_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the hand-written kernels, from
    the committed `ncu --set full` capture of this workload (profiles/ncu_traffic.json, written by
    scripts/ncu_summary.py); {} when there is none."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return {}
    with open(path) as fp:
        return json.load(fp).get("kernels", {})


def stage_rooflines(timings, steps, n, d, nnz, stats, peak, count_mode=None, world=1):
    """Algorithmic bytes per launch (SURVEY.md 8d / DESIGN.md) over the measured mean duration."""
    traffic = ncu_traffic()

    def entry(stage, kernels, nbytes, note, bound_by):
        ms, calls = timings[stage]
        if calls == 0 or ms <= 0:
            return None
        dur = ms / calls
        ach = nbytes / (dur * 1e-3) / 1e9
        tr = [traffic.get(k, {}).get("dram_bytes") for k in kernels]
        return {"kernel": " + ".join(kernels), "stage": stage, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": sum(tr) if all(t is not None for t in tr) else None,
                "ms_per_launch": dur, "algorithmic_bytes": nbytes, "limited_by": bound_by, "note": note}

    cand, unsettled = stats["candidate_edges"], stats["candidate_edges"] - stats["skipped"]
    out = [
        entry("search_count", ["grid_rows_f32_kernel"], 8 * d * n + 16 * n + 4 * n + 4 * nnz,
              "coords in (f64 for the exact recheck + float4 cell-local copy) + 4*N degrees out + 4*nnz hits out (gapped rows)",
              "issue slots: ~450 distance tests per point; the compulsory bytes are small (see DESIGN.md 4)"),
        entry("search_fill", ["compact_rows_kernel"], 8 * nnz + 16 * (n + 1), "4*nnz in + 4*nnz out + 16*(N+1) offsets", "HBM"),
        entry("comp_sample", ["comp_sample_kernel"], 8 * (n + 1) + 4 * n + 4 * n,
              "row_ptr + parents + first-above positions (the accepted edges are linked in place); rows read by the "
              "executed CN checks are not counted", "memory latency (dependent row look-ups, union-find chains)"),
        entry("comp_edges", ["comp_scan_kernel", "comp_worklist_kernel"],
              2 * nnz + 8 * (n + 1) + 8 * n + 4 * cand + 16 * unsettled,
              "upper half of every core row (about 2*nnz bytes) + row_ptr + parents + first-above positions + one 4-byte "
              "parent look-up per candidate edge + 8 B written and read per work-list entry; rows re-read by the "
              "executed CN checks are not counted", "memory latency (scan) / issue slots (work list)"),
    ]
    if count_mode:
        ms, nb, edges, pack_ms = count_mode
        kname = "cn_count_packed_kernel" if pack_ms is not None else "cn_count_kernel"
        tr = traffic.get(kname, {}).get("dram_bytes")
        out.append({"kernel": kname, "stage": "cn_count (count mode: every candidate edge, no early exit; not part "
                    "of the timed fit step)" + (f"; rows sharded over {world} GPUs, all bytes over the slowest rank's time, "
                                                f"peak = {world} x one GPU" if world > 1 else ""),
                    "bound": "hbm", "achieved": nb / (ms * 1e-3) / 1e9, "peak": peak * world, "unit": "GB/s",
                    "frac": nb / (ms * 1e-3) / 1e9 / (peak * world), "traffic": tr, "ms_per_launch": ms, "algorithmic_bytes": nb,
                    "limited_by": "issue slots (one AND+POPC per streamed 32-index word)" if pack_ms is not None else
                    "issue slots (one shared-memory probe per streamed entry)", "cn_checks_per_s": edges / (ms * 1e-3),
                    "pack_rows_ms_once_per_csr": pack_ms,
                    "note": "no-reuse model of SURVEY 8d: 4*sum over candidate edges of (deg_i + deg_j) + 8 B per edge; the rows "
                            "are in fact served from L1/L2 (N(i) sits in a shared-memory bitmap), so `traffic` is close to the "
                            "compulsory 4*nnz in + 4*nnz out and `frac` can exceed 1: it measures the kernel against what a "
                            "reuse-free streaming implementation could reach, not DRAM utilisation"})
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
        rows = torch.repeat_interleave(torch.arange(b, e, device=dev), deg[b:e])
        cols = ci[s0:s1].long()
        m = (cols > rows) & core[rows] & core[cols]
        total += int((deg[rows][m] + deg[cols][m]).sum())
        edges += int(m.sum())
    return 4 * total + 8 * edges, edges


def run_b200(args):
    import torch
    import torch.distributed as dist

    from commonnnclustering_b200 import _lib, cluster

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = workload(args.workload, args.n)
    x = w["x"]
    n, d = x.shape
    r2 = w["r"] ** 2
    nmc, sim = w["c"] + 1, w["c"] + 2
    peak, peak_src = peaks()

    ctx = _lib.Context(local, torch_allocator=True)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream)

    if world > 1:
        from commonnnclustering_b200 import dist as cdist

        runner = cdist.ShardedFit(ctx, dist.group.WORLD)
    else:
        runner = None

    xd = torch.from_numpy(x).cuda()
    labd = torch.empty(n, dtype=torch.int64, device="cuda")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step_device():
        if runner is None:
            ctx.fit_points(xd, r2, nmc, sim, labd, member_cutoff=w["member_cutoff"])
        else:
            runner.fit_points(xd, r2, nmc, sim, labd, member_cutoff=w["member_cutoff"])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()

    # ---- timed region 1: inputs resident in HBM, CUDA events on the launching stream -------------
    ctx.set_timing(True)
    ctx.timings(reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
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
    lab_dev = labd.cpu().numpy()

    # ---- timed region 2: end to end through the C-ABI on pinned HOST buffers ----------------------
    xh = torch.empty((n, d), dtype=torch.float64).pin_memory()
    xh.copy_(torch.from_numpy(x))
    lh = torch.empty(n, dtype=torch.int64).pin_memory()
    xh_np, lh_np = xh.numpy(), lh.numpy()

    def step_host():
        if runner is None:
            ctx.fit_points(xh_np, r2, nmc, sim, lh_np, member_cutoff=w["member_cutoff"])
        else:
            runner.fit_points(xh_np, r2, nmc, sim, lh_np, member_cutoff=w["member_cutoff"])

    for _ in range(3):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_host()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    assert np.array_equal(lh_np, lab_dev), "host-buffer and device-buffer paths disagree"

    # ---- count mode (SimilarityChecker._get for every candidate edge): the intersection kernel on its own -------
    # every rank counts the candidate edges of its own row block (the rows it needs are in its halo CSR); the job's
    # rate is all bytes over the slowest rank's time
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
    count_mode = (float(tmax.item()), float(agg[0].item()), float(agg[1].item()), pack_ms)
    del counts

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the user-facing call: Clustering.fit (host wrapper included), single GPU only -----------
    api = None
    if world == 1:
        cl = cluster.Clustering(xh_np)
        cl.fitter._ctx = ctx
        ts = []
        for i in range(3 + min(args.steps, 5)):
            t0 = time.perf_counter()
            cl.fit(w["r"], w["c"], member_cutoff=w["member_cutoff"], v=False)
            ts.append(time.perf_counter() - t0)
        api = {"call": "Clustering(x).fit(r, c, member_cutoff)", "value": n / float(np.mean(ts[3:])), "unit": UNIT,
               "fitter_points_per_s": n / cl.summary[-1].execution_time}
        assert np.array_equal(cl.labels, lab_dev)

    roofs = stage_rooflines(timings, args.steps, n, d, nnz, stats, peak, count_mode, world)
    fit_roofs = [e for e in roofs if not e["stage"].startswith("cn_count")]
    top = max(fit_roofs, key=lambda e: e["ms_per_launch"]) if fit_roofs else None
    stage_ms = {k: v[0] / args.steps for k, v in timings.items() if v[1]}
    t_cn = (timings["comp_sample"][0] + timings["comp_edges"][0]) / args.steps * 1e-3
    line = {
        "metric": METRIC, "value": n * args.steps / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64 (fp32 filter + exact f64 recheck of borderline pairs) / int32 indices / int64 labels",
        "data": "synthetic",
        "config": {"workload": w["name"], "l2_flush_between_steps": True, "parallelism": f"row-sharded x{world}",
                   "n_points": n, "n_dim": d, "nnz": nnz, "max_degree": max_deg},
        "e2e": {"value": n * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(x.nbytes),
                "d2h_bytes_per_step": int(n * 8) * world, "ms_per_step": 1e3 * e2e_s / args.steps,
                "call": "cnnb200_fit_points (C-ABI, pinned host buffers)" if world == 1 else
                        "dist.ShardedFit.fit_points over the C-ABI, pinned host buffers: every rank uploads 1/G of the "
                        "coordinates, NCCL all-gather over NVLink, every rank reads the full labels back"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": ({k: top[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic")} | {
            "kernel": top["kernel"], "limited_by": top["limited_by"], "peak_source": peak_src}) if top else None,
        "kernels": roofs,
        "stage_ms_per_step": stage_ms,
        "cn_checks": {"candidate_edges": stats["candidate_edges"], "checks_executed": stats["checks"],
                      "accepted": stats["accepted"], "skipped_already_connected": stats["skipped"],
                      "edges_settled_per_s": stats["candidate_edges"] / t_cn if t_cn > 0 else None,
                      "checks_executed_per_s": stats["checks"] / t_cn if t_cn > 0 else None},
        "n_clusters": int(lab_dev.max()),
    }
    if api:
        line["api"] = api

    if world == 1 and not args.no_cpu_baseline:
        state = cpu_reference_setup(w, os.cpu_count() or 1)
        t_fit, lab_cpu = cpu_reference_step(w, state)
        line["cpu_baseline"] = {
            "value": n / t_fit, "unit": UNIT, "cores": 1, "kind": state["kind"],
            "sample": (f"full workload N={n}, ONE fit: the reference's fitter.fit (Record.execution_time) on precomputed "
                       f"sorted neighbourhoods, {t_fit:.2f} s; its neighbour search is excluded (oracle C cell list, "
                       f"{state['t_neighbours']:.1f} s on {os.cpu_count()} threads)"),
            "labels_equal_gpu": bool(np.array_equal(lab_cpu, lab_dev)),
        }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--n", type=int, default=None, help="override the number of points (debugging only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
