"""Parameter sweeps and the hierarchy helpers around them.

The reference scans a ``(radius_cutoff, cnn_cutoff)`` grid by calling ``Clustering.fit`` once per
cell (scan loop of ``docsrc/benchmark/scans.ipynb``; ``HierarchicalFitterRepeat._fit`` does the same
per step, ``_fit.pyx:1058-1187``).  Every cell repeats the radius search and the common-neighbour
checks.  On the device the work factors:

* one radius search per ``radius_cutoff``;
* one count-mode pass per ``radius_cutoff`` (``cnnb200_cn_count``: |N(i) ∩ N(j)| of every candidate
  edge, the reference's ``SimilarityChecker._get``), computed for the smallest member cutoff of the
  scan -- it answers every ``cnn_cutoff`` at once;
* per cell only the component pass over the counts (``cnnb200_components_from_counts``), the
  numbering and ``Labels.sort_by_size``.

Labels of every cell are identical to what ``Clustering.fit(r, c, member_cutoff)`` returns for that cell.
"""
import time

import numpy as np

from . import _lib
from .report import Record, Summary


def fit_sweep(x, radius_cutoffs, cnn_cutoffs, member_cutoff=None, max_clusters=None, metric="euclidean_r",
              ctx=None, return_labels=True, labels_out=None):
    """All cells of ``radius_cutoffs x cnn_cutoffs`` on coordinates ``x`` (numpy or CUDA tensor, f64 [N, D]).

    Returns ``(labels, summary)``: ``labels`` int64 ``[len(radius_cutoffs), len(cnn_cutoffs), N]`` on the host
    (``None`` when ``return_labels`` is false) and a ``Summary`` with one ``Record`` per cell in row-major
    order -- the same fields the reference's scan loop records (``n_clusters``, ``ratio_largest``,
    ``ratio_noise``; ``execution_time`` is the cell's share of the device time).
    """
    import torch

    from . import _types

    if ctx is None:
        ctx = _lib.default_context()
    # the metric decides how a radius becomes the cutoff the search compares with (`adjust_radius`) and may carry
    # state the context needs (periods); a name stands for the metric without state
    if isinstance(metric, str):
        if metric not in ("euclidean_r", "euclidean"):
            raise ValueError(f"fit_sweep: pass a Metric object for {metric!r} (it carries state), or use euclidean / euclidean_r")
        metric = _types.MetricEuclideanReduced() if metric == "euclidean_r" else _types.MetricEuclidean()
    if isinstance(metric, _types.MetricPrecomputed):
        raise ValueError("fit_sweep works on coordinates: a distance matrix has no radius search to share between cells")
    if isinstance(metric, _types.MetricEuclideanPeriodicReduced):
        ctx.set_periodic_bounds(metric.bounds)
    dev = torch.device("cuda", ctx.device)
    xd = x if hasattr(x, "is_cuda") else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
    xd = xd.to(dev)
    n = xd.shape[0]
    radius_cutoffs = [float(r) for r in radius_cutoffs]
    cnn_cutoffs = [int(c) for c in cnn_cutoffs]
    assert all(c >= 0 for c in cnn_cutoffs)
    n_r, n_c = len(radius_cutoffs), len(cnn_cutoffs)
    if return_labels:
        if labels_out is None:  # pinned: 8 bytes x N per cell come back over PCIe
            labels_out = torch.empty((n_r, n_c, n), dtype=torch.int64, pin_memory=True).numpy()
        assert labels_out.shape == (n_r, n_c, n) and labels_out.dtype == np.int64
    lab_dev = torch.empty(n, dtype=torch.int64, device=dev)
    summary = Summary()
    c_min = min(cnn_cutoffs)
    for ir, r in enumerate(radius_cutoffs):
        t0 = time.perf_counter()
        nnz = ctx.neighbours_from_points(xd, metric.adjust_radius(r), metric=metric.name)
        counts = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
        ctx.cn_count(c_min + 1, counts)          # self-counting neighbourhoods: member cutoff c + 1
        torch.cuda.synchronize(dev)
        t_shared = (time.perf_counter() - t0) / n_c
        for ic, c in enumerate(cnn_cutoffs):
            t0 = time.perf_counter()
            ctx.components_begin(c + 1, c + 2)   # FitterExtInterface.make_parameters, _fit.pyx:129-161
            ctx.components_from_counts(counts)
            ctx.components_labels_sorted(lab_dev, 1, member_cutoff, max_clusters)   # numbering + size sort as one ranking
            if return_labels:
                torch.from_numpy(labels_out[ir, ic]).copy_(lab_dev, non_blocking=True)
            # the record's three numbers come out of the device-side size sort: no pass over the labels
            n_kept, n_largest, n_noise = ctx.label_summary()
            summary.append(Record(n, r, c, member_cutoff, max_clusters, n_kept, n_largest / n, n_noise / n,
                                  t_shared + time.perf_counter() - t0))
        del counts
    torch.cuda.synchronize(dev)  # the last label copies
    return (labels_out if return_labels else None), summary
