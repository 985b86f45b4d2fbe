"""ctypes binding of ``cnn_oracle.c`` (TEST INFRASTRUCTURE, see ``oracle/__init__.py``)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cnn_oracle.c")
LIB = os.path.join(HERE, "libcnn_oracle.so")

_lib = None
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(
            ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"],
            check=True,
        )
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.cnn_oracle_distance.restype = C.c_double
        L.cnn_oracle_distance.argtypes = [_f64p, C.c_int64, C.c_int64, C.c_int64, C.c_int]
        L.cnn_oracle_adjust_radius.restype = C.c_double
        L.cnn_oracle_adjust_radius.argtypes = [C.c_double, C.c_int]
        for name in ("brute", "grid"):
            f = getattr(L, f"cnn_oracle_neighbours_{name}_count")
            f.argtypes = [_f64p, C.c_int64, C.c_int64, C.c_double, C.c_int, _i64p]
            f = getattr(L, f"cnn_oracle_neighbours_{name}_fill")
            f.argtypes = [_f64p, C.c_int64, C.c_int64, C.c_double, C.c_int, _i64p, _i64p]
        L.cnn_oracle_neighbours_brute_count.restype = None
        L.cnn_oracle_neighbours_brute_fill.restype = None
        L.cnn_oracle_neighbours_upper.restype = None
        L.cnn_oracle_neighbours_upper.argtypes = [_f64p, C.c_int64, C.c_int64, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                                  C.c_void_p]
        L.cnn_oracle_neighbours_periodic_count.restype = None
        L.cnn_oracle_neighbours_periodic_count.argtypes = [_f64p, C.c_int64, C.c_int64, C.c_double, _f64p, _i64p]
        L.cnn_oracle_neighbours_periodic_fill.restype = None
        L.cnn_oracle_neighbours_periodic_fill.argtypes = [_f64p, C.c_int64, C.c_int64, C.c_double, _f64p, _i64p, _i64p]
        for name in ("check_sorted", "check_contains"):
            f = getattr(L, f"cnn_oracle_{name}")
            f.restype = C.c_int
            f.argtypes = [_i64p, C.c_int64, _i64p, C.c_int64, C.c_int64]
        L.cnn_oracle_count_sorted.restype = C.c_int64
        L.cnn_oracle_count_sorted.argtypes = [_i64p, C.c_int64, _i64p, C.c_int64]
        L.cnn_oracle_fit_bfs.restype = C.c_int
        L.cnn_oracle_fit_bfs.argtypes = [_i64p, _i64p, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                         _i64p, _u8p, _i64p]
        L.cnn_oracle_sort_by_size.restype = C.c_int
        L.cnn_oracle_sort_by_size.argtypes = [_i64p, C.c_int64, C.c_int64, C.c_int64]
        L.cnn_oracle_csr_to_padded.restype = None
        L.cnn_oracle_csr_to_padded.argtypes = [_i64p, _i64p, C.c_int64, C.c_int64, _i64p, _i64p]
        _lib = L
    return _lib


METRICS = {"euclidean_r": 0, "euclidean": 1}


def set_threads(n: int) -> None:
    """Number of OpenMP threads for the neighbour search (the BFS is always single-threaded)."""
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        omp = C.CDLL("libgomp.so.1")
        omp.omp_set_num_threads(int(n))
    except OSError:
        pass


def adjust_radius(r: float, metric: str = "euclidean_r") -> float:
    return lib().cnn_oracle_adjust_radius(float(r), METRICS[metric])


def neighbours(x, radius_cutoff: float, metric: str = "euclidean_r", method: str = "auto"):
    """CSR (row_ptr int64[n+1], col_idx int64[nnz]) of {j: dist(i,j) <= radius_cutoff}.

    ``radius_cutoff`` is already metric-adjusted (r*r for ``euclidean_r``)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    n, d = x.shape
    L = lib()
    if method == "auto":
        method = "grid" if (d <= 3 and n > 4000) else "brute"
    deg = np.zeros(n, dtype=np.int64)
    m = METRICS[metric]
    cnt = getattr(L, f"cnn_oracle_neighbours_{method}_count")
    fil = getattr(L, f"cnn_oracle_neighbours_{method}_fill")
    rc = cnt(x, n, d, float(radius_cutoff), m, deg)
    assert not rc, rc
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=row_ptr[1:])
    col_idx = np.empty(max(int(row_ptr[-1]), 1), dtype=np.int64)
    rc = fil(x, n, d, float(radius_cutoff), m, row_ptr, col_idx)
    assert not rc, rc
    return row_ptr, col_idx[: int(row_ptr[-1])]


def neighbours_large(x, radius_cutoff: float, metric: str = "euclidean_r"):
    """The brute-force neighbour sets for any dimension at sizes where two full N^2 passes are too slow (C3, N = 2e5,
    D = 64): every unordered pair once, the sum abandoned as soon as it exceeds the cutoff (exact, see
    ``cnn_oracle_neighbours_upper``), then mirrored.  Same CSR as ``neighbours(..., method="brute")``."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    n, d = x.shape
    L = lib()
    m = METRICS[metric]
    up_deg = np.zeros(n, dtype=np.int64)
    L.cnn_oracle_neighbours_upper(x, n, d, float(radius_cutoff), m, up_deg.ctypes.data, None, None)
    up_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(up_deg, out=up_ptr[1:])
    up_idx = np.empty(max(int(up_ptr[-1]), 1), dtype=np.int64)
    L.cnn_oracle_neighbours_upper(x, n, d, float(radius_cutoff), m, None, up_ptr.ctypes.data, up_idx.ctypes.data)
    up_idx = up_idx[: int(up_ptr[-1])]
    rows = np.repeat(np.arange(n, dtype=np.int64), up_deg)
    # both directions of every pair plus the diagonal (dist(i, i) = 0 <= cutoff for any cutoff >= 0)
    r_all = np.concatenate([rows, up_idx, np.arange(n, dtype=np.int64)])
    c_all = np.concatenate([up_idx, rows, np.arange(n, dtype=np.int64)])
    order = np.lexsort((c_all, r_all))
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(r_all, minlength=n), out=row_ptr[1:])
    return row_ptr, np.ascontiguousarray(c_all[order])


def check(a, b, c: int, kind: str = "sorted") -> bool:
    a = np.ascontiguousarray(a, dtype=np.int64)
    b = np.ascontiguousarray(b, dtype=np.int64)
    f = lib().cnn_oracle_check_sorted if kind == "sorted" else lib().cnn_oracle_check_contains
    return bool(f(a, a.size, b, b.size, int(c)))


def count_common(a, b) -> int:
    a = np.ascontiguousarray(a, dtype=np.int64)
    b = np.ascontiguousarray(b, dtype=np.int64)
    return int(lib().cnn_oracle_count_sorted(a, a.size, b, b.size))


def make_parameters(r: float, c: int, metric="euclidean_r", selfcounting=True, current_start=1):
    """FitterExtInterface.make_parameters, _fit.pyx:129-161."""
    return {
        "radius_cutoff": adjust_radius(r, metric) if metric is not None else float(r),
        "similarity_cutoff": c + 2 if selfcounting else c,
        "n_member_cutoff": c + 1 if selfcounting else c,
        "current_start": current_start,
    }


def fit_bfs(row_ptr, col_idx, n_member_cutoff, similarity_cutoff, current_start=1,
            labels=None, consider=None, return_stats=False):
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    col_idx = np.ascontiguousarray(col_idx, dtype=np.int64)
    if col_idx.size == 0:
        col_idx = np.zeros(1, dtype=np.int64)
    n = row_ptr.size - 1
    if labels is None:
        labels = np.zeros(n, dtype=np.int64)
    if consider is None:
        consider = np.ones(n, dtype=np.uint8)
    stats = np.zeros(2, dtype=np.int64)
    rc = lib().cnn_oracle_fit_bfs(row_ptr, col_idx, n, int(n_member_cutoff), int(similarity_cutoff),
                                  int(current_start), labels, consider, stats)
    assert rc == 0, rc
    return (labels, stats) if return_stats else labels


def sort_by_size(labels, member_cutoff=None, max_clusters=None):
    labels = np.ascontiguousarray(labels, dtype=np.int64)
    rc = lib().cnn_oracle_sort_by_size(
        labels, labels.size, -1 if member_cutoff is None else int(member_cutoff),
        -1 if max_clusters is None else int(max_clusters))
    assert rc == 0, rc
    return labels


def csr_to_padded(row_ptr, col_idx):
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    col_idx = np.ascontiguousarray(col_idx, dtype=np.int64)
    n = row_ptr.size - 1
    max_deg = int(np.diff(row_ptr).max()) if n else 0
    padded = np.zeros((n, max(max_deg, 1)), dtype=np.int64)
    counts = np.zeros(n, dtype=np.int64)
    if col_idx.size == 0:
        col_idx = np.zeros(1, dtype=np.int64)
    lib().cnn_oracle_csr_to_padded(row_ptr, col_idx, n, padded.shape[1], padded, counts)
    return padded, counts


def fit_csr(row_ptr, col_idx, c: int, member_cutoff=None, max_clusters=None, sort=True,
            selfcounting=True, current_start=1, return_stats=False):
    """Clustering.fit on precomputed neighbourhoods (cluster.pyx:256-392, lookup recipe)."""
    p = make_parameters(0.0, c, metric=None, selfcounting=selfcounting, current_start=current_start)
    out = fit_bfs(row_ptr, col_idx, p["n_member_cutoff"], p["similarity_cutoff"], current_start,
                  return_stats=return_stats)
    labels, stats = out if return_stats else (out, None)
    if sort:
        sort_by_size(labels, member_cutoff, max_clusters)
    return (labels, stats) if return_stats else labels


def fit_points(x, r: float, c: int, member_cutoff=None, max_clusters=None, sort=True,
               metric="euclidean_r", method="auto", current_start=1):
    """Clustering.fit with the default "coordinates" recipe (cluster.pyx:256-392)."""
    p = make_parameters(r, c, metric=metric, current_start=current_start)
    row_ptr, col_idx = neighbours(x, p["radius_cutoff"], metric, method)
    labels = fit_bfs(row_ptr, col_idx, p["n_member_cutoff"], p["similarity_cutoff"], current_start)
    if sort:
        sort_by_size(labels, member_cutoff, max_clusters)
    return labels


def neighbours_other(x, x_other, radius_cutoff: float, metric: str = "euclidean_r"):
    """NeighboursGetterExtBruteForce._get_other (_types.pyx:1712-1731) for every other point: CSR of the clustered
    points within the cutoff, ascending.  NumPy restatement of the reference's f64 expression (sequential sum of
    squared differences, no FMA): for small cases."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    xo = np.ascontiguousarray(x_other, dtype=np.float64)
    rows = []
    for p in range(xo.shape[0]):
        s = np.zeros(x.shape[0])
        for k in range(x.shape[1]):
            t = xo[p, k] - x[:, k]
            s = s + t * t
        if metric == "euclidean":
            s = np.sqrt(s)
        rows.append(np.flatnonzero(s <= radius_cutoff))
    row_ptr = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum([r.size for r in rows], out=row_ptr[1:])
    return row_ptr, (np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64)).astype(np.int64)


def predict_csr(row_ptr, col_idx, o_row_ptr, o_col_idx, labels, n_member_cutoff, similarity_cutoff, clusters=None,
                other_labels=None, other_consider=None):
    """PredictorFirstmatch.predict (_fit.pyx:1231-1291), loop for loop."""
    labels = np.asarray(labels)
    m = o_row_ptr.size - 1
    if other_labels is None:
        other_labels = np.zeros(m, dtype=np.int64)
    if other_consider is None:
        other_consider = np.ones(m, dtype=np.uint8)
    if clusters is None:
        clusters = set(np.unique(labels).tolist()) - {0}
    for p in range(m):
        if other_consider[p] == 0:
            continue
        na = o_col_idx[o_row_ptr[p]:o_row_ptr[p + 1]]
        if not na.size > n_member_cutoff:
            continue
        for member in na:
            label = int(labels[member])
            if label not in clusters:
                continue
            nb = col_idx[row_ptr[member]:row_ptr[member + 1]]
            if not nb.size > n_member_cutoff:
                continue
            if similarity_cutoff == 0 or count_common(na, nb) >= similarity_cutoff:
                other_consider[p] = 0
                other_labels[p] = label
                break
    return other_labels, other_consider


def predict_points(x, labels, x_other, r: float, c: int, clusters=None, metric="euclidean_r"):
    """Clustering.predict with brute-force getters on coordinates (cluster.pyx:560-653)."""
    p = make_parameters(r, c, metric=metric, current_start=0)
    row_ptr, col_idx = neighbours(x, p["radius_cutoff"], metric, "brute")
    o_row_ptr, o_col_idx = neighbours_other(x, x_other, p["radius_cutoff"], metric)
    return predict_csr(row_ptr, col_idx, o_row_ptr, o_col_idx, labels, p["n_member_cutoff"], p["similarity_cutoff"],
                       clusters)[0]


def neighbours_periodic(x, radius_cutoff: float, bounds):
    """Brute-force CSR under MetricExtEuclideanPeriodicReduced (_types.pyx:2290-2318); radius_cutoff = r*r."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    bounds = np.ascontiguousarray(bounds, dtype=np.float64)
    n, d = x.shape
    assert bounds.shape == (d,)
    L = lib()
    deg = np.zeros(n, dtype=np.int64)
    L.cnn_oracle_neighbours_periodic_count(x, n, d, float(radius_cutoff), bounds, deg)
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=row_ptr[1:])
    col_idx = np.empty(max(int(row_ptr[-1]), 1), dtype=np.int64)
    L.cnn_oracle_neighbours_periodic_fill(x, n, d, float(radius_cutoff), bounds, row_ptr, col_idx)
    return row_ptr, col_idx[: int(row_ptr[-1])]


def fit_points_periodic(x, r: float, c: int, bounds, member_cutoff=None, max_clusters=None, sort=True):
    p = make_parameters(r, c, metric=None)
    row_ptr, col_idx = neighbours_periodic(x, r * r, bounds)
    labels = fit_bfs(row_ptr, col_idx, p["n_member_cutoff"], p["similarity_cutoff"], 1)
    if sort:
        sort_by_size(labels, member_cutoff, max_clusters)
    return labels


def neighbours_from_distances(dist, radius_cutoff: float):
    """CSR of {j: dist[i, j] <= radius_cutoff}, rows ascending: NeighboursGetterExtBruteForce._get
    (_types.pyx:1692-1710) over MetricExtPrecomputed._calc_distance (_types.pyx:2108-2127) -- the metric
    returns ``input_data[i, j]`` and leaves the radius unchanged (``_adjust_radius``, :2126-2127)."""
    dist = np.ascontiguousarray(dist, dtype=np.float64)
    n = dist.shape[0]
    assert dist.shape == (n, n)
    with np.errstate(invalid="ignore"):
        hit = dist <= float(radius_cutoff)
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(hit.sum(axis=1), out=row_ptr[1:])
    col_idx = np.nonzero(hit)[1].astype(np.int64)
    return row_ptr, col_idx


def fit_distances(dist, r: float, c: int, member_cutoff=None, max_clusters=None, sort=True):
    """Clustering.fit with the "distances" recipe (recipes.py:329-339; self-counting brute-force getter)."""
    p = make_parameters(r, c, metric=None)
    row_ptr, col_idx = neighbours_from_distances(dist, p["radius_cutoff"])
    labels = fit_bfs(row_ptr, col_idx, p["n_member_cutoff"], p["similarity_cutoff"], 1)
    if sort:
        sort_by_size(labels, member_cutoff, max_clusters)
    return labels
