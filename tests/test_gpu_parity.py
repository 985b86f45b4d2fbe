"""Parity of the CUDA fit path (through the C-ABI) against the oracle, the golden fixtures the
reference itself produced, and size-independent properties at BASELINE.json's full sizes.

Everything here is bit-exact (integer labels / indices / counts); no tolerances."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import cases  # noqa: E402
from commonnnclustering_b200 import _lib, _types, cluster, datasets  # noqa: E402
from oracle import port  # noqa: E402

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(HERE, "golden", "ref_labels.npz"))


def fit_via_api(x, r, c, recipe=None, **kw):
    metric = kw.pop("metric", None)
    extra = {"fitter__ngetter__dgetter__metric": metric} if metric else {}
    method = kw.pop("method", None)
    cl = cluster.Clustering(x, recipe=recipe, **extra)
    if method:
        cl.fitter._neighbours_getter.method = method
    cl.fit(r, c, v=False, **kw)
    return cl


# ---- golden fixtures produced by the reference -----------------------------------------------------
@pytest.mark.parametrize("name", list(cases.CASES))
def test_clustering_fit_equals_reference_golden(name):
    case = cases.CASES[name]
    x = case["data"]()
    kw = dict(case["fit"])
    cl = fit_via_api(x, kw.pop("r"), kw.pop("c"), **kw)
    np.testing.assert_array_equal(cl.labels, GOLD[name].astype(np.int64))
    assert cl.labels.dtype == np.int64
    assert not cl.root._labels.consider.any()  # consider[:] = 0 after a fit (_fit.pyx:434-437)


@pytest.mark.parametrize("name", ["blobs1500_r010_c20", "md5000_r04_c5", "lattice3d", "lattice1d", "blobs1500_euclid"])
def test_tiled_search_equals_reference_golden(name):
    case = cases.CASES[name]
    x = case["data"]()
    kw = dict(case["fit"])
    cl = fit_via_api(x, kw.pop("r"), kw.pop("c"), method="tiled", **kw)
    np.testing.assert_array_equal(cl.labels, GOLD[name].astype(np.int64))


@pytest.mark.parametrize("name", ["feat64d_2000", "gauss5d", "gauss8d_euclid", "blobs1500_r010_c20", "lattice3d"])
def test_tensor_core_search_equals_reference_golden(name):
    """tcgen05 filter + f64 boundary recheck: labels identical to the reference's."""
    case = cases.CASES[name]
    x = case["data"]()
    kw = dict(case["fit"])
    cl = fit_via_api(x, kw.pop("r"), kw.pop("c"), method="tensor", **kw)
    np.testing.assert_array_equal(cl.labels, GOLD[name].astype(np.int64))


def test_tensor_core_neighbour_sets_are_exact(ctx):
    """Neighbour sets of the tensor-core path equal the exact f64 all-pairs sets, including data with
    many distances exactly on the cutoff, large norms (far from the origin) and odd sizes."""
    rng = np.random.default_rng(21)
    sets = [
        (datasets.features_64d(6000), 0.9, "euclidean_r"),
        (cases.lattice(1000, 6, 31), 0.5, "euclidean_r"),              # exact ties at the cutoff
        (cases.lattice(700, 4, 32), 0.375, "euclidean"),
        (rng.standard_normal((1500, 33)) + 50.0, 7.0, "euclidean_r"),   # |x|^2 ~ 8e4 >> r^2: wide band
        (rng.standard_normal((129, 85)), 12.0, "euclidean_r"),          # largest supported dimension
        (rng.standard_normal((1, 10)), 1.0, "euclidean_r"),
    ]
    for x, r, metric in sets:
        rc = port.adjust_radius(r, metric)
        ctx.neighbours_from_points(x, rc, metric=metric, method="tiled")
        want = ctx.neighbours_export()
        ctx.neighbours_from_points(x, rc, metric=metric, method="tensor")
        got = ctx.neighbours_export()
        np.testing.assert_array_equal(got[0], want[0])
        np.testing.assert_array_equal(got[1], want[1])
    assert ctx.stats()["rechecks"] >= 0
    with pytest.raises(_lib.CnnB200Error):
        ctx.neighbours_from_points(rng.standard_normal((10, 86)), 1.0, method="tensor")  # > 85 dims


def test_tensor_core_fp16_split_edge_cases(ctx):
    """The fp16 hi/lo split of the tensor-core filter: every number of 64-wide K atoms (D = 4 .. 85, odd D: the scalar
    split kernel), instruction boundaries that cut through the hi.hi block (D % 16 != 0), coordinates scaled by 1e-20 /
    1e+20 (the power-of-two scaling), one dimension 1e6 times wider than the others (small coordinates become fp16
    subnormals after the scaling) and a radius far beyond the data (every pair a neighbour)."""
    rng = np.random.default_rng(77)
    sets = []
    for d in (4, 5, 21, 22, 43, 64, 85):
        x = rng.standard_normal((700 + d, d))
        r = float(np.sqrt(2.0 * d) * 0.8)
        sets.append((x, r))
    x = rng.standard_normal((900, 12))
    sets += [(x * 1e-20, 3.5e-20), (x * 1e20, 3.5e20), (x + 1e3, 3.5)]
    wide = rng.standard_normal((800, 10))
    wide[:, 0] *= 1e6
    sets.append((wide, 2.5e5))
    sets.append((rng.standard_normal((300, 7)), 1e3))     # R >> every distance
    sets.append((np.zeros((200, 9)), 0.0))                # all points identical, r = 0: 0 <= 0
    for x, r in sets:
        x = np.ascontiguousarray(x)
        ctx.neighbours_from_points(x, r * r, metric="euclidean_r", method="tiled")
        want = ctx.neighbours_export()
        ctx.neighbours_from_points(x, r * r, metric="euclidean_r", method="tensor")
        got = ctx.neighbours_export()
        np.testing.assert_array_equal(got[0], want[0])
        np.testing.assert_array_equal(got[1], want[1])
    x = rng.standard_normal((50, 6))
    x[7, 3] = np.nan
    with pytest.raises(ValueError):
        ctx.neighbours_from_points(x, 1.0, method="tensor")


def test_tensor_core_reordered_ties_and_pair_list_regrowth():
    """Pivot-sorted order + tile pruning (n >= 4096) on data full of exact ties, with a pair list that is far too
    small at first (the pass is repeated once with the size it asked for), with and without the reordering."""
    x = cases.lattice(5000, 6, 33)
    rc = 0.5 ** 2
    want = None
    for env in ({"CNNB200_TENSOR_PAIR_CAP": "1000"}, {"CNNB200_TENSOR_REORDER": "0"}, {}):
        os.environ.update(env)
        try:
            c = _lib.Context(0)
        finally:
            for k in env:
                del os.environ[k]
        if want is None:
            c.neighbours_from_points(x, rc, method="tiled")
            want = c.neighbours_export()
        c.neighbours_from_points(x, rc, method="tensor")
        got = c.neighbours_export()
        np.testing.assert_array_equal(got[0], want[0])
        np.testing.assert_array_equal(got[1], want[1])
        lab = np.empty(x.shape[0], dtype=np.int64)
        c.fit_points(x, rc, 6, 7, lab, method="tensor")
        np.testing.assert_array_equal(lab, port.fit_points(x, 0.5, 5))


def test_readme_doctest():
    cl = cluster.Clustering(datasets.README_POINTS)
    cl.fit(1.5, 1, v=False)
    assert cl.labels.tolist() == [1, 1, 1, 1, 1, 0, 0, 0, 0, 2, 2, 2]
    rec = cl.summary[-1]
    assert (rec.n_points, rec.n_clusters, rec.cnn_cutoff) == (12, 2, 1)
    assert abs(rec.ratio_noise - 4 / 12) < 1e-15 and rec.execution_time > 0
    assert cl.root._labels.meta["params"] == {1: (1.5, 1), 2: (1.5, 1)}


# ---- neighbour sets ------------------------------------------------------------------------------------
@pytest.mark.parametrize("maker,r,metric", [
    (lambda: datasets.blobs_2d(3000), 0.1, "euclidean_r"),
    (lambda: datasets.md_like_3d(4000), 0.5, "euclidean_r"),
    (lambda: cases.lattice(900, 3, 11), 0.25, "euclidean_r"),   # distances exactly on the cutoff
    (lambda: cases.lattice(900, 2, 12), 0.375, "euclidean"),
    (lambda: cases.lattice(300, 1, 13), 0.25, "euclidean_r"),
    (lambda: cases.gauss(500, 3, 5), 0.0, "euclidean_r"),        # radius 0: only self / duplicates
])
@pytest.mark.parametrize("method", ["grid", "tiled"])
def test_neighbour_sets_equal_brute_force(ctx, maker, r, metric, method):
    x = maker()
    rc = port.adjust_radius(r, metric)
    want_rp, want_ci = port.neighbours(x, rc, metric, method="brute")
    nnz = ctx.neighbours_from_points(x, rc, metric=metric, method=method)
    rp, ci = ctx.neighbours_export()
    assert nnz == want_rp[-1]
    np.testing.assert_array_equal(rp, want_rp)
    np.testing.assert_array_equal(ci, want_ci)


@pytest.mark.parametrize("name,maker,r,metric", [
    # coordinates far from the origin: the fp32 filter works on cell-local coordinates, so a large offset
    # must change nothing (the f64 test itself loses bits: many pairs land exactly on the cutoff)
    ("offset3d", lambda: cases.gauss(3000, 3, 21) * 0.2 + 1.0e6, 0.05, "euclidean_r"),
    ("offset2d", lambda: cases.gauss(3000, 2, 22) * 0.2 - 3.0e7, 0.03, "euclidean"),
    # lattice with every distance a tie, shifted off the origin
    ("lattice_shifted", lambda: cases.lattice(1000, 3, 23) + 1234.5, 0.25, "euclidean_r"),
    # duplicates and a dense clump inside a sparse box (coarse cells, long candidate runs)
    ("duplicates", lambda: np.repeat(cases.gauss(400, 3, 24), 5, axis=0), 0.3, "euclidean_r"),
    ("clump", lambda: np.concatenate([cases.gauss(2500, 3, 25) * 1e-3, cases.gauss(50, 3, 26) * 1e3]), 1e-3,
     "euclidean_r"),
    # tiny and huge scales
    ("tiny", lambda: cases.gauss(2000, 3, 27) * 1e-9, 3e-10, "euclidean_r"),
    ("huge", lambda: cases.gauss(2000, 2, 28) * 1e9, 2e8, "euclidean"),
    ("one_dim", lambda: cases.gauss(5000, 1, 29), 0.01, "euclidean_r"),
])
def test_filter_kernel_neighbour_sets_are_exact(ctx, name, maker, r, metric):
    """The single-pass cell-list kernel tests in fp32 and re-decides borderline pairs in f64
    (grid.cu, grid_rows_f32_kernel): its neighbour sets must equal the reference's f64 test."""
    x = np.ascontiguousarray(maker(), dtype=np.float64)
    rc = port.adjust_radius(r, metric)
    want_rp, want_ci = port.neighbours(x, rc, metric, method="brute")
    ctx.neighbours_from_points(x, rc, metric=metric, method="grid")
    rp, ci = ctx.neighbours_export()
    np.testing.assert_array_equal(rp, want_rp)
    np.testing.assert_array_equal(ci, want_ci)


def test_filter_kernel_equals_f64_kernel_on_c2_sample(ctx):
    """Same CSR from the fp32-filter kernel and from the all-f64 kernel (CNNB200_FILTER=0)."""
    import os
    x = datasets.md_like_3d(200_000)
    ctx.neighbours_from_points(x, 0.12 ** 2)
    a = ctx.neighbours_export()
    assert ctx.stats()["distance_evals"] > 0
    os.environ["CNNB200_FILTER"] = "0"
    try:
        ctx2 = _lib.Context(0)
    finally:
        del os.environ["CNNB200_FILTER"]
    ctx2.neighbours_from_points(x, 0.12 ** 2)
    b = ctx2.neighbours_export()
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])
    # two-pass form of the filter kernel (count, scan, fill): what very large inputs use
    os.environ["CNNB200_SINGLE_PASS_MB"] = "0"
    try:
        ctx3 = _lib.Context(0)
    finally:
        del os.environ["CNNB200_SINGLE_PASS_MB"]
    ctx3.neighbours_from_points(x, 0.12 ** 2)
    c = ctx3.neighbours_export()
    np.testing.assert_array_equal(a[0], c[0])
    np.testing.assert_array_equal(a[1], c[1])


def test_neighbour_search_high_dim_and_odd_sizes(ctx):
    for n, d in [(1, 3), (2, 2), (63, 7), (65, 17), (257, 33), (130, 64)]:
        x = cases.gauss(n, d, n + d)
        r = 0.9 * np.sqrt(d)
        want = port.neighbours(x, r * r, method="brute")
        ctx.neighbours_from_points(x, r * r, method="auto")
        got = ctx.neighbours_export()
        np.testing.assert_array_equal(got[0], want[0])
        np.testing.assert_array_equal(got[1], want[1])


# ---- precomputed neighbourhoods (lookup path) ------------------------------------------------------------
def test_lookup_recipes_equal_oracle():
    x = datasets.md_like_3d(6000)
    rp, ci = port.neighbours(x, 0.3 ** 2)
    want = port.fit_csr(rp, ci, 8, member_cutoff=5)
    # reference container layout: list of lists -> padded int64 + counts
    nb = [ci[rp[i]:rp[i + 1]] for i in range(x.shape[0])]
    cl = fit_via_api(nb, 0.3, 8, recipe="sorted_neighbourhoods", member_cutoff=5)
    np.testing.assert_array_equal(cl.labels, want)
    # CSR input, int64 and int32 indices
    for dt in (np.int64, np.int32):
        cl = fit_via_api((rp, ci.astype(dt)), 0.3, 8, recipe="neighbourhoods_csr", member_cutoff=5)
        np.testing.assert_array_equal(cl.labels, want)


GOLD_RECIPES = np.load(os.path.join(HERE, "golden", "ref_labels_recipes.npz"))


@pytest.mark.parametrize("name", list(cases.RECIPE_CASES))
def test_other_recipes_equal_reference_golden(name):
    """"distances" (N x N matrix through ``cnnb200_neighbours_from_distances``) and "neighbourhoods" (lists in any
    order, sorted on the device by ``cnnb200_neighbours_sort_rows``) through ``Clustering`` with the reference's own
    recipe names, against labels the reference produced."""
    case = cases.RECIPE_CASES[name]
    data = case["data"]()
    kw = dict(case["fit"])
    cl = fit_via_api(data, kw.pop("r"), kw.pop("c"), recipe=case["recipe"], **kw)
    np.testing.assert_array_equal(cl.labels, GOLD_RECIPES[name].astype(np.int64))


def test_distance_matrix_neighbour_sets(ctx):
    # sets equal the oracle's thresholding, for host and device matrices, odd sizes, NaN/inf entries, r on a value
    import torch

    rng = np.random.default_rng(3)
    for n in (1, 31, 33, 1024, 1025, 2500):
        dm = rng.random((n, n))
        dm = np.minimum(dm, dm.T)
        np.fill_diagonal(dm, 0.0)
        if n > 40:
            dm[5, 7] = dm[7, 5] = np.nan
            dm[1, 2] = dm[2, 1] = np.inf
            dm[3, 4] = dm[4, 3] = -np.inf
        r = float(dm[n // 2, n // 3]) if n > 1 else 0.0   # an entry exactly ON the cutoff: "<=" keeps it
        rp, ci = port.neighbours_from_distances(dm, r)
        for arr in (dm, torch.from_numpy(dm).cuda()):
            ctx.neighbours_from_distances(arr, r)
            got_rp, got_ci = ctx.neighbours_export()
            np.testing.assert_array_equal(got_rp, rp)
            np.testing.assert_array_equal(got_ci, ci)
    with pytest.raises(ValueError):
        ctx.neighbours_from_distances(np.zeros((3, 4)), 1.0)
    with pytest.raises(ValueError):
        ctx.neighbours_from_distances(np.zeros((3, 3), dtype=np.float32), 1.0)


def test_condensed_distances_equal_square_matrix(ctx):
    """Condensed distances (``InputDataExtDistancesLinearMemoryview``'s layout, _types.pyx:1112-1150) give the neighbour
    sets of the square matrix they stand for -- host and device vectors, sizes around the 32 / 1024 block edges, NaN /
    inf entries, a cutoff ON a value, a negative cutoff -- and the same labels through ``Clustering``."""
    import torch

    from commonnnclustering_b200 import _types

    rng = np.random.default_rng(5)
    for n in (1, 2, 31, 33, 64, 1023, 1025, 2500):
        dm = rng.random((n, n))
        dm = np.minimum(dm, dm.T)
        np.fill_diagonal(dm, 0.0)
        if n > 40:
            dm[5, 7] = dm[7, 5] = np.nan
            dm[1, 2] = dm[2, 1] = np.inf
            dm[3, 40] = dm[40, 3] = -np.inf
        iu = np.triu_indices(n, k=1)
        cond = np.ascontiguousarray(dm[iu])
        r = float(dm[n // 2, n // 3]) if n > 2 else 0.5
        rp, ci = port.neighbours_from_distances(dm, r)
        for arr in (cond, torch.from_numpy(cond).cuda()):
            ctx.neighbours_from_condensed(arr, n, r)
            got_rp, got_ci = ctx.neighbours_export()
            np.testing.assert_array_equal(got_rp, rp)
            np.testing.assert_array_equal(got_ci, ci)
        ctx.neighbours_from_condensed(cond, n, -1.0)     # d(a, a) = 0 > -1: only the -inf entries pass
        assert ctx.neighbours_info()[1] == int(port.neighbours_from_distances(dm, -1.0)[0][-1])
    with pytest.raises(ValueError):
        ctx.neighbours_from_condensed(np.zeros(4), 4, 1.0)
    # through the component interface: InputDataDistancesLinear + MetricPrecomputed
    x = datasets.blobs_2d(1500)
    dm = np.sqrt(((x[:, None, :] - x[None, :, :]) ** 2).sum(-1))
    data = _types.InputDataDistancesLinear(np.ascontiguousarray(dm[np.triu_indices(1500, k=1)]))
    assert data.n_points == 1500 and data.get_distance(7, 3) == dm[3, 7] and data.get_distance(4, 4) == 0.0
    cl = cluster.Clustering(data, recipe="distances")
    cl.fit(0.12, 8, v=False)
    want = cluster.Clustering(dm, recipe="distances")
    want.fit(0.12, 8, v=False)
    np.testing.assert_array_equal(cl.labels, want.labels)
    np.testing.assert_array_equal(cl.labels, port.fit_distances(dm, 0.12, 8))
    sub = data.get_subset(np.arange(0, 1500, 3))
    np.testing.assert_array_equal(sub.data, dm[::3, ::3][np.triu_indices(500, k=1)])


def test_new_input_paths_edge_cases(ctx):
    # distance matrix with no neighbours at all (not even the diagonal), a negative radius, a single point
    dm = np.full((40, 40), np.inf)
    ctx.neighbours_from_distances(dm, 1.0)
    assert ctx.neighbours_info()[:2] == (40, 0)
    assert ctx.neighbours_sort_rows() is False
    lab = np.ones(40, dtype=np.int64)
    ctx.fit(1, 2, 1, lab)
    assert not lab.any()
    ctx.neighbours_from_distances(np.zeros((40, 40)), -1.0)
    assert ctx.neighbours_info()[1] == 0
    ctx.neighbours_from_distances(np.zeros((1, 1)), 0.0)
    assert ctx.neighbours_info()[:2] == (1, 1)
    with pytest.raises(ValueError):
        ctx.neighbours_from_distances(np.zeros((2, 2)), float("nan"))
    # periodic cell list asked for explicitly on a handful of points, on duplicates, and with r = 0
    rng = np.random.default_rng(1)
    for n, r in ((1, 0.1), (7, 0.2), (50, 0.0), (300, 0.15)):
        x = np.ascontiguousarray(rng.random((n, 2)) * 3.0 - 1.0)
        if n == 50:
            x[25:] = x[:25]          # exact duplicates: distance 0 <= 0
        bounds = np.array([1.0, 1.5])
        ctx.set_periodic_bounds(bounds)
        want = port.neighbours_periodic(x, r * r, bounds)
        ctx.neighbours_from_points(x, r * r, metric="euclidean_periodic_r", method="grid")
        got = ctx.neighbours_export()
        np.testing.assert_array_equal(got[0], want[0])
        np.testing.assert_array_equal(got[1], want[1])


def test_sort_rows_on_arrival(ctx):
    # unsorted CSR / padded input: sorted once on the device; sorted input is left alone; a borrowed device CSR is
    # never written to
    import torch

    x = datasets.md_like_3d(4000)
    rp, ci = port.neighbours(x, 0.4 ** 2)
    want = port.fit_csr(rp, ci, 6, member_cutoff=5)
    rng = np.random.default_rng(0)
    sh = ci.copy()
    for i in range(x.shape[0]):
        rng.shuffle(sh[rp[i]:rp[i + 1]])
    ctx.neighbours_from_csr(rp, ci)
    assert ctx.neighbours_sort_rows() is False
    ctx.neighbours_from_csr(rp, sh)                      # lists in any order are sorted by the ingest itself
    assert ctx.neighbours_sort_rows() is False
    got_rp, got_ci = ctx.neighbours_export()
    np.testing.assert_array_equal(got_ci, ci)
    lab = np.zeros(x.shape[0], dtype=np.int64)
    ctx.fit(6 + 1, 6 + 2, 1, lab)
    ctx.sort_by_size(lab, member_cutoff=5)
    np.testing.assert_array_equal(lab, want)
    d_rp, d_sh = torch.from_numpy(rp).cuda(), torch.from_numpy(sh.astype(np.int32)).cuda()
    keep = d_sh.clone()
    ctx.neighbours_from_csr(d_rp, d_sh, borrow=True)
    assert ctx.neighbours_sort_rows() is False
    assert torch.equal(d_sh, keep)
    np.testing.assert_array_equal(ctx.neighbours_export()[1], ci)
    # through the API: list of shuffled lists, the reference's default lookup getter (is_sorted=False)
    nb = [sh[rp[i]:rp[i + 1]] for i in range(x.shape[0])]
    cl = fit_via_api(nb, 0.4, 6, recipe="neighbourhoods", member_cutoff=5)
    np.testing.assert_array_equal(cl.labels, want)


def test_lookup_without_selfcounting(ctx):
    # lists without the point itself: n_member_cutoff = c, similarity_cutoff = c (_fit.pyx:143-151)
    x = datasets.blobs_2d(2000)
    rp, ci = port.neighbours(x, 0.12 ** 2)
    keep = ci != np.repeat(np.arange(x.shape[0]), np.diff(rp))
    ci2 = ci[keep]
    rp2 = np.zeros_like(rp)
    np.cumsum(np.diff(rp) - 1, out=rp2[1:])
    for c in (0, 3, 10):
        want = port.fit_csr(rp2, ci2, c, selfcounting=False)
        ctx.neighbours_from_csr(rp2, ci2)
        lab = np.zeros(x.shape[0], dtype=np.int64)
        ctx.fit(c, c, 1, lab)
        ctx.sort_by_size(lab)
        np.testing.assert_array_equal(lab, want)


def test_ingest_rejects_bad_input(ctx):
    rp = np.array([0, 2, 3], dtype=np.int64)
    with pytest.raises(ValueError):
        ctx.neighbours_from_csr(rp, np.array([0, 5, 1], dtype=np.int64))  # index out of range
    with pytest.raises(ValueError):
        ctx.neighbours_from_csr(np.array([0, 3, 2], dtype=np.int64), np.array([0, 1, 1], dtype=np.int64))
    with pytest.raises(ValueError):
        ctx.neighbours_from_padded(np.zeros((2, 2), dtype=np.int64), np.array([3, 0], dtype=np.int64))
    with pytest.raises(ValueError):
        ctx.neighbours_from_points(np.array([[0.0, np.nan]]), 1.0)
    with pytest.raises(_lib.CnnB200Error):
        _lib.Context(0).fit(1, 2, 1, np.zeros(3, dtype=np.int64))  # no neighbourhoods yet


# ---- similarity checker ----------------------------------------------------------------------------------
@pytest.mark.parametrize("a,b,c,expected", [
    # reference test/unit/test_check_similarity.py:41-48 (sorted cases)
    ([], [], 0, True),
    ([], [], 1, False),
    ([1, 2, 3], [2, 5], 1, True),
    ([1, 2, 3], [2, 5, 8, 9], 2, False),
    ([1, 2, 3, 9, 10], [2, 5, 8, 9, 10], 3, True),
])
def test_check_truth_table(a, b, c, expected):
    chk = _types.SimilarityCheckerCommonNN()
    assert chk.check(a, b, _types.ClusterParameters(0.0, c)) == expected


def test_common_neighbour_counts_equal_oracle(ctx):
    x = datasets.blobs_2d(2500)
    rp, ci = port.neighbours(x, 0.15 ** 2)
    ctx.neighbours_from_csr(rp, ci)
    cutoff = 6
    counts, n_edges = ctx.cn_count(cutoff)
    deg = np.diff(rp)
    rows = np.repeat(np.arange(x.shape[0]), deg)
    is_edge = (ci > rows) & (deg[rows] > cutoff) & (deg[ci] > cutoff)
    assert n_edges == int(is_edge.sum())
    assert (counts[~is_edge] == 0xFFFFFFFF).all()
    idx = np.flatnonzero(is_edge)
    pick = idx[:: max(1, idx.size // 3000)]
    for k in pick:
        i, j = rows[k], ci[k]
        assert counts[k] == port.count_common(ci[rp[i]:rp[i + 1]], ci[rp[j]:rp[j + 1]])
    # a row block only (row-sharded runs): the block's slots equal the full pass, the others are left alone
    import torch

    part = torch.full((ci.shape[0],), 7, dtype=torch.int32, device="cuda")
    b, e = 700, 1900
    _, n_block = ctx.cn_count(cutoff, part, rows=(b, e))
    part = part.cpu().numpy().view(np.uint32)
    in_block = (rows >= b) & (rows < e)
    np.testing.assert_array_equal(part[in_block], counts[in_block])
    assert (part[~in_block] == 7).all() and n_block == int((is_edge & in_block).sum())
    # long lists (reference benchmark case equal_c50: identical lists, docsrc/benchmark/bm_similarity.ipynb)
    chk = _types.SimilarityCheckerCommonNN()
    a = np.arange(20000)
    assert chk.get(a, a) == 20000
    assert chk.get(a[::2], a[1::2]) == 0


# ---- fitter: frozen labels / consider masks / numbering ----------------------------------------------------
def test_partial_consider_mask_and_current_start(ctx):
    x = datasets.blobs_2d(2500)
    rp, ci = port.neighbours(x, 0.1 ** 2)
    rng = np.random.default_rng(5)
    consider = (rng.random(x.shape[0]) < 0.7).astype(np.uint8)
    base = np.where(consider == 0, 3, 0).astype(np.int64)  # "frozen" points carry label 3
    want = port.fit_bfs(rp, ci, 6, 7, current_start=4, labels=base.copy(), consider=consider.copy())
    ctx.neighbours_from_csr(rp, ci)
    lab, con = base.copy(), consider.copy()
    ctx.fit(6, 7, 4, lab, con)
    np.testing.assert_array_equal(lab, want)
    assert not con.any()


def test_frozen_labels_through_clustering_api():
    x = datasets.blobs_2d(2000)
    cl = cluster.Clustering(x)
    cl.fit(0.1, 10, v=False)
    first = cl.labels.copy()
    cl.root._labels.meta["frozen"] = True
    cl.root._labels.consider[:] = (first == 0).astype(np.uint8)  # recluster the noise, keep the rest
    cl.fit(0.2, 3, v=False, sort_by_size=False)
    rp, ci = port.neighbours(x, 0.2 ** 2)
    want = port.fit_bfs(rp, ci, 4, 5, current_start=int(first.max()) + 1, labels=first.copy(),
                        consider=(first == 0).astype(np.uint8))
    np.testing.assert_array_equal(cl.labels, want)


@pytest.mark.parametrize("kwargs,expected", [
    # reference test/unit/test_types.py:98-117
    ({"member_cutoff": 1}, [1, 1, 3, 3, 0, 1, 2, 2, 2, 1, 4]),
    ({}, [1, 1, 3, 3, 0, 1, 2, 2, 2, 1, 0]),
    ({"member_cutoff": 2, "max_clusters": 2}, [1, 1, 0, 0, 0, 1, 2, 2, 2, 1, 0]),
])
def test_sort_by_size_known_answers(kwargs, expected):
    lab = _types.Labels(np.array([1, 1, 2, 2, 0, 1, 3, 3, 3, 1, 5], dtype=np.int64))
    lab.meta["params"] = {1: "a", 2: "b", 3: "c", 5: "e"}
    lab.sort_by_size(**kwargs)
    assert lab.labels.tolist() == expected
    # meta["params"] is re-keyed like the labels (_types.pyx:268-273)
    want = {e: v for o, e, v in [(1, expected[0], "a"), (2, expected[2], "b"), (3, expected[6], "c"), (5, expected[10], "e")] if e}
    assert lab.meta["params"] == want


def test_sort_by_size_random_equals_oracle(ctx):
    rng = np.random.default_rng(3)
    for trial in range(10):
        n = int(rng.integers(1, 50000))
        lab = rng.integers(0, max(2, int(rng.integers(2, 400))), n).astype(np.int64)
        if trial % 3 == 0:
            lab[rng.random(n) < 0.5] = 7  # one dominant label: heavy same-address atomics
        mc = [None, 1, 5, 50][trial % 4]
        mx = [None, 3, None, 10][trial % 4]
        want = port.sort_by_size(lab.copy(), mc, mx)
        got = lab.copy()
        ctx.sort_by_size(got, mc, mx)
        np.testing.assert_array_equal(got, want)
    zero = np.zeros(10, dtype=np.int64)
    ctx.sort_by_size(zero)
    assert not zero.any()


def test_random_configs_equal_oracle():
    rng = np.random.default_rng(11)
    for trial in range(25):
        n = int(rng.integers(1, 3000))
        d = int(rng.choice([1, 2, 3, 4, 6]))
        x = rng.standard_normal((n, d)) if trial % 3 else cases.lattice(n, d, trial)
        r = float(rng.uniform(0.1, 0.8)) * (1.0 if d <= 3 else np.sqrt(d) / 1.5)
        c = int(rng.integers(0, 12))
        opts = [dict(), dict(member_cutoff=5), dict(sort_by_size=False), dict(member_cutoff=1, max_clusters=3)][trial % 4]
        o = dict(opts)
        want = port.fit_points(x, r, c, sort=o.pop("sort_by_size", True), **o)
        got = fit_via_api(x, r, c, **opts).labels
        np.testing.assert_array_equal(got, want, err_msg=f"trial {trial} n={n} d={d} r={r} c={c} {opts}")


def test_vendored_scan_records():
    """All sampled cells of the reference's vendored scan records (docsrc/benchmark/records)."""
    with open(os.path.join(HERE, "golden", "records_sample.json")) as fp:
        sample = json.load(fp)
    for name, recs in sample.items():
        x = cases.RECORD_SETS[name]()
        cl = cluster.Clustering(x)
        for rec in recs:
            cl.fit(rec["radius_cutoff"], rec["cnn_cutoff"], member_cutoff=rec["member_cutoff"], v=False)
            got = cl.summary[-1]
            assert got.n_clusters == rec["n_clusters"], (name, rec)
            assert abs(got.ratio_noise - rec["ratio_noise"]) < 1e-12, (name, rec)
            assert abs(got.ratio_largest - rec["ratio_largest"]) < 1e-12, (name, rec)


# ---- device-resident inputs --------------------------------------------------------------------------------
def test_device_resident_inputs_equal_host_path(ctx):
    import torch

    x = datasets.md_like_3d(30000)
    want = np.empty(x.shape[0], dtype=np.int64)
    ctx.fit_points(x, 0.2 ** 2, 11, 12, want, member_cutoff=10)
    xd = torch.from_numpy(x).cuda()
    lab = torch.empty(x.shape[0], dtype=torch.int64, device="cuda")
    ctx.set_stream(torch.cuda.current_stream())
    ctx.fit_points(xd, 0.2 ** 2, 11, 12, lab, member_cutoff=10)
    np.testing.assert_array_equal(lab.cpu().numpy(), want)
    np.testing.assert_array_equal(want, port.fit_points(x, 0.2, 10, member_cutoff=10))
    ctx.set_stream(None)


def test_torch_allocator_hook():
    import torch

    c2 = _lib.Context(0, torch_allocator=True)
    x = datasets.blobs_2d(3000)
    lab = np.empty(3000, dtype=np.int64)
    before = torch.cuda.memory_allocated()
    c2.fit_points(x, 0.01, 11, 12, lab)
    assert torch.cuda.memory_allocated() > before  # scratch came from torch's caching allocator
    np.testing.assert_array_equal(lab, port.fit_points(x, 0.1, 10))
    c2.close()


# ---- full-size configurations: properties that do not need the oracle's answer ---------------------------------
def test_c2_full_size_properties_and_oracle(ctx):
    """BASELINE configs[1]: N = 1e6 3-D MD-like, r = 0.09, c = 20, member_cutoff = 10."""
    n = 1_000_000
    x = datasets.md_like_3d(n)
    p = datasets.C2_PARAMS
    r, c = p["radius_cutoff"], p["cnn_cutoff"]
    lab = np.empty(n, dtype=np.int64)
    ctx.fit_points(x, r * r, c + 1, c + 2, lab, member_cutoff=p["member_cutoff"])
    _, nnz, max_deg = ctx.neighbours_info()
    assert (nnz, max_deg) == (72_661_402, 484)      # SURVEY 8(d), measured with the reference
    assert int(lab.max()) == 167
    # idempotence: relabelling canonical labels is the identity
    again = lab.copy()
    ctx.sort_by_size(again, p["member_cutoff"])
    np.testing.assert_array_equal(again, lab)
    # sizes are non-increasing in the label and every kept cluster has >= member_cutoff members
    sizes = np.bincount(lab)[1:]
    assert (np.diff(sizes) <= 0).all() and sizes.min() >= p["member_cutoff"]
    # invariance under a permutation of the input points (labels follow the points; numbering is
    # by size, ties by first occurrence, so compare partitions)
    perm = np.random.default_rng(1).permutation(n)
    lab_p = np.empty(n, dtype=np.int64)
    ctx.fit_points(np.ascontiguousarray(x[perm]), r * r, c + 1, c + 2, lab_p, member_cutoff=p["member_cutoff"])
    assert ((lab_p == 0) == (lab[perm] == 0)).all()
    pairs = np.unique(np.stack([lab[perm], lab_p]), axis=1)
    assert pairs.shape[1] == 168 and np.unique(pairs[0]).size == 168 and np.unique(pairs[1]).size == 168
    np.testing.assert_array_equal(np.bincount(lab_p), np.bincount(lab))
    # and the oracle itself (grid neighbour search + BFS, ~10 s on one core)
    np.testing.assert_array_equal(lab, port.fit_points(x, r, c, member_cutoff=p["member_cutoff"]))


def test_c1_through_lookup_equals_coordinates():
    """BASELINE configs[0] by both routes the reference offers (coordinates / sorted neighbourhoods)."""
    x = datasets.blobs_2d(10_000)
    a = fit_via_api(x, 0.3, 10).labels
    rp, ci = port.neighbours(x, 0.3 ** 2)
    b = fit_via_api((rp, ci), 0.3, 10, recipe="neighbourhoods_csr").labels
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(a, GOLD["c1_full"].astype(np.int64))


# ---- parameter sweeps and hierarchy --------------------------------------------------------
def test_sweep_cells_equal_oracle(ctx):
    from commonnnclustering_b200 import sweep

    x = datasets.blobs_2d(3000)
    rs, cs = [0.08, 0.12, 0.2], [0, 3, 10, 25]
    labels, summary = sweep.fit_sweep(x, rs, cs, member_cutoff=5, ctx=ctx)
    assert labels.shape == (3, 4, 3000) and len(summary) == 12
    k = 0
    for ir, r in enumerate(rs):
        for ic, c in enumerate(cs):
            want = port.fit_points(x, r, c, member_cutoff=5)
            np.testing.assert_array_equal(labels[ir, ic], want, err_msg=f"cell r={r} c={c}")
            rec = summary[k]
            k += 1
            assert (rec.radius_cutoff, rec.cnn_cutoff, rec.n_clusters) == (r, c, int(want.max()))
            assert abs(rec.ratio_noise - float((want == 0).mean())) < 1e-15


def test_sweep_with_the_periodic_metric_equals_the_oracle(ctx):
    """fit_sweep takes the metric OBJECT: the periodic metric squares the radius (`adjust_radius`) and brings its bounds
    along (a sweep with metric.name only searched with the unsquared radius and stale bounds)."""
    from commonnnclustering_b200 import _types, sweep

    rng = np.random.default_rng(12)
    bounds = [2.0, 3.0]
    x = np.ascontiguousarray(rng.random((4000, 2)) * np.array(bounds))
    x[:1500] = x[:1500] * 0.15 + np.array([1.9, 2.9])       # a dense patch that wraps around both boundaries
    metric = _types.MetricEuclideanPeriodicReduced(bounds)
    rs, cs = [0.05, 0.08], [2, 8]
    labels, summary = sweep.fit_sweep(x, rs, cs, member_cutoff=4, metric=metric, ctx=ctx)
    for ir, r in enumerate(rs):
        for ic, c in enumerate(cs):
            want = port.fit_points_periodic(x, r, c, bounds, member_cutoff=4)
            np.testing.assert_array_equal(labels[ir, ic], want, err_msg=f"cell r={r} c={c}")
    with pytest.raises(ValueError):
        sweep.fit_sweep(x, rs, cs, metric="euclidean_periodic_r", ctx=ctx)   # a name cannot carry the bounds


def test_sweep_3d_equals_one_fit_per_cell(ctx):
    from commonnnclustering_b200 import sweep

    x = datasets.md_like_3d(120_000)
    rs, cs = [0.12, 0.2], [2, 20, 60]
    labels, _ = sweep.fit_sweep(x, rs, cs, member_cutoff=10, ctx=ctx)
    single = np.empty(x.shape[0], dtype=np.int64)
    for ir, r in enumerate(rs):
        for ic, c in enumerate(cs):
            ctx.fit_points(x, r * r, c + 1, c + 2, single, member_cutoff=10)
            np.testing.assert_array_equal(labels[ir, ic], single, err_msg=f"cell r={r} c={c}")


def test_hierarchical_recluster_equals_reference():
    """fit -> isolate -> recluster the children -> reel, against the reference doing the same."""
    from oracle import ref

    if not ref.available():
        pytest.skip("oracle/_ref not built")
    rmod = ref.module()
    x = datasets.blobs_2d(1500)
    out = []
    for mod in (cluster, rmod):
        cl = mod.Clustering(x)
        cl.fit(0.25, 5, member_cutoff=10, v=False)
        cl.isolate()
        for label, child in cl.root._children.items():
            if child._input_data.n_points >= 100:
                cl.fit(0.12, 12, member_cutoff=5, v=False, bundle=child)
        cl.reel()
        out.append(np.array(cl.labels))
    np.testing.assert_array_equal(out[0], out[1])
    assert out[0].max() > 2


# ---- prediction --------------------------------------------------------------------------------
@pytest.mark.parametrize("maker,r,c,r2,c2,method", [
    (lambda: datasets.blobs_2d(2400), 0.12, 8, 0.15, 5, "auto"),          # cell list, 2-D
    (lambda: datasets.md_like_3d(3000), 0.45, 6, 0.5, 4, "auto"),         # cell list, 3-D
    (lambda: cases.gauss(1600, 5, 41) * 0.6, 0.9, 6, 1.0, 4, "tiled"),    # any dimension: all-pairs f64
    (lambda: cases.lattice(1800, 2, 42), 0.25, 3, 0.25, 2, "auto"),       # exact ties
])
def test_predict_equals_oracle(maker, r, c, r2, c2, method):
    pts = maker()
    x, x_other = np.ascontiguousarray(pts[::2]), np.ascontiguousarray(pts)   # the reference's own test layout: stride 2
    cl = cluster.Clustering(x, fitter__ngetter__method=method, predictor__ngetter__method=method) if method != "auto" \
        else cluster.Clustering(x)
    cl.fit(r, c, member_cutoff=5, v=False)
    other = cluster.Clustering(x_other)
    cl.predict(other, r2, c2)
    want = port.predict_points(x, cl.labels, x_other, r2, c2)
    np.testing.assert_array_equal(other.labels, want)
    assert want.max() > 0 and (other.root._labels.consider[want > 0] == 0).all()
    assert other.root._labels.meta["frozen"] and other.root._labels.meta["origin"] == "predict"
    # restricted to one cluster, on frozen labels: a second call only fills points that are still considered
    other2 = cluster.Clustering(x_other)
    cl.predict(other2, r2, c2, clusters=[1])
    want1 = port.predict_points(x, cl.labels, x_other, r2, c2, clusters={1})
    np.testing.assert_array_equal(other2.labels, want1)


def test_predict_from_neighbourhoods_equals_oracle():
    """The reference's own predict test layout (test/integration/test_procedures.py:219-277): both sets given as
    neighbourhoods, the other set's neighbourhoods being those of its points IN the clustered set."""
    pts = datasets.blobs_2d(2000)
    x, x_other = np.ascontiguousarray(pts[::2]), np.ascontiguousarray(pts)
    r, c = 0.15, 4
    rp, ci = port.neighbours(x, r * r)
    orp, oci = port.neighbours_other(x, x_other, r * r)
    labels = port.fit_points(x, 0.12, 8, member_cutoff=5)
    cl = cluster.Clustering(_types.InputDataNeighbourhoodsCSR(rp, ci), recipe="neighbourhoods_csr",
                            predictor="b200", predictor__ngetter=("lookup", (), {"is_sorted": True}))
    cl.labels = labels.copy()
    other = cluster.Clustering(_types.InputDataNeighbourhoodsCSR(orp, oci), recipe="neighbourhoods_csr")
    cl.predict(other, r, c)
    want, _ = port.predict_csr(rp, ci, orp, oci, labels, c + 1, c + 2)
    np.testing.assert_array_equal(other.labels, want)


def test_predict_equals_reference():
    from oracle import ref

    if not ref.available():
        pytest.skip("oracle/_ref not built")
    rmod = ref.module()
    from cnnclustering import _fit as rfit, _types as rtypes, recipes as rrecipes

    pts = datasets.blobs_2d(1200)
    x, x_other = np.ascontiguousarray(pts[::2]), np.ascontiguousarray(pts)
    recipe = dict(rrecipes.get_registered_recipe("coordinates"))
    # types instead of names: the reference's name table has no entry for the `*_other` / `neighbour_neighbours` keywords
    recipe.update({"predictor": rfit.PredictorFirstmatch, "predictor.ngetter": rtypes.NeighboursGetterExtBruteForce,
                   "predictor.ngetter.dgetter": rtypes.DistanceGetterExtMetric,
                   "predictor.ngetter.dgetter.metric": rtypes.MetricExtEuclideanReduced,
                   "predictor.ongetter": rtypes.NeighboursGetterExtBruteForce,
                   "predictor.ongetter.dgetter": rtypes.DistanceGetterExtMetric,
                   "predictor.ongetter.dgetter.metric": rtypes.MetricExtEuclideanReduced,
                   "predictor.na": rtypes.NeighboursExtVectorCPPUnorderedSet,
                   "predictor.nb": rtypes.NeighboursExtVectorCPPUnorderedSet,
                   "predictor.checker": rtypes.SimilarityCheckerExtSwitchContains})
    out = []
    for mod, kw in ((cluster, {}), (rmod, {"recipe": recipe})):
        cl = mod.Clustering(x, **kw)
        cl.fit(0.15, 8, member_cutoff=5, v=False)
        other = mod.Clustering(x_other, **kw)
        cl.predict(other._bundle, 0.18, 5)
        out.append(np.array(other.labels))
    np.testing.assert_array_equal(out[0], out[1])
    assert out[0].max() > 0


# ---- periodic metric -----------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["dihedral2d", "mixed3d_many_periods", "ring1d", "ring1d_long_rows", "lattice3d_ties",
                                  "short_period"])
def test_periodic_cell_list_neighbour_sets_are_exact(ctx, name):
    """The periodic cell list (images as a filter + the reference's arithmetic on the original coordinates) against the
    oracle's brute force under MetricExtEuclideanPeriodicReduced: identical rows, for coordinates outside [0, L),
    negative ones, a non-periodic dimension, distances exactly on the cutoff across the boundary, and a period too
    short for the image construction (falls through to the all-pairs kernel)."""
    rng = np.random.default_rng(17)
    n = 6000
    if name == "dihedral2d":      # angles in [-pi, pi), as MD dihedrals come
        x = np.c_[rng.vonmises(2.9, 1.5, n), rng.vonmises(-3.0, 2.0, n)]
        bounds, r = np.array([2 * np.pi, 2 * np.pi]), 0.12
    elif name == "mixed3d_many_periods":   # y is not periodic; x and z spill over several periods on both sides
        x = np.c_[rng.random(n) * 50.0 - 20.0, rng.standard_normal(n), rng.random(n) * 21.0 - 7.0]
        bounds, r = np.array([10.0, 0.0, 7.0]), 0.6
    elif name == "ring1d":
        x = (rng.random((n, 1)) * 3.0) - 1.0
        bounds, r = np.array([1.0]), 0.004
    elif name == "ring1d_long_rows":   # 2400 neighbours per point: rows with images beyond the shared-memory staging
        x = (rng.random((n, 1)) * 3.0) - 1.0
        bounds, r = np.array([1.0]), 0.2
    elif name == "lattice3d_ties":  # lattice spacing 1/8, period 3, r = 3/8: many pairs exactly ON the cutoff, also across faces
        x = rng.integers(0, 24, (n, 3)) / 8.0 + rng.integers(-2, 3, (n, 3)) * 3.0
        bounds, r = np.array([3.0, 3.0, 3.0]), 0.375
    else:                          # period 0.5 < 2 r: served by the all-pairs kernel
        x = rng.random((n, 2))
        bounds, r = np.array([0.5, 1.0]), 0.26
    x = np.ascontiguousarray(x, dtype=np.float64)
    ctx.set_periodic_bounds(bounds)
    want_rp, want_ci = port.neighbours_periodic(x, r * r, bounds)
    for method in ("auto", "grid", "tiled"):
        ctx.neighbours_from_points(x, r * r, metric="euclidean_periodic_r", method=method)
        rp, ci = ctx.neighbours_export()
        np.testing.assert_array_equal(rp, want_rp)
        np.testing.assert_array_equal(ci, want_ci)
    # and the labels through the API
    cl = cluster.Clustering(x, fitter__ngetter__dgetter__metric=(_types.MetricEuclideanPeriodicReduced, (bounds,), {}))
    cl.fit(r, 4, member_cutoff=3, v=False)
    np.testing.assert_array_equal(cl.labels, port.fit_points_periodic(x, r, 4, bounds, member_cutoff=3))


def test_periodic_metric_equals_oracle_and_reference(ctx):
    rng = np.random.default_rng(3)
    bounds = np.array([2 * np.pi, 0.0, 2 * np.pi])
    x = np.ascontiguousarray(np.c_[rng.vonmises(3.0, 2.0, 2500) + np.pi, rng.standard_normal(2500) * 0.3,
                                   rng.random(2500) * 2 * np.pi])
    ctx.set_periodic_bounds(bounds)
    ctx.neighbours_from_points(x, 0.35 ** 2, metric="euclidean_periodic_r")
    rp, ci = ctx.neighbours_export()
    want_rp, want_ci = port.neighbours_periodic(x, 0.35 ** 2, bounds)
    np.testing.assert_array_equal(rp, want_rp)
    np.testing.assert_array_equal(ci, want_ci)
    cl = cluster.Clustering(x, fitter__ngetter__dgetter__metric=(_types.MetricEuclideanPeriodicReduced, (bounds,), {}))
    cl.fit(0.35, 8, member_cutoff=5, v=False)
    np.testing.assert_array_equal(cl.labels, port.fit_points_periodic(x, 0.35, 8, bounds, member_cutoff=5))
    with pytest.raises(ValueError):
        ctx.neighbours_from_points(x, 0.1, metric="euclidean_periodic_r", method="tensor")
    from oracle import ref

    if ref.available():
        rmod = ref.module()
        from cnnclustering import _types as rtypes, recipes as rrecipes

        recipe = dict(rrecipes.get_registered_recipe("coordinates"))
        recipe["fitter.ngetter.dgetter.metric"] = (rtypes.MetricExtEuclideanPeriodicReduced, (bounds,), {})
        rcl = rmod.Clustering(x[:1200], recipe=recipe)
        rcl.fit(0.45, 6, member_cutoff=5, v=False)
        ours = cluster.Clustering(x[:1200], fitter__ngetter__dgetter__metric=(_types.MetricEuclideanPeriodicReduced, (bounds,), {}))
        ours.fit(0.45, 6, member_cutoff=5, v=False)
        np.testing.assert_array_equal(ours.labels, np.array(rcl.labels))


@pytest.mark.gpu
def test_fused_numbering_and_size_sort(ctx):
    """cnnb200_fit_points ranks the components in one CTA when few survive member_cutoff and falls back to the
    general chain (numbering + sort_by_size over all labels) otherwise; both equal the oracle."""
    # 20 000 clusters of two points each: more survivors than one CTA ranks -> fallback
    base = np.arange(20_000, dtype=np.float64)[:, None] * 10.0
    x = np.concatenate([base, base + 0.1]).reshape(-1, 1)
    rng = np.random.default_rng(5)
    x = np.ascontiguousarray(x[rng.permutation(x.shape[0])])
    want = port.fit_points(x, 0.5, 0, member_cutoff=None)
    got = np.zeros(x.shape[0], dtype=np.int64)
    nc = ctx.fit_points(x, 0.25, 1, 2, got)
    assert nc == 20_000 and np.array_equal(got, want) and got.max() == 20_000
    # 6 000 such clusters: ranked by the one CTA (ties in size broken by first occurrence)
    x6 = np.ascontiguousarray(x[np.isin((x[:, 0] // 10).astype(np.int64), np.arange(6000))])
    want = port.fit_points(x6, 0.5, 0, member_cutoff=None)
    got = np.zeros(x6.shape[0], dtype=np.int64)
    assert ctx.fit_points(x6, 0.25, 1, 2, got) == 6000 and np.array_equal(got, want)
    # ties in size are broken by first occurrence; max_clusters and member_cutoff cut the tail
    x = datasets.md_like_3d(60_000)
    for mc, mx in [(None, None), (10, None), (5, 7), (1, 3), (50, 1)]:
        want = port.fit_points(x, 0.2, 5, member_cutoff=mc, max_clusters=mx)
        got = np.zeros(x.shape[0], dtype=np.int64)
        ctx.fit_points(x, 0.04, 6, 7, got, member_cutoff=mc, max_clusters=mx)
        assert np.array_equal(got, want), (mc, mx)
        kept, largest, noise = ctx.label_summary()
        cnt = np.bincount(want)
        assert (kept, largest, noise) == (len(cnt) - 1, int(cnt[1:].max()) if len(cnt) > 1 else 0, int(cnt[0]))


@pytest.mark.gpu
def test_predict_with_lists_in_any_order(ctx):
    """The 'neighbourhoods' recipe hands the predictor lists in ANY order (NeighboursGetterExtLookup, is_sorted = False,
    recipes.py:340-348): the label comes from the first matching member IN LIST ORDER (_fit.pyx:1264-1291), whatever
    the order, and malformed lists are rejected instead of read out of bounds."""
    rng = np.random.default_rng(11)
    pts = datasets.blobs_2d(1600)
    x, x_other = np.ascontiguousarray(pts[::2]), np.ascontiguousarray(pts[1::2])
    r, c = 0.15, 4
    rp, ci = port.neighbours(x, r * r)
    orp, oci = port.neighbours_other(x, x_other, r * r)
    labels = port.fit_points(x, 0.12, 8, member_cutoff=5)
    # shuffle every list of both sets
    def shuffled(rp_, ci_):
        out = ci_.copy()
        for a, b in zip(rp_[:-1], rp_[1:]):
            out[a:b] = rng.permutation(out[a:b])
        return out
    ci_s, oci_s = shuffled(rp, ci), shuffled(orp, oci)
    # expected: the reference's loop with set intersections (order-free counts), members walked in list order
    rows = [set(ci[a:b].tolist()) for a, b in zip(rp[:-1], rp[1:])]
    want = np.zeros(x_other.shape[0], dtype=np.int64)
    for p in range(x_other.shape[0]):
        na = oci_s[orp[p]:orp[p + 1]]
        if not na.size > c + 1:
            continue
        sa = set(na.tolist())
        for member in na:
            if labels[member] == 0 or not len(rows[member]) > c + 1:
                continue
            if len(sa & rows[member]) >= c + 2:
                want[p] = labels[member]
                break
    ctx.neighbours_from_csr(rp, ci_s)          # sorted on arrival
    ctx.neighbours_other_from_csr(orp, oci_s)
    got = np.zeros(x_other.shape[0], dtype=np.int64)
    consider = np.ones(x_other.shape[0], dtype=np.uint8)
    label_ok = np.ones(int(labels.max()) + 1, dtype=np.uint8)
    label_ok[0] = 0
    ctx.predict(c + 1, c + 2, labels, label_ok, got, consider)
    np.testing.assert_array_equal(got, want)
    assert want.max() > 0 and (consider[want > 0] == 0).all()
    # the first match in list order differs from the smallest index for some point (otherwise the test is vacuous)
    ctx.neighbours_other_from_csr(orp, oci)
    got_sorted = np.zeros_like(got)
    ctx.predict(c + 1, c + 2, labels, label_ok, got_sorted, np.ones_like(consider))
    assert (got_sorted > 0).sum() == (got > 0).sum()
    # malformed predictand lists
    bad = oci.copy()
    bad[5] = x.shape[0]
    with pytest.raises(ValueError):
        ctx.neighbours_other_from_csr(orp, bad)
    bad_rp = orp.copy()
    bad_rp[3] = bad_rp[4] + 1
    with pytest.raises(ValueError):
        ctx.neighbours_other_from_csr(bad_rp, oci)


@pytest.mark.gpu
def test_unsorted_lists_are_sorted_on_every_ingest_path(ctx):
    """Ascending rows are a REQUIREMENT of the intersection kernels; no ingest path trusts the caller for it."""
    rng = np.random.default_rng(3)
    x = datasets.md_like_3d(4000)
    rp, ci = port.neighbours(x, 0.3 * 0.3)
    want = port.fit_csr(rp, ci, 5, member_cutoff=None)
    ci_s = ci.copy()
    for a, b in zip(rp[:-1], rp[1:]):
        ci_s[a:b] = rng.permutation(ci_s[a:b])
    import torch
    for form in ("host64", "host32", "device_borrowed"):
        if form == "host64":
            ctx.neighbours_from_csr(rp, ci_s)
        elif form == "host32":
            ctx.neighbours_from_csr(rp, ci_s.astype(np.int32))
        else:
            rp_d, ci_d = torch.from_numpy(rp).cuda(), torch.from_numpy(ci_s.astype(np.int32)).cuda()
            keep = ci_d.clone()
            ctx.neighbours_from_csr(rp_d, ci_d, borrow=True)
        got = np.zeros(x.shape[0], dtype=np.int64)
        ctx.fit(6, 7, 1, got, None)
        ctx.sort_by_size(got, None)
        assert np.array_equal(got, want), form
        if form == "device_borrowed":
            assert torch.equal(ci_d, keep)     # the caller's buffer is not modified
    padded, counts = port.csr_to_padded(rp, ci_s)
    ctx.neighbours_from_padded(padded, counts)
    got = np.zeros(x.shape[0], dtype=np.int64)
    ctx.fit(6, 7, 1, got, None)
    ctx.sort_by_size(got, None)
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_hierarchical_repeat_equals_reference(ctx):
    """Clustering.fit_hierarchical with HierarchicalFitterRepeat against the reference's own
    (_fit.pyx:1018-1184, cluster.pyx:394-411): the same tree -- label array, children and reference indices of every
    bundle."""
    from oracle import ref

    if not ref.available():
        pytest.skip("oracle/_ref not built")
    rmod = ref.module()
    x = datasets.blobs_2d(1500)
    kw = dict(radius_cutoff=[0.5, 0.25, 0.12, 0.08], cnn_cutoff=[5, 8, 8, 10], member_cutoff=5, v=False)
    theirs = rmod.Clustering(
        x, hierarchical_fitter="repeat", hierarchical_fitter__fitter="bfs", hierarchical_fitter__fitter__ngetter="brute_force",
        hierarchical_fitter__fitter__na="vuset", hierarchical_fitter__fitter__nb="vuset",
        hierarchical_fitter__fitter__checker="switch", hierarchical_fitter__fitter__queue="fifo",
        hierarchical_fitter__fitter__ngetter__dgetter="metric", hierarchical_fitter__fitter__ngetter__dgetter__metric="euclidean_r")
    theirs.fit_hierarchical(**kw)
    ours = cluster.Clustering(x, hierarchical_fitter="repeat", hierarchical_fitter__fitter="b200_staged",
                              hierarchical_fitter__fitter__ngetter="radius_b200",
                              hierarchical_fitter__fitter__ngetter__dgetter="metric",
                              hierarchical_fitter__fitter__ngetter__dgetter__metric="euclidean_r")
    ours._hierarchical_fitter._fitter._ctx = ctx
    ours.fit_hierarchical(**kw)
    nodes = [0]

    def same(a, b):
        nodes[0] += 1
        la = None if a._labels is None else np.asarray(a._labels.labels)
        lb = None if b._labels is None else np.asarray(b._labels.labels)
        assert (la is None) == (lb is None)
        if la is not None:
            np.testing.assert_array_equal(la, lb)
        ca, cb = a._children or {}, b._children or {}
        assert list(ca) == list(cb)
        for k in ca:
            np.testing.assert_array_equal(ca[k].parent_indices, np.asarray(cb[k]._reference_indices.parent))
            np.testing.assert_array_equal(ca[k].root_indices, np.asarray(cb[k]._reference_indices.root))
            same(ca[k], cb[k])

    same(ours._bundle, theirs._bundle)
    assert nodes[0] > 5
    # the default hierarchical fitter wraps the clustering's own fitter; sequences may be given as scalars
    cl = cluster.Clustering(x)
    cl.fitter._ctx = ctx
    cl.fit_hierarchical(radius_cutoff=0.25, cnn_cutoff=[8, 20], member_cutoff=5, v=False)
    assert cl.root._children and cl.root._labels is not None
    # the device grouping primitive on its own: order and rank against NumPy
    rng = np.random.default_rng(2)
    prev, cur = rng.integers(0, 7, 5000), rng.integers(0, 9, 5000)
    order, rank = ctx.labels_group(prev, cur)
    np.testing.assert_array_equal(order, np.lexsort((np.arange(5000), cur, prev)))
    want_rank = np.empty(5000, dtype=np.int64)
    for p in range(7):
        idx = np.flatnonzero(prev == p)
        want_rank[idx] = np.arange(idx.size)
    np.testing.assert_array_equal(rank, want_rank[order])


@pytest.mark.gpu
def test_symmetry_check_of_precomputed_lists(ctx):
    x = datasets.md_like_3d(3000)
    rp, ci = port.neighbours(x, 0.3 ** 2)
    ctx.neighbours_from_csr(rp, ci)
    assert ctx.neighbours_asymmetry() == 0
    # kNN-style lists: drop the last entry of every row with more than three entries
    keep = np.ones(ci.shape[0], dtype=bool)
    keep[(rp[1:] - 1)[np.diff(rp) > 3]] = False
    rp2 = np.zeros_like(rp)
    np.cumsum(np.add.reduceat(keep.astype(np.int64), rp[:-1]), out=rp2[1:])
    ci2 = ci[keep]
    ctx.neighbours_from_csr(rp2, ci2)
    rows = np.repeat(np.arange(x.shape[0]), np.diff(rp2))
    have = set(zip(rows.tolist(), ci2.tolist()))
    want = sum((j, i) not in have for i, j in have)
    assert want > 0 and ctx.neighbours_asymmetry() == want
    g = _types.NeighboursGetterLookup(is_sorted=True, check_symmetric=True)
    with pytest.raises(ValueError):
        g.get_all(_types.InputDataNeighbourhoodsCSR(rp2, ci2), ctx=ctx)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["local", "long_rows", "no_locality"])
def test_early_exit_checks_equal_oracle(ctx, kind):
    """The two early-exit forms of SimilarityChecker._check that the fit runs -- the per-lane merge of the sampling
    round (with its skip-ahead for long lists) and the whole-warp window form of the work list (bitmap window, or the
    global binary search when a row has no index locality) -- on arbitrary pairs against |N(i) ∩ N(j)| >= c."""
    rng = np.random.default_rng(21)
    if kind == "long_rows":
        x, r = datasets.blobs_2d(3000), 0.35        # hundreds of entries per row
    else:
        x, r = datasets.md_like_3d(6000), 0.45
    if kind == "no_locality":
        x = np.ascontiguousarray(x[rng.permutation(x.shape[0])])   # rows span the whole index range: no window
        x = np.concatenate([x, x + 100.0, x + 200.0, x + 300.0, x + 400.0, x + 500.0, x + 600.0, x + 700.0, x + 800.0,
                            x + 900.0, x + 1000.0, x + 1100.0])     # 72 000 points: spans beyond the 65 536-bit window
        x = np.ascontiguousarray(x[rng.permutation(x.shape[0])])
    rp, ci = port.neighbours(x, r * r)
    n = x.shape[0]
    ctx.neighbours_from_csr(rp, ci)
    rows = np.repeat(np.arange(n), np.diff(rp))
    pick = rng.choice(ci.shape[0], size=3000, replace=False)
    pairs = np.concatenate([np.stack([rows[pick], ci[pick]], 1),                       # neighbours
                            rng.integers(0, n, (500, 2))]).astype(np.int32)              # arbitrary rows
    want_cnt = np.array([np.intersect1d(ci[rp[i]:rp[i + 1]], ci[rp[j]:rp[j + 1]], assume_unique=True).size
                         for i, j in pairs])
    if kind == "long_rows":
        assert np.diff(rp).max() > 400
    for c in (0, 1, 3, 12, 40, 200):
        got = ctx.cn_check_pairs(pairs, c)
        want = (want_cnt >= c) | (c == 0)
        np.testing.assert_array_equal(got & 1, want.astype(np.uint8), err_msg=f"per-lane merge, c={c}")
        np.testing.assert_array_equal(got >> 1, want.astype(np.uint8), err_msg=f"window form, c={c}")
