"""Pins the CPU oracle (oracle/cnn_oracle.c) against the reference's own known answers and
against label arrays produced by the reference itself (tests/golden/ref_labels.npz)."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import cases  # noqa: E402
from oracle import port, ref  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "ref_labels.npz"))


def test_readme_labels():
    # reference README.md:75-93
    x = cases.datasets.README_POINTS
    assert port.fit_points(x, 1.5, 1).tolist() == [1, 1, 1, 1, 1, 0, 0, 0, 0, 2, 2, 2]


@pytest.mark.parametrize("a,b,c,is_sorted,expected", [
    # reference test/unit/test_check_similarity.py:41-48
    ([], [], 0, True, True),
    ([], [], 1, True, False),
    ([1, 2, 3], [2, 5], 1, True, True),
    ([1, 2, 3], [2, 5, 8, 9], 2, True, False),
    ([1, 2, 3, 9, 10], [2, 5, 8, 9, 10], 3, True, True),
    ([3, 2, 1, 18, 9, 10], [2, 5, 9, 8, 10, 11, 18], 3, False, True),
])
def test_check_truth_table(a, b, c, is_sorted, expected):
    assert port.check(a, b, c, kind="contains") == expected
    if is_sorted:
        assert port.check(a, b, c, kind="sorted") == expected


@pytest.mark.parametrize("kwargs,expected", [
    # reference test/unit/test_types.py:98-117
    ({"member_cutoff": 1}, [1, 1, 3, 3, 0, 1, 2, 2, 2, 1, 4]),
    ({}, [1, 1, 3, 3, 0, 1, 2, 2, 2, 1, 0]),
    ({"member_cutoff": 2, "max_clusters": 2}, [1, 1, 0, 0, 0, 1, 2, 2, 2, 1, 0]),
])
def test_sort_by_size_known_answers(kwargs, expected):
    lab = np.array([1, 1, 2, 2, 0, 1, 3, 3, 3, 1, 5], dtype=np.int64)
    assert port.sort_by_size(lab, **kwargs).tolist() == expected


def test_sort_by_size_tie_order():
    # ties broken by first occurrence
    lab = np.array([2, 2, 1, 1, 3, 3, 3, 0, 5], dtype=np.int64)
    assert port.sort_by_size(lab).tolist() == [2, 2, 3, 3, 1, 1, 1, 0, 0]


def test_enough_is_strict():
    # reference test/unit/test_types.py:206-219: 4 members -> enough(3) True, enough(4) False.
    # A point with exactly n_member_cutoff+1 neighbours is core, with n_member_cutoff it is not.
    row_ptr = np.array([0, 4, 8, 12, 16], dtype=np.int64)
    col = np.tile(np.arange(4), 4).astype(np.int64)
    assert port.fit_bfs(row_ptr, col, 3, 0).tolist() == [1, 1, 1, 1]
    assert port.fit_bfs(row_ptr, col, 4, 0).tolist() == [0, 0, 0, 0]


def test_metric_matches_reference_formula():
    # reference test/unit/test_metric.py:105-157 (12 significant digits); here exact
    rng = np.random.default_rng(0)
    x = rng.standard_normal((20, 7))
    L = port.lib()
    for i, j in [(0, 1), (3, 3), (5, 19)]:
        s = 0.0
        for k in range(7):
            s += (x[i, k] - x[j, k]) ** 2
        assert L.cnn_oracle_distance(x, 7, i, j, 0) == s
        assert L.cnn_oracle_distance(x, 7, i, j, 1) == np.sqrt(s)
    assert port.adjust_radius(0.3) == 0.3 * 0.3
    assert port.adjust_radius(0.3, "euclidean") == 0.3


def test_make_parameters():
    # reference _fit.pyx:129-161 for the default coordinates recipe
    p = port.make_parameters(0.3, 10)
    assert p == {"radius_cutoff": 0.3 * 0.3, "similarity_cutoff": 12, "n_member_cutoff": 11, "current_start": 1}


@pytest.mark.parametrize("name", list(cases.CASES))
def test_port_equals_reference_golden(name):
    case = cases.CASES[name]
    x = case["data"]()
    kw = dict(case["fit"])
    sort = kw.pop("sort_by_size", True)
    method = "brute" if x.shape[0] <= 2500 else "auto"
    got = port.fit_points(x, kw.pop("r"), kw.pop("c"), sort=sort, method=method, **kw)
    np.testing.assert_array_equal(got, GOLD[name].astype(np.int64))


GOLD_RECIPES = np.load(os.path.join(HERE, "golden", "ref_labels_recipes.npz"))


def _lists_to_csr(lists):
    rows = [np.sort(np.asarray(a, dtype=np.int64)) for a in lists]
    row_ptr = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum([len(a) for a in rows], out=row_ptr[1:])
    return row_ptr, np.concatenate(rows)


@pytest.mark.parametrize("name", list(cases.RECIPE_CASES))
def test_port_equals_reference_golden_other_recipes(name):
    """The reference's "distances" recipe (thresholded N x N matrix) and "neighbourhoods" recipe (lists in any
    order: the labels only depend on the sets) against labels the reference itself produced."""
    case = cases.RECIPE_CASES[name]
    data = case["data"]()
    kw = dict(case["fit"])
    r, c = kw.pop("r"), kw.pop("c")
    if case["recipe"] == "distances":
        got = port.fit_distances(data, r, c, **kw)
    else:
        got = port.fit_csr(*_lists_to_csr(data), c, **kw)
    np.testing.assert_array_equal(got, GOLD_RECIPES[name].astype(np.int64))


def test_grid_and_brute_neighbour_sets_agree():
    for x, r in [(cases.lattice(600, 2, 3), 0.25), (cases.datasets.md_like_3d(3000), 0.5), (cases.lattice(200, 1, 5), 0.125)]:
        a = port.neighbours(x, r * r, method="brute")
        b = port.neighbours(x, r * r, method="grid")
        np.testing.assert_array_equal(a[0], b[0])
        np.testing.assert_array_equal(a[1], b[1])


def test_vendored_scan_records():
    """docsrc/benchmark/records/*.json of the reference (distilled sample): n_clusters,
    ratio_largest, ratio_noise for (r, c) cells, member_cutoff=20."""
    with open(os.path.join(HERE, "golden", "records_sample.json")) as fp:
        sample = json.load(fp)
    for name, recs in sample.items():
        x = cases.RECORD_SETS[name]()
        for rec in recs[::4]:  # every 4th cell keeps the CPU suite short; the GPU suite runs all
            lab = port.fit_points(x, rec["radius_cutoff"], rec["cnn_cutoff"], member_cutoff=rec["member_cutoff"])
            counts = np.bincount(lab)
            assert int(np.count_nonzero(counts[1:])) == rec["n_clusters"], (name, rec)
            assert abs(counts[0] / x.shape[0] - rec["ratio_noise"]) < 1e-12, (name, rec)
            largest = counts[1:].max() / x.shape[0] if counts.size > 1 else 0.0
            assert abs(largest - rec["ratio_largest"]) < 1e-12, (name, rec)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_port_equals_live_reference_on_random_configs():
    rng = np.random.default_rng(7)
    for trial in range(12):
        n = int(rng.integers(50, 500))
        d = int(rng.choice([1, 2, 3, 5]))
        x = rng.standard_normal((n, d)) if trial % 3 else cases.lattice(n, d, trial)
        r = float(rng.uniform(0.2, 0.9))
        c = int(rng.integers(0, 8))
        opts = [dict(), dict(member_cutoff=5), dict(sort_by_size=False), dict(member_cutoff=1, max_clusters=3)][trial % 4]
        want = ref.fit_points(x, r, c, **opts)
        o = dict(opts)
        sort = o.pop("sort_by_size", True)
        got = port.fit_points(x, r, c, sort=sort, **o)
        np.testing.assert_array_equal(got, want)
        rp, ci = port.neighbours(x, r * r)
        got2 = ref.fit_neighbourhoods(rp, ci, c, **opts)
        np.testing.assert_array_equal(got2, want)


def test_predict_port_equals_reference():
    """The oracle's restatement of PredictorFirstmatch.predict (_fit.pyx:1231-1291) against the reference itself."""
    from oracle import ref

    if not ref.available():
        pytest.skip("oracle/_ref not built")
    rmod = ref.module()
    from cnnclustering import _fit as rfit, _types as rtypes, recipes as rrecipes

    from commonnnclustering_b200 import datasets

    pts = datasets.blobs_2d(700)
    x, x_other = np.ascontiguousarray(pts[::2]), np.ascontiguousarray(pts)
    recipe = dict(rrecipes.get_registered_recipe("coordinates"))
    recipe.update({"predictor": rfit.PredictorFirstmatch, "predictor.ngetter": rtypes.NeighboursGetterExtBruteForce,
                   "predictor.ngetter.dgetter": rtypes.DistanceGetterExtMetric,
                   "predictor.ngetter.dgetter.metric": rtypes.MetricExtEuclideanReduced,
                   "predictor.ongetter": rtypes.NeighboursGetterExtBruteForce,
                   "predictor.ongetter.dgetter": rtypes.DistanceGetterExtMetric,
                   "predictor.ongetter.dgetter.metric": rtypes.MetricExtEuclideanReduced,
                   "predictor.na": rtypes.NeighboursExtVectorCPPUnorderedSet,
                   "predictor.nb": rtypes.NeighboursExtVectorCPPUnorderedSet,
                   "predictor.checker": rtypes.SimilarityCheckerExtSwitchContains})
    cl = rmod.Clustering(x, recipe=recipe)
    cl.fit(0.2, 6, member_cutoff=5, v=False)
    other = rmod.Clustering(x_other, recipe=recipe)
    cl.predict(other._bundle, 0.25, 4)
    labels = np.array(cl.labels)
    assert np.array_equal(labels, port.fit_points(x, 0.2, 6, member_cutoff=5))
    got = port.predict_points(x, labels, x_other, 0.25, 4)
    assert np.array_equal(got, np.array(other.labels)) and got.max() > 0
    other1 = rmod.Clustering(x_other, recipe=recipe)
    cl.predict(other1._bundle, 0.25, 4, clusters=[1])
    assert np.array_equal(port.predict_points(x, labels, x_other, 0.25, 4, clusters={1}), np.array(other1.labels))


def test_periodic_metric_port_equals_reference():
    """oracle: MetricExtEuclideanPeriodicReduced (_types.pyx:2290-2318) against the reference's own fit."""
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    rmod = ref.module()
    from cnnclustering import _types as rtypes, recipes as rrecipes

    rng = np.random.default_rng(3)
    bounds = np.array([2 * np.pi, 0.0, 2 * np.pi])
    x = np.ascontiguousarray(np.c_[rng.vonmises(3.0, 2.0, 900) + np.pi, rng.standard_normal(900) * 0.3,
                                   rng.random(900) * 2 * np.pi])
    recipe = dict(rrecipes.get_registered_recipe("coordinates"))
    recipe["fitter.ngetter.dgetter.metric"] = (rtypes.MetricExtEuclideanPeriodicReduced, (bounds,), {})
    cl = rmod.Clustering(x, recipe=recipe)
    for r, c in [(0.4, 5), (0.7, 12)]:
        cl.fit(r, c, member_cutoff=3, v=False)
        assert np.array_equal(np.array(cl.labels), port.fit_points_periodic(x, r, c, bounds, member_cutoff=3))
    # wrap-around actually matters on this data
    assert port.neighbours_periodic(x, 0.49, bounds)[0][-1] > port.neighbours(x, 0.49)[0][-1]
