"""The package's host logic behind `Clustering(...)`, exercised WITHOUT a GPU: every recipe, input type, getter, the
staged and the fused fitter, frozen labels, prediction and the hierarchy drivers run through a stand-in context
(tests/standin.py: the context's contract answered by the oracle), and the labels are compared with the golden arrays
the reference itself produced (tests/golden/) and with the live reference (`oracle/_ref`).  What is checked here is the
glue -- which call is made with which radius / cutoffs / flags for which input --, not the device: the `-m gpu` twins
of these tests (tests/test_gpu_parity.py) make the same comparisons with the real context."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, HERE)

import cases  # noqa: E402
from standin import StandInContext  # noqa: E402

from commonnnclustering_b200 import _fit, _types, cluster, datasets  # noqa: E402
from oracle import port, ref  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "ref_labels.npz"))
GOLD_RECIPES = np.load(os.path.join(HERE, "golden", "ref_labels_recipes.npz"))
SMALL = [name for name in cases.CASES if not name.startswith("c1_")]      # the 10^4-point cases take minutes on the CPU


def clustering(data, recipe=None, **kw):
    """A Clustering whose fitter and predictor talk to a stand-in context (returned with it)."""
    cl = cluster.Clustering(data, recipe=recipe, **kw)
    ctx = StandInContext()
    cl.fitter._ctx = ctx
    if cl._predictor is not None:
        cl._predictor._ctx = ctx
    return cl, ctx


@pytest.mark.parametrize("name", SMALL)
def test_fit_through_the_api_equals_reference_golden(name):
    case = cases.CASES[name]
    kw = dict(case["fit"])
    r, c = kw.pop("r"), kw.pop("c")
    metric = kw.pop("metric", None)
    cl, ctx = clustering(case["data"](), **({"fitter__ngetter__dgetter__metric": metric} if metric else {}))
    cl.fit(r, c, v=False, **kw)
    np.testing.assert_array_equal(cl.labels, GOLD[name].astype(np.int64))
    assert cl.labels.dtype == np.int64 and not cl.root._labels.consider.any()
    # a fresh fit with size sorting is ONE fused call; without it the staged route
    assert ctx.calls == (["fit_points"] if kw.get("sort_by_size", True) else ["neighbours_from_points", "fit"])
    rec = cl.summary[-1]
    want = GOLD[name]
    assert rec.n_clusters == len(set(want.tolist()) - {0}) and abs(rec.ratio_noise - float((want == 0).mean())) < 1e-15


@pytest.mark.parametrize("name", list(cases.RECIPE_CASES))
def test_other_recipes_through_the_api_equal_reference_golden(name):
    case = cases.RECIPE_CASES[name]
    kw = dict(case["fit"])
    cl, ctx = clustering(case["data"](), recipe=case["recipe"])
    cl.fit(kw.pop("r"), kw.pop("c"), v=False, **kw)
    np.testing.assert_array_equal(cl.labels, GOLD_RECIPES[name].astype(np.int64))
    first = {"distances": "neighbours_from_distances", "neighbourhoods": "neighbours_from_padded"}[case["recipe"]]
    assert ctx.calls[0] == first and ctx.calls[-1] == "fit_sorted"


def test_lookup_recipes_and_input_types():
    """Padded matrix, CSR and sorted / unsorted getters give the labels of the coordinates route."""
    x = datasets.blobs_2d(1200)
    r, c = 0.12, 6
    want = port.fit_points(x, r, c, member_cutoff=4)
    rp, ci = port.neighbours(x, r * r)
    lists = [ci[rp[i]:rp[i + 1]] for i in range(x.shape[0])]
    rng = np.random.default_rng(1)
    shuffled = [rng.permutation(row) for row in lists]
    for data, recipe in ((lists, "sorted_neighbourhoods"), (shuffled, "neighbourhoods"), ((rp, ci), "neighbourhoods_csr")):
        cl, ctx = clustering(data, recipe=recipe)
        cl.fit(r, c, member_cutoff=4, v=False)
        np.testing.assert_array_equal(cl.labels, want, err_msg=recipe)
    # the radius plays no role for precomputed lists; lists that do not count the point itself use c / c as cutoffs
    bare = [row[row != i] for i, row in enumerate(lists)]
    cl, ctx = clustering(bare, recipe="sorted_neighbourhoods", fitter__ngetter=("lookup", (), {"is_sorted": True,
                                                                                                  "is_selfcounting": False}))
    cl.fit(123.0, c, member_cutoff=4, v=False)
    rp2 = np.zeros(len(bare) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in bare], out=rp2[1:])
    np.testing.assert_array_equal(cl.labels, port.fit_csr(rp2, np.concatenate(bare), c, member_cutoff=4, selfcounting=False))


def test_condensed_distances_and_periodic_metric_routes():
    from scipy.spatial.distance import pdist, squareform

    x = datasets.blobs_2d(700)
    v = pdist(x)
    cl, ctx = clustering(_types.InputDataDistancesLinear(v), recipe="distances")
    cl.fit(0.1, 5, v=False)
    np.testing.assert_array_equal(cl.labels, port.fit_distances(squareform(v), 0.1, 5))
    assert ctx.calls[0] == "neighbours_from_condensed"
    rng = np.random.default_rng(3)
    bounds = np.array([2 * np.pi, 0.0, 2 * np.pi])
    y = np.ascontiguousarray(np.c_[rng.vonmises(3.0, 2.0, 900) + np.pi, rng.standard_normal(900) * 0.3,
                                   rng.random(900) * 2 * np.pi])
    cl, ctx = clustering(y, fitter__ngetter__dgetter__metric=(_types.MetricEuclideanPeriodicReduced, (bounds,), {}))
    cl.fit(0.4, 5, member_cutoff=3, v=False)
    np.testing.assert_array_equal(cl.labels, port.fit_points_periodic(y, 0.4, 5, bounds, member_cutoff=3))
    assert ctx.calls == ["set_periodic_bounds", "fit_points"]          # the bounds travel before the search, r is squared


def test_frozen_labels_and_consider_mask():
    x = datasets.blobs_2d(2000)
    cl, ctx = clustering(x)
    cl.fit(0.1, 10, v=False)
    first = cl.labels.copy()
    cl.root._labels.meta["frozen"] = True
    cl.root._labels.consider[:] = (first == 0).astype(np.uint8)          # recluster the noise, keep the rest
    cl.fit(0.2, 3, v=False, sort_by_size=False)
    rp, ci = port.neighbours(x, 0.2 ** 2)
    want = port.fit_bfs(rp, ci, 4, 5, current_start=int(first.max()) + 1, labels=first.copy(),
                        consider=(first == 0).astype(np.uint8))
    np.testing.assert_array_equal(cl.labels, want)
    new = set(want.tolist()) - set(first.tolist())
    assert new and all(cl.root._labels.meta["params"][k] == (0.2, 3) for k in new)
    assert all(cl.root._labels.meta["params"][k] == (0.1, 10) for k in set(first.tolist()) - {0})


def test_cnn_offset_and_purge():
    x = datasets.blobs_2d(1000)
    cl, _ = clustering(x)
    cl.fit(0.1, 12, cnn_offset=2, v=False)
    np.testing.assert_array_equal(cl.labels, port.fit_points(x, 0.1, 10))
    assert cl.summary[-1].cnn_cutoff == 10                               # the reference records the reduced value
    cl.root._labels.meta["frozen"] = True
    cl.fit(0.1, 5, purge=True, v=False)                                   # purge: frozen labels are thrown away
    np.testing.assert_array_equal(cl.labels, port.fit_points(x, 0.1, 5))


@pytest.mark.parametrize("maker,r,c,r2,c2", [
    (lambda: datasets.blobs_2d(1600), 0.12, 8, 0.15, 5),
    (lambda: cases.lattice(1200, 2, 42), 0.25, 3, 0.25, 2),             # exact ties
])
def test_predict_through_the_api_equals_oracle(maker, r, c, r2, c2):
    pts = maker()
    x, x_other = np.ascontiguousarray(pts[::2]), np.ascontiguousarray(pts)   # the reference's own test layout
    cl, _ = clustering(x)
    cl.fit(r, c, member_cutoff=5, v=False)
    other = cluster.Clustering(x_other)
    cl.predict(other, r2, c2)
    want = port.predict_points(x, cl.labels, x_other, r2, c2)
    np.testing.assert_array_equal(other.labels, want)
    assert want.max() > 0 and (other.root._labels.consider[want > 0] == 0).all()
    assert other.root._labels.meta["frozen"] and other.root._labels.meta["origin"] == "predict"
    other2 = cluster.Clustering(x_other)
    cl.predict(other2, r2, c2, clusters=[1])
    np.testing.assert_array_equal(other2.labels, port.predict_points(x, cl.labels, x_other, r2, c2, clusters={1}))


def test_predict_from_neighbourhoods_in_any_order():
    pts = datasets.blobs_2d(1600)
    x, x_other = np.ascontiguousarray(pts[::2]), np.ascontiguousarray(pts)
    r, c = 0.15, 4
    rp, ci = port.neighbours(x, r * r)
    orp, oci = port.neighbours_other(x, x_other, r * r)
    rng = np.random.default_rng(9)
    oci_shuffled = np.concatenate([rng.permutation(oci[orp[p]:orp[p + 1]]) for p in range(x_other.shape[0])])
    labels = port.fit_points(x, 0.12, 8, member_cutoff=5)
    for other_idx in (oci, oci_shuffled):
        cl, _ = clustering(_types.InputDataNeighbourhoodsCSR(rp, ci), recipe="neighbourhoods_csr", predictor="b200",
                           predictor__ngetter=("lookup", (), {"is_sorted": True}))
        cl.labels = labels.copy()
        other = cluster.Clustering(_types.InputDataNeighbourhoodsCSR(orp, other_idx), recipe="neighbourhoods_csr")
        cl.predict(other, r, c)
        # the reference's loop with order-free counts, members walked in LIST order (_fit.pyx:1264-1291)
        rows = [set(ci[a:b].tolist()) for a, b in zip(rp[:-1], rp[1:])]
        want = np.zeros(x_other.shape[0], dtype=np.int64)
        for p in range(x_other.shape[0]):
            na = other_idx[orp[p]:orp[p + 1]]
            if not na.size > c + 1:
                continue
            for member in na:
                if labels[member] != 0 and len(rows[member]) > c + 1 and len(set(na.tolist()) & rows[member]) >= c + 2:
                    want[p] = labels[member]
                    break
        np.testing.assert_array_equal(other.labels, want)
        if other_idx is oci:
            np.testing.assert_array_equal(want, port.predict_csr(rp, ci, orp, oci, labels, c + 1, c + 2)[0])


def test_similarity_checker_per_pair_forms():
    """The reference's truth table (test/unit/test_check_similarity.py:41-48) through the per-pair wrappers."""
    chk, ctx = _types.SimilarityCheckerCommonNN(), StandInContext()
    table = [([], [], 0, True), ([], [], 1, False), ([1, 2, 3], [2, 5], 1, True), ([1, 2, 3], [2, 5, 9, 8], 2, False),
             ([1, 2, 3, 9, 10], [2, 5, 9, 10], 3, True)]
    for a, b, c, expected in table:
        params = _types.ClusterParameters(0.0, c, c, 0, 1)
        assert chk.check(a, b, params, ctx=ctx) is expected
        assert chk.get(a, b, params, ctx=ctx) == len(set(a) & set(b))


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_fused_and_staged_fitters_equal_live_reference():
    """`Clustering(x)` (fused fitter) and `fitter="b200_staged"` against the reference, several fits on one object."""
    rmod = ref.module()   # puts oracle/_ref on sys.path
    from cnnclustering import recipes as rrecipes

    x = datasets.md_like_3d(1500)
    theirs = rmod.Clustering(x, recipe=dict(rrecipes.get_registered_recipe("coordinates")))
    fused, _ = clustering(x)
    staged, ctx = clustering(x, fitter="b200_staged")
    assert isinstance(fused.fitter, _fit.FitterB200Fused) and not isinstance(staged.fitter, _fit.FitterB200Fused)
    for r, c, kw in [(0.4, 5, {}), (0.3, 8, {"member_cutoff": 10}), (0.5, 3, {"max_clusters": 2, "member_cutoff": 1})]:
        for cl in (theirs, fused, staged):
            cl.fit(r, c, v=False, **kw)
        for cl in (fused, staged):
            np.testing.assert_array_equal(cl.labels, np.asarray(theirs.labels))
            a, b = cl.summary[-1], theirs.summary[-1]
            assert (a.n_clusters, a.ratio_largest, a.ratio_noise) == (b.n_clusters, b.ratio_largest, b.ratio_noise)
            assert cl.root._labels.meta["params"] == theirs._bundle._labels.meta["params"]
    assert "fit_points" not in ctx.calls and "sort_by_size" in ctx.calls


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_random_fit_sequences_equal_live_reference():
    """Forty random sequences of calls on one object -- fresh fits, fits on frozen labels, with and without size sorting,
    cut-offs, offsets, purges -- on small sets with ties: labels, records and parameter tables follow the reference's."""
    rmod = ref.module()   # puts oracle/_ref on sys.path
    from cnnclustering import recipes as rrecipes

    rng = np.random.default_rng(2026)
    for trial in range(40):
        n = int(rng.integers(40, 300))
        x = cases.lattice(n, int(rng.integers(1, 4)), 100 + trial) if trial % 3 == 0 else \
            np.ascontiguousarray(rng.standard_normal((n, int(rng.integers(1, 6)))) * 0.5)
        theirs = rmod.Clustering(x, recipe=dict(rrecipes.get_registered_recipe("coordinates")))
        mine, _ = clustering(x, **({"fitter": "b200_staged"} if trial % 2 else {}))
        for step in range(int(rng.integers(1, 4))):
            kw = {"v": False}
            if rng.random() < 0.5:
                kw["member_cutoff"] = int(rng.integers(1, 8))
            if rng.random() < 0.3:
                kw["max_clusters"] = int(rng.integers(1, 4))
            if rng.random() < 0.3:
                kw["sort_by_size"] = False
            if rng.random() < 0.2:
                kw["cnn_offset"] = 1
            if rng.random() < 0.2:
                kw["purge"] = True
            r = float(rng.choice([0.125, 0.25, 0.3, 0.5, 0.8]))
            c = int(rng.integers(1 if "cnn_offset" in kw else 0, 7))
            if step and rng.random() < 0.6:          # keep what was found, recluster the rest
                for cl in (mine, theirs):
                    cl._bundle._labels.meta["frozen"] = True
                    lab = np.asarray(cl.labels)
                    cl._bundle._labels.consider[:] = (lab == 0).astype(np.uint8)
            for cl in (mine, theirs):
                cl.fit(r, c, **kw)
            what = f"trial {trial} step {step}: n={n} d={x.shape[1]} r={r} c={c} {kw}"
            np.testing.assert_array_equal(mine.labels, np.asarray(theirs.labels), err_msg=what)
            a, b = mine.summary[-1], theirs.summary[-1]
            assert (a.n_clusters, a.ratio_largest, a.ratio_noise, a.cnn_cutoff) == \
                (b.n_clusters, b.ratio_largest, b.ratio_noise, b.cnn_cutoff), what
            assert mine._bundle._labels.meta["params"] == theirs._bundle._labels.meta["params"], what
            assert mine._bundle._labels.consider.tolist() == np.asarray(theirs._bundle._labels.consider).tolist(), what


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_random_predictions_equal_live_reference():
    """`Clustering.predict` (cluster.pyx:560-653, PredictorFirstmatch) on twelve random configurations: a first call
    restricted to some clusters, a second one on the same (now frozen) labels with other parameters."""
    rmod = ref.module()   # puts oracle/_ref on sys.path
    from cnnclustering import _fit as rfit, _types as rtypes, recipes as rrecipes

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
    rng = np.random.default_rng(404)
    seen = 0
    for trial in range(12):
        pts = datasets.blobs_2d(int(rng.integers(400, 900))) if trial % 2 else cases.lattice(500, 2, 300 + trial)
        x, x_other = np.ascontiguousarray(pts[::2]), np.ascontiguousarray(pts[1::2])
        r, c = (0.15, int(rng.integers(3, 9))) if trial % 2 else (0.25, int(rng.integers(1, 4)))
        out = []
        for mod in ("ours", "theirs"):
            if mod == "ours":
                cl, _ = clustering(x)
                other = cluster.Clustering(x_other)
            else:
                cl, other = rmod.Clustering(x, recipe=dict(recipe)), rmod.Clustering(x_other, recipe=dict(recipe))
            cl.fit(r, c, member_cutoff=3, v=False)
            n_found = int(np.asarray(cl.labels).max())
            first = [1] if n_found >= 1 else None
            cl.predict(other._bundle, r * 1.2, max(c - 1, 0), clusters=first)
            a = np.asarray(other.labels).copy()
            cl.predict(other._bundle, r * 1.5, max(c - 2, 0))          # frozen: only the still-considered points move
            out.append((a, np.asarray(other.labels).copy(), np.asarray(other._bundle._labels.consider).copy()))
        for got, want in zip(out[0], out[1]):
            np.testing.assert_array_equal(got, want, err_msg=f"trial {trial}")
        seen += int(out[0][1].max() > 0)
    assert seen >= 6
