"""Host-side mirror of the reference interface: parameters, labels container, recipes, hooks."""
import numpy as np
import pytest

from commonnnclustering_b200 import _fit, _types, cluster, datasets, recipes


def test_cluster_parameters_defaults():
    # reference test/unit/test_types/test_create_params_0_1_1_.txt
    p = _types.ClusterParameters(0.1, 1, 1.)
    assert p.to_dict() == {"radius_cutoff": 0.1, "similarity_cutoff": 1, "similarity_cutoff_continuous": 1.0,
                           "n_member_cutoff": 1, "current_start": 1}


def test_make_parameters_coordinates_recipe():
    f = _fit.FitterB200()
    p = f.make_parameters(0.3, 10, 1)
    assert p.radius_cutoff == 0.3 * 0.3 and p.similarity_cutoff == 12 and p.n_member_cutoff == 11
    f = _fit.FitterB200(_types.NeighboursGetterRadius(_types.DistanceGetterMetric(_types.MetricEuclidean())))
    assert f.make_parameters(0.3, 10, 4).radius_cutoff == 0.3
    f = _fit.FitterB200(_types.NeighboursGetterLookup(is_selfcounting=False))
    p = f.make_parameters(0.3, 10, 1)
    assert (p.radius_cutoff, p.similarity_cutoff, p.n_member_cutoff) == (0.3, 10, 10)


def test_labels_container():
    lab = _types.Labels.from_sequence([1, 1, 2, 2, 0, 1])
    assert lab.n_points == 6 and lab.set == {0, 1, 2}
    assert dict(lab.mapping) == {0: [4], 1: [0, 1, 5], 2: [2, 3]}
    assert list(lab.mapping) == [1, 2, 0]           # keys in order of first occurrence, like the reference's loop
    assert lab.consider.tolist() == [1] * 6
    with pytest.raises(ValueError):
        _types.Labels(np.zeros(3, dtype=np.int64), consider=np.ones(2, dtype=np.uint8))
    a = np.zeros(4, dtype=np.int64)
    assert _types.Labels(a).labels is a  # a view: in-place writes by a fitter are visible


def test_builder_default_recipe():
    b = recipes.Builder()
    f = b.make_component("fitter")
    assert isinstance(f, _fit.FitterB200Fused)
    assert isinstance(f._neighbours_getter, _types.NeighboursGetterRadius)
    assert isinstance(f._neighbours_getter.metric, _types.MetricEuclideanReduced)
    b = recipes.Builder(fitter__ngetter__dgetter__metric="euclidean")
    assert isinstance(b.make_component("fitter")._neighbours_getter.metric, _types.MetricEuclidean)
    b = recipes.Builder("sorted_neighbourhoods")
    f = b.make_component("fitter")
    assert isinstance(f._neighbours_getter, _types.NeighboursGetterLookup) and f._neighbours_getter.is_selfcounting


def test_builder_reference_recipes():
    """All four registered recipes of the reference (recipes.py:316-358) build under their own names, with the
    parameters the reference's fitters would derive (_fit.pyx:129-161)."""
    for name in ("coordinates", "distances", "neighbourhoods", "sorted_neighbourhoods"):
        assert name in recipes.REGISTERED_RECIPES
    f = recipes.Builder("distances").make_component("fitter")
    assert isinstance(f._neighbours_getter.metric, _types.MetricPrecomputed)
    p = f.make_parameters(0.3, 5, 1)     # precomputed metric: radius unchanged; self-counting getter
    assert (p.radius_cutoff, p.similarity_cutoff, p.n_member_cutoff) == (0.3, 7, 6)
    dm = _types.InputDataComponents(np.array([[0.0, 2.0], [2.0, 0.0]]))
    assert _types.MetricPrecomputed().calc_distance(0, 1, dm) == 2.0
    f = recipes.Builder("neighbourhoods").make_component("fitter")
    assert isinstance(f._neighbours_getter, _types.NeighboursGetterLookup) and not f._neighbours_getter.is_sorted
    f = recipes.Builder("sorted_neighbourhoods").make_component("fitter")
    assert f._neighbours_getter.is_sorted
    assert _types.NeighboursGetterLookup().is_sorted is False      # the reference's default (_types.pyx:1794)
    p = f.make_parameters(0.3, 5, 1)     # no metric: radius unchanged
    assert (p.radius_cutoff, p.similarity_cutoff, p.n_member_cutoff) == (0.3, 7, 6)


def test_prepare_points_from_parts():
    # reference test/unit/test_hooks.py
    (a,), kw = recipes.prepare_points_from_parts([[0, 0], [1, 1]])
    assert a.shape == (2, 2) and kw["meta"]["edges"] == [2] and a.dtype == np.float64 and a.flags.c_contiguous
    (a,), kw = recipes.prepare_points_from_parts([[[0, 0], [1, 1]], [[2, 2]]])
    assert a.shape == (3, 2) and kw["meta"]["edges"] == [2, 1]
    (a,), kw = recipes.prepare_points_from_parts([1, 2, 3])
    assert a.shape == (1, 3)


def test_prepare_neighbourhoods_padding():
    (padded, counts), _ = recipes.prepare_neighbourhoods([[0, 1], [0, 1, 2], [2]])
    assert padded.tolist() == [[0, 1, 0], [0, 1, 2], [2, 0, 0]] and counts.tolist() == [2, 3, 1]


def test_clustering_construction():
    c = cluster.Clustering(datasets.README_POINTS)
    assert c.root._input_data.n_points == 12 and c.labels is None
    assert cluster.Clustering(datasets.README_POINTS, bundle_kwargs={"alias": "traj-7"}).root.alias == "traj-7" and c.root.alias == "root"
    assert isinstance(c.fitter, _fit.Fitter)
    with pytest.raises(TypeError):
        c.fitter = object()


def test_datasets_are_deterministic():
    a, b = datasets.md_like_3d(2000), datasets.md_like_3d(2000)
    assert np.array_equal(a, b) and a.shape == (2000, 3)
    # the state chain is sequential by definition; the vectorised form must equal a plain loop
    rng = np.random.default_rng(0)
    rng.random((12, 3)); rng.uniform(0.25, 0.6, 12)
    n = 5000
    switch = rng.random(n) < 1e-3
    new_state = rng.integers(0, 12, n)
    state = np.zeros(n, dtype=np.int64)
    prev = 0
    for i in range(n):
        prev = new_state[i] if switch[i] else prev
        state[i] = prev
    idx = np.where(switch, np.arange(n), -1)
    last = np.maximum.accumulate(idx)
    assert np.array_equal(state, np.where(last >= 0, new_state[np.maximum(last, 0)], 0))
    assert datasets.features_64d(500).shape == (500, 64)


# ---- hierarchy helpers (host logic): isolate / reel --------------------------------------------------------
def _toy_hierarchy(cluster_mod, labels_cls, x, parent_labels, child_labels_by_label):
    cl = cluster_mod.Clustering(x)
    cl.labels = labels_cls(np.array(parent_labels, dtype=np.int64))
    cl.root._labels.meta["params"] = {int(k): (1.0, 1) for k in set(parent_labels) if k != 0}
    cl.isolate()
    for label, child_labels in child_labels_by_label.items():
        child = cl.root._children[label]
        child._labels = labels_cls(np.array(child_labels, dtype=np.int64))
        child._labels.meta["params"] = {int(k): (0.5, 2) for k in set(child_labels) if k != 0}
    return cl


def test_isolate_and_reel_known_answer():
    from commonnnclustering_b200 import cluster
    from commonnnclustering_b200._types import Labels

    x = np.arange(24, dtype=np.float64).reshape(12, 2)
    parent = [2, 2, 1, 0, 1, 2, 1, 0, 3, 3, 1, 2]
    cl = _toy_hierarchy(cluster, Labels, x, parent, {1: [1, 1, 0, 2], 2: [0, 1, 1, 2]})
    # children in order of first occurrence; members ascending; data subsets attached
    assert list(cl.root.children) == [2, 1, 3]
    assert cl.root.children[1].parent_indices.tolist() == [2, 4, 6, 10]
    assert cl.root.children[2].parent_indices.tolist() == [0, 1, 5, 11]
    np.testing.assert_array_equal(cl.root.children[3].input_data, x[[8, 9]])
    cl.reel()
    # child 2 first (labels + 3), then child 1 (labels + current max 5); child 3 has no labels
    assert cl.labels.tolist() == [0, 4, 6, 0, 6, 4, 0, 0, 3, 3, 7, 5]
    assert cl.root._labels.meta["origin"] == "reel"
    assert set(cl.root._labels.meta["params"]) == {3, 4, 5, 6, 7}


def test_isolate_and_reel_equal_reference():
    from oracle import ref

    if not ref.available():
        pytest.skip("oracle/_ref not built")
    from commonnnclustering_b200 import cluster
    from commonnnclustering_b200._types import Labels

    rmod = ref.module()
    from cnnclustering import _types as rtypes

    rng = np.random.default_rng(5)
    x = rng.standard_normal((400, 2))
    parent = rng.integers(0, 6, 400)
    kids = {}
    for label in range(1, 6):
        m = int((parent == label).sum())
        if label != 4:
            kids[label] = rng.integers(0, 4, m).tolist()
    ours = _toy_hierarchy(cluster, Labels, x, parent.tolist(), kids)
    theirs = _toy_hierarchy(rmod, rtypes.Labels, x, parent.tolist(), kids)
    assert list(ours.root._children) == list(theirs.root._children)
    for label in ours.root._children:
        np.testing.assert_array_equal(ours.root._children[label].parent_indices,
                                      theirs.root._children[label]._reference_indices.parent)
    ours.reel()
    theirs.reel()
    np.testing.assert_array_equal(ours.labels, theirs.labels)
    assert ours.root._labels.meta["params"] == theirs.root._labels.meta["params"]


def test_get_subset_of_every_input_type():
    """Bundle.isolate relies on get_subset: coordinates carry their part lengths along, neighbourhood inputs are
    restricted to the subset and renumbered (_bundle.pyx:321-350, _types.pyx:1199-1217)."""
    from commonnnclustering_b200 import _types

    x = np.arange(20, dtype=np.float64).reshape(10, 2)
    comp = _types.InputDataComponents(x, meta={"edges": [4, 3, 3]})
    sub = comp.get_subset([0, 1, 5, 9])
    assert sub.meta["edges"] == [2, 1, 1] and np.array_equal(sub.to_components_array(), x[[0, 1, 5, 9]])
    rp = np.array([0, 3, 5, 8, 10, 12], dtype=np.int64)
    ci = np.array([0, 1, 2, 0, 1, 0, 2, 4, 3, 4, 2, 4], dtype=np.int64)
    idx = np.array([0, 2, 4])
    want_rows = [[0, 1], [0, 1, 2], [1, 2]]            # members {0, 2, 4} renumbered to 0, 1, 2
    csr = _types.InputDataNeighbourhoodsCSR(rp, ci).get_subset(idx)
    got_rp, got_ci = csr.to_csr()
    assert [got_ci[a:b].tolist() for a, b in zip(got_rp[:-1], got_rp[1:])] == want_rows
    padded = np.zeros((5, 3), dtype=np.int64)
    counts = np.diff(rp)
    for i in range(5):
        padded[i, :counts[i]] = ci[rp[i]:rp[i + 1]]
    nb = _types.InputDataNeighbourhoods(padded, counts).get_subset(idx)
    assert [row.tolist() for row in nb.data] == want_rows and nb.n_neighbours.tolist() == [2, 3, 2]


def test_reel_running_maximum_equals_the_per_child_pass():
    """reel shifts every child's labels by the parent's CURRENT largest label (cluster.pyx:502-558 takes a max over the
    parent per child); the histogram that replaces those passes must give the same labels, also when a child is all
    noise (the parent's largest label then disappears) and over two levels."""
    from commonnnclustering_b200 import cluster
    from commonnnclustering_b200._types import Labels

    rng = np.random.default_rng(8)

    def naive(parent_labels, children):
        out = parent_labels.copy()
        for label, (idx, child_labels) in children.items():
            n_clusters = int(out.max())
            out[idx] = np.where(child_labels == 0, 0, child_labels + n_clusters)
        return out

    for trial in range(20):
        n = 400
        lab = rng.integers(0, 7, n)
        x = rng.random((n, 2))
        cl = cluster.Clustering(x)
        cl.labels = lab.copy()
        cl.isolate()
        children = {}
        for label, child in cl.root._children.items():
            m = child.parent_indices.shape[0]
            kind = rng.integers(0, 3)
            cl_lab = np.zeros(m, dtype=np.int64) if kind == 0 else rng.integers(0, 1 + rng.integers(1, 5), m)
            child._labels = Labels(cl_lab.copy())
            children[label] = (child.parent_indices, cl_lab)
        want = naive(lab.astype(np.int64), children)
        cl.reel()
        np.testing.assert_array_equal(cl.labels, want)


def test_condensed_distances_input_type():
    """InputDataDistancesLinear mirrors InputDataExtDistancesLinearMemoryview (_types.pyx:1112-1150): n_points from the
    vector length, `_get_distance`'s index arithmetic, subsets again condensed; registered under "distances_linear"."""
    rng = np.random.default_rng(4)
    n = 23
    dm = rng.random((n, n))
    dm = np.minimum(dm, dm.T)
    np.fill_diagonal(dm, 0.0)
    cond = dm[np.triu_indices(n, k=1)]
    data = _types.InputDataDistancesLinear(cond)
    assert data.n_points == n and data.meta["access_distances"] and data.data.shape == (n * (n - 1) // 2,)
    for a, b in ((0, 1), (5, 3), (22, 0), (21, 22), (7, 7)):
        assert data.get_distance(a, b) == dm[a, b]
    idx = np.array([1, 4, 5, 19, 22])
    sub = data.get_subset(idx)
    assert sub.n_points == 5 and np.array_equal(sub.data, dm[np.ix_(idx, idx)][np.triu_indices(5, k=1)])
    assert _types.InputDataDistancesLinear(np.zeros(0)).n_points == 1     # one point: no pairs
    with pytest.raises(ValueError):
        _types.InputDataDistancesLinear(np.zeros(4))                      # not a triangular number
    with pytest.raises(ValueError):
        _types.InputDataDistancesLinear(np.zeros((3, 3)))
    assert recipes.COMPONENT_NAME_TYPE_MAP["input_data"]["distances_linear"] is _types.InputDataDistancesLinear
    # pinning starts with the second fit (a child of isolate is usually fitted once)
    big = _types.InputDataComponents(np.zeros((50_000, 3)))
    assert big.pin() is False and big._pin_calls == 1


def test_window_plan_of_the_sharded_exchange():
    """dist.plan_windows: every rank's window covers the rows it built, all windows have one length and stay in [0, n)."""
    from commonnnclustering_b200 import dist as cdist

    n = 1000
    ranges = np.array([[0, 300], [250, 520], [480, 790], [760, 1000]])
    length, starts = cdist.plan_windows(ranges, n)
    assert length == 310 and starts.tolist() == [0, 250, 480, 690]
    assert np.all(starts <= ranges[:, 0]) and np.all(starts + length >= ranges[:, 1]) and np.all(starts + length <= n)
    length, starts = cdist.plan_windows(np.array([[0, 0], [0, 1000]]), n)      # an empty block next to a full one
    assert length == n and starts.tolist() == [0, 0]
    length, starts = cdist.plan_windows(np.array([[5, 5], [9, 9]]), 10)        # nothing built anywhere
    assert length == 1 and np.all(starts + length <= 10)
