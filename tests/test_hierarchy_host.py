"""Host logic of `Clustering.fit` / `fit_hierarchical` / `isolate` / `reel` without a GPU: the device is replaced by a
stand-in that answers the SAME context contract from the oracle (tests/standin.py), so what runs here is the package's
own Python -- `cluster.Clustering`, `FitterB200`, `Labels.sort_by_size`'s bookkeeping, `_fit.HierarchicalFitterRepeat` --
and its result is compared with the reference itself (`oracle/_ref`; cluster.pyx:256-413, :502-558, _fit.pyx:1018-1184).
The GPU twins of these tests (tests/test_gpu_parity.py::test_hierarchical_repeat_equals_reference, ...) run the same
comparisons with the real context."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from standin import StandInContext  # noqa: E402

from commonnnclustering_b200 import _fit, cluster, datasets  # noqa: E402
from oracle import ref  # noqa: E402

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")


def ours_on_stand_in(x, hierarchical=False):
    """The package's own staged fitter (and hierarchical fitter) over a stand-in context."""
    ctx = StandInContext()
    fitter = _fit.FitterB200(ctx=ctx)
    kw = {"hierarchical_fitter": _fit.HierarchicalFitterRepeat(fitter)} if hierarchical else {}
    return cluster.Clustering(x, fitter=fitter, **kw)


REF_KW = dict(fitter="bfs", fitter__ngetter="brute_force", fitter__na="vuset", fitter__nb="vuset", fitter__checker="switch",
              fitter__queue="fifo", fitter__ngetter__dgetter="metric", fitter__ngetter__dgetter__metric="euclidean_r")


def reference_clustering(rmod, x, prefix="", **extra):
    """The reference's Clustering on a COPY of its default recipe: keyword arguments would be merged into the
    registered recipe itself (recipes.py:193-194) and leak into every later construction in the process."""
    from cnnclustering import recipes as rrecipes

    recipe = dict(rrecipes.get_registered_recipe("coordinates"))
    recipe.update({prefix + k: v for k, v in REF_KW.items()})
    recipe.update(extra)
    return rmod.Clustering(x, recipe=recipe)


def same_tree(a, b, count):
    count[0] += 1
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
        same_tree(ca[k], cb[k], count)


@needs_ref
def test_fit_records_and_frozen_labels_equal_reference():
    rmod = ref.module()
    x = datasets.blobs_2d(1200)
    ours = ours_on_stand_in(x)
    theirs = reference_clustering(rmod, x)
    for cl in (ours, theirs):
        cl.fit(0.3, 10, member_cutoff=5, v=False)
    np.testing.assert_array_equal(ours.labels, np.asarray(theirs.labels))
    a, b = ours.summary[-1], theirs.summary[-1]
    for field in ("n_points", "radius_cutoff", "cnn_cutoff", "member_cutoff", "max_clusters", "n_clusters", "ratio_largest",
                  "ratio_noise"):
        assert getattr(a, field) == getattr(b, field), field
    assert ours._bundle._labels.meta["params"] == theirs._bundle._labels.meta["params"]
    # a second fit on frozen labels: only the noise is reclustered, new clusters are numbered after the old ones
    for cl in (ours, theirs):
        cl._bundle._labels.meta["frozen"] = True
        cl.fit(0.45, 3, v=False, sort_by_size=False)
    np.testing.assert_array_equal(ours.labels, np.asarray(theirs.labels))
    assert ours._bundle._labels.meta["params"] == theirs._bundle._labels.meta["params"]
    assert ours.summary[-1].n_clusters == theirs.summary[-1].n_clusters


@needs_ref
def test_hierarchical_repeat_host_logic_equals_reference():
    rmod = ref.module()
    x = datasets.blobs_2d(1500)
    kw = dict(radius_cutoff=[0.5, 0.25, 0.12, 0.08], cnn_cutoff=[5, 8, 8, 10], member_cutoff=5, v=False)
    theirs = reference_clustering(rmod, x, prefix="hierarchical_fitter__", hierarchical_fitter="repeat")
    theirs.fit_hierarchical(**kw)
    ours = ours_on_stand_in(x, hierarchical=True)
    ours.fit_hierarchical(**kw)
    count = [0]
    same_tree(ours._bundle, theirs._bundle, count)
    assert count[0] > 5
    # parameters recorded per cluster (info=True) follow the reference's, level by level
    def params(b):
        return None if b._labels is None else b._labels.meta.get("params")
    assert params(ours._bundle) == params(theirs._bundle)
    for k, child in (ours._bundle._children or {}).items():
        assert params(child) == params(theirs._bundle._children[k])
    # reel: the tree folded back into one label array
    ours.reel()
    theirs.reel()
    np.testing.assert_array_equal(ours.labels, np.asarray(theirs.labels))


@needs_ref
def test_isolate_recluster_reel_host_logic_equals_reference():
    rmod = ref.module()
    x = datasets.blobs_2d(1500)
    ours = ours_on_stand_in(x)
    theirs = reference_clustering(rmod, x)
    for cl in (ours, theirs):
        cl.fit(0.45, 5, member_cutoff=5, v=False)
        cl.isolate()
    assert list(ours.children) == list(theirs.children)
    for k in ours.children:
        for cl in (ours, theirs):
            cl.fit(0.2, 8, member_cutoff=5, v=False, bundle=cl.children[k])
        np.testing.assert_array_equal(ours.children[k].labels, np.asarray(theirs.children[k].labels))
    ours.reel()
    theirs.reel()
    np.testing.assert_array_equal(ours.labels, np.asarray(theirs.labels))
    assert ours._bundle._labels.meta["params"] == theirs._bundle._labels.meta["params"]


def test_stand_in_context_follows_the_device_contract():
    """labels_group: order by (prev, cur, index), rank within the previous cluster; sort_by_size: map old -> new."""
    ctx = StandInContext()
    prev = np.array([1, 2, 1, 0, 2, 1], dtype=np.int64)
    cur = np.array([2, 1, 1, 3, 1, 2], dtype=np.int64)
    order, rank = ctx.labels_group(prev, cur)
    assert order.tolist() == [3, 2, 0, 5, 1, 4]
    assert rank.tolist() == [0, 1, 0, 2, 0, 1]
    lab = np.array([3, 3, 1, 2, 2, 2, 0], dtype=np.int64)
    reassign = ctx.sort_by_size(lab, 1, None, return_map=True, max_label=3)
    assert lab.tolist() == [2, 2, 3, 1, 1, 1, 0] and reassign.tolist() == [0, 3, 1, 2]


@needs_ref
@pytest.mark.parametrize("trim", ["trim_shrinking_leafs", "trim_trivial_leafs"])
def test_trimmed_hierarchies_equal_reference(trim):
    """`trim_shrinking_leafs` / `trim_trivial_leafs` (cluster.pyx:435-500) on the same repeated-fit hierarchy."""
    from cnnclustering import _bundle as rbundle

    rmod = ref.module()
    x = datasets.blobs_2d(1500)
    # a ladder of ever harder parameters: clusters shrink for a while, split, and finally dissolve into noise
    kw = dict(radius_cutoff=[0.5, 0.08, 0.075, 0.07, 0.065, 0.05], cnn_cutoff=[5, 10, 10, 10, 10, 60], member_cutoff=5,
              v=False)
    theirs = reference_clustering(rmod, x, prefix="hierarchical_fitter__", hierarchical_fitter="repeat")
    ours = ours_on_stand_in(x, hierarchical=True)
    for cl in (ours, theirs):
        cl.fit_hierarchical(**kw)
    before = sum(b._labels is not None for b in cluster.bfs_leafs(ours._bundle))
    for cl in (ours, theirs):
        getattr(cl, trim)()
    count = [0]
    same_tree(ours._bundle, theirs._bundle, count)
    after = sum(b._labels is not None for b in cluster.bfs_leafs(ours._bundle))
    assert after < before, "the ladder was meant to leave something to trim"
    # the breadth-first walk visits the same bundles in the same order
    assert [b.alias.split(" - ")[-1] for b in cluster.bfs_leafs(ours._bundle)][1:] == \
        [str(k) for b in rbundle.bfs_leafs(theirs._bundle) for k in (b._children or {})]
    ours.reel()
    theirs.reel()
    np.testing.assert_array_equal(ours.labels, np.asarray(theirs.labels))


def test_to_dtrajs_follows_the_parts_of_the_input():
    rng = np.random.default_rng(0)
    parts = [rng.random((5, 2)), rng.random((3, 2)), rng.random((4, 2))]
    cl = cluster.Clustering(parts)
    assert cl.to_dtrajs() == []
    cl.labels = np.arange(12) % 3
    got = cl.to_dtrajs()
    assert [len(p) for p in got] == [5, 3, 4] and np.concatenate(got).tolist() == (np.arange(12) % 3).tolist()
    flat = cluster.Clustering(np.concatenate(parts))
    flat.labels = np.arange(12) % 3
    assert len(flat.to_dtrajs()) == 1


@needs_ref
def test_sort_by_size_rekeys_params_and_children_like_the_reference():
    """`Labels.sort_by_size(..., bundle=...)` (_types.pyx:212-286): labels, `meta["params"]` and the bundle's children are
    re-keyed together; clusters that fall under `member_cutoff` / beyond `max_clusters` lose their entry."""
    from cnnclustering import _types as rtypes

    from commonnnclustering_b200 import _types

    rmod = ref.module()
    rng = np.random.default_rng(8)
    raw = rng.choice(np.arange(0, 8), size=400, p=[0.1, 0.05, 0.3, 0.02, 0.2, 0.01, 0.25, 0.07]).astype(np.int64)
    for kw in ({}, {"member_cutoff": 10}, {"max_clusters": 3}, {"member_cutoff": 30, "max_clusters": 4}):
        out = []
        for mod, labels_type, extra in ((cluster, _types.Labels, {"ctx": StandInContext()}), (rmod, rtypes.Labels, {})):
            cl = mod.Clustering(rng.random((400, 2)))
            lab = labels_type(raw.copy())
            lab.meta["params"] = {int(k): (0.1 * k, k) for k in np.unique(raw) if k != 0}
            cl._bundle._labels = lab
            cl._bundle._children = {int(k): f"child of {k}" for k in np.unique(raw) if k != 0}
            lab.sort_by_size(bundle=cl._bundle, **kw, **extra)
            out.append((np.asarray(lab.labels).tolist(), lab.meta["params"], cl._bundle._children))
        assert out[0][0] == out[1][0], kw
        assert out[0][1] == out[1][1], kw
        assert out[0][2] == out[1][2] and list(out[0][2]) == list(out[1][2]), kw


@needs_ref
def test_random_ladders_equal_reference():
    """Fifteen random parameter ladders (2-5 steps, scalar or sequence arguments, with and without size sorting and
    offsets): the trees, what the two trims leave of them, and the reeled labels follow the reference's."""
    rmod = ref.module()
    rng = np.random.default_rng(77)
    for trial in range(15):
        x = datasets.blobs_2d(int(rng.integers(300, 700)))
        steps = int(rng.integers(2, 6))
        radii = sorted((float(v) for v in rng.choice([0.5, 0.4, 0.3, 0.25, 0.2, 0.15, 0.1, 0.08], steps, replace=False)),
                       reverse=True)
        kw = dict(radius_cutoff=radii, cnn_cutoff=[int(v) for v in np.sort(rng.integers(2, 12, steps))], v=False)
        if trial % 4 == 0:
            kw["cnn_cutoff"] = int(kw["cnn_cutoff"][0])             # a scalar is repeated for every step
        if trial % 3 == 0:
            kw["member_cutoff"] = int(rng.integers(2, 10))
        if trial % 5 == 0:
            kw["sort_by_size"] = False
        if trial % 7 == 0:
            kw["cnn_offset"] = 1
        for trim in (None, "trim_shrinking_leafs", "trim_trivial_leafs"):
            theirs = reference_clustering(rmod, x, prefix="hierarchical_fitter__", hierarchical_fitter="repeat")
            ours = ours_on_stand_in(x, hierarchical=True)
            for cl in (ours, theirs):
                cl.fit_hierarchical(**kw)
                if trim:
                    getattr(cl, trim)()
            same_tree(ours._bundle, theirs._bundle, [0])
            for cl in (ours, theirs):
                cl.reel()
            np.testing.assert_array_equal(ours.labels, np.asarray(theirs.labels), err_msg=f"trial {trial} {trim} {kw}")
