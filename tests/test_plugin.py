"""The GPU fitter dropped into the UNMODIFIED reference (oracle/_ref) through its own recipe registry."""
import numpy as np
import pytest

from oracle import ref

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
def test_registration_with_the_reference_registry():
    mod = ref.module()
    from commonnnclustering_b200 import plugin

    cls = plugin.register()
    import cnnclustering._fit as ref_fit
    import cnnclustering.recipes as ref_recipes

    assert ref_recipes.COMPONENT_NAME_TYPE_MAP["fitter"]["b200"] is cls
    assert "coordinates_b200" in ref_recipes.REGISTERED_RECIPES
    f = cls()
    assert isinstance(f, ref_fit.Fitter)
    p = f.make_parameters(0.3, 10, 1)
    # the numbers the reference's own "coordinates" recipe produces
    want = mod.Clustering(np.zeros((3, 2))).fitter.make_parameters(0.3, 10, 1)
    assert p.to_dict() == want.to_dict()
    cl = mod.Clustering(np.zeros((3, 2)), recipe="coordinates_b200")
    assert isinstance(cl.fitter, cls)
    cl2 = mod.Clustering(np.zeros((3, 2)))
    cl2.fitter = cls()  # the property setter type-checks against the reference's Fitter ABC


@needs_ref
@pytest.mark.gpu
def test_reference_clustering_runs_on_the_gpu_fitter():
    from commonnnclustering_b200 import datasets, plugin

    mod = ref.module()
    plugin.register()
    for x, r, c, kw in [(datasets.README_POINTS, 1.5, 1, {}),
                        (datasets.blobs_2d(3000), 0.1, 10, {"member_cutoff": 5}),
                        (datasets.md_like_3d(4000), 0.4, 5, {"sort_by_size": False}),
                        (datasets.blobs_2d(2000), 0.08, 5, {"member_cutoff": 1, "max_clusters": 3})]:
        cpu = mod.Clustering(x)
        cpu.fit(r, c, v=False, **kw)
        gpu = mod.Clustering(x, recipe="coordinates_b200")
        gpu.fit(r, c, v=False, **kw)
        np.testing.assert_array_equal(np.asarray(gpu.labels), np.asarray(cpu.labels))
        assert gpu.summary[-1].n_clusters == cpu.summary[-1].n_clusters
        assert gpu.summary[-1].execution_time > 0
    # frozen labels: a second fit continues numbering after the existing clusters (cluster.pyx:316-326)
    x = datasets.blobs_2d(2000)
    outs = []
    for recipe in ("coordinates", "coordinates_b200"):
        cl = mod.Clustering(x, recipe=recipe)
        cl.fit(0.1, 10, v=False)
        first = np.asarray(cl.labels).copy()
        cl.root._labels.meta["frozen"] = True
        cl.root._labels.consider[:] = (first == 0).astype(np.uint8)
        cl.fit(0.2, 3, v=False, sort_by_size=False)
        outs.append(np.asarray(cl.labels).copy())
    np.testing.assert_array_equal(outs[0], outs[1])


@needs_ref
@pytest.mark.gpu
def test_reference_neighbourhood_recipes_run_on_the_gpu_fitter():
    """`sorted_neighbourhoods_b200`: the reference's own input type and preparation hook (padded int64 matrix), the GPU
    fitter; lists in ascending order and in any order, against the reference's two neighbourhood recipes."""
    from commonnnclustering_b200 import datasets, plugin
    from oracle import port

    mod = ref.module()
    plugin.register()
    rng = np.random.default_rng(4)
    for x, r, c, kw in [(datasets.blobs_2d(2500), 0.1, 10, {"member_cutoff": 5}),
                        (datasets.md_like_3d(3000), 0.4, 5, {})]:
        rp, ci = port.neighbours(x, r * r)
        nb = [ci[a:b] for a, b in zip(rp[:-1], rp[1:])]
        shuffled = [rng.permutation(row) for row in nb]
        cpu = mod.Clustering(nb, recipe="sorted_neighbourhoods")
        cpu.fit(r, c, v=False, **kw)
        cpu_any = mod.Clustering(shuffled, recipe="neighbourhoods")
        cpu_any.fit(r, c, v=False, **kw)
        np.testing.assert_array_equal(np.asarray(cpu_any.labels), np.asarray(cpu.labels))
        for lists in (nb, shuffled):
            gpu = mod.Clustering(lists, recipe="sorted_neighbourhoods_b200")
            gpu.fit(r, c, v=False, **kw)
            np.testing.assert_array_equal(np.asarray(gpu.labels), np.asarray(cpu.labels))
        padded, counts = plugin._padded_view(gpu._bundle._input_data)
        assert padded is not None and padded.shape[0] == len(nb) and int(counts.sum()) == int(rp[-1])


@needs_ref
def test_plugin_host_logic_on_the_stand_in_context():
    """The plug-in fitter inside the UNMODIFIED reference, without a GPU: its context is the oracle-backed stand-in
    (tests/standin.py), so this checks what the plug-in hands over -- the reference's own buffers, squared radius,
    raised cutoffs, the padded matrix as it is, the consider mask of frozen labels -- not the device."""
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from standin import StandInContext

    from commonnnclustering_b200 import datasets, plugin
    from oracle import port

    mod = ref.module()
    plugin.register()

    def on_stand_in(data, recipe):
        cl = mod.Clustering(data, recipe=recipe)
        cl.fitter._ctx = StandInContext()
        return cl

    for x, r, c, kw in [(datasets.README_POINTS, 1.5, 1, {}),
                        (datasets.blobs_2d(1500), 0.1, 10, {"member_cutoff": 5}),
                        (datasets.md_like_3d(1500), 0.4, 5, {"sort_by_size": False}),
                        (datasets.blobs_2d(1500), 0.08, 5, {"member_cutoff": 1, "max_clusters": 3})]:
        cpu = mod.Clustering(x)
        cpu.fit(r, c, v=False, **kw)
        ours = on_stand_in(x, "coordinates_b200")
        ours.fit(r, c, v=False, **kw)
        np.testing.assert_array_equal(np.asarray(ours.labels), np.asarray(cpu.labels))
        assert ours.fitter.ctx.calls == ["neighbours_from_points", "fit"]
        assert ours.summary[-1].n_clusters == cpu.summary[-1].n_clusters
    # frozen labels: the second fit hands the consider mask over and continues the numbering
    x = datasets.blobs_2d(1500)
    outs = []
    for cl in (mod.Clustering(x), on_stand_in(x, "coordinates_b200")):
        cl.fit(0.1, 10, v=False)
        first = np.asarray(cl.labels).copy()
        cl.root._labels.meta["frozen"] = True
        cl.root._labels.consider[:] = (first == 0).astype(np.uint8)
        cl.fit(0.2, 3, v=False, sort_by_size=False)
        outs.append(np.asarray(cl.labels).copy())
    np.testing.assert_array_equal(outs[0], outs[1])
    # neighbourhoods: the reference's padded matrix goes to the context as it is, lists in any order included
    rng = np.random.default_rng(4)
    x, r, c = datasets.blobs_2d(1500), 0.1, 10
    rp, ci = port.neighbours(x, r * r)
    nb = [ci[a:b] for a, b in zip(rp[:-1], rp[1:])]
    cpu = mod.Clustering(nb, recipe="sorted_neighbourhoods")
    cpu.fit(r, c, member_cutoff=5, v=False)
    for lists in (nb, [rng.permutation(row) for row in nb]):
        ours = on_stand_in(lists, "sorted_neighbourhoods_b200")
        ours.fit(r, c, member_cutoff=5, v=False)
        np.testing.assert_array_equal(np.asarray(ours.labels), np.asarray(cpu.labels))
        assert ours.fitter.ctx.calls == ["neighbours_from_padded", "fit"]
