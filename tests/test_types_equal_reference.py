"""The host-side mirror types (`commonnnclustering_b200._types`) next to the reference's own (`cnnclustering._types`,
built into `oracle/_ref`): same constructor arguments, same values from the same accessors -- `Labels`, `ClusterParameters`,
the input data types, the metrics and the distance getter.  Differences that are deliberate are asserted as such."""
import numpy as np
import pytest

from commonnnclustering_b200 import _types as ours
from oracle import ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module")
def theirs():
    ref.module()
    from cnnclustering import _types

    return _types


def both(a, b, f, raises=()):
    """f applied to our object and to the reference's: the value of each; an exception only where one is expected
    (`raises`: type names), and then its type name."""
    out = []
    for obj in (a, b):
        try:
            v = f(obj() if callable(obj) else obj)
            out.append(v.tolist() if isinstance(v, np.ndarray) else v)
        except Exception as e:  # noqa: BLE001 -- the error behaviour is part of the comparison
            if type(e).__name__ not in raises:
                raise
            out.append(type(e).__name__)
    return out


def test_labels(theirs):
    raw = np.array([1, 2, 0, 1, 5, 5, 5], dtype=np.int64)
    for f in (lambda t: repr(t.Labels.from_sequence([1, 2, 0])),
              lambda t: str(t.Labels(raw.copy())),
              lambda t: list(t.Labels(raw.copy()).mapping.items()),         # key order included
              lambda t: t.Labels(raw.copy()).set,
              lambda t: t.Labels(raw.copy()).n_points,
              lambda t: t.Labels(raw.copy()).consider.tolist(),
              lambda t: t.Labels(raw.copy()).consider_set,
              lambda t: t.Labels(raw.copy()).meta,
              lambda t: t.Labels.from_length(3).labels.tolist(),
              lambda t: t.Labels(raw.copy(), meta={"origin": "x"}).meta):
        a, b = both(ours, theirs, f)
        assert a == b
    a, b = both(ours, theirs, lambda t: t.Labels(raw.copy(), consider=np.ones(3, dtype=np.uint8)), raises=("ValueError",))
    assert a == b == "ValueError"


def test_cluster_parameters(theirs):
    for f in (lambda t: repr(t.ClusterParameters(0.1, 2)), lambda t: t.ClusterParameters(0.1, 2).to_dict(),
              lambda t: str(t.ClusterParameters(0.1, 2, 3, 4, 5)),
              lambda t: (t.ClusterParameters(0.3, 7, 7.5, 8, 2).radius_cutoff, t.ClusterParameters(0.3, 7, 7.5, 8, 2).n_member_cutoff,
                         t.ClusterParameters(0.3, 7, 7.5, 8, 2).current_start)):
        a, b = both(ours, theirs, f)
        assert a == b


def test_input_data_types(theirs):
    x = np.arange(12, dtype=np.float64).reshape(6, 2)
    mk = (lambda: ours.InputDataComponents(x, meta={"edges": [4, 2]}),
          lambda: theirs.InputDataExtComponentsMemoryview(x, meta={"edges": [4, 2]}))
    for f in (lambda o: (o.n_points, o.n_dim), lambda o: dict(o.meta), lambda o: o.get_component(2, 1),
              lambda o: np.asarray(o.to_components_array()), lambda o: np.asarray(o.data),
              lambda o: [np.asarray(p).tolist() for p in o.by_parts()],
              lambda o: np.asarray(o.get_subset([1, 4, 5]).data), lambda o: np.asarray(o.get_subset([5, 1]).data)):
        a, b = both(*mk, f)
        assert a == b
    # deliberate: our subset carries the part lengths along (the reference recomputes them in Bundle.isolate)
    assert mk[0]().get_subset([1, 4, 5]).meta["edges"] == [1, 2] and "edges" not in mk[1]().get_subset([1, 4, 5]).meta
    padded = np.array([[0, 1, 0], [0, 1, 2], [2, 0, 0]], dtype=np.int64)
    counts = np.array([2, 3, 1], dtype=np.int64)
    mk = (lambda: ours.InputDataNeighbourhoods(padded, counts),
          lambda: theirs.InputDataExtNeighbourhoodsMemoryview(padded, counts))
    for f in (lambda o: (o.n_points, o.n_dim), lambda o: dict(o.meta), lambda o: np.asarray(o.n_neighbours),
              lambda o: o.get_n_neighbours(1), lambda o: o.get_neighbour(1, 2),
              lambda o: [np.asarray(r).tolist() for r in o.data]):
        a, b = both(*mk, f)
        assert a == b
    # deliberate: the reference's get_subset of this type raises (its constructor call lacks the counts); ours works
    a, b = both(*mk, lambda o: [np.asarray(r).tolist() for r in o.get_subset([0, 1]).data], raises=("TypeError",))
    assert a == [[0, 1], [0, 1]] and b == "TypeError"


def test_metrics_and_distance_getter(theirs):
    x = np.random.default_rng(0).random((6, 3))
    io, ir = ours.InputDataComponents(x), theirs.InputDataExtComponentsMemoryview(x)
    bounds = np.array([0.5, 0.0, 0.25])
    pairs = [(ours.MetricEuclideanReduced(), theirs.MetricExtEuclideanReduced()),
             (ours.MetricEuclidean(), theirs.MetricExtEuclidean()),
             (ours.MetricEuclideanPeriodicReduced(bounds), theirs.MetricExtEuclideanPeriodicReduced(bounds))]
    for mo, mr in pairs:
        assert mo.adjust_radius(0.3) == mr.adjust_radius(0.3)
        for a, b in [(1, 4), (0, 0), (5, 2)]:
            assert mo.calc_distance(a, b, io) == mr.calc_distance(a, b, ir)        # bit for bit
        assert ours.DistanceGetterMetric(mo).get_single(1, 4, io) == theirs.DistanceGetterExtMetric(mr).get_single(1, 4, ir)
    dm = np.abs(x[:, None, 0] - x[None, :, 0])
    assert ours.MetricPrecomputed().calc_distance(1, 4, ours.InputDataComponents(dm)) == \
        theirs.MetricExtPrecomputed().calc_distance(1, 4, theirs.InputDataExtComponentsMemoryview(dm))
    assert ours.MetricPrecomputed().adjust_radius(0.3) == theirs.MetricExtPrecomputed().adjust_radius(0.3) == 0.3
    lo, lr = ours.NeighboursGetterLookup(), theirs.NeighboursGetterExtLookup()
    assert (lo.is_sorted, lo.is_selfcounting) == (lr.is_sorted, lr.is_selfcounting) == (False, True)


def test_records_print_and_tabulate_like_the_reference(theirs):
    """`Record.__str__` (what `fit(v=True)` prints) and `Summary.to_DataFrame` (report.py:47-88, :142-176)."""
    from cnnclustering import report as rreport

    from commonnnclustering_b200 import report

    rows = [dict(n_points=100, radius_cutoff=0.1, cnn_cutoff=5, member_cutoff=None, max_clusters=None, n_clusters=3,
                 ratio_largest=0.5, ratio_noise=0.1, execution_time=0.01),
            dict(n_points=100, radius_cutoff=0.2, cnn_cutoff=5, member_cutoff=2, max_clusters=4, n_clusters=1,
                 ratio_largest=0.9, ratio_noise=0.0, execution_time=None),
            dict(n_points=10 ** 7, radius_cutoff=2, cnn_cutoff=50, member_cutoff=10, max_clusters=None, n_clusters=167,
                 ratio_largest=1 / 3, ratio_noise=2 / 3, execution_time=3725.0625)]
    a = report.Summary([report.Record(**r) for r in rows])
    b = rreport.Summary([rreport.Record(**r) for r in rows])
    for x, y in zip(a, b):
        assert str(x) == str(y) and repr(x) == repr(y) and x.to_dict() == y.to_dict()
    fa, fb = a.to_DataFrame(), b.to_DataFrame()
    assert fa.dtypes.to_dict() == fb.dtypes.to_dict() and fa.equals(fb)
    with pytest.raises(TypeError):
        a.append("not a record")
