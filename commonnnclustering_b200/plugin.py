"""Drop-in registration of the B200 fit path with the reference package (``cnnclustering`` 0.5.0).

The reference assembles its fitter from a recipe (``recipes.Builder``, ``recipes.py:182-313``) and
``Clustering.fit`` then calls exactly two methods on it (``cluster.pyx:328-343``):

    cluster_params = fitter.make_parameters(radius_cutoff, cnn_cutoff, current_start)
    fitter.fit(input_data, labels, cluster_params)          # mutates labels.labels / .consider in place

``register()`` adds a fitter with that contract, backed by ``libcnnb200.so``, to the reference's
registries -- no reference file is edited:

    import cnnclustering.cluster as cluster
    import commonnnclustering_b200.plugin as b200
    b200.register()
    clustering = cluster.Clustering(points, recipe="coordinates_b200")     # or fitter="b200"
    clustering.fit(radius_cutoff=0.3, cnn_cutoff=10)                        # runs on the GPU
    clustering.labels                                                       # identical to the CPU result

Everything outside ``fitter.fit`` (label bookkeeping, ``sort_by_size``, ``Record``, hierarchy) stays
the reference's own code.
"""
import numpy as np

from . import _lib
from ._types import MetricEuclidean, MetricEuclideanReduced

_METRICS = {"euclidean_r": MetricEuclideanReduced, "euclidean": MetricEuclidean}


def _padded_view(input_data):
    """(padded int64 [n, max_deg], counts int64 [n]) of an ``InputDataExtNeighbourhoodsMemoryview`` without touching a
    row in Python: ``n_neighbours`` is the counts array, and the rows its ``data`` accessor hands out are views into
    the padded matrix -- the matrix is the 2-D array at the end of the first row's ``base`` chain."""
    counts = getattr(input_data, "n_neighbours", None)
    rows = getattr(input_data, "data", None)
    if counts is None or not rows:
        return None, None
    base = rows[0]
    while base is not None and not (isinstance(base, np.ndarray) and base.ndim == 2):
        base = getattr(base, "base", None)
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    if base is None or base.dtype != np.int64 or not base.flags.c_contiguous or base.shape[0] != counts.shape[0]:
        return None, None
    return base, counts


def make_fitter_type(ref_types, ref_fit):
    """Builds the plug-in fitter class against an imported reference (its ABC and parameter type)."""

    class FitterB200Plugin:
        """``FitterExtBFS`` replacement (reference ``_fit.pyx:356-504``) running on one B200."""

        def __init__(self, metric="euclidean_r", method="auto", device=0, is_selfcounting=True, **_ignored):
            self.metric = _METRICS[metric]() if isinstance(metric, str) else metric
            self.method = method
            self.is_selfcounting = bool(is_selfcounting)
            self._device = device
            self._ctx = None

        @property
        def ctx(self):
            if self._ctx is None:
                self._ctx = _lib.default_context(self._device)
            return self._ctx

        def make_parameters(self, radius_cutoff, cnn_cutoff, current_start):
            """Same numbers as ``FitterExtInterface.make_parameters`` (``_fit.pyx:129-161``) yields for
            the reference's "coordinates" recipe, returned as the reference's own parameter type."""
            self._user_radius = radius_cutoff
            radius_cutoff = self.metric.adjust_radius(radius_cutoff)
            n_member_cutoff = cnn_cutoff
            if self.is_selfcounting:
                n_member_cutoff += 1
                cnn_cutoff += 2
            return ref_types.ClusterParameters(radius_cutoff, cnn_cutoff, cnn_cutoff, n_member_cutoff, current_start)

        def fit(self, input_data, labels, cluster_params):
            ctx = self.ctx
            if input_data.meta.get("access_coords", False):
                x = np.ascontiguousarray(input_data.to_components_array(), dtype=np.float64)
                ctx.neighbours_from_points(x, cluster_params.radius_cutoff, metric=self.metric.name, method=self.method)
            elif input_data.meta.get("access_neighbours", False):
                padded, counts = _padded_view(input_data)
                if padded is not None:
                    # the reference's own zero-padded int64 matrix (_types.pyx:1159-1184), ingested as it is
                    ctx.neighbours_from_padded(padded, counts)
                else:                    # any other container: its row accessor, one concatenation
                    rows = input_data.data
                    counts = np.fromiter((len(r) for r in rows), dtype=np.int64, count=len(rows))
                    row_ptr = np.zeros(len(rows) + 1, dtype=np.int64)
                    np.cumsum(counts, out=row_ptr[1:])
                    col = np.concatenate(rows).astype(np.int64) if row_ptr[-1] else np.zeros(0, dtype=np.int64)
                    ctx.neighbours_from_csr(row_ptr, col)
                # lists in any order are sorted by the ingest (the labels do not depend on the order within a list)
            else:
                raise TypeError("FitterB200Plugin needs coordinates or neighbourhoods as input data")
            lab = labels.labels          # NumPy view on the reference's buffer: written in place
            consider = labels.consider
            fresh = bool(consider.all()) and not lab.any()
            ctx.fit(cluster_params.n_member_cutoff, cluster_params.similarity_cutoff, cluster_params.current_start,
                    lab, None if fresh else consider)
            if fresh:
                consider[:] = 0

    ref_fit.Fitter.register(FitterB200Plugin)
    return FitterB200Plugin


_registered = {}


def register():
    """Registers ``fitter="b200"`` and the recipe ``"coordinates_b200"`` with the importable reference."""
    import cnnclustering._fit as ref_fit
    import cnnclustering._types as ref_types
    import cnnclustering.recipes as ref_recipes

    key = id(ref_recipes)
    if key in _registered:
        return _registered[key]
    cls = make_fitter_type(ref_types, ref_fit)
    ref_recipes.COMPONENT_NAME_TYPE_MAP["fitter"]["b200"] = cls
    ref_recipes.REGISTERED_RECIPES["coordinates_b200"] = {
        "input_data": "components_mview",
        "preparation_hook": "points_from_parts",
        "fitter": "b200",
    }
    ref_recipes.REGISTERED_RECIPES["sorted_neighbourhoods_b200"] = {
        "input_data": "neighbourhoods_mview",
        "preparation_hook": "neighbourhoods",
        "fitter": "b200",
    }
    _registered[key] = cls
    return cls
