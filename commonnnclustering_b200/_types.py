"""Component types of the B200 fit path -- the host-side mirror of the reference's ``_types``.

Same roles and names as the reference (``src/cnnclustering/_types.pyx``): ``ClusterParameters``,
``Labels``, InputData, NeighboursGetter, Metric, SimilarityChecker.  The reference drives them one
point / one pair at a time through virtual calls; here each role exposes a BULK method that runs
on the device through the C-ABI (``include/cnnb200.h``):

    NeighboursGetter.get_all(input_data, cluster_params)    -> Neighbourhoods (CSR in HBM)
    SimilarityChecker.get_all(neighbourhoods, cluster_params) -> common-neighbour count per edge
    Fitter.fit(input_data, labels, cluster_params)            (in ``_fit.py``)

The per-item reference methods that are cheap to answer from the bulk result (``get``, ``check``,
``calc_distance``) are provided too so that the reference's unit tests translate one to one.
"""
from collections import defaultdict
from typing import Optional

import numpy as np

from . import _lib

P_AINDEX = np.int64    # _primitive_types.pyx:4
P_AVALUE = np.float64  # _primitive_types.pyx:5
P_ABOOL = np.uint8     # _primitive_types.pyx:7


class ClusterParameters:
    """Reference: ``_types.pyx:22-85``."""

    def __init__(self, radius_cutoff: float, similarity_cutoff: int = 0,
                 similarity_cutoff_continuous: float = 0., n_member_cutoff: Optional[int] = None,
                 current_start: int = 1):
        if n_member_cutoff is None:
            n_member_cutoff = similarity_cutoff
        self.radius_cutoff = float(radius_cutoff)
        self.similarity_cutoff = int(similarity_cutoff)
        self.similarity_cutoff_continuous = float(similarity_cutoff_continuous)
        self.n_member_cutoff = int(n_member_cutoff)
        self.current_start = int(current_start)

    def to_dict(self):
        return {
            "radius_cutoff": self.radius_cutoff,
            "similarity_cutoff": self.similarity_cutoff,
            "similarity_cutoff_continuous": self.similarity_cutoff_continuous,
            "n_member_cutoff": self.n_member_cutoff,
            "current_start": self.current_start,
        }

    def __repr__(self):
        return f"{self.to_dict()!r}"

    __str__ = __repr__


class Labels:
    """Cluster label assignments.  Reference: ``_types.pyx:88-286``.

    ``labels`` (int64) and ``consider`` (uint8) are NumPy arrays that fitters mutate in place, as
    in the reference (``Labels.labels`` is a view, ``_types.pyx:143-144``)."""

    def __init__(self, labels, *, consider=None, meta=None):
        self._labels = np.ascontiguousarray(labels, dtype=P_AINDEX) if not (
            isinstance(labels, np.ndarray) and labels.dtype == P_AINDEX and labels.flags.c_contiguous) else labels
        if consider is None:
            self._consider = np.ones_like(self._labels, dtype=P_ABOOL)
        else:
            self._consider = consider if (
                isinstance(consider, np.ndarray) and consider.dtype == P_ABOOL and consider.flags.c_contiguous
            ) else np.ascontiguousarray(consider, dtype=P_ABOOL)
            if self._labels.shape[0] != self._consider.shape[0]:
                raise ValueError("'labels' and 'consider' must have the same length")
        self.meta = {} if meta is None else meta
        self._consider_set = set()

    @property
    def mapping(self):
        return self.to_mapping()

    @property
    def set(self):
        return self.to_set()

    @property
    def labels(self):
        return self._labels

    @property
    def consider(self):
        return self._consider

    @property
    def consider_set(self):
        return self._consider_set

    @consider_set.setter
    def consider_set(self, value):
        self._consider_set = value

    @property
    def n_points(self):
        return self._labels.shape[0]

    def __repr__(self):
        return f"{type(self).__name__}({list(self.labels)!s})"

    def __str__(self):
        return f"{self.labels!s}"

    @classmethod
    def from_sequence(cls, labels, *, consider=None, meta=None):
        labels = np.array(labels, order="C", dtype=P_AINDEX)
        if consider is not None:
            consider = np.array(consider, order="C", dtype=P_ABOOL)
        return cls(labels, consider=consider, meta=meta)

    @classmethod
    def from_length(cls, n: int, meta=None):
        return cls(np.zeros(n, order="C", dtype=P_AINDEX), meta=meta)

    def to_mapping(self):
        """label -> list of point indices, keys in the order in which the labels first occur (what the reference's
        loop over the points yields, ``_types.pyx:187-198``); vectorised."""
        mapping = defaultdict(list)
        if self._labels.size:
            order = np.argsort(self._labels, kind="stable")
            sl = self._labels[order]
            bounds = np.flatnonzero(np.diff(sl)) + 1
            chunks = np.split(order, bounds)           # each ascending: the sort is stable
            for chunk in sorted(chunks, key=lambda c: c[0]):
                mapping[int(self._labels[chunk[0]])] = chunk.tolist()
        return mapping

    def to_set(self):
        return set(np.unique(self._labels).tolist())

    def sort_by_size(self, member_cutoff: Optional[int] = None, max_clusters: Optional[int] = None,
                     bundle=None, ctx=None):
        """Re-number clusters by size, in place, on the device.  Reference: ``_types.pyx:212-286``."""
        if self._labels.size == 0:
            return
        ctx = ctx or _lib.default_context()
        max_label = int(self._labels.max())
        if max_label <= 0:
            return  # only noise: `if frequencies:` is false in the reference, nothing happens
        reassign = ctx.sort_by_size(self._labels, member_cutoff, max_clusters, return_map=True, max_label=max_label)
        params = self.meta.get("params", {})
        self.meta["params"] = {
            int(reassign[k]): v for k, v in params.items()
            if 0 <= k <= max_label and reassign[k] != 0
        }
        if bundle is not None and getattr(bundle, "_children", None) is not None:
            processed = {}
            for old_label, child in bundle._children.items():
                new_label = int(reassign[old_label]) if 0 <= old_label <= max_label else 0
                if new_label != 0:
                    processed[new_label] = child
            bundle._children = processed


def _pin_arrays(owner, arrays) -> bool:
    """cudaHostRegister on each (non-empty) array; released when `owner` is collected."""
    import weakref

    ok = False
    for a in arrays:
        if isinstance(a, np.ndarray) and a.nbytes and _lib.host_register(a):
            weakref.finalize(owner, _lib.host_unregister, a.ctypes.data)
            ok = True
    return ok


# ---- InputData ------------------------------------------------------------------------------------
class InputData:
    """Role marker (reference ABC ``InputData``, ``_types.pyx:305-436``)."""
    n_points: int
    n_dim: int
    meta: dict


class InputDataComponents(InputData):
    """Point coordinates, f64 row-major ``[n, d]`` (host NumPy array or CUDA tensor).

    Reference: ``InputDataExtComponentsMemoryview``, ``_types.pyx:1034-1109``."""

    def __init__(self, data, *, meta=None):
        if isinstance(data, np.ndarray) or not hasattr(data, "is_cuda"):
            data = np.ascontiguousarray(data, dtype=P_AVALUE)
            if data.ndim == 1:
                data = data.reshape(1, -1)
        else:
            import torch

            if data.dtype != torch.float64 or not data.is_contiguous():
                data = data.to(torch.float64).contiguous()
        if data.ndim != 2:
            raise ValueError("coordinates must be a 2-D array [n_points, n_dim]")
        self._data = data
        self.n_points, self.n_dim = int(data.shape[0]), int(data.shape[1])
        _meta = {"access_coords": True}
        if meta is not None:
            _meta.update(meta)
        self.meta = _meta

    @property
    def data(self):
        return self.to_components_array()

    def to_components_array(self):
        return self._data if isinstance(self._data, np.ndarray) else self._data.cpu().numpy()

    @property
    def device_or_host_array(self):
        return self._data

    def pin(self) -> bool:
        """Page-locks the host array in place so that repeated fits on these coordinates copy them at PCIe speed;
        nothing is copied or cached, later changes to the array are seen as before.  Released with the object.
        The FIRST fit copies from pageable memory: registering and releasing 24 MB costs 3-4 ms, more than the one
        pageable copy it would save -- the children of ``isolate`` are typically fitted once (C5: nine child fits,
        30 ms of registrations)."""
        if not isinstance(self._data, np.ndarray) or self._data.nbytes < (1 << 20):
            return False   # below a megabyte the registration costs more than the pageable copies it saves
        self._pin_calls = getattr(self, "_pin_calls", 0) + 1
        if self._pin_calls < 2:
            return False
        if getattr(self, "_pinned", None) is None:
            import weakref

            self._pinned = _lib.host_register(self._data)
            if self._pinned:
                weakref.finalize(self, _lib.host_unregister, self._data.ctypes.data)
        return self._pinned

    def get_component(self, point: int, dimension: int) -> float:
        return float(self._data[point, dimension])

    def get_subset(self, indices):
        """The points ``indices`` (ascending) as a new input data object; the part lengths ``meta["edges"]`` follow
        the points (what the reference's ``Bundle.isolate`` computes for a child, ``_bundle.pyx:321-350``)."""
        idx = np.asarray(indices, dtype=np.int64)
        sub = self._data[idx] if isinstance(self._data, np.ndarray) else self._data[idx.tolist()]
        meta = {}
        edges = self.meta.get("edges", None)
        if edges:
            bounds = np.cumsum(np.asarray(edges, dtype=np.int64))
            part = np.searchsorted(bounds, idx, side="right")
            meta["edges"] = np.bincount(part, minlength=len(edges))[:len(edges)].astype(int).tolist()
        return type(self)(sub, meta=meta)

    def by_parts(self):
        """Yield the parts delimited by ``meta["edges"]`` (``_types.pyx:1089-1109``)."""
        if self.n_points > 0:
            edges = self.meta.get("edges", None)
            if not edges:
                edges = [self.n_points]
            arr = self.to_components_array()
            start = 0
            for end in np.cumsum(edges):
                yield arr[start:end, :]
                start = end
        else:
            yield from ()


class InputDataDistancesLinear(InputData):
    """Pairwise distances in condensed form: the ``N (N - 1) / 2`` values ``d(a, b)``, ``a < b``, row by row (what
    ``scipy.spatial.distance.pdist`` returns).  Reference: ``InputDataExtDistancesLinearMemoryview``,
    ``_types.pyx:1112-1150``; used with ``MetricPrecomputed``.  Half the bytes of the square matrix, read once
    (``cnnb200_neighbours_from_condensed``)."""

    def __init__(self, data, *, n_points=None, meta=None):
        if isinstance(data, np.ndarray):
            data = np.ascontiguousarray(data, dtype=np.float64)
        if len(data.shape) != 1:
            raise ValueError("condensed distances are a 1-D sequence")
        self._data = data
        # the reference's own expression (_types.pyx:1121-1122)
        self.n_points = int(0.5 * (np.sqrt(8 * data.shape[0] + 1) + 1)) if n_points is None else int(n_points)
        if self.n_points * (self.n_points - 1) // 2 != data.shape[0]:
            raise ValueError(f"{data.shape[0]} values are not the condensed distances of {self.n_points} points")
        self.n_dim = 0
        _meta = {"access_distances": True}
        if meta is not None:
            _meta.update(meta)
        self.meta = _meta

    @property
    def data(self):
        return self._data if isinstance(self._data, np.ndarray) else self._data.cpu().numpy()

    @property
    def device_or_host_array(self):
        return self._data

    def pin(self) -> bool:
        """Page-locks the host vector in place from the second fit on (see InputDataComponents.pin)."""
        if not isinstance(self._data, np.ndarray) or self._data.nbytes < (1 << 20):
            return False
        self._pin_calls = getattr(self, "_pin_calls", 0) + 1
        if self._pin_calls < 2:
            return False
        if getattr(self, "_pinned", None) is None:
            self._pinned = _pin_arrays(self, [self._data])
        return self._pinned

    def get_distance(self, point_a: int, point_b: int) -> float:
        """``_get_distance`` of the reference (``_types.pyx:1129-1146``)."""
        if point_a == point_b:
            return 0.0
        a, b = (point_b, point_a) if point_a > point_b else (point_a, point_b)
        return float(self.data[a * (self.n_points - 1) - (a * a - a) // 2 + (b - a - 1)])

    def get_subset(self, indices):
        """The distances among the points ``indices`` (ascending), again in condensed form."""
        idx = np.asarray(indices, dtype=np.int64)
        m = idx.shape[0]
        ia, ib = np.triu_indices(m, k=1)
        a, b = idx[ia], idx[ib]
        return type(self)(self.data[a * (self.n_points - 1) - (a * a - a) // 2 + (b - a - 1)], n_points=m)


class InputDataNeighbourhoods(InputData):
    """Precomputed neighbourhoods in the reference's padded layout: int64 ``[n, max_deg]`` plus
    int64 counts ``[n]``.  Reference: ``InputDataExtNeighbourhoodsMemoryview``, ``_types.pyx:1153-1217``."""

    def __init__(self, data, n_neighbours, *, meta=None):
        self._data = np.ascontiguousarray(data, dtype=P_AINDEX)
        self._n_neighbours = np.ascontiguousarray(n_neighbours, dtype=P_AINDEX)
        if self._data.ndim != 2 or self._n_neighbours.shape != (self._data.shape[0],):
            raise ValueError("need data [n, max_deg] and n_neighbours [n]")
        self.n_points = int(self._data.shape[0])
        self.n_dim = 0
        _meta = {"access_neighbours": True}
        if meta is not None:
            _meta.update(meta)
        self.meta = _meta

    @property
    def data(self):
        return [row[:m] for row, m in zip(self._data, self._n_neighbours)]

    @property
    def n_neighbours(self):
        return self._n_neighbours

    @property
    def padded(self):
        return self._data

    def get_n_neighbours(self, point: int) -> int:
        return int(self._n_neighbours[point])

    def get_neighbour(self, point: int, member: int) -> int:
        return int(self._data[point, member])

    def pin(self) -> bool:
        """Page-locks the padded matrix and the counts in place, once (see InputDataComponents.pin)."""
        if getattr(self, "_pinned", None) is None:
            self._pinned = _pin_arrays(self, [self._data, self._n_neighbours])
        return self._pinned

    def to_csr(self):
        """(row_ptr int64 [n+1], col_idx int64 [nnz]) of the valid entries."""
        row_ptr = np.zeros(self.n_points + 1, dtype=P_AINDEX)
        np.cumsum(self._n_neighbours, out=row_ptr[1:])
        mask = np.arange(self._data.shape[1])[None, :] < self._n_neighbours[:, None]
        return row_ptr, np.ascontiguousarray(self._data[mask])

    def get_subset(self, indices):
        """The neighbourhoods of the points ``indices`` restricted to those points, members renumbered to their
        position in the subset.  (The reference keeps the old member numbers and drops the counts,
        ``_types.pyx:1199-1217`` -- SURVEY 8f notes the defect; child reclustering needs the renumbered form.)"""
        row_ptr, col_idx = _subset_csr(*self.to_csr(), np.asarray(indices, dtype=np.int64), self.n_points)
        counts = np.diff(row_ptr)
        padded = np.zeros((counts.shape[0], max(int(counts.max()) if counts.size else 0, 1)), dtype=P_AINDEX)
        rows = np.repeat(np.arange(counts.shape[0]), counts)
        padded[rows, np.arange(col_idx.shape[0]) - np.repeat(row_ptr[:-1], counts)] = col_idx
        return type(self)(padded, counts)


class InputDataNeighbourhoodsCSR(InputData):
    """Precomputed neighbourhoods in CSR form (row_ptr int64 ``[n+1]``, col_idx int32/int64
    ``[nnz]``; host arrays or CUDA tensors).  No reference twin: the padded layout costs
    8*n*max_deg bytes, which is what makes the reference's large runs memory bound."""

    def __init__(self, row_ptr, col_idx, *, meta=None):
        self.row_ptr = row_ptr
        self.col_idx = col_idx
        self.n_points = int(row_ptr.shape[0]) - 1
        self.n_dim = 0
        _meta = {"access_neighbours": True}
        if meta is not None:
            _meta.update(meta)
        self.meta = _meta

    def pin(self) -> bool:
        """Page-locks host CSR arrays in place, once (see InputDataComponents.pin); device tensors need nothing."""
        if getattr(self, "_pinned", None) is None:
            self._pinned = _pin_arrays(self, [a for a in (self.row_ptr, self.col_idx) if isinstance(a, np.ndarray)])
        return self._pinned

    def to_csr(self):
        rp, ci = self.row_ptr, self.col_idx
        if hasattr(rp, "is_cuda"):
            rp, ci = rp.cpu().numpy(), ci.cpu().numpy()
        return np.ascontiguousarray(rp, dtype=P_AINDEX), np.ascontiguousarray(ci, dtype=P_AINDEX)

    def get_subset(self, indices):
        """Rows of ``indices`` restricted to those points, members renumbered to their position in the subset."""
        return type(self)(*_subset_csr(*self.to_csr(), np.asarray(indices, dtype=np.int64), self.n_points))


def _subset_csr(row_ptr, col_idx, indices, n_points):
    """CSR of the sub-graph on ``indices`` (ascending, unique): rows of those points, members outside dropped, the
    rest renumbered; vectorised."""
    new_id = np.full(n_points, -1, dtype=np.int64)
    new_id[indices] = np.arange(indices.shape[0])
    counts = (row_ptr[1:] - row_ptr[:-1])[indices]
    starts = row_ptr[:-1][indices]
    flat = np.repeat(starts - (np.cumsum(counts) - counts), counts) + np.arange(int(counts.sum()))
    members = new_id[col_idx[flat]]
    keep = members >= 0
    rows = np.repeat(np.arange(indices.shape[0]), counts)[keep]
    sub_ptr = np.zeros(indices.shape[0] + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=indices.shape[0]), out=sub_ptr[1:])
    return sub_ptr, np.ascontiguousarray(members[keep])


class Neighbourhoods:
    """Handle on the CSR neighbourhood structure held in HBM by a context (the bulk analogue of
    the reference's per-point ``Neighbours`` scratch containers, ``_types.pyx:1285-1576``)."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.n_points, self.nnz, self.max_n_neighbours = ctx.neighbours_info()
        self._host = None

    def to_csr(self):
        """(row_ptr, col_idx) int64 in the caller's index space, rows ascending."""
        if self._host is None:
            self._host = self.ctx.neighbours_export()
        return self._host

    def get(self, index: int):
        row_ptr, col_idx = self.to_csr()
        return col_idx[row_ptr[index]:row_ptr[index + 1]]

    def enough(self, index: int, member_cutoff: int) -> bool:
        """``Neighbours.enough``: strictly more than ``member_cutoff`` (``_types.pyx:1408-1411``)."""
        row_ptr, _ = self.to_csr()
        return bool(row_ptr[index + 1] - row_ptr[index] > member_cutoff)


# ---- Metric -----------------------------------------------------------------------------------------
class Metric:
    name = None

    def adjust_radius(self, radius_cutoff: float) -> float:
        raise NotImplementedError


class MetricEuclideanReduced(Metric):
    """Squared Euclidean distance; ``adjust_radius(r) = r*r``.  Reference:
    ``MetricExtEuclideanReduced``, ``_types.pyx:2250-2287`` (recipe key ``euclidean_r``)."""
    name = "euclidean_r"

    def adjust_radius(self, radius_cutoff: float) -> float:
        r = float(radius_cutoff)
        return r * r

    def calc_distance(self, index_a: int, index_b: int, input_data) -> float:
        a = input_data.to_components_array()
        total = 0.0
        for k in range(a.shape[1]):
            t = float(a[index_a, k]) - float(a[index_b, k])
            total += t * t
        return total


class MetricEuclidean(Metric):
    """Euclidean distance; radius unchanged.  Reference: ``MetricExtEuclidean``, ``_types.pyx:2170-2207``."""
    name = "euclidean"

    def adjust_radius(self, radius_cutoff: float) -> float:
        return float(radius_cutoff)

    def calc_distance(self, index_a: int, index_b: int, input_data) -> float:
        return float(np.sqrt(MetricEuclideanReduced().calc_distance(index_a, index_b, input_data)))


class MetricEuclideanPeriodicReduced(Metric):
    """Squared Euclidean distance with per-dimension periods ``bounds`` (0 = not periodic).
    Reference: ``MetricExtEuclideanPeriodicReduced``, ``_types.pyx:2290-2348``."""
    name = "euclidean_periodic_r"

    def __init__(self, bounds):
        self.bounds = np.ascontiguousarray(bounds, dtype=P_AVALUE)

    def adjust_radius(self, radius_cutoff: float) -> float:
        return float(radius_cutoff) ** 2

    def calc_distance(self, index_a: int, index_b: int, input_data) -> float:
        x = input_data.to_components_array()
        total = 0.0
        for k in range(x.shape[1]):
            bound = float(self.bounds[k])
            distance = abs(float(x[index_a, k]) - float(x[index_b, k]))
            if bound > 0:
                distance = distance % bound
                if distance > bound / 2:
                    distance = bound - distance
            total += distance * distance
        return total


class MetricPrecomputed(Metric):
    """The input data ARE the distances: ``distance(a, b) = data[a, b]``; radius unchanged.
    Reference: ``MetricExtPrecomputed``, ``_types.pyx:2108-2127`` (recipe key ``precomputed``, the
    ``distances`` recipe)."""
    name = "precomputed"

    def adjust_radius(self, radius_cutoff: float) -> float:
        return float(radius_cutoff)

    def calc_distance(self, index_a: int, index_b: int, input_data) -> float:
        return float(input_data.to_components_array()[index_a, index_b])


class DistanceGetterMetric:
    """Reference: ``DistanceGetterExtMetric``, ``_types.pyx:1966-2017`` (pure indirection)."""

    def __init__(self, metric=None):
        self._metric = metric if metric is not None else MetricEuclideanReduced()

    @classmethod
    def get_builder_kwargs(cls):
        return [("metric", None)]

    def get_single(self, point_a: int, point_b: int, input_data) -> float:
        return self._metric.calc_distance(point_a, point_b, input_data)


# ---- NeighboursGetter ------------------------------------------------------------------------------
class NeighboursGetter:
    is_sorted = True
    is_selfcounting = True

    def get_all(self, input_data, cluster_params, ctx=None) -> Neighbourhoods:
        raise NotImplementedError

    def get(self, index: int, input_data, cluster_params, ctx=None):
        """Neighbours of one point (reference per-point ``get``); answered from the bulk result."""
        return self.get_all(input_data, cluster_params, ctx).get(index)


class NeighboursGetterRadius(NeighboursGetter):
    """Exact radius-neighbour search on coordinates for all points at once.

    Bulk replacement of ``NeighboursGetterExtBruteForce`` (``_types.pyx:1658-1735``): same neighbour
    sets (self included, ascending), found with a cell list for ``n_dim <= 3`` (``method="grid"``)
    or a tiled all-pairs kernel (``"tiled"`` exact f64, ``"tensor"`` tcgen05 filter + f64 recheck)."""

    def __init__(self, distance_getter=None, method: str = "auto"):
        self._distance_getter = distance_getter if distance_getter is not None else DistanceGetterMetric()
        if method not in _lib.METHOD_IDS:
            raise ValueError(f"unknown method {method!r}")
        self.method = method

    @classmethod
    def get_builder_kwargs(cls):
        return [("distance_getter", None)]

    @property
    def metric(self):
        return self._distance_getter._metric

    def prepare(self, ctx):
        """Hands metric state the C library cannot take through the metric id (periods) to the context."""
        if isinstance(self.metric, MetricEuclideanPeriodicReduced):
            ctx.set_periodic_bounds(self.metric.bounds)

    def get_all(self, input_data, cluster_params, ctx=None) -> Neighbourhoods:
        ctx = ctx or _lib.default_context()
        if isinstance(self.metric, MetricPrecomputed):
            # one thresholding pass over the N x N matrix (``cnnb200_neighbours_from_distances``); a host matrix is
            # page-locked in place once so that every further fit (parameter scans) streams it at PCIe speed
            if hasattr(input_data, "pin"):
                input_data.pin()
            if isinstance(input_data, InputDataDistancesLinear):
                ctx.neighbours_from_condensed(input_data.device_or_host_array, input_data.n_points, cluster_params.radius_cutoff)
            else:
                ctx.neighbours_from_distances(input_data.device_or_host_array, cluster_params.radius_cutoff)
            return Neighbourhoods(ctx)
        self.prepare(ctx)
        ctx.neighbours_from_points(input_data.device_or_host_array, cluster_params.radius_cutoff,
                                   metric=self.metric.name, method=self.method)
        return Neighbourhoods(ctx)


class NeighboursGetterGrid(NeighboursGetterRadius):
    def __init__(self, distance_getter=None):
        super().__init__(distance_getter, "grid")


class NeighboursGetterTiled(NeighboursGetterRadius):
    def __init__(self, distance_getter=None):
        super().__init__(distance_getter, "tiled")


class NeighboursGetterTensor(NeighboursGetterRadius):
    def __init__(self, distance_getter=None):
        super().__init__(distance_getter, "tensor")


class NeighboursGetterLookup(NeighboursGetter):
    """Precomputed neighbourhoods.  Bulk replacement of ``NeighboursGetterExtLookup``
    (``_types.pyx:1791-1832``): uploads the padded matrix or CSR once and converts it to int32 CSR.
    ``is_sorted=True`` (the reference's ``sorted_neighbourhoods`` recipe, ``recipes.py:349-357``) promises
    ascending rows; with ``is_sorted=False`` (the ``neighbourhoods`` recipe, ``recipes.py:340-348``, the
    reference's default for the getter, ``_types.pyx:1794``) the rows are checked on the device and sorted
    there when needed -- the labels do not depend on the order inside a list.

    The lists must be SYMMETRIC (``j in N(i)`` exactly when ``i in N(j)``), as radius neighbourhoods are: candidate
    edges are read as the entries ``j > i`` of row ``i``, and the union-find that replaces the reference's BFS equals
    it only on an undirected graph.  kNN-style lists violate this; ``check_symmetric=True`` counts
    the one-sided entries on the device (``cnnb200_neighbours_asymmetry``) and raises when there are any."""

    def __init__(self, is_sorted: bool = False, is_selfcounting: bool = True, check_symmetric: bool = False):
        self.is_sorted = bool(is_sorted)
        self.is_selfcounting = bool(is_selfcounting)
        self.check_symmetric = bool(check_symmetric)

    def get_all(self, input_data, cluster_params=None, ctx=None) -> Neighbourhoods:
        ctx = ctx or _lib.default_context()
        input_data.pin()
        if isinstance(input_data, InputDataNeighbourhoodsCSR):
            ctx.neighbours_from_csr(input_data.row_ptr, input_data.col_idx,
                                    borrow=hasattr(input_data.col_idx, "is_cuda"))
        elif isinstance(input_data, InputDataNeighbourhoods):
            ctx.neighbours_from_padded(input_data.padded, input_data.n_neighbours)
        else:
            raise TypeError("NeighboursGetterLookup needs InputDataNeighbourhoods or InputDataNeighbourhoodsCSR")
        if not self.is_sorted:
            ctx.neighbours_sort_rows()
        if self.check_symmetric:
            one_sided = ctx.neighbours_asymmetry()
            if one_sided:
                raise ValueError(f"neighbourhoods are not symmetric: {one_sided} entries (i, j) without (j, i)")
        return Neighbourhoods(ctx)


# ---- SimilarityChecker -----------------------------------------------------------------------------
class SimilarityChecker:
    pass


class SimilarityCheckerCommonNN(SimilarityChecker):
    """Common-nearest-neighbour criterion over sorted CSR rows, evaluated by warp-cooperative
    intersection on the device.

    Bulk replacement of ``SimilarityCheckerExtContains`` / ``...SwitchContains`` /
    ``...Screensorted`` (``_types.pyx:2496-2753``): ``check`` = at least ``similarity_cutoff``
    common neighbours (early exit), ``get`` = the full count."""

    def get_all(self, neighbourhoods: Neighbourhoods, cluster_params, out=None):
        """uint32 count per CSR slot in the device's CSR order (0xFFFFFFFF where the slot is not a
        candidate edge i<j between two points with more than ``n_member_cutoff`` neighbours)."""
        return neighbourhoods.ctx.cn_count(cluster_params.n_member_cutoff, out=out)

    # per-pair forms of the reference, for small inputs / tests: two ascending index lists
    def get(self, neighbours_a, neighbours_b, cluster_params=None, ctx=None) -> int:
        ctx = ctx or _lib.default_context()
        a = np.unique(np.asarray(neighbours_a, dtype=np.int64))
        b = np.unique(np.asarray(neighbours_b, dtype=np.int64))
        if a.size == 0 or b.size == 0:
            return 0
        # a two-row neighbourhood structure over a private index space: row 0 = a, row 1 = b, where
        # both rows additionally contain each other (0 and 1 are fresh indices) so that (0, 1) is an edge
        both = np.unique(np.concatenate([a, b]))
        ra = np.searchsorted(both, a) + 2
        rb = np.searchsorted(both, b) + 2
        n = both.size + 2
        rows = [np.concatenate([[1], ra]), np.concatenate([[0], rb])] + [np.empty(0, np.int64)] * both.size
        row_ptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum([len(r) for r in rows], out=row_ptr[1:])
        col_idx = np.concatenate(rows).astype(np.int64)
        ctx.neighbours_from_csr(row_ptr, col_idx)
        counts, _ = ctx.cn_count(0)
        return int(counts[0])

    def check(self, neighbours_a, neighbours_b, cluster_params, ctx=None) -> bool:
        """``SimilarityChecker.check`` of the reference for two index lists, answered by the EARLY-EXIT kernels the fit
        uses (per-lane merge and whole-warp window form must agree), not by the count."""
        c = cluster_params.similarity_cutoff
        if c == 0:
            return True  # _types.pyx:2589-2590
        ctx = ctx or _lib.default_context()
        a = np.unique(np.asarray(neighbours_a, dtype=np.int64))
        b = np.unique(np.asarray(neighbours_b, dtype=np.int64))
        if a.size == 0 or b.size == 0:
            return False
        both = np.unique(np.concatenate([a, b]))
        rows = [np.searchsorted(both, a), np.searchsorted(both, b)] + [np.empty(0, np.int64)] * max(both.size - 2, 0)
        row_ptr = np.zeros(len(rows) + 1, dtype=np.int64)
        np.cumsum([len(r) for r in rows], out=row_ptr[1:])
        ctx.neighbours_from_csr(row_ptr, np.concatenate(rows).astype(np.int64))
        bits = int(ctx.cn_check_pairs([[0, 1]], c)[0])
        if bits not in (0, 3):
            raise RuntimeError("the two early-exit forms of the common-neighbour check disagree")
        return bits == 3
