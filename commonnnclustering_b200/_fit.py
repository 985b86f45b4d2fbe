"""Fitter of the B200 fit path -- host-side mirror of the reference's ``_fit`` module.

``FitterB200`` takes the place of ``FitterExtBFS`` (reference ``_fit.pyx:356-504``): the same
``make_parameters(radius_cutoff, cnn_cutoff, current_start)`` and ``fit(input_data, labels,
cluster_params)`` contract (``Clustering.fit`` calls exactly these two, ``cluster.pyx:328-343``),
the same in-place effect on ``labels.labels`` / ``labels.consider``, bit-identical label values.
"""
import numpy as np

from . import _lib
from ._types import (ClusterParameters, InputDataComponents, Labels, NeighboursGetter,
                     NeighboursGetterLookup, NeighboursGetterRadius, SimilarityCheckerCommonNN)


class Fitter:
    """Role marker (reference ABC ``Fitter``, ``_fit.pyx:43-85``)."""

    def make_parameters(self, radius_cutoff: float, cnn_cutoff: int, current_start: int) -> ClusterParameters:
        """Reference: ``FitterExtInterface.make_parameters``, ``_fit.pyx:129-161``."""
        try:
            used_metric = self._neighbours_getter._distance_getter._metric
        except AttributeError:
            pass
        else:
            radius_cutoff = used_metric.adjust_radius(radius_cutoff)
        n_member_cutoff = cnn_cutoff
        try:
            is_selfcounting = self._neighbours_getter.is_selfcounting
        except AttributeError:
            pass
        else:
            if is_selfcounting:
                n_member_cutoff += 1
                cnn_cutoff += 2
        return ClusterParameters(radius_cutoff, cnn_cutoff, cnn_cutoff, n_member_cutoff, current_start)


class FitterB200(Fitter):
    """Density-connected components of the common-nearest-neighbour graph, on one B200.

    Components (reference builder keywords, ``_fit.pyx:398-406``): ``neighbours_getter`` (bulk
    radius search or lookup), ``similarity_checker`` (bulk common-neighbour criterion).  The
    reference's ``neighbours`` / ``neighbour_neighbours`` scratch containers and ``queue`` have no
    device analogue (CSR rows and a concurrent union-find replace them)."""

    def __init__(self, neighbours_getter: NeighboursGetter = None, similarity_checker=None, ctx=None, **_ignored):
        self._neighbours_getter = neighbours_getter if neighbours_getter is not None else NeighboursGetterRadius()
        self._similarity_checker = similarity_checker if similarity_checker is not None else SimilarityCheckerCommonNN()
        self._ctx = ctx
        self.last_neighbourhoods = None

    @classmethod
    def get_builder_kwargs(cls):
        return [("neighbours_getter", None), ("similarity_checker", None)]

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = _lib.default_context()
        return self._ctx

    def fit(self, input_data, labels: Labels, cluster_params: ClusterParameters, reuse_neighbourhoods=False) -> None:
        """Writes cluster numbers ``current_start, current_start+1, ...`` (0 = noise) into
        ``labels.labels`` in place and clears ``labels.consider`` -- exactly the reference's
        ``FitterExtBFS.fit`` (``_fit.pyx:408-504``)."""
        ctx = self.ctx
        lab = labels.labels
        if lab.shape[0] != input_data.n_points:
            raise ValueError("labels and input data differ in length")
        if not (reuse_neighbourhoods and self.last_neighbourhoods is not None):
            self.last_neighbourhoods = self._neighbours_getter.get_all(input_data, cluster_params, ctx)
        consider = labels.consider
        fresh = bool(consider.all())
        if fresh and not lab.any():
            # fresh fit: nothing to upload, labels are fully overwritten
            ctx.fit(cluster_params.n_member_cutoff, cluster_params.similarity_cutoff,
                    cluster_params.current_start, lab, None)
            consider[:] = 0
        else:
            ctx.fit(cluster_params.n_member_cutoff, cluster_params.similarity_cutoff,
                    cluster_params.current_start, lab, consider)
        labels.meta["b200"] = dict(ctx.stats(), n_points=int(lab.shape[0]), nnz=self.last_neighbourhoods.nnz)


class FitterB200Fused(FitterB200):
    """Same result as ``FitterB200`` followed by ``Labels.sort_by_size`` but as ONE C-ABI call that keeps the labels on
    the device between the two steps (``cnnb200_fit_points`` on coordinates, ``cnnb200_fit_sorted`` on neighbourhoods
    that were ingested first).  Used by ``Clustering.fit`` for fresh fits with size sorting."""

    def fit_sorted(self, input_data, labels: Labels, cluster_params: ClusterParameters, member_cutoff=None,
                   max_clusters=None) -> int:
        ctx = self.ctx
        g = self._neighbours_getter
        if isinstance(input_data, InputDataComponents) and isinstance(g, NeighboursGetterRadius) \
                and g.metric.name in _lib.METRIC_IDS:
            g.prepare(ctx)
            n_clusters = ctx.fit_points(
                input_data.device_or_host_array, cluster_params.radius_cutoff, cluster_params.n_member_cutoff,
                cluster_params.similarity_cutoff, labels.labels, cluster_params.current_start,
                metric=g.metric.name, method=g.method, sort_by_size=True,
                member_cutoff=member_cutoff, max_clusters=max_clusters)
        else:
            self.last_neighbourhoods = g.get_all(input_data, cluster_params, ctx)
            n_clusters = ctx.fit_sorted(cluster_params.n_member_cutoff, cluster_params.similarity_cutoff, labels.labels,
                                        cluster_params.current_start, member_cutoff, max_clusters)
        labels.consider[:] = 0
        labels.meta["b200"] = dict(ctx.stats(), n_points=int(labels.labels.shape[0]))
        return n_clusters


class HierarchicalFitter:
    """Role marker (reference ABC ``HierarchicalFitter``, ``_fit.pyx:86-126``).  Its ``make_parameters`` has the same
    body as ``Fitter.make_parameters`` and looks at ``self._neighbours_getter`` -- an attribute
    ``HierarchicalFitterRepeat`` does not have, so for that class NEITHER the radius is adjusted NOR are the cutoffs
    raised for self-counting neighbourhoods: the wrapped fitter receives ``(r, c, c, c, start)`` as they were given.
    Kept as it is (bit-exact parity)."""

    make_parameters = Fitter.make_parameters


class HierarchicalFitterRepeat(HierarchicalFitter):
    """A hierarchy from repeated flat fits on the ROOT data with a sequence of ``(radius_cutoff, cnn_cutoff)`` pairs
    (reference ``HierarchicalFitterRepeat``, ``_fit.pyx:1018-1184``): after every step each cluster of the previous
    step ("terminal" bundle) receives the current labels of its members as its own label array and is isolated into
    children, which become the terminal bundles of the next step.

    The per-step fit runs on the device; so does the grouping of the points by (previous label, current label)
    (``cnnb200_labels_group``: two radix sorts instead of the reference's per-point ``parent_labels_map`` loop and
    one ``Labels.mapping`` pass per parent), after which every bundle is a few NumPy slices."""

    def __init__(self, fitter=None):
        self._fitter = fitter if fitter is not None else FitterB200()

    def __str__(self):
        return f"{type(self).__name__}(fitter={self._fitter})"

    @classmethod
    def get_builder_kwargs(cls):
        return [("fitter", None)]

    def fit(self, bundle, **kwargs):
        self._fit(bundle, kwargs["radius_cutoff"], kwargs["cnn_cutoff"], kwargs.get("member_cutoff"),
                  kwargs.get("max_clusters"), kwargs.get("cnn_offset"), kwargs.get("sort_by_size", True),
                  kwargs.get("info", True), kwargs.get("v", True))

    def _fit(self, bundle, radius_cutoff, cnn_cutoff, member_cutoff=None, max_clusters=None, cnn_offset=None,
             sort_by_size=True, info=True, v=True):
        import weakref
        from collections.abc import Iterable

        from .cluster import Bundle

        radius_cutoff = [float(x) for x in (radius_cutoff if isinstance(radius_cutoff, Iterable) else [radius_cutoff])]
        cnn_cutoff = [int(x) for x in (cnn_cutoff if isinstance(cnn_cutoff, Iterable) else [cnn_cutoff])]
        if len(radius_cutoff) == 1:
            radius_cutoff *= len(cnn_cutoff)
        if len(cnn_cutoff) == 1:
            cnn_cutoff *= len(radius_cutoff)
        n_steps = len(radius_cutoff)
        assert n_steps == len(cnn_cutoff)
        if cnn_offset is None:
            cnn_offset = 0
        ctx = self._fitter.ctx
        n_points = bundle._input_data.n_points
        previous = np.ones(n_points, dtype=np.int64)
        terminal = {1: bundle}
        for step in range(n_steps):
            if v:
                print(f"Running step {step:<5} (r = {radius_cutoff[step]}, c = {cnn_cutoff[step]})")
            current = Labels(np.zeros(n_points, order="C", dtype=np.int64))
            cluster_params = self.make_parameters(radius_cutoff[step], cnn_cutoff[step] - cnn_offset, 1)
            self._fitter.fit(bundle._input_data, current, cluster_params)
            if sort_by_size:
                current.sort_by_size(member_cutoff, max_clusters, ctx=ctx)
            cur = current.labels
            # all (parent, child) groups of the step at once; points of parent 0 (noise of the previous step) drop out
            order, rank = ctx.labels_group(previous, cur)
            p_sorted, c_sorted = previous[order], cur[order]
            first = int(np.searchsorted(p_sorted, 1, side="left"))
            new_terminal = {}
            p_starts = first + np.flatnonzero(np.r_[True, p_sorted[first + 1:] != p_sorted[first:-1]]) if first < n_points \
                else np.zeros(0, dtype=np.int64)
            p_ends = np.r_[p_starts[1:], n_points]
            for ps, pe in zip(p_starts.tolist(), p_ends.tolist()):
                p_label = int(p_sorted[ps])
                parent = terminal[p_label]
                lab = np.empty(pe - ps, dtype=np.int64)
                lab[rank[ps:pe]] = c_sorted[ps:pe]        # the current labels of the parent's members in point order
                parent._labels = Labels(lab)
                if info:
                    params = {int(k): (radius_cutoff[step], cnn_cutoff[step]) for k in np.unique(lab).tolist() if k != 0}
                    parent._labels.meta.update({"params": params, "reference": weakref.proxy(parent), "origin": "fit"})
                # Bundle.isolate(isolate_input_data=False) from the same sort: children in order of first occurrence
                parent._children = {}
                cs = ps + np.flatnonzero(np.r_[True, c_sorted[ps + 1:pe] != c_sorted[ps:pe - 1]])
                ce = np.r_[cs[1:], pe]
                groups = [(int(rank[a]), a, b) for a, b in zip(cs.tolist(), ce.tolist()) if c_sorted[a] != 0]
                for _, a, b in sorted(groups):
                    c_label = int(c_sorted[a])
                    child = Bundle(parent=parent, alias=f"{parent.alias} - {c_label}")
                    child.parent_indices = rank[a:b].astype(np.intp)      # ascending: the sorts are stable
                    child.root_indices = child.parent_indices if parent.root_indices is None \
                        else parent.root_indices[child.parent_indices]
                    parent._children[c_label] = child
                    new_terminal[c_label] = child
            terminal = new_terminal
            previous = cur


class Predictor:
    """Role marker (reference ABC ``Predictor``, ``_fit.pyx:164-216``)."""

    make_parameters = Fitter.make_parameters


class PredictorB200(Predictor):
    """Cluster labels for points outside the clustered set (reference ``PredictorFirstmatch``,
    ``_fit.pyx:1189-1291``): a predictand point with enough neighbours in the clustered set gets the label of the
    first of those neighbours (ascending index) that belongs to a considered cluster, has enough neighbours itself
    and passes the common-neighbour check against the point's neighbour list."""

    def __init__(self, neighbours_getter: NeighboursGetter = None, neighbours_getter_other=None, similarity_checker=None,
                 ctx=None, **_ignored):
        self._neighbours_getter = neighbours_getter if neighbours_getter is not None else NeighboursGetterRadius()
        self._neighbours_getter_other = neighbours_getter_other if neighbours_getter_other is not None else self._neighbours_getter
        self._similarity_checker = similarity_checker if similarity_checker is not None else SimilarityCheckerCommonNN()
        self._ctx = ctx

    @classmethod
    def get_builder_kwargs(cls):
        return [("neighbours_getter", None), ("neighbours_getter_other", "neighbours_getter"), ("similarity_checker", None)]

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = _lib.default_context()
        return self._ctx

    def predict(self, input_data, predictand_input_data, labels: Labels, predictand_labels: Labels,
                cluster_params: ClusterParameters) -> None:
        ctx = self.ctx
        self._neighbours_getter.get_all(input_data, cluster_params, ctx)
        if isinstance(predictand_input_data, InputDataComponents):
            ctx.neighbours_other_from_points(predictand_input_data.to_components_array())
        else:  # neighbourhoods of the predictand points IN the clustered set
            row_ptr, col_idx = predictand_input_data.to_csr()
            ctx.neighbours_other_from_csr(row_ptr, col_idx)
        lab = labels.labels
        max_label = int(lab.max()) if lab.size else 0
        label_ok = np.zeros(max_label + 1, dtype=np.uint8)
        for k in predictand_labels.consider_set:
            if 0 <= k <= max_label:
                label_ok[k] = 1
        ctx.predict(cluster_params.n_member_cutoff, cluster_params.similarity_cutoff, lab, label_ok,
                    predictand_labels.labels, predictand_labels.consider)
