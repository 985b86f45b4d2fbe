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
    """Same result as ``FitterB200`` followed by ``Labels.sort_by_size`` but as ONE C-ABI call
    (``cnnb200_fit_points``) that keeps the labels on the device between the two steps.  Used by
    ``Clustering.fit`` for fresh fits on coordinates."""

    def fit_sorted(self, input_data: InputDataComponents, labels: Labels, cluster_params: ClusterParameters,
                   member_cutoff=None, max_clusters=None) -> int:
        ctx = self.ctx
        g = self._neighbours_getter
        g.prepare(ctx)
        n_clusters = ctx.fit_points(
            input_data.device_or_host_array, cluster_params.radius_cutoff, cluster_params.n_member_cutoff,
            cluster_params.similarity_cutoff, labels.labels, cluster_params.current_start,
            metric=g.metric.name, method=g.method, sort_by_size=True,
            member_cutoff=member_cutoff, max_clusters=max_clusters)
        labels.consider[:] = 0
        labels.meta["b200"] = dict(ctx.stats(), n_points=int(labels.labels.shape[0]))
        return n_clusters


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
