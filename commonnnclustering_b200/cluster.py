"""User API -- mirror of the reference's ``cluster.Clustering`` for the fit path
(``cluster.pyx:102-147`` construction, ``:256-392`` ``fit``, ``:155-177`` ``labels``)."""
import time
import weakref
from collections.abc import Iterable, MutableMapping, MutableSequence

import numpy as np

from . import _lib, recipes
from ._fit import Fitter, FitterB200Fused, HierarchicalFitter, Predictor
from ._types import InputData, InputDataComponents, Labels
from .report import Record, Summary


def _label_buffer(n: int, pinned: bool):
    """Zeroed int64 labels; page-locked (torch's caching host allocator) when the fit writes them with one D2H copy."""
    if pinned and n > 0:
        try:
            import torch

            if torch.cuda.is_available():
                return torch.zeros(n, dtype=torch.int64, pin_memory=True).numpy()
        except Exception:
            pass
    return np.zeros(n, order="C", dtype=np.int64)


def _label_path(label):
    """A child address as a list of integers: 3, [1, 1, 2] and "1.1.2" are all accepted (reference ``get_child``,
    ``_bundle.pyx:174-218``)."""
    if isinstance(label, str):
        label = label.split(".")
    if isinstance(label, Iterable):
        return [int(part) for part in label]
    return [int(label)]


class Bundle:
    """Input data + labels + summary of one clustering and its place in a hierarchy (reference
    ``_bundle.pyx:20-273``): ``children`` maps a cluster label to the Bundle of that cluster's points,
    ``parent_indices`` / ``root_indices`` say where a child's points sit in its parent / the root
    (``ReferenceIndices``, ``_types.pyx:289-305``).  As in the reference the public ``labels`` / ``input_data`` hand out
    the ARRAYS (``Labels.labels``, ``InputData.data``); the objects themselves are ``_labels`` / ``_input_data``."""

    def __init__(self, input_data=None, labels=None, summary=None, parent=None, alias=None, children=None, meta=None,
                 hierarchy_level=0):
        self.input_data = input_data
        self.labels = labels
        self.summary = summary
        self.children = children
        self._parent = parent
        self.alias = "root" if alias is None else str(alias)
        self.hierarchy_level = parent.hierarchy_level + 1 if parent is not None else hierarchy_level
        self.meta = meta if meta is not None else {}
        self.parent_indices = None
        self.root_indices = None

    parent = property(lambda self: self._parent)

    @property
    def labels(self):
        return None if self._labels is None else self._labels.labels

    @labels.setter
    def labels(self, value):
        if value is not None and not isinstance(value, Labels):
            value = Labels.from_sequence(value)
        self._labels = value

    @property
    def input_data(self):
        return None if self._input_data is None else self._input_data.data

    @input_data.setter
    def input_data(self, value):
        if value is not None and not isinstance(value, InputData):
            raise TypeError(f"Can't use object of type {type(value).__name__} as input data. "
                            f"Expected type {InputData.__name__}.")
        self._input_data = value

    @property
    def children(self):
        return self._children

    @children.setter
    def children(self, value):
        if value is None:
            value = {}
        if not isinstance(value, MutableMapping):
            raise TypeError("Expected a mutable mapping")
        self._children = value

    @property
    def summary(self):
        return self._summary

    @summary.setter
    def summary(self, value):
        if value is None:
            value = Summary()
        if not isinstance(value, MutableSequence):
            raise TypeError("Expected a mutable sequence")
        self._summary = value

    @property
    def hierarchy_level(self):
        return self._hierarchy_level

    @hierarchy_level.setter
    def hierarchy_level(self, value):
        self._hierarchy_level = int(value)

    def info(self) -> str:
        """The reference's short description (``_bundle.pyx:148-172``)."""
        meta = self._input_data.meta if self._input_data is not None else {}
        kinds = [kind for kind, flag in (("coordinates", "access_coords"), ("distances", "access_distances"),
                                         ("neighbours", "access_neighbours")) if meta.get(flag, False)]
        lines = {"alias": repr(self.alias), "hierarchy_level": self._hierarchy_level,
                 "access": ", ".join(kinds or ["unknown"]),
                 "points": self._input_data.n_points if self._input_data is not None else None,
                 "children": len(self._children)}
        return "\n".join(f"{k}: {v}" for k, v in lines.items())

    def get_child(self, label):
        """The child with this label; a list or a dotted string walks down the hierarchy."""
        node = self
        for step in _label_path(label):
            try:
                node = node._children[step]
            except KeyError:
                raise KeyError(f"Clustering {node.alias!r} has no child with label {step}") from None
        return node

    def add_child(self, label, bundle=None):
        """Attach ``bundle`` (a new empty one by default) under the address ``label``; an existing child is replaced."""
        *above, last = _label_path(label)
        parent = self.get_child(above) if above else self
        if bundle is None:
            bundle = type(self)(parent=parent)
        if not isinstance(bundle, Bundle):
            raise TypeError("Expected a Bundle")
        parent._children[last] = bundle

    def __getitem__(self, key):
        return self.get_child(key)

    def __setitem__(self, key, value):
        self.add_child(key, value)

    def isolate(self, purge: bool = True, isolate_input_data: bool = True, ctx=None) -> None:
        """Child bundles from the cluster labels (reference ``_bundle.pyx:275-352``): one child per
        label != 0 holding the indices of its members (ascending) and, optionally, that subset of the
        input data.  The member lists come from one stable sort by label instead of a Python loop -- on the
        device (``cnnb200_labels_group``) when a context is given and the set is large."""
        if purge or self._children is None:
            self._children = {}
        lab = self._labels.labels
        if ctx is not None and lab.shape[0] >= 100_000 and lab.size and lab.min() >= 0:
            order, _ = ctx.labels_group(lab, np.zeros(lab.shape[0], dtype=np.int64))
        else:
            order = np.argsort(lab, kind="stable")
        sorted_lab = lab[order]
        starts = np.flatnonzero(np.r_[True, sorted_lab[1:] != sorted_lab[:-1]])
        ends = np.r_[starts[1:], sorted_lab.shape[0]]
        # children in order of first occurrence of their label, as the reference's label mapping yields them
        # (reel() shifts child labels by the parent's running maximum, so the order is observable)
        for s, e in sorted(zip(starts, ends), key=lambda se: order[se[0]]):
            label = int(sorted_lab[s])
            if label == 0:
                continue
            parent_indices = order[s:e].astype(np.intp)   # ascending: the argsort is stable
            child = Bundle(parent=self, alias=f"{self.alias} - {label}")
            child.parent_indices = parent_indices
            child.root_indices = parent_indices if self.root_indices is None else self.root_indices[parent_indices]
            if isolate_input_data:
                child._input_data = self._input_data.get_subset(parent_indices)
            self._children[label] = child


def bfs_leafs(root: Bundle):
    """All bundles of a hierarchy, breadth first, the root included (reference ``_bundle.pyx:403-414``)."""
    level = [root]
    yield root
    while level:
        level = [child for node in level for child in (node._children or {}).values()]
        yield from level


class Clustering:
    """``Clustering(data).fit(radius_cutoff, cnn_cutoff, member_cutoff)`` on a B200."""

    def __init__(self, data=None, *, fitter=None, hierarchical_fitter=None, predictor=None, bundle_kwargs=None, recipe=None,
                 **recipe_kwargs):
        builder = recipes.Builder(recipe, **recipe_kwargs)
        bundle_kwargs = dict(bundle_kwargs or {})     # e.g. alias, summary, meta of the root bundle (cluster.pyx:102-131)
        self._hierarchical_fitter = hierarchical_fitter if isinstance(hierarchical_fitter, HierarchicalFitter) else None
        if hierarchical_fitter is not None and self._hierarchical_fitter is None:
            builder.set("hierarchical_fitter", hierarchical_fitter)
        self._predictor = predictor if isinstance(predictor, Predictor) else None
        if predictor is not None and self._predictor is None:
            builder.set("predictor", predictor)
        if data is None:
            self._bundle = None
        elif isinstance(data, Bundle):
            self._bundle = data
        elif isinstance(data, InputData):
            self._bundle = Bundle(data, **bundle_kwargs)
        else:
            self._bundle = Bundle(builder.make_input_data(data), **bundle_kwargs)
        if isinstance(fitter, Fitter):
            self._fitter = fitter
        else:
            if fitter is not None:
                builder.set("fitter", fitter)
            f = builder.make_component("fitter")
            self._fitter = None if f is object else f
        if self._predictor is None:
            pr = builder.make_component("predictor")
            self._predictor = None if pr is object else pr
        if self._hierarchical_fitter is None:
            hf = builder.make_component("hierarchical_fitter")
            self._hierarchical_fitter = None if hf is object else hf

    @property
    def root(self):
        return self._bundle

    @property
    def input_data(self):
        return None if self._bundle is None or self._bundle._input_data is None else self._bundle._input_data.data

    @property
    def labels(self):
        """Cluster label assignments as a NumPy array (reference ``cluster.pyx:155-177``)."""
        if self._bundle is None or self._bundle._labels is None:
            return None
        return self._bundle._labels.labels

    @labels.setter
    def labels(self, value):
        if self._bundle is None:
            raise ValueError("Can't set labels because there is no root bundle")
        self._bundle.labels = value

    @property
    def summary(self):
        return self._bundle._summary

    @property
    def children(self):
        return self._bundle._children

    def __str__(self):
        parts = {"input_data": self._bundle._input_data if self._bundle is not None else None, "fitter": self._fitter,
                 "hfitter": self._hierarchical_fitter, "predictor": self._predictor}
        return f"{type(self).__name__}({', '.join(f'{k}={v}' for k, v in parts.items())})"

    def __getitem__(self, value):
        """``clustering[3]``, ``clustering["1.1.2"]``: a bundle further down the hierarchy (``cluster.pyx:227-228``)."""
        return self._bundle.get_child(value)

    def _fit(self, cluster_params, bundle=None) -> None:
        """Low-level fit (reference ``cluster.pyx:230-254``): the parameters go to the fitter as they are, the labels
        are not initialised."""
        if bundle is None:
            bundle = self._bundle
        self._fitter.fit(bundle._input_data, bundle._labels, cluster_params)

    def isolate(self, purge: bool = True, isolate_input_data: bool = True, bundle=None) -> None:
        """Reference ``Clustering.isolate`` (``cluster.pyx:394-413``)."""
        target = bundle if bundle is not None else self._bundle
        ctx = None
        if target._labels is not None and target._labels.labels.shape[0] >= 100_000 and self._fitter is not None:
            try:          # large sets: the grouping by label runs on the device
                ctx = getattr(self._fitter, "ctx", None)
            except _lib.CnnB200Error:
                ctx = None    # no device here: the NumPy grouping (isolate is host logic and must not need one)
        target.isolate(purge, isolate_input_data, ctx=ctx)

    def fit_hierarchical(self, purge=True, bundle=None, **fitter_kwargs) -> None:
        """Reference ``Clustering.fit_hierarchical`` (``cluster.pyx:394-411``): the keyword arguments are those of the
        hierarchical fitter (``radius_cutoff`` / ``cnn_cutoff`` sequences, ``member_cutoff``, ``max_clusters``,
        ``cnn_offset``, ``sort_by_size``, ``info``, ``v`` for ``HierarchicalFitterRepeat``)."""
        if bundle is None:
            bundle = self._bundle
        if purge or bundle._children is None:
            bundle._children = {}
        if self._hierarchical_fitter is None:   # default: repeated flat fits with this clustering's fitter
            from ._fit import HierarchicalFitterRepeat

            self._hierarchical_fitter = HierarchicalFitterRepeat(self._fitter)
        self._hierarchical_fitter.fit(bundle, **fitter_kwargs)

    def trim_shrinking_leafs(self, bundle=None) -> None:
        """Cuts the tails of a hierarchy in which a cluster only loses members (reference ``cluster.pyx:435-470``): a
        bundle stays when its parent split into several clusters (it is one of them), when its own labels hold more
        than one cluster, or when something below it is kept for one of these reasons; every other bundle loses its
        labels and children."""
        def visit(node, parent_splits: bool) -> bool:
            splits = keeps_below = False
            if node._children:
                splits = len(node._labels.to_set() - {0}) > 1
                kept = [visit(child, splits) for child in node._children.values()]   # every child is visited
                keeps_below = any(kept)
            keep = parent_splits or splits or keeps_below
            if not keep:
                node._labels = None
                node._children = {}
            return keep

        visit(self._bundle if bundle is None else bundle, True)

    def trim_trivial_leafs(self, bundle=None) -> None:
        """Bundles whose labels are all noise lose their labels and children (reference ``cluster.pyx:472-500``)."""
        pending = [self._bundle if bundle is None else bundle]
        while pending:
            node = pending.pop()
            if node._labels is None:
                continue
            if node._labels.to_set() == {0}:
                node._labels = None
                node._children = {}
            else:
                pending.extend((node._children or {}).values())

    def to_dtrajs(self, bundle=None) -> list:
        """The labels cut into the parts the input data came in (``meta["edges"]``), e.g. one label sequence per
        trajectory (reference ``cluster.pyx:1166-1188``)."""
        bundle = self._bundle if bundle is None else bundle
        labels = bundle.labels
        if labels is None:
            return []
        edges = bundle._input_data.meta.get("edges") if bundle._input_data is not None else None
        if edges is None:
            return [labels]
        parts = np.split(labels, np.cumsum(edges))
        return parts[:-1] if len(parts[-1]) == 0 else parts

    def reel(self, depth=None, bundle=None) -> None:
        """Wrap up the label assignments of lower hierarchy levels (reference ``cluster.pyx:502-558``):
        depth first, every child's non-zero labels are shifted by the parent's current maximum and
        written over the parent's labels at the child's ``parent_indices``; parameters move along."""
        if bundle is None:
            bundle = self._bundle
        if depth is not None:
            assert depth > 0

        def _reel(parent, depth):
            if not parent._children:
                return
            if depth is not None:
                depth -= 1
            parent._labels.meta["origin"] = "reel"
            parent_labels = parent._labels.labels
            # the reference asks for the parent's largest label before every child (an O(n) pass per child); kept exact
            # with a histogram that follows the assignments: a child's points are ALL points of its label, so the label
            # leaves the parent when the child's labels are written over it
            counts = np.bincount(parent_labels[parent_labels >= 0]).astype(np.int64) if parent_labels.size else np.zeros(1, np.int64)
            negative = bool(parent_labels.size and parent_labels.min() < 0)
            top = counts.shape[0] - 1
            for label, child in parent._children.items():
                if child._labels is None:
                    continue
                if depth is None or depth > 0:
                    _reel(child, depth)
                if negative:
                    n_clusters = int(parent_labels.max())
                else:
                    while top > 0 and counts[top] == 0:
                        top -= 1
                    n_clusters = top
                child_labels = child._labels.labels
                new = np.where(child_labels == 0, 0, child_labels + n_clusters)
                if not negative:
                    old_vals = parent_labels[child.parent_indices]
                    counts -= np.bincount(old_vals, minlength=counts.shape[0])
                    add = np.bincount(new)
                    if add.shape[0] > counts.shape[0]:
                        counts = np.concatenate([counts, np.zeros(add.shape[0] - counts.shape[0], np.int64)])
                    counts[:add.shape[0]] += add
                    top = max(top, add.shape[0] - 1)
                parent_labels[child.parent_indices] = new
                params_parent = parent._labels.meta.setdefault("params", {})
                params_parent.pop(label, None)
                for old_label, p in child._labels.meta.get("params", {}).items():
                    params_parent[old_label + n_clusters] = p

        _reel(bundle, depth)

    def predict(self, other, radius_cutoff: float, cnn_cutoff: int, clusters=None, cnn_offset=None, info: bool = True,
                record: bool = True, record_time: bool = True, v: bool = True, purge: bool = False, bundle=None) -> None:
        """Cluster labels for the points of ``other`` (a Bundle or a Clustering) from this clustering's labels:
        same arguments and effects as the reference's ``Clustering.predict`` (``cluster.pyx:560-653``)."""
        if bundle is None:
            bundle = self._bundle
        if isinstance(other, Clustering):
            other = other._bundle
        if cnn_offset is None:
            cnn_offset = 0
        cnn_cutoff -= cnn_offset
        assert cnn_cutoff >= 0
        if (other._labels is None) or purge or (not other._labels.meta.get("frozen", False)):
            other._labels = Labels(np.zeros(other._input_data.n_points, order="C", dtype=np.int64))
        cluster_params = self._predictor.make_parameters(radius_cutoff, cnn_cutoff, 0)
        if clusters is None:
            clusters = bundle._labels.to_set() - {0}
        other._labels.consider_set = set(clusters)
        t0 = time.perf_counter()
        self._predictor.predict(bundle._input_data, other._input_data, bundle._labels, other._labels, cluster_params)
        execution_time = time.perf_counter() - t0 if record_time else None
        if info:
            params = {k: (radius_cutoff, cnn_cutoff) for k in clusters if k != 0}
            old_params = other._labels.meta.get("params", {})
            old_params.update(params)
            other._labels.meta.update({"params": old_params, "reference": weakref.proxy(bundle), "origin": "predict"})
        other._labels.meta["frozen"] = True
        del execution_time

    def fit_sweep(self, radius_cutoffs, cnn_cutoffs, member_cutoff=None, max_clusters=None, return_labels=True):
        """Every cell of a ``radius_cutoffs x cnn_cutoffs`` scan, labels identical to one ``fit`` per cell
        (see ``sweep.fit_sweep``); the records are appended to ``summary``."""
        from . import sweep

        g = self._fitter._neighbours_getter
        labels, summary = sweep.fit_sweep(self._bundle._input_data.device_or_host_array, radius_cutoffs, cnn_cutoffs,
                                          member_cutoff, max_clusters, metric=g.metric, ctx=self._fitter.ctx,
                                          return_labels=return_labels)
        for rec in summary:
            self._bundle._summary.append(rec)
        return labels, summary

    @property
    def fitter(self):
        return self._fitter

    @fitter.setter
    def fitter(self, value):
        if value is not None and not isinstance(value, Fitter):
            raise TypeError(f"Can't use object of type {type(value).__name__} as fitter.")
        self._fitter = value

    def fit(self, radius_cutoff: float, cnn_cutoff: int, member_cutoff: int = None, max_clusters: int = None,
            cnn_offset: int = None, sort_by_size: bool = True, info: bool = True, record: bool = True,
            record_time: bool = True, v: bool = True, purge: bool = False, bundle=None) -> None:
        """Same arguments and effects as the reference's ``Clustering.fit`` (``cluster.pyx:256-392``)."""
        if bundle is None:
            bundle = self._bundle
        if cnn_offset is None:
            cnn_offset = 0
        cnn_cutoff -= cnn_offset
        assert cnn_cutoff >= 0
        n = bundle._input_data.n_points
        # a fresh fit on coordinates with size sorting can keep the labels on the device
        # between FitterExtBFS.fit and Labels.sort_by_size: one C-ABI call, one D2H copy
        may_fuse = sort_by_size and isinstance(self._fitter, FitterB200Fused)
        if (bundle._labels is None) or purge or (not bundle._labels.meta.get("frozen", False)):
            bundle._labels = Labels(_label_buffer(n, pinned=may_fuse))
            old_label_set = set()
            current_start = 1
            fresh = True
        else:
            old_label_set = bundle._labels.to_set()
            current_start = max(old_label_set) + 1
            fresh = False
        cluster_params = self._fitter.make_parameters(radius_cutoff, cnn_cutoff, current_start)

        fused = fresh and may_fuse
        if fused and hasattr(bundle._input_data, "pin"):
            bundle._input_data.pin()
        t0 = time.perf_counter()
        if fused:
            n_found = self._fitter.fit_sorted(bundle._input_data, bundle._labels, cluster_params,
                                              member_cutoff, max_clusters)
        else:
            self._fitter.fit(bundle._input_data, bundle._labels, cluster_params)
        execution_time = time.perf_counter() - t0 if record_time else None

        if fused:  # the size sort on the device already knows what the record and the parameter table need
            n_kept, n_largest, n_noise = self._fitter.ctx.label_summary()
            del n_found
        if info:
            if fused:
                params = {k: (radius_cutoff, cnn_cutoff) for k in range(1, n_kept + 1)}
            else:
                new_label_set = bundle._labels.to_set()
                params = {k: (radius_cutoff, cnn_cutoff) for k in new_label_set - old_label_set if k != 0}
            old_params = bundle._labels.meta.get("params", {})
            old_params.update(params)
            bundle._labels.meta.update({"params": old_params, "reference": weakref.proxy(bundle), "origin": "fit"})
        if sort_by_size and not fused:
            bundle._labels.sort_by_size(member_cutoff, max_clusters, ctx=getattr(self._fitter, "ctx", None))
        if record:
            if fused:
                rec = Record(n, radius_cutoff, cnn_cutoff, member_cutoff, max_clusters, n_kept, n_largest / n,
                             n_noise / n, execution_time)
            else:
                lab = bundle._labels.labels
                counts = np.bincount(lab) if n else np.zeros(1, dtype=np.int64)
                n_noise = int(counts[0]) if counts.size else 0
                rest = counts[1:]
                rec = Record(n, radius_cutoff, cnn_cutoff, member_cutoff, max_clusters,
                             int(np.count_nonzero(rest)), (int(rest.max()) if rest.size else 0) / n,
                             n_noise / n, execution_time)
            if v:
                print(rec)
            bundle._summary.append(rec)
