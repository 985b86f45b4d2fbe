"""Recipes: string keys -> component types, as in the reference's ``recipes.py`` (registry
``recipes.py:316-445``), and a builder that assembles components from them (a recipe tree; the reference's builder,
``recipes.py:182-313``, defines the key grammar it accepts) -- for the component set of the B200 fit path."""
import numpy as np

from . import _fit, _types


def prepare_pass(data):
    """Reference: ``recipes.py:37-49``."""
    return (data,), {}


def prepare_points_from_parts(data):
    """Points, or a sequence of parts of points, -> one C-contiguous f64 array ``[n, d]`` plus the
    part lengths in ``meta["edges"]``.  Reference: ``recipes.py:52-147``."""
    if hasattr(data, "is_cuda"):
        return (data,), {"meta": {"edges": [int(data.shape[0])]}}
    try:
        arr = np.asarray(data, dtype=np.float64)
        ragged = arr.dtype == object
    except ValueError:
        ragged = True
    if ragged or arr.ndim == 3:
        parts = [np.atleast_2d(np.asarray(p, dtype=np.float64)) for p in data]
        parts = [p for p in parts if p.size]
        edges = [p.shape[0] for p in parts]
        arr = np.vstack(parts) if parts else np.zeros((0, 0))
    else:
        if arr.ndim == 1:
            arr = arr.reshape(1, -1)
        if arr.ndim != 2:
            raise ValueError("cannot interpret data as points [n, d] or parts of points")
        edges = [arr.shape[0]]
    return (np.ascontiguousarray(arr, dtype=np.float64),), {"meta": {"edges": edges}}


def prepare_neighbourhoods(data):
    """Sequence of neighbour-index sequences -> zero-padded int64 matrix + counts, vectorised.
    Reference: ``recipes.py:150-179`` (a Python loop with one ``np.pad`` per row)."""
    counts = np.fromiter((len(s) for s in data), dtype=np.int64, count=len(data))
    pad_to = int(counts.max()) if counts.size else 0
    padded = np.zeros((len(data), max(pad_to, 1)), dtype=np.int64)
    if counts.sum():
        flat = np.concatenate([np.asarray(s, dtype=np.int64) for s in data])
        rows = np.repeat(np.arange(len(data)), counts)
        starts = np.cumsum(counts) - counts
        cols = np.arange(flat.size) - np.repeat(starts, counts)
        padded[rows, cols] = flat
    return (padded, counts), {"meta": {}}


COMPONENT_NAME_TYPE_MAP = {
    "input_data": {
        "components_b200": _types.InputDataComponents,
        "components_mview": _types.InputDataComponents,      # reference key, same layout
        "neighbourhoods_b200": _types.InputDataNeighbourhoods,
        "neighbourhoods_mview": _types.InputDataNeighbourhoods,
        "neighbourhoods_csr": _types.InputDataNeighbourhoodsCSR,
        "distances_linear": _types.InputDataDistancesLinear,   # condensed distances (InputDataExtDistancesLinearMemoryview)
    },
    "preparation_hook": {
        "pass": prepare_pass,
        "points_from_parts": prepare_points_from_parts,
        "neighbourhoods": prepare_neighbourhoods,
    },
    "neighbours_getter": {
        "radius_b200": _types.NeighboursGetterRadius,
        "grid": _types.NeighboursGetterGrid,
        "tiled": _types.NeighboursGetterTiled,
        "tensor": _types.NeighboursGetterTensor,
        "brute_force": _types.NeighboursGetterRadius,        # reference key: same sets, bulk search
        "lookup": _types.NeighboursGetterLookup,
    },
    "distance_getter": {"metric": _types.DistanceGetterMetric},
    "metric": {
        "precomputed": _types.MetricPrecomputed,
        "euclidean": _types.MetricEuclidean,
        "euclidean_r": _types.MetricEuclideanReduced,
        "euclidean_reduced": _types.MetricEuclideanReduced,
        "euclidean_periodic_r": _types.MetricEuclideanPeriodicReduced,
        "euclidean_periodic_reduced": _types.MetricEuclideanPeriodicReduced,
    },
    "similarity_checker": {
        "commonnn_b200": _types.SimilarityCheckerCommonNN,
        "contains": _types.SimilarityCheckerCommonNN,        # reference keys: one device kernel
        "switch": _types.SimilarityCheckerCommonNN,
        "screen": _types.SimilarityCheckerCommonNN,
    },
    "fitter": {"b200": _fit.FitterB200Fused, "b200_staged": _fit.FitterB200, "bfs": _fit.FitterB200Fused},
    "predictor": {"b200": _fit.PredictorB200, "firstmatch": _fit.PredictorB200},
    "hierarchical_fitter": {"repeat": _fit.HierarchicalFitterRepeat},
}

COMPONENT_ALT_KW_MAP = {
    "input": "input_data", "data": "input_data", "n": "neighbours", "na": "neighbours",
    "nb": "neighbour_neighbours", "getter": "neighbours_getter", "ogetter": "neighbours_getter_other",
    "ngetter": "neighbours_getter", "ongetter": "neighbours_getter_other", "dgetter": "distance_getter",
    "checker": "similarity_checker", "q": "queue", "prep": "preparation_hook", "hfitter": "hierarchical_fitter",
}

# component keywords that share the type table of another keyword (reference ``recipes.py``, COMPONENT_KW_TYPE_ALIAS_MAP)
COMPONENT_KW_TYPE_ALIAS_MAP = {"neighbours_getter_other": "neighbours_getter", "neighbour_neighbours": "neighbours"}

REGISTERED_RECIPES = {
    "none": {},
    "coordinates": {
        "input_data": "components_b200",
        "preparation_hook": "points_from_parts",
        "fitter": "b200",
        "fitter.ngetter": "radius_b200",
        "fitter.checker": "commonnn_b200",
        "fitter.ngetter.dgetter": "metric",
        "fitter.ngetter.dgetter.metric": "euclidean_r",
        "predictor": "b200",
        "predictor.ngetter": "radius_b200",
        "predictor.checker": "commonnn_b200",
        "predictor.ngetter.dgetter": "metric",
        "predictor.ngetter.dgetter.metric": "euclidean_r",
    },
    "distances": {   # reference recipes.py:329-339: the input is an N x N distance matrix
        "input_data": "components_b200",
        "preparation_hook": "points_from_parts",
        "fitter": "b200",
        "fitter.ngetter": "radius_b200",
        "fitter.checker": "commonnn_b200",
        "fitter.ngetter.dgetter": "metric",
        "fitter.ngetter.dgetter.metric": "precomputed",
    },
    "neighbourhoods": {   # reference recipes.py:340-348: neighbour lists in any order
        "input_data": "neighbourhoods_b200",
        "preparation_hook": "neighbourhoods",
        "fitter": "b200",
        "fitter.ngetter": ("lookup", (), {"is_sorted": False}),
        "fitter.checker": "commonnn_b200",
        "predictor": "b200",
        "predictor.ngetter": ("lookup", (), {"is_sorted": False}),
        "predictor.checker": "commonnn_b200",
    },
    "sorted_neighbourhoods": {
        "input_data": "neighbourhoods_b200",
        "preparation_hook": "neighbourhoods",
        "fitter": "b200",
        "fitter.ngetter": ("lookup", (), {"is_sorted": True}),
        "fitter.checker": "commonnn_b200",
        "predictor": "b200",
        "predictor.ngetter": ("lookup", (), {"is_sorted": True}),
        "predictor.checker": "commonnn_b200",
    },
    "neighbourhoods_csr": {
        "input_data": "neighbourhoods_csr",
        "preparation_hook": "pass",
        "fitter": "b200",
        "fitter.ngetter": ("lookup", (), {"is_sorted": True}),
        "fitter.checker": "commonnn_b200",
    },
}


def get_registered_recipe(key):
    return dict(REGISTERED_RECIPES[key.lower()])


class _Node:
    """One component of a recipe: how to make it (``details``) and the recipe entries below it."""
    __slots__ = ("details", "present", "below")

    def __init__(self):
        self.details = None
        self.present = False
        self.below = {}


class Builder:
    """Assembles components from a recipe (the role of the reference's ``recipes.Builder``, ``recipes.py:182-313``,
    with the same key grammar: dotted paths such as ``fitter.ngetter.dgetter.metric``, ``__`` instead of the dot in
    keyword arguments, the short aliases of ``COMPONENT_ALT_KW_MAP``; an entry is a registered name, a type, a
    ``(type-or-name, args, kwargs)`` triple or ``None``).

    The flat recipe is turned into a TREE once -- one node per component, children keyed by the constructor keyword
    they fill -- and a component is made by walking its subtree: each type says through ``get_builder_kwargs()`` which
    keywords it takes and which sibling keyword may stand in (``neighbours_getter_other`` falls back to
    ``neighbours_getter``)."""

    default_recipe = "coordinates"

    def __init__(self, recipe=None, **kwargs):
        if recipe is None:
            recipe = self.default_recipe
        entries = dict(get_registered_recipe(recipe) if isinstance(recipe, str) else recipe)
        entries.update(kwargs)
        self._root = _Node()
        for key, details in entries.items():
            self.set(key, details)

    @staticmethod
    def _path(key):
        return [COMPONENT_ALT_KW_MAP.get(part, part) for part in key.replace("__", ".").split(".")]

    def set(self, key, details):
        """Adds or replaces one recipe entry (``key`` in any spelling the grammar allows)."""
        node = self._root
        for part in self._path(key):
            node = node.below.setdefault(part, _Node())
        node.details, node.present = details, True

    @property
    def recipe(self):
        """The normalised flat form (full keywords, dots) -- what the reference's builder keeps as ``recipe``."""
        flat = {}

        def walk(node, prefix):
            for kw, sub in node.below.items():
                path = prefix + [kw]
                if sub.present:
                    flat[".".join(path)] = sub.details
                walk(sub, path)

        walk(self._root, [])
        return flat

    @staticmethod
    def _resolve(table_kw, details):
        """-> (callable, args, kwargs) for an entry of the type table ``table_kw``."""
        table = COMPONENT_NAME_TYPE_MAP.get(COMPONENT_KW_TYPE_ALIAS_MAP.get(table_kw, table_kw), {})
        args, kwargs = (), {}
        if isinstance(details, tuple):
            details, args, kwargs = details
            kwargs = dict(kwargs)
        if isinstance(details, str):
            details = table[details]
        return details, tuple(args), kwargs

    def _construct(self, kw, node, scope):
        if node.details is None:
            return None
        ctype, args, kwargs = self._resolve(kw, node.details)
        for sub_kw, stand_in in (ctype.get_builder_kwargs() if hasattr(ctype, "get_builder_kwargs") else ()):
            sub = node.below.get(sub_kw)
            if (sub is None or not sub.present) and stand_in is not None:
                sub = node.below.get(stand_in)
            if sub is not None and sub.present:
                kwargs[sub_kw] = self._construct(sub_kw, sub, node)
        return ctype(*args, **kwargs)

    def make_component(self, component_kw, alternative=None, prev_kw=None):
        """The component registered under ``component_kw`` (below ``prev_kw``, a list of keywords, when given), or the
        sentinel ``object`` when the recipe has no entry for it -- ``None`` is a legal component."""
        node = self._root
        for part in prev_kw or ():
            node = node.below.get(part)
            if node is None:
                return object
        found = node.below.get(COMPONENT_ALT_KW_MAP.get(component_kw, component_kw))
        if (found is None or not found.present) and alternative is not None:
            found = node.below.get(alternative)
        if found is None or not found.present:
            return object
        return self._construct(component_kw, found, node)

    def make_input_data(self, data, preparation_hook=None, details=None):
        """Runs the preparation hook on ``data`` and hands the result to the input-data type of the recipe."""
        hook = preparation_hook
        if hook is None:
            hook_node = self._root.below.get("preparation_hook")
            hook = hook_node.details if hook_node is not None and hook_node.present else prepare_pass
        if isinstance(hook, str):
            hook = COMPONENT_NAME_TYPE_MAP["preparation_hook"][hook]
        if details is None:
            data_node = self._root.below.get("input_data")
            if data_node is None or not data_node.present:
                raise LookupError("Input data type details not found")
            details = data_node.details
        ctype, args, kwargs = self._resolve("input_data", details)
        data_args, data_kwargs = hook(data)
        if ctype is _types.InputDataNeighbourhoodsCSR and len(data_args) == 1:
            data_args = tuple(data_args[0])          # (row_ptr, col_idx) handed over as one pair
        meta = dict(data_kwargs.get("meta") or {})
        meta.update(kwargs.get("meta", {}))
        kwargs.update(data_kwargs)
        kwargs["meta"] = meta
        return ctype(*data_args, *args, **kwargs)
