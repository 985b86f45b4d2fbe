"""Recipes: string keys -> component types, as in the reference's ``recipes.py`` (registry
``recipes.py:316-445``, ``Builder`` :182-313), for the component set of the B200 fit path."""
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
}

COMPONENT_ALT_KW_MAP = {
    "input": "input_data", "data": "input_data", "n": "neighbours", "na": "neighbours",
    "nb": "neighbour_neighbours", "getter": "neighbours_getter", "ogetter": "neighbours_getter_other",
    "ngetter": "neighbours_getter", "ongetter": "neighbours_getter_other", "dgetter": "distance_getter",
    "checker": "similarity_checker", "q": "queue", "prep": "preparation_hook",
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
        "fitter": "b200_staged",
        "fitter.ngetter": "radius_b200",
        "fitter.checker": "commonnn_b200",
        "fitter.ngetter.dgetter": "metric",
        "fitter.ngetter.dgetter.metric": "precomputed",
    },
    "neighbourhoods": {   # reference recipes.py:340-348: neighbour lists in any order
        "input_data": "neighbourhoods_b200",
        "preparation_hook": "neighbourhoods",
        "fitter": "b200_staged",
        "fitter.ngetter": ("lookup", (), {"is_sorted": False}),
        "fitter.checker": "commonnn_b200",
        "predictor": "b200",
        "predictor.ngetter": ("lookup", (), {"is_sorted": False}),
        "predictor.checker": "commonnn_b200",
    },
    "sorted_neighbourhoods": {
        "input_data": "neighbourhoods_b200",
        "preparation_hook": "neighbourhoods",
        "fitter": "b200_staged",
        "fitter.ngetter": ("lookup", (), {"is_sorted": True}),
        "fitter.checker": "commonnn_b200",
        "predictor": "b200",
        "predictor.ngetter": ("lookup", (), {"is_sorted": True}),
        "predictor.checker": "commonnn_b200",
    },
    "neighbourhoods_csr": {
        "input_data": "neighbourhoods_csr",
        "preparation_hook": "pass",
        "fitter": "b200_staged",
        "fitter.ngetter": ("lookup", (), {"is_sorted": True}),
        "fitter.checker": "commonnn_b200",
    },
}


def get_registered_recipe(key):
    return dict(REGISTERED_RECIPES[key.lower()])


class Builder:
    """Assembles components from a recipe.  Reference: ``recipes.Builder``, ``recipes.py:182-313``
    (same key grammar: dotted paths, ``__`` as separator in keyword arguments, short aliases)."""

    default_recipe = "coordinates"

    def __init__(self, recipe=None, **kwargs):
        if recipe is None:
            recipe = self.default_recipe
        if isinstance(recipe, str):
            recipe = get_registered_recipe(recipe)
        recipe = dict(recipe)
        recipe.update(kwargs)
        recipe = {k.replace("__", "."): v for k, v in recipe.items()}
        self.recipe = {
            ".".join(COMPONENT_ALT_KW_MAP.get(kw, kw) for kw in k.split(".")): v for k, v in recipe.items()
        }

    def make_component(self, component_kw, alternative=None, prev_kw=None):
        prev_kw = prev_kw or []
        full_kw = ".".join([*prev_kw, component_kw])
        if alternative is not None:
            alternative = full_kw.rsplit(".", 1)[0] + f".{alternative}"
        if full_kw in self.recipe:
            details = self.recipe[full_kw]
        elif alternative is not None and alternative in self.recipe:
            details = self.recipe[alternative]
        else:
            return object
        if details is None:
            return None
        args, kwargs = (), {}
        type_kw = COMPONENT_KW_TYPE_ALIAS_MAP.get(component_kw, component_kw)
        if isinstance(details, str):
            ctype = COMPONENT_NAME_TYPE_MAP[type_kw][details]
        elif isinstance(details, tuple):
            ctype, args, kwargs = details
            kwargs = dict(kwargs)
            if isinstance(ctype, str):
                ctype = COMPONENT_NAME_TYPE_MAP[type_kw][ctype]
        else:
            ctype = details
        if hasattr(ctype, "get_builder_kwargs"):
            for kw, alt in ctype.get_builder_kwargs():
                comp = self.make_component(kw, alt, prev_kw=full_kw.split("."))
                if comp is not object:
                    kwargs[kw] = comp
        return ctype(*args, **kwargs)

    def make_input_data(self, data, preparation_hook=None, details=None):
        if preparation_hook is None:
            preparation_hook = self.recipe.get("preparation_hook", prepare_pass)
        if isinstance(preparation_hook, str):
            preparation_hook = COMPONENT_NAME_TYPE_MAP["preparation_hook"][preparation_hook]
        if details is None:
            try:
                details = self.recipe["input_data"]
            except KeyError:
                raise LookupError("Input data type details not found")
        args, kwargs = (), {}
        if isinstance(details, str):
            ctype = COMPONENT_NAME_TYPE_MAP["input_data"][details]
        elif isinstance(details, tuple):
            ctype, args, kwargs = details
            kwargs = dict(kwargs)
            if isinstance(ctype, str):
                ctype = COMPONENT_NAME_TYPE_MAP["input_data"][ctype]
        else:
            ctype = details
        data_args, data_kwargs = preparation_hook(data)
        if ctype is _types.InputDataNeighbourhoodsCSR and len(data_args) == 1:
            data_args = tuple(data_args[0])
        data_kwargs.setdefault("meta", {})
        data_kwargs["meta"].update(kwargs.get("meta", {}))
        kwargs.update(data_kwargs)
        return ctype(*data_args, *args, **kwargs)
