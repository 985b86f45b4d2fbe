"""Access to the UNMODIFIED reference built into ``oracle/_ref`` (TEST INFRASTRUCTURE).

``available()`` is False when ``oracle/_ref`` has not been built (e.g. a fresh checkout without
``/root/reference``); callers then fall back to the pinned C port.
"""
import contextlib
import io
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_cluster = None


def available() -> bool:
    from . import build_ref

    return build_ref.built()


def module():
    """The reference's ``cnnclustering.cluster`` module."""
    global _cluster
    if _cluster is None:
        import importlib
        import marshal
        import types

        pkg_dir = os.path.join(REF_DIR, "cnnclustering")

        def load_pure(name, fullname, is_pkg=False):
            mod = types.ModuleType(fullname)
            mod.__file__ = os.path.join(pkg_dir, name + ".code")
            if is_pkg:
                mod.__path__ = [pkg_dir]
                mod.__package__ = fullname
            else:
                mod.__package__ = fullname.rpartition(".")[0]
            sys.modules[fullname] = mod
            with open(mod.__file__, "rb") as fp:
                exec(marshal.load(fp), mod.__dict__)
            return mod

        with contextlib.redirect_stdout(io.StringIO()):  # "Optional dependency ... matplotlib"
            # the pure-Python parts of the reference travel as marshalled code objects (build_ref.py);
            # the extension modules are found through the package's __path__ as usual
            pkg = load_pure("__init__", "cnnclustering", is_pkg=True)
            pkg.report = load_pure("report", "cnnclustering.report")
            pkg.recipes = load_pure("recipes", "cnnclustering.recipes")
            _cluster = importlib.import_module("cnnclustering.cluster")
    return _cluster


def fit_points(x, r, c, member_cutoff=None, max_clusters=None, sort_by_size=True, recipe="coordinates",
               return_time=False, **recipe_kwargs):
    """``Clustering(x).fit(r, c, ...)`` with the default brute-force recipe."""
    cl = module().Clustering(np.asarray(x, dtype=np.float64), recipe=recipe, **recipe_kwargs)
    cl.fit(r, c, member_cutoff=member_cutoff, max_clusters=max_clusters, sort_by_size=sort_by_size,
           v=False, record=True, record_time=True)
    labels = np.array(cl.labels, dtype=np.int64)
    if return_time:
        return labels, cl.summary[-1].execution_time
    return labels


def prepare_sorted_neighbourhoods(row_ptr, col_idx):
    """Padded int64 matrix + counts as ``recipes.prepare_neighbourhoods`` (recipes.py:150-179)
    would build them, produced vectorised (the reference's version is a per-row Python loop)."""
    from . import port

    return port.csr_to_padded(row_ptr, col_idx)


def fit_neighbourhoods(row_ptr, col_idx, c, member_cutoff=None, max_clusters=None, sort_by_size=True,
                       r=0.0, return_time=False, fit_only=False):
    """``sorted_neighbourhoods`` recipe (recipes.py:349-357) on precomputed sorted lists.

    ``return_time`` returns ``Record.execution_time`` (time of ``fitter.fit`` only,
    cluster.pyx:332-337)."""
    mod = module()
    from cnnclustering import _types  # after module() put _ref on sys.path

    padded, counts = prepare_sorted_neighbourhoods(row_ptr, col_idx)
    data = _types.InputDataExtNeighbourhoodsMemoryview(padded, counts)
    cl = mod.Clustering(data, recipe="sorted_neighbourhoods")
    t0 = time.perf_counter()
    cl.fit(r, c, member_cutoff=member_cutoff, max_clusters=max_clusters, sort_by_size=sort_by_size,
           v=False, record=True, record_time=True, info=not fit_only)
    wall = time.perf_counter() - t0
    labels = np.array(cl.labels, dtype=np.int64)
    if return_time:
        return labels, cl.summary[-1].execution_time, wall
    return labels
