"""Golden cases: seeded inputs + fit arguments.  ``make_golden.py`` runs the reference on them,
the tests run the oracle port and the CUDA path on the same inputs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from commonnnclustering_b200 import datasets  # noqa: E402


def moons(n=1500):
    from sklearn.datasets import make_moons
    from sklearn.preprocessing import StandardScaler

    x, _ = make_moons(n_samples=n, noise=0.05, random_state=8)
    return np.ascontiguousarray(StandardScaler().fit_transform(x))


def lattice(n=600, d=2, seed=3):
    """Points on a 1/8 lattice with duplicates: many distances exactly ON the cutoff."""
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(rng.integers(0, 24, (n, d)) / 8.0)


def gauss(n, d, seed):
    return np.ascontiguousarray(np.random.default_rng(seed).standard_normal((n, d)))


# records sets of the reference benchmark (docsrc/benchmark/helper_base.py:135-173, scans.ipynb)
def rec_no_structure():
    return datasets.uniform_2d(5000, 2021)


def rec_blobs():
    from sklearn.datasets import make_blobs
    from sklearn.preprocessing import StandardScaler

    x, _ = make_blobs(n_samples=5000, n_features=2, random_state=8)
    return np.ascontiguousarray(StandardScaler().fit_transform(x))


def rec_varied():
    from sklearn.datasets import make_blobs
    from sklearn.preprocessing import StandardScaler

    x, _ = make_blobs(n_samples=5000, n_features=2, random_state=170, cluster_std=[1.0, 2.5, 0.5])
    return np.ascontiguousarray(StandardScaler().fit_transform(x))


def rec_aniso():
    from sklearn.datasets import make_blobs
    from sklearn.preprocessing import StandardScaler

    x, _ = make_blobs(n_samples=5000, n_features=2, random_state=170)
    x = StandardScaler().fit_transform(x)  # scans.ipynb: scaled first, then sheared
    return np.ascontiguousarray(np.dot(x, [[0.6, -0.6], [-0.4, 0.8]]))


def rec_circles():
    from sklearn.datasets import make_circles
    from sklearn.preprocessing import StandardScaler

    x, _ = make_circles(n_samples=5000, factor=.5, noise=.05, random_state=8)
    return np.ascontiguousarray(StandardScaler().fit_transform(x))


RECORD_SETS = {"no_structure": rec_no_structure, "blobs": rec_blobs, "varied": rec_varied,
               "aniso": rec_aniso, "circles": rec_circles}

CASES = {
    "readme": dict(data=lambda: datasets.README_POINTS, fit=dict(r=1.5, c=1)),
    "blobs1500_r025_c20_m20": dict(data=lambda: datasets.blobs_2d(1500), fit=dict(r=0.25, c=20, member_cutoff=20)),
    "blobs1500_r010_c20": dict(data=lambda: datasets.blobs_2d(1500), fit=dict(r=0.10, c=20)),
    "blobs1500_r005_c5": dict(data=lambda: datasets.blobs_2d(1500), fit=dict(r=0.05, c=5)),
    "blobs1500_r019_c37": dict(data=lambda: datasets.blobs_2d(1500), fit=dict(r=0.19, c=37)),
    "blobs1500_c0": dict(data=lambda: datasets.blobs_2d(1500), fit=dict(r=0.08, c=0)),
    "blobs1500_unsorted": dict(data=lambda: datasets.blobs_2d(1500), fit=dict(r=0.1, c=10, sort_by_size=False)),
    "blobs1500_max3": dict(data=lambda: datasets.blobs_2d(1500), fit=dict(r=0.08, c=5, member_cutoff=1, max_clusters=3)),
    "blobs1500_euclid": dict(data=lambda: datasets.blobs_2d(1500), fit=dict(r=0.1, c=10, metric="euclidean")),
    "moons1500": dict(data=moons, fit=dict(r=0.2, c=5)),
    "md5000_r04_c5": dict(data=lambda: datasets.md_like_3d(5000), fit=dict(r=0.4, c=5)),
    "md5000_r025_c10_m10": dict(data=lambda: datasets.md_like_3d(5000), fit=dict(r=0.25, c=10, member_cutoff=10)),
    "lattice2d": dict(data=lambda: lattice(600, 2, 3), fit=dict(r=0.25, c=3)),
    "lattice3d": dict(data=lambda: lattice(800, 3, 4), fit=dict(r=0.375, c=4, member_cutoff=5)),
    "lattice1d": dict(data=lambda: lattice(200, 1, 5), fit=dict(r=0.125, c=2)),
    "gauss5d": dict(data=lambda: gauss(800, 5, 1), fit=dict(r=1.0, c=5)),
    "gauss8d_euclid": dict(data=lambda: gauss(700, 8, 2), fit=dict(r=2.0, c=8, metric="euclidean")),
    "feat64d_2000": dict(data=lambda: datasets.features_64d(2000), fit=dict(r=0.9, c=5, member_cutoff=5)),
    "c1_full": dict(data=lambda: datasets.blobs_2d(10_000), fit=dict(r=0.3, c=10)),
    "c1_r019_c37": dict(data=lambda: datasets.blobs_2d(10_000), fit=dict(r=0.19, c=37)),
}


# ---- the reference's other input recipes: "distances" (N x N matrix, MetricExtPrecomputed) and
# "neighbourhoods" (neighbour lists in ANY order) ------------------------------------------------------
def cityblock_matrix(x):
    """Manhattan distances: not a metric the coordinate path offers, so only the matrix route can produce these labels."""
    x = np.asarray(x, dtype=np.float64)
    return np.ascontiguousarray(np.abs(x[:, None, :] - x[None, :, :]).sum(axis=2))


def euclid_matrix_with_holes(x, seed=11):
    """Euclidean distances with NaN / inf entries (never neighbours), kept symmetric: the device path settles each
    candidate edge once from its lower end and therefore needs ``j in N(i) <=> i in N(j)``."""
    x = np.asarray(x, dtype=np.float64)
    dm = np.sqrt(((x[:, None, :] - x[None, :, :]) ** 2).sum(axis=2))
    rng = np.random.default_rng(seed)
    n = dm.shape[0]
    i, j = rng.integers(0, n, 4000), rng.integers(0, n, 4000)
    dm[i[:2500], j[:2500]] = np.nan
    dm[j[:2500], i[:2500]] = np.nan
    dm[i[2500:], j[2500:]] = np.inf
    dm[j[2500:], i[2500:]] = np.inf
    return np.ascontiguousarray(dm)


def shuffled_neighbour_lists(x, r, seed=5):
    """Neighbour lists (self included) of a point set, every list in random order."""
    x = np.asarray(x, dtype=np.float64)
    d2 = ((x[:, None, :] - x[None, :, :]) ** 2).sum(axis=2)
    rng = np.random.default_rng(seed)
    out = []
    for row in d2 <= r * r:
        nb = np.nonzero(row)[0]
        rng.shuffle(nb)
        out.append(nb.astype(np.int64))
    return out


RECIPE_CASES = {
    "dist_cityblock_blobs800": dict(recipe="distances", data=lambda: cityblock_matrix(datasets.blobs_2d(800)),
                                    fit=dict(r=0.15, c=8)),
    "dist_holes_md1000": dict(recipe="distances", data=lambda: euclid_matrix_with_holes(datasets.md_like_3d(1000)),
                              fit=dict(r=0.6, c=5, member_cutoff=5)),
    "dist_odd_n1031": dict(recipe="distances", data=lambda: cityblock_matrix(gauss(1031, 3, 9)),
                           fit=dict(r=0.9, c=6)),
    "nbh_unsorted_md2000": dict(recipe="neighbourhoods", data=lambda: shuffled_neighbour_lists(datasets.md_like_3d(2000), 0.45),
                                fit=dict(r=0.45, c=6, member_cutoff=5)),
    "nbh_unsorted_blobs1200": dict(recipe="neighbourhoods", data=lambda: shuffled_neighbour_lists(datasets.blobs_2d(1200), 0.12),
                                   fit=dict(r=0.12, c=10)),
}
