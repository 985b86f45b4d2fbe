"""Synthetic workloads of the BASELINE.json configurations.

Pure NumPy (C1 additionally uses scikit-learn's generator, as the reference's own tests and
benchmarks do: test/integration/test_procedures.py:131-139, docsrc/benchmark/helper_base.py:142-149).
The draw order of every generator is part of its definition -- do not reorder.
"""
import numpy as np

# (radius_cutoff, cnn_cutoff, member_cutoff) per configuration
C1_PARAMS = dict(radius_cutoff=0.3, cnn_cutoff=10, member_cutoff=None)
C2_PARAMS = dict(radius_cutoff=0.09, cnn_cutoff=20, member_cutoff=10)
C3_PARAMS = dict(radius_cutoff=0.6, cnn_cutoff=20, member_cutoff=10)
C4_PARAMS = dict(radius_cutoff=0.037, cnn_cutoff=10, member_cutoff=10)
C5_RADII = (0.045, 0.050, 0.056, 0.063, 0.071, 0.079, 0.089, 0.100)
C5_CNN = (2, 5, 10, 15, 20, 30, 50, 80)


def blobs_2d(n: int = 10_000) -> np.ndarray:
    """C1: three 2-D Gaussian blobs with different spread, standard-scaled."""
    from sklearn.datasets import make_blobs
    from sklearn.preprocessing import StandardScaler

    x, _ = make_blobs(n_samples=n, n_features=2, centers=3, cluster_std=[1.0, 2.5, 0.5], random_state=170)
    return np.ascontiguousarray(StandardScaler().fit_transform(x), dtype=np.float64)


def md_like_3d(n: int = 1_000_000, seed: int = 0) -> np.ndarray:
    """C2/C4/C5: metastable jump process between 12 Gaussian wells in 3-D plus 5 % uniform background."""
    rng = np.random.default_rng(seed)
    centres = rng.random((12, 3)) * 10
    sigma = rng.uniform(0.25, 0.6, 12)
    switch = rng.random(n) < 1e-3
    new_state = rng.integers(0, 12, n)
    # state[i] = new_state[i] if switch[i] else state[i-1], state[-1] = 0
    idx = np.where(switch, np.arange(n), -1)
    last = np.maximum.accumulate(idx)
    state = np.where(last >= 0, new_state[np.maximum(last, 0)], 0)
    x = centres[state] + rng.standard_normal((n, 3)) * sigma[state][:, None]
    bg = rng.random(n) < 0.05
    x[bg] = rng.random((int(bg.sum()), 3)) * 12 - 1
    return np.ascontiguousarray(x, dtype=np.float64)


def features_64d(n: int = 200_000, seed: int = 0, d: int = 64) -> np.ndarray:
    """C3: 16 clusters of intrinsic dimension 6 embedded in d = 64 dimensions plus small isotropic noise."""
    rng = np.random.default_rng(seed)
    centres = rng.uniform(-2, 2, (16, d))
    lab = rng.integers(0, 16, n)
    x = np.empty((n, d), dtype=np.float64)
    for k in range(16):
        q, _ = np.linalg.qr(rng.standard_normal((d, 6)))
        nk = int((lab == k).sum())
        z = rng.standard_normal((nk, 6)) * np.linspace(0.6, 0.2, 6)
        x[lab == k] = centres[k] + z @ q.T
    x += rng.standard_normal((n, d)) * 0.03
    return np.ascontiguousarray(x)


def uniform_2d(n: int, seed: int = 2021) -> np.ndarray:
    """The reference benchmark's "no_structure" set (docsrc/benchmark/helper_base.py:135-139)."""
    from sklearn.preprocessing import StandardScaler

    x = np.random.default_rng(seed).random((n, 2))
    return np.ascontiguousarray(StandardScaler().fit_transform(x), dtype=np.float64)


README_POINTS = np.array(
    [[0, 0], [1, 1], [1, 0], [0, -1], [0.5, -0.5], [2, 1.5], [2.5, -0.5], [4, 2], [4.5, 2.5], [5, -1],
     [5.5, -0.5], [5.5, -1.5]], dtype=np.float64)
README_LABELS = np.array([1, 1, 1, 1, 1, 0, 0, 0, 0, 2, 2, 2], dtype=np.int64)
