"""Host-side model of the periodic cell list (csrc/periodic.cu): every point wrapped into [0, L), plus its images within
r' of a periodic face; the non-periodic search with radius r' on that augmented set is a FILTER, the reference's own
arithmetic on the ORIGINAL coordinates decides.  What can go wrong silently is the filter: the reference takes
`fmod(|a - b|, L)` of the rounded difference, the device wraps a and b separately, so r' carries a rounding margin
(`box.rp`, read from the source here).  Checked on the CPU against the oracle's brute force under
MetricExtEuclideanPeriodicReduced (_types.pyx:2290-2318): every neighbour pair of the reference is met by the filter
(coordinates many periods outside the box, negative ones, offsets of 1e6-1e9 periods where the roundings of the wrap are
as large as they get, pairs exactly on the cutoff across a face), no point meets another point AND one of its images,
and a control without the margin does lose pairs.  No GPU involved;
tests/test_gpu_parity.py::test_periodic_cell_list_neighbour_sets_are_exact checks the real kernels."""
import os
import re

import numpy as np
import pytest

from oracle import port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "commonnnclustering_b200", "csrc", "periodic.cu")).read()


def filter_radius(r, absmax, lmax, with_margin=True):
    m = re.search(r"box\.rp = r \* \(1\.0 \+ ([0-9.e-]+)\) \+ ([0-9.]+) \* eps \* \(([0-9.]+) \* absmax \+ ([0-9.]+) \* Lmax\);",
                  SRC)
    assert m, "box.rp changed its form: update this test together with periodic.cu"
    rel, k, ka, kl = (float(g) for g in m.groups())
    if not with_margin:
        return r
    return r * (1.0 + rel) + k * 2.220446049250313e-16 * (ka * absmax + kl * lmax)


def wrap(v, L):
    """pg_wrap: fmod, negative remainders moved up by one period."""
    if not L > 0.0:
        return v.copy()
    w = np.fmod(v, L)
    return np.where(w < 0.0, w + L, w)


def augmented(x, bounds, rp):
    """pg_image_count / pg_image_fill: (coordinates [m, d], the point every entry stands for [m])."""
    n, d = x.shape
    w = np.stack([wrap(x[:, k], bounds[k]) for k in range(d)], axis=1)
    pts, orig = [w], [np.arange(n)]
    shifts = []
    for k in range(d):
        L = bounds[k]
        lower = (w[:, k] < rp) if L > 0 else np.zeros(n, bool)          # an image at w + L
        upper = (w[:, k] >= L - rp) if L > 0 else np.zeros(n, bool)     # an image at w - L
        shifts.append([(np.ones(n, bool), 0.0), (lower, L), (upper, -L)])
    for combo in np.ndindex(*([3] * d)):
        if not any(combo):
            continue
        need = np.ones(n, bool)
        delta = np.zeros(d)
        for k, o in enumerate(combo):
            need &= shifts[k][o][0]
            delta[k] = shifts[k][o][1]
        idx = np.flatnonzero(need)
        if idx.size:
            pts.append(w[idx] + delta)
            orig.append(idx)
    return np.concatenate(pts), np.concatenate(orig)


def filter_pairs(x, bounds, rp):
    """(i, o) for every entry of the filter rows of the real points: augmented entries within r' (the cell list's f64
    expression, squared), mapped to the points they stand for."""
    n, d = x.shape
    aug, orig = augmented(x, bounds, rp)
    s = np.zeros((n, aug.shape[0]))
    for k in range(d):
        t = aug[:n, None, k] - aug[None, :, k]
        s = s + t * t
    i, q = np.nonzero(s <= rp * rp)
    return i, orig[q], aug.shape[0] - n


def reference_pairs(x, r, bounds):
    row_ptr, col_idx = port.neighbours_periodic(x, r * r, bounds)
    return np.repeat(np.arange(x.shape[0]), np.diff(row_ptr)), col_idx


def across_a_face(rng, n, offset):
    """Pairs EXACTLY r = 0.125 apart across the periodic face of a box of period 1 (binary fractions), many periods out."""
    base = rng.integers(0, 64, (n, 2)) / 64.0
    base[: n // 2, 0] = rng.integers(0, 4, n // 2) / 64.0                # a strip at the lower face ...
    base[n // 2:, 0] = 1.0 - rng.integers(1, 9, n - n // 2) / 64.0       # ... and one at the upper face
    return base + rng.integers(-3, 4, (n, 2)) + offset


CASES = {
    # name: (points, periods (0 = not periodic), r)
    "box_2d": lambda rng: (rng.random((900, 2)) * np.array([2.0, 3.0]), [2.0, 3.0], 0.2),
    "many_periods_mixed_3d": lambda rng: (np.c_[rng.random(900) * 50.0 - 20.0, rng.standard_normal(900),
                                                rng.random(900) * 21.0 - 7.0], [10.0, 0.0, 7.0], 1.1),
    "ring_1d_negative": lambda rng: (rng.random((800, 1)) * 3.0 - 4.0, [1.0], 0.01),
    "ties_across_the_face": lambda rng: (across_a_face(rng, 800, 0.0), [1.0, 1.0], 0.125),
    "ties_1e6_periods_out": lambda rng: (across_a_face(rng, 800, 1.0e6), [1.0, 1.0], 0.125),
    "far_out_1e9_irrational_period": lambda rng: (rng.random((800, 2)) * 40.0 + 1.0e9, [2 * np.pi, 2 * np.pi], 0.35),
    "dihedrals": lambda rng: (np.c_[rng.vonmises(2.9, 1.5, 900), rng.vonmises(-3.0, 2.0, 900)], [2 * np.pi] * 2, 0.15),
}


@pytest.mark.parametrize("name", list(CASES))
def test_filter_meets_every_neighbour_pair_of_the_reference(name):
    x, bounds, r = CASES[name](np.random.default_rng(31))
    x = np.ascontiguousarray(x, dtype=np.float64)
    bounds = np.asarray(bounds, dtype=np.float64)
    rp = filter_radius(r, float(np.abs(x).max()), float(bounds.max()))
    assert all(2.0 * rp < L for L in bounds if L > 0)                   # the construction serves this request
    n = x.shape[0]
    fi, fo, n_images = filter_pairs(x, bounds, rp)
    assert n_images > 0
    got = set((fi * n + fo).tolist())
    assert len(got) == fi.size, "a point meets another point and one of its images (or two images of it)"
    ri, rj = reference_pairs(x, r, bounds)
    want = set((ri * n + rj).tolist())
    assert want <= got, f"{len(want - got)} neighbour pairs of the reference never reach the exact test"
    assert len(want) > n                                                # more than the diagonal
    if "ties" in name:
        # the construction of the case worked: pairs across the face at exactly the cutoff are neighbours
        w = np.stack([wrap(x[:, k], bounds[k]) for k in range(2)], axis=1)
        cross = np.abs(w[ri, 0] - w[rj, 0]) > 0.5
        assert cross.any()


def test_the_model_notices_a_missing_margin():
    """Control: a in [0, r) and b = a - r < 0, exactly r apart for the reference (it subtracts the coordinates as they
    are).  The device moves b up by one period (a rounding at 2^-53) and meets the pair through an image: with r' = r
    about half of these pairs never reach the exact test, with the margin all do."""
    rng = np.random.default_rng(5)
    r = 0.125
    a = r * rng.random(600)
    x = np.ascontiguousarray(np.r_[a, a - r].reshape(-1, 1))
    bounds = np.array([1.0])
    n = x.shape[0]
    ri, rj = reference_pairs(x, r, bounds)
    want = set((ri * n + rj).tolist())
    assert all((k * n + k + 600) in want for k in range(600))           # every constructed pair IS a neighbour pair
    fi, fo, _ = filter_pairs(x, bounds, filter_radius(r, 0.0, 0.0, with_margin=False))
    assert len(want - set((fi * n + fo).tolist())) > 100
    fi, fo, _ = filter_pairs(x, bounds, filter_radius(r, float(np.abs(x).max()), 1.0))
    assert want <= set((fi * n + fo).tolist())
