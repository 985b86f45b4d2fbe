"""Host-side check of the fp32 filter of the cell-list search ("Exactness of the fp32 filter (K2b)"):
the grid geometry, the cell-local fp32 copies on lattices A / B, the per-group shifts to the centre of the candidate box
and the expanded form `w = |c|^2 - 2 q.c`, `t = w - fl32(R - |q|^2)` of `grid_mask_kernel` are emulated in NumPy, for
EVERY group a query could end up in (first cell 0..3 to its left, last cell up to 3 further on: a superset of the groups
the kernel forms), and every pair the filter would decide by the sign of `t` is compared with the reference's own f64
expression.  The band's constants are read from csrc/grid.cu, so the test and the kernel cannot drift apart.  No GPU
involved; tests/test_gpu_parity.py::test_filter_kernel_neighbour_sets_are_exact checks the real kernel against the
oracle.  (fmaf is emulated through extended precision: a double rounding can move a result by one fp32 ulp in rare
cases, far inside the six-fold head-room of the band.)"""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "commonnnclustering_b200", "csrc", "grid.cu")).read()
F32 = np.float32


def band_terms():
    """[(what, exponent)] of `eps` in filter_params, plus the cell-edge factor and the group span."""
    m = re.search(r"const double eps = ldexp\(h \* r, (-\d+)\) \+ ldexp\(h \* h, (-\d+)\) \+ ldexp\(R, (-\d+)\) \+ "
                  r"ldexp\(h \* h, (-\d+)\);", SRC)
    assert m, "the band in filter_params changed its form: update this test and DESIGN.md together"
    grow = re.search(r"g\.h = r_true \* \(1\.0 \+ ([0-9.e-]+)\);", SRC)
    span = re.search(r"#define GF_XSPAN (\d+)", SRC)
    coarsen = re.search(r"g\.h \*= ([0-9.]+);", SRC)
    assert grow and span and coarsen
    return [int(g) for g in m.groups()], float(grow.group(1)), int(span.group(1)), float(coarsen.group(1))


def fma32(a, b, c):
    return (a.astype(np.longdouble) * b.astype(np.longdouble) + c.astype(np.longdouble)).astype(F32)


def geometry(x, r_true):
    """cnn_grid_build: bounding box, cell edge (coarsened for sparse boxes), cell coordinates."""
    _, grow, _, coarsen = band_terms()
    n, d = x.shape
    lo, hi = x.min(axis=0), x.max(axis=0)
    h = r_true * (1.0 + grow)
    max_cells = max(8.0 * n, 1048576.0)
    while True:
        dims = np.floor((hi - lo) / h) + 1
        if np.all(dims <= 1048576.0) and np.prod(dims) <= max_cells:
            break
        h *= coarsen
    dims = dims.astype(np.int64)
    cell = np.clip(np.floor((x - lo) / h).astype(np.int64), 0, dims - 1)
    return lo, h, dims, cell


def origin(lo, h, c):
    return lo + c.astype(np.float64) * h          # cell_origin: __dadd_rn(lo, __dmul_rn(c, h))


def emulate(x, rc, sqrt_metric=False):
    """-> (decided_hit, decided_miss, exact_hit, margin) over all (group, query, candidate) triples."""
    exps, _, xspan, _ = band_terms()
    n, d = x.shape
    R = rc * rc if sqrt_metric else rc
    r = np.sqrt(R)
    lo, h, dims, cell = geometry(x, r)
    eps = sum(np.ldexp(v, e) for v, e in zip((h * r, h * h, R, h * h), exps))
    band = np.nextafter(F32(eps * (1.0 + 2.0 ** -20)), F32(np.inf))
    cx = cell[:, 0]
    # the fp32 cell-local copy (scatter_kernel / gather_kernel)
    fa = (x[:, 0] - origin(lo[0], h, cx & ~15)).astype(F32)
    fb = (x[:, 0] - origin(lo[0], h, ((cx + 8) & ~15) - 8)).astype(F32)
    fy = (x[:, 1] - origin(lo[1], h, cell[:, 1])).astype(F32) if d > 1 else np.zeros(n, F32)
    fz = (x[:, 2] - origin(lo[2], h, cell[:, 2])).astype(F32) if d > 2 else np.zeros(n, F32)
    hh = F32(0.5) * F32(h)
    # the reference's f64 expression on the original coordinates (`within`)
    s = np.zeros((n, n))
    for k in range(d):
        t = x[:, None, k] - x[None, :, k]
        s = s + t * t if k else t * t
    exact = (np.sqrt(s) if sqrt_metric else s) <= rc
    near_rows = np.ones((n, n), bool)
    for k in range(1, d):
        near_rows &= np.abs(cell[:, None, k] - cell[None, :, k]) <= 1
    out = []
    worst = 0.0
    for df in range(xspan + 1):                  # the group's first cell lies df cells left of the query's
        for sp in range(df, xspan + 1):          # and its last cell sp cells right of the first
            first, last = cx - df, cx - df + sp
            ok = (first >= 0) & (last <= dims[0] - 1)
            cx0, cx1 = np.maximum(first - 1, 0), np.minimum(last + 1, dims[0] - 1)
            use_b = (cx0 >> 4) != (cx1 >> 4)
            lat = np.where(use_b, ((cx0 + 8) & ~15) - 8, cx0 & ~15)
            xc = origin(lo[0], h, cx0) + (0.5 * (cx1 - cx0 + 1).astype(np.float64)) * h
            shx = (origin(lo[0], h, lat) - xc).astype(F32)
            pair = near_rows & ok[:, None] & (cx[None, :] >= cx0[:, None]) & (cx[None, :] <= cx1[:, None])
            qi, cj = np.nonzero(pair)
            if qi.size == 0:
                continue
            # the stored copy must refer to the lattice origin the group's shift assumes
            own = np.where(use_b[qi], ((cx[cj] + 8) & ~15) - 8, cx[cj] & ~15)
            assert np.array_equal(own, lat[qi]), "a candidate's lattice origin differs from its group's"
            qx = np.where(use_b, fb, fa) + shx                       # f32 + f32
            qy = fy - hh if d > 1 else np.zeros(n, F32)
            qz = fz - hh if d > 2 else np.zeros(n, F32)
            qn = (qx.astype(np.float64) ** 2 + qy.astype(np.float64) ** 2) + qz.astype(np.float64) ** 2
            T = (R - qn).astype(F32)
            ccx = np.where(use_b[qi], fb[cj], fa[cj]) + shx[qi]
            zero = np.zeros(qi.size, F32)
            if d > 1:
                shy = ((origin(lo[1], h, cell[cj, 1]) - origin(lo[1], h, cell[qi, 1])) - 0.5 * h).astype(F32)
                ccy = fy[cj] + shy
            else:
                ccy = zero
            if d > 2:
                shz = ((origin(lo[2], h, cell[cj, 2]) - origin(lo[2], h, cell[qi, 2])) - 0.5 * h).astype(F32)
                ccz = fz[cj] + shz
            else:
                ccz = zero
            cw = fma32(ccz, ccz, fma32(ccy, ccy, ccx * ccx))
            q2x, q2y, q2z = F32(-2) * qx[qi], F32(-2) * qy[qi], F32(-2) * qz[qi]
            w = fma32(q2z, ccz, fma32(q2y, ccy, fma32(q2x, ccx, cw)))
            tt = w - T[qi]
            decided = np.abs(tt) > band
            ex = exact[qi, cj]
            out.append((decided & (tt < 0), decided & ~(tt < 0), ex))
            # how much of the band the arithmetic really uses, on the pairs within reach of the cutoff
            true_t = s[qi, cj].astype(np.longdouble) - np.longdouble(R)
            close = s[qi, cj] <= 4.0 * R                          # within twice the radius
            if close.any():
                worst = max(worst, float(np.abs(tt[close].astype(np.longdouble) - true_t[close]).max()) / float(band))
    hit, miss, ex = (np.concatenate(c) for c in zip(*out))
    return hit, miss, ex, worst


def lattice(rng, n, d, step, extent):
    return rng.integers(0, extent, (n, d)).astype(np.float64) * step


CASES = {
    # name: (points, radius_cutoff as the search receives it, sqrt metric)
    "uniform_3d": lambda rng: (rng.random((1800, 3)) * 4.0, 0.25 ** 2, False),
    "uniform_2d": lambda rng: (rng.random((1800, 2)) * 10.0, 0.3 ** 2, False),
    "uniform_1d": lambda rng: (rng.random((1500, 1)) * 60.0, 0.2 ** 2, False),
    "wide_x_crosses_lattice_periods": lambda rng: (rng.random((2000, 3)) * np.array([40.0, 1.5, 1.5]), 0.5 ** 2, False),
    "offset_1e7": lambda rng: (rng.random((1500, 3)) * 3.0 + 1.0e7, 0.25 ** 2, False),
    "negative_offset_1e6_2d": lambda rng: (rng.random((1500, 2)) * 8.0 - 1.0e6, 0.3 ** 2, False),
    "lattice_ties_3d": lambda rng: (lattice(rng, 1800, 3, 0.125, 20), 0.375 ** 2, False),
    "lattice_ties_2d_sqrt": lambda rng: (lattice(rng, 1500, 2, 0.25, 40), 0.75, True),
    "scale_1e-9": lambda rng: (rng.random((1500, 3)) * 4.0e-9, (0.25e-9) ** 2, False),
    "scale_1e9": lambda rng: (rng.random((1500, 3)) * 4.0e9, (0.25e9) ** 2, False),
    "euclidean_sqrt_3d": lambda rng: (rng.random((1500, 3)) * 4.0, 0.25, True),
    "duplicates": lambda rng: (np.repeat(rng.random((300, 3)) * 2.0, 5, axis=0), 0.2 ** 2, False),
    # two far clusters: the cell count exceeds its budget and the cells are coarsened well beyond the radius
    "coarsened_cells": lambda rng: (np.r_[rng.random((800, 3)) * 0.3, rng.random((800, 3)) * 0.3 + 60.0], 0.02 ** 2,
                                    False),
}


@pytest.mark.parametrize("name", list(CASES))
def test_fp32_filter_decides_only_what_it_can(name):
    x, rc, sqrt_metric = CASES[name](np.random.default_rng(7))
    hit, miss, exact, worst = emulate(np.ascontiguousarray(x, dtype=np.float64), rc, sqrt_metric)
    assert not np.any(hit & ~exact), "a pair the filter accepts outright is not a neighbour"
    assert not np.any(miss & exact), "a pair the filter rejects outright is a neighbour"
    assert hit.any() and miss.any()                      # both sides of the cutoff are exercised
    borderline = 1.0 - (hit | miss).mean()
    assert borderline < (0.2 if "ties" in name else 0.01)  # and the band stays a band (ties sit ON the cutoff)
    assert worst < 0.5, f"the arithmetic uses {worst:.2f} of the band within reach of the cutoff: the head-room is gone"


def test_coarsened_case_really_coarsens():
    x, rc, _ = CASES["coarsened_cells"](np.random.default_rng(7))
    _, h, dims, _ = geometry(x, np.sqrt(rc))
    assert h > 3.0 * np.sqrt(rc) and np.prod(dims) <= max(8.0 * len(x), 1048576.0)


def test_the_emulation_notices_a_band_that_is_too_narrow(monkeypatch):
    """Control: with the band 2^-12 of its size the pairs that sit ON the cutoff are decided by fp32 roundings."""
    import sys

    me = sys.modules[__name__]
    real = band_terms()
    monkeypatch.setattr(me, "band_terms", lambda: ([e - 12 for e in real[0]], *real[1:]))
    x, rc, sqrt_metric = CASES["lattice_ties_3d"](np.random.default_rng(7))
    hit, miss, exact, worst = emulate(np.ascontiguousarray(x, dtype=np.float64), rc, sqrt_metric)
    assert np.any(hit & ~exact) or np.any(miss & exact)
    assert worst > 1.0


EXTREME = {
    # a bounding box 10^11 radii wide: the cells are coarsened to the 2^20-per-dimension cap, (x - lo) loses its low bits
    "two_clusters_1e9_apart": lambda rng: (np.r_[rng.random((600, 2)) * 0.2, rng.random((600, 2)) * 0.2 + 1.0e9], 0.01 ** 2,
                                           False),
    "negative_far_corner": lambda rng: (np.r_[rng.random((800, 3)) * 0.3, [[-3.0e8, -3.0e8, -3.0e8]]], 0.02 ** 2, False),
}


@pytest.mark.parametrize("name", list(CASES) + list(EXTREME))
def test_candidate_windows_hold_every_neighbour(name):
    """The cell list only ever tests pairs whose cells differ by at most one along every axis (cell edge = radius times
    1 + 1e-6, read from grid.cu): no pair the reference's f64 expression accepts may lie outside, whatever the rounding
    of (x - lo) / h does -- also when the box is so large that the cells are coarsened to the per-dimension cap."""
    x, rc, sqrt_metric = (CASES.get(name) or EXTREME[name])(np.random.default_rng(7))
    x = np.ascontiguousarray(x, dtype=np.float64)
    r = rc if sqrt_metric else np.sqrt(rc)
    _, h, dims, cell = geometry(x, r)
    assert np.all(dims <= 1048576) and h >= r
    n, d = x.shape
    s = np.zeros((n, n))
    for k in range(d):
        t = x[:, None, k] - x[None, :, k]
        s = s + t * t if k else t * t
    i, j = np.nonzero((np.sqrt(s) if sqrt_metric else s) <= rc)
    assert i.size > n                                                   # more than the diagonal
    assert np.abs(cell[i] - cell[j]).max() <= 1
