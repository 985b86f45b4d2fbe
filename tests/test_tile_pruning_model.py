"""Host-side model of the tile pruning of the tensor-core search (csrc/tensor.cu: `tile_ball_kernel`, `balls_apart`,
`tile_skip_kernel`, `tile_skip_split_kernel`): points sorted by (nearest pivot, distance to it), one bounding ball per
128-row tile -- two when the tile straddles a change of pivot --, a tile PAIR is skipped when its balls are further
apart than the radius.  The first pass looks at first balls only; the split pass takes a skip back unless ALL
combinations of the two tiles' balls are apart.  Checked here on the CPU: no pair of points within the radius ever sits
in a skipped tile pair (the search would lose a neighbour silently), the pruning really prunes, and two balls prune more
than one.  The safety margins are READ from tensor.cu.  No GPU involved; tests/test_gpu_fullsize.py compares the real
kernel's neighbour sets with the oracle's brute force entry for entry."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "commonnnclustering_b200", "csrc", "tensor.cu")).read()


def constants():
    tt = re.search(r"#define TT (\d+)", SRC)
    piv = re.search(r"#define N_PIVOTS (\d+)", SRC)
    grow = re.search(r"const double rr = rmax\[0\] \* \(1\.0 \+ ([0-9.e-]+)\) \+ ([0-9.e-]+);", SRC)
    shrink = re.search(r"return sqrt\(s\) \* \(1\.0 - ([0-9.e-]+)\) - ra - rb > reach;", SRC)
    reach = re.findall(r"r_true \* \(1\.0 \+ ([0-9.e-]+)\)", SRC)
    assert tt and piv and grow and shrink and len(reach) >= 2 and len(set(reach)) == 1, \
        "the pruning's margins changed their form: update this test together with tensor.cu"
    return int(tt.group(1)), int(piv.group(1)), float(grow.group(1)), float(grow.group(2)), float(shrink.group(1)), \
        float(reach[0])


def pivot_order(x, n_piv):
    """pivot_assign_kernel + key sort: pivots are evenly spaced input points, key = (nearest pivot, distance to it)."""
    n = x.shape[0]
    piv = x[(np.arange(n_piv) * (n / n_piv)).astype(np.int64)]
    d2 = ((x[:, None, :] - piv[None, :, :]) ** 2).sum(axis=-1)
    arg = d2.argmin(axis=1)
    order = np.lexsort((d2[np.arange(n), arg], arg))
    return order, arg[order]


def tile_balls(x, pivot_of, tt, grow, tiny, split=True):
    """tile_ball_kernel: centre = mean of the rows, radius = largest distance to it, rounded up; rows from the first
    change of pivot on get a second ball (radius2 = -1: there is none)."""
    n, d = x.shape
    n_tiles = -(-n // tt)
    c1, c2 = np.zeros((n_tiles, d)), np.zeros((n_tiles, d))
    r1, r2 = np.zeros(n_tiles), -np.ones(n_tiles)
    for t in range(n_tiles):
        rows = x[t * tt:(t + 1) * tt]
        pv = pivot_of[t * tt:(t + 1) * tt]
        change = np.flatnonzero(pv != pv[0])
        sp = int(change[0]) if (split and change.size) else rows.shape[0]
        c1[t] = rows[:sp].mean(axis=0)
        r1[t] = np.sqrt(((rows[:sp] - c1[t]) ** 2).sum(axis=1)).max() * (1.0 + grow) + tiny
        if sp < rows.shape[0]:
            c2[t] = rows[sp:].mean(axis=0)
            r2[t] = np.sqrt(((rows[sp:] - c2[t]) ** 2).sum(axis=1)).max() * (1.0 + grow) + tiny
    return c1, r1, c2, r2


def apart(ca, ra, cb, rb, reach, shrink):
    dist = np.sqrt(((ca[:, None, :] - cb[None, :, :]) ** 2).sum(axis=-1))
    return dist * (1.0 - shrink) - ra[:, None] - rb[None, :] > reach


def skip_matrix(c1, r1, c2, r2, r_true, shrink, margin):
    reach = r_true * (1.0 + margin)
    skip = apart(c1, r1, c1, r1, reach, shrink)                       # tile_skip_kernel: first balls only
    has2 = r2 >= 0.0
    # tile_skip_split_kernel, S = a tile with a second ball, X = any tile
    ok = apart(c1, r1, c1, r1, reach, shrink) & apart(c2, r2, c1, r1, reach, shrink)          # [S, X]
    both = apart(c1, r1, c2, r2, reach, shrink) & apart(c2, r2, c2, r2, reach, shrink)
    ok = np.where(has2[None, :], ok & both, ok)
    take_back = has2[:, None] & ~ok
    skip &= ~take_back
    skip &= ~take_back.T
    return skip


def clustered(rng, n, d, k, spread):
    centres = rng.standard_normal((k, d)) * 4.0
    return centres[rng.integers(0, k, n)] + rng.standard_normal((n, d)) * spread


CASES = {
    "blobs_16d": lambda rng: (clustered(rng, 4000, 16, 12, 0.5), 2.4),
    "blobs_64d": lambda rng: (clustered(rng, 3000, 64, 8, 0.35), 3.6),
    "blobs_5d_touching": lambda rng: (clustered(rng, 4000, 5, 20, 0.8), 1.2),
    "uniform_8d": lambda rng: (rng.random((3000, 8)) * 3.0, 0.9),
    "lattice_ties_6d": lambda rng: (rng.integers(0, 5, (3000, 6)).astype(np.float64) * 0.5, 0.5),
    "odd_size_last_tile_short": lambda rng: (clustered(rng, 2000 + 37, 12, 6, 0.5), 2.0),
}


@pytest.mark.parametrize("n_piv", [None, 8])     # the kernel's pivot count (a few rows per pivot at these sizes) and a
@pytest.mark.parametrize("name", list(CASES))    # count that gives whole tiles per pivot, as at the real sizes
def test_no_neighbour_pair_sits_in_a_skipped_tile_pair(name, n_piv):
    tt, n_piv_kernel, grow, tiny, shrink, margin = constants()
    n_piv = n_piv or n_piv_kernel
    x, r = CASES[name](np.random.default_rng(21))
    order, pivot_of = pivot_order(x, n_piv)
    xs = np.ascontiguousarray(x[order])
    n = xs.shape[0]
    skip = skip_matrix(*tile_balls(xs, pivot_of, tt, grow, tiny), r, shrink, margin)
    assert not skip.diagonal().any()
    assert np.array_equal(skip, skip.T)
    # the reference's expression for every pair, then the tile pair it falls into
    s = np.zeros((n, n))
    for k in range(xs.shape[1]):
        t = xs[:, None, k] - xs[None, :, k]
        s = s + t * t
    i, j = np.nonzero(s <= r * r)
    assert not skip[i // tt, j // tt].any(), "a pair within the radius lies in a tile pair the search would skip"
    if name.startswith("blobs") and n_piv == 8:
        assert skip.mean() > (0.1 if "touching" in name else 0.3)   # clustered data: the pruning prunes


def test_a_straddling_tile_gets_two_balls_and_they_help():
    tt, _, grow, tiny, shrink, margin = constants()
    x, r = CASES["blobs_16d"](np.random.default_rng(21))
    order, pivot_of = pivot_order(x, 8)
    xs = np.ascontiguousarray(x[order])
    two = tile_balls(xs, pivot_of, tt, grow, tiny)
    one = tile_balls(xs, pivot_of, tt, grow, tiny, split=False)
    assert (two[3] >= 0).sum() >= 5                   # tiles that straddle a change of pivot
    assert skip_matrix(*two, r, shrink, margin).sum() > skip_matrix(*one, r, shrink, margin).sum()


def test_the_model_notices_a_ball_that_is_too_small():
    """Control: balls at 60 % of their radius skip tile pairs that hold neighbours."""
    tt, _, grow, tiny, shrink, margin = constants()
    x, r = CASES["blobs_5d_touching"](np.random.default_rng(21))
    order, pivot_of = pivot_order(x, 8)
    xs = np.ascontiguousarray(x[order])
    c1, r1, c2, r2 = tile_balls(xs, pivot_of, tt, grow, tiny)
    skip = skip_matrix(c1, 0.6 * r1, c2, np.where(r2 >= 0, 0.6 * r2, r2), r, shrink, margin)
    d2 = ((xs[:, None, :] - xs[None, :, :]) ** 2).sum(axis=-1)
    i, j = np.nonzero(d2 <= r * r)
    assert skip[i // tt, j // tt].any()
