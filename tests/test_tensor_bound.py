"""Host-side check of the tensor-core filter's error bound ("Exactness of the tensor-core filter"): the fp16
hi/lo split, the [lo.hi | hi.lo | hi.hi] accumulation in f32 and the epilogue's f32 roundings are emulated in NumPy and
the classification "neighbour for certain / certainly not" is compared with the exact f64 distances.  The bound's
constants are read from csrc/tensor.cu, so the test and the kernel cannot drift apart.  No GPU involved: what the
hardware adds beyond this emulation (the order and rounding mode inside one MMA instruction) is what the factor 17 per
instruction and the factor 2 of head-room are for; tests/test_gpu_* check the real thing against the oracle."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "commonnnclustering_b200", "csrc", "tensor.cu")).read()


def bound_constants():
    m = re.search(r"double coef = ([0-9.]+) \* ldexp\(1\.0, -22\) \+ ([0-9.]+) \* ldexp\(1\.0, -23\) \* "
                  r"\(n_hh \* ([0-9.]+) \+ n_small \* ([0-9.e-]+)\) \+ ldexp\(1\.0, -20\);", SRC)
    assert m, "the bound in cnn_tensor_search changed its form: update this test and DESIGN.md together"
    return [float(g) for g in m.groups()]


def emulate(x, r):
    """-> (sure, maybe, d2_exact_scaled, R_s, band fraction) for all pairs i < j."""
    c_split, c_acc, w_hh, w_small = bound_constants()
    n, d = x.shape
    xc = x - x.mean(axis=0)
    maxabs = np.abs(xc).max()
    e = 14 - int(np.frexp(maxabs)[1]) if maxabs > 0 else 0
    X = np.ldexp(xc, e)
    hi = X.astype(np.float16)
    lo = (X - hi.astype(np.float64)).astype(np.float16)
    assert np.all(np.isfinite(hi.astype(np.float64)))
    A = np.concatenate([lo, hi, hi], axis=1).astype(np.float32)
    B = np.concatenate([hi, lo, hi], axis=1).astype(np.float32)
    acc = np.zeros((n, n), dtype=np.float32)
    for k0 in range(0, 3 * d, 16):                      # one K = 16 step at a time, f32 accumulator in between
        acc = (acc + A[:, k0:k0 + 16] @ B[:, k0:k0 + 16].T).astype(np.float32)
    norm = (X * X).sum(axis=1)
    nf = norm.astype(np.float32)
    M = float(nf.max()) * (1.0 + 2.0 ** -20)
    kt = -(-3 * d // 64) * 64
    n_instr, n_small = kt // 16, (2 * d) // 16
    n_hh = n_instr - n_small
    coef = c_split * 2.0 ** -22 + c_acc * 2.0 ** -23 * (n_hh * w_hh + n_small * w_small) + 2.0 ** -20
    Rs = np.ldexp(r * r, 2 * e)
    eps = M * coef + Rs * 2.0 ** -22 + 1e-300
    thr_lo = np.nextafter(np.float32(Rs - eps), np.float32(-np.inf))
    thr_hi = np.nextafter(np.float32(Rs + eps), np.float32(np.inf))
    w = (nf[None, :] - np.float32(2.0) * acc).astype(np.float32)     # fma(-2, acc, nc): one rounding
    t_lo = (thr_lo - nf)[:, None].astype(np.float32)
    t_hi = (thr_hi - nf)[:, None].astype(np.float32)
    sure, maybe = w <= t_lo, w <= t_hi
    Xl = X.astype(np.longdouble)
    d2 = ((Xl[:, None, :] - Xl[None, :, :]) ** 2).sum(axis=-1)
    iu = np.triu_indices(n, k=1)
    return sure[iu], maybe[iu], d2[iu], np.longdouble(Rs), float((maybe[iu] & ~sure[iu]).mean())


CASES = {
    "gauss_d5": lambda rng: (rng.standard_normal((500, 5)), 1.8),
    "gauss_d21": lambda rng: (rng.standard_normal((400, 21)), 5.5),
    "gauss_d64": lambda rng: (rng.standard_normal((400, 64)), 10.5),
    "gauss_d85": lambda rng: (rng.standard_normal((300, 85)), 12.5),
    "offset_50_sigma": lambda rng: (rng.standard_normal((400, 33)) + 50.0, 7.5),
    "scaled_1e-20": lambda rng: (rng.standard_normal((400, 12)) * 1e-20, 4.0e-20),
    "one_wide_dimension": lambda rng: (rng.standard_normal((400, 10)) * np.array([1e6] + [1.0] * 9), 1.0e6),
    "lattice_ties": lambda rng: (rng.integers(0, 4, (400, 6)).astype(np.float64) * 0.25, 0.5),
}


@pytest.mark.parametrize("name", list(CASES))
def test_filter_never_misclassifies_outside_its_band(name):
    x, r = CASES[name](np.random.default_rng(99))
    sure, maybe, d2, Rs, band = emulate(np.ascontiguousarray(x), r)
    assert not np.any(sure & (d2 > Rs)), "a pair taken as a neighbour for certain is not one"
    assert not np.any(~maybe & (d2 <= Rs)), "a pair ruled out for certain is a neighbour"
    assert sure.any() and (~maybe).any()              # the case exercises both sides of the cutoff
    assert band < 0.25                                # and the band stays a band
