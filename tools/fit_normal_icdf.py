#!/usr/bin/env python
"""Fit the float64 polynomials used by mbt_normal_from_bits (include/mbt_math.h).

The half-normal quantile is evaluated in the form popularised by M. Giles
("Approximating the erfinv function", GPU Computing Gems 2011):

    erfinv(v) = v * P(w),   w = -log((1 - v) * (1 + v)),

with P a polynomial in (w - w0) for small w and in (sqrt(w) - s0) for the tails.
The float32 coefficients in mbt_math.h are Giles' published single-precision set;
the float64 sets are produced HERE by Chebyshev interpolation against
scipy.special.erfinv / erfcinv, so the header carries no borrowed table.

Run:  python tools/fit_normal_icdf.py   -> prints C initialisers + max errors.
"""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P
from scipy import special


def target(w):
    """g(w) = erfinv(v)/v with v = sqrt(1 - exp(-w)), evaluated stably."""
    w = np.asarray(w, dtype=np.float64)
    one_minus_v2 = np.exp(-w)
    v = np.sqrt(-np.expm1(-w))
    small = w < 1.0
    out = np.empty_like(w)
    out[small] = special.erfinv(v[small]) / v[small]
    big = ~small
    out[big] = special.erfcinv(one_minus_v2[big] / (1.0 + v[big])) / v[big]
    return out


def fit(lo, hi, centre, use_sqrt, deg):
    """Chebyshev-interpolate g on t in [lo,hi] (t = w or sqrt(w)); return monomial coeffs in (t-centre)."""
    def f(x):  # x in [-1,1]
        t = 0.5 * (hi - lo) * x + 0.5 * (hi + lo)
        w = t * t if use_sqrt else t
        return target(w)
    cheb = C.chebinterpolate(f, deg)
    # to monomial in x, then substitute x = (t - mid)/half, re-centre at `centre`
    mono_x = C.cheb2poly(cheb)
    mid, half = 0.5 * (hi + lo), 0.5 * (hi - lo)
    # p(t) = sum a_k ((t-centre) + (centre-mid))^k / half^k
    shift = P.Polynomial([centre - mid, 1.0]) / half
    poly = P.Polynomial([0.0])
    for k, a in enumerate(mono_x):
        poly = poly + a * shift ** k
    return poly.coef


def horner(coef, x):
    acc = np.full_like(x, coef[-1])
    for c in coef[-2::-1]:
        acc = acc * x + c
    return acc


SEGMENTS = [
    # name, t-range, centre, sqrt?, degree
    ("CENTRAL", (0.0, 6.25), 3.125, False, 24),
    ("MID", (2.5, 4.0), 3.25, True, 20),
    ("TAIL", (4.0, 4.75), 4.375, True, 14),
]

if __name__ == "__main__":
    for name, (lo, hi), centre, use_sqrt, deg in SEGMENTS:
        coef = fit(lo, hi, centre, use_sqrt, deg)
        t = np.linspace(lo + 1e-9 if lo == 0 else lo, hi, 200001)
        w = t * t if use_sqrt else t
        approx = horner(coef, t - centre)
        err = np.max(np.abs(approx / target(w) - 1.0))
        print(f"// {name}: t in [{lo},{hi}] centre {centre} sqrt={use_sqrt} deg {deg} max rel err {err:.3e}")
        print("{" + ", ".join(f"{c:.17e}" for c in coef[::-1]) + "},")
