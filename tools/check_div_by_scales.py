#!/usr/bin/env python
"""Exhaustive check of csrc/postnet.cuh::div_by_scales (the multi-scale post-network kernel's v / n_scales):
    c = RN(1/n);  q0 = RN(v*c);  r = fma(-q0, n, v);  q = fma(r, c, q0)
equals the correctly rounded float32 quotient for EVERY float32 significand (two binades: scaling by 2 is exact away from
over/underflow, which the kernel excludes) and n = 2..9.  numpy emulation: the residual r is exact in float64 (asserted),
the final FMA is evaluated in extended precision; mismatches against the float64 quotient and a random sample are
re-checked with exact rational arithmetic.   usage: python tools/check_div_by_scales.py [n ...]"""
import sys
from fractions import Fraction
import numpy as np


def exact_quotient(v, d):
    fr = Fraction(float(v)) / d
    f = np.float32(float(fr))
    cands = [np.nextafter(f, np.float32(-np.inf)), f, np.nextafter(f, np.float32(np.inf))]
    return np.float32(min(cands, key=lambda t: (abs(Fraction(float(t)) - fr), int(np.float32(t).view(np.uint32)) & 1)))


def check(d, binades=(127, 128), sample=2000):
    m = np.arange(1 << 23, dtype=np.uint32)
    wrong = 0
    for e in binades:
        v = ((np.uint32(e) << 23) | m).view(np.float32)
        c = np.float32(1.0) / np.float32(d)
        q0 = v * c
        r = v.astype(np.float64) - q0.astype(np.float64) * float(d)
        assert np.all(r == r.astype(np.float32).astype(np.float64)), "the residual must be an exact float32"
        q = ((r.astype(np.longdouble) * np.longdouble(c)) + q0.astype(np.longdouble)).astype(np.float32)
        ref = (v.astype(np.float64) / float(d)).astype(np.float32)
        doubtful = list(np.nonzero(q != ref)[0][:1000]) + list(np.random.default_rng(d).integers(0, 1 << 23, sample))
        wrong += sum(1 for i in doubtful if exact_quotient(v[i], d) != q[i]) + max(0, int((q != ref).sum()) - 1000)
    return wrong


if __name__ == "__main__":
    ns = [int(a) for a in sys.argv[1:]] or list(range(2, 10))
    bad = {n: check(n) for n in ns}
    print("wrong quotients per n:", bad)
    sys.exit(1 if any(bad.values()) else 0)
