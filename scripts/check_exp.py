"""Accuracy of the branch-free, table-driven exp(min(x, 10)) used by the kernels (csrc/common.cuh trunc_exp), emulated
in NumPy (without FMA, so slightly pessimistic) against a 50-digit reference: maximum error 0.9 ulp on [-700, 10]."""
import decimal
import math

import numpy as np

decimal.getcontext().prec = 50
LN2 = decimal.Decimal(2).ln()
_c = LN2 / 32
_m, _e = math.frexp(float(_c))
HI = math.ldexp(math.floor(_m * 2 ** 32), _e - 32)          # ln2 / 32, upper 32 bits: t * HI is exact for |t| < 2^21
LO = float(_c - decimal.Decimal(HI))
INV = float(32 / LN2)
T = np.array([float(decimal.Decimal(2) ** (decimal.Decimal(j) / 32)) for j in range(32)])
COEF = [1.0 / math.factorial(k) for k in range(7)]


def exp_le10(x):
    x = np.minimum(x, 10.0)
    shift = 6755399441055744.0
    t = (x * INV + shift) - shift
    mi = t.astype(np.int64)
    r = (x - t * HI) - t * LO
    p = np.full_like(x, COEF[6])
    for k in range(5, 0, -1):
        p = p * r + COEF[k]
    p = p * r
    tj = T[mi & 31]
    return (tj + tj * p) * np.exp2(np.maximum(mi >> 5, -1022).astype(float))


if __name__ == "__main__":
    print("HI %.20e LO %.20e INV %.20e" % (HI, LO, INV))
    rng = np.random.default_rng(0)
    x = np.concatenate([np.linspace(-700, 10, 20001), rng.uniform(-12, 10, 20000), rng.uniform(-1, 1, 5000)])
    ref = [decimal.Decimal(float(v)).exp() for v in x]
    err = np.array([abs((decimal.Decimal(float(g)) - r) / r) for g, r in zip(exp_le10(x), ref)], dtype=float)
    print("max rel err %.3e (%.2f ulp), mean %.3f ulp" % (err.max(), err.max() / 2.22e-16, err.mean() / 2.22e-16))
