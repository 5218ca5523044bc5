"""Accuracy of the branch-free exp(min(x, 10)) used by the kernels (csrc/common.cuh trunc_exp), emulated in NumPy
(without FMA, so slightly pessimistic): maximum error 1 ulp on [-708, 10]."""
import math
import numpy as np


def exp_le10(x):
    x = np.maximum(np.minimum(x, 10.0), -708.0)
    shift = 6755399441055744.0
    t = (x * 1.4426950408889634 + shift) - shift
    r = x + t * (-6.93147180369123816490e-01)
    r = r + t * (-1.90821492927058770002e-10)
    p = np.full_like(x, 1.0 / math.factorial(13))
    for k in range(12, -1, -1):
        p = p * r + 1.0 / math.factorial(k)
    return p * np.exp2(t)


if __name__ == "__main__":
    x = np.concatenate([np.linspace(-708, 10, 2000001), np.random.default_rng(0).uniform(-12, 10, 2000000)])
    rel = np.abs(exp_le10(x) - np.exp(x)) / np.exp(x)
    print("max rel err %.3e (%.2f ulp), mean %.3f ulp" % (rel.max(), rel.max() / 2.22e-16, rel.mean() / 2.22e-16))
