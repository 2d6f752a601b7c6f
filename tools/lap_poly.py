#!/usr/bin/env python
"""Where the six coefficients of the contract's Laplace-kernel polynomial come from (oracle/oracle_math.h::om_lap, csrc/dmath.cuh::DM_LAP_C1..C6):
p(f) = 1 + c1 f + ... + c6 f^6 (constant term fixed at 1, so k(0) = 1 exactly) INTERPOLATES 2^f at the seven Chebyshev extrema of [-1/2, 1/2]
(0, +-1/4, +-sqrt(3)/4, +-1/2; the node 0 is satisfied by the constant term, the other six fix c1..c6).  Maximum relative error 3.9e-9 in exact arithmetic, within
a few per cent of the minimax error and trivially reproducible; the float32 Horner evaluation dominates the total error (0.93 ulp measured over every float of the
interval, tests/test_cpu_oracle.py through the C oracle).  Prints the coefficients and their float32 roundings.
usage: lap_poly.py"""
import numpy as np


def levelled_exp2(deg=6, a=-0.5, b=0.5, grid=400001):
    """(c1..c_deg, max relative error on a dense grid, residual of the linear solve)"""
    x = 0.5 * (a + b) + 0.5 * (b - a) * np.cos(np.pi * np.arange(deg + 1) / deg)[::-1]
    x = x[np.abs(x) > 1e-12]                                   # the node 0 holds by construction
    A = np.array([[xi ** j for j in range(1, deg + 1)] for xi in x])
    c = np.linalg.solve(A, 2.0 ** x - 1.0)
    xs = np.linspace(a, b, grid)
    err = (1.0 + sum(c[j - 1] * xs ** j for j in range(1, deg + 1)) - 2.0 ** xs) / 2.0 ** xs
    return c, float(np.max(np.abs(err))), float(np.max(np.abs(A @ c - (2.0 ** x - 1.0))))


if __name__ == "__main__":
    c, emax, res = levelled_exp2()
    print("max relative error on [-1/2, 1/2]: %.3e" % emax)
    for j, v in enumerate(c, 1):
        print("c%d = %.17g   float32: %.9g" % (j, v, np.float32(v)))
