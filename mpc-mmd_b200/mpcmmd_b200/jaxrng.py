"""Host-side (NumPy) restatement of the slice of the JAX 0.3.23 PRNG protocol the reference's SCENE GENERATOR uses
(synthetic_dynamic_obs/obs_data_generate_dynamic.py:112-116,136-148): `PRNGKey`, `split`, random bits, `normal` and
`choice(..., replace=False)`.  The solver's own draws happen on the device (csrc/drng.cuh); this module only prepares
episode inputs, like the reference's driver does before it calls `compute_cem_*`.

Protocol (jax/_src/prng.py, jax/_src/random.py of 0.3.23):
  * Threefry-2x32, 20 rounds; `bits(key, n)`: counters iota(n) padded to even, first / second half are the two words;
  * `split(key, m)` = `bits(key, 2m).reshape(m, 2)`;
  * `uniform` = bitcast((bits >> 9) | 0x3F800000) - 1, scaled, max(minval, .); `normal` = sqrt(2) * erf_inv(uniform(nextafter(-1, 0), 1)),
    erf_inv = XLA's float32 polynomial (Giles), w = -log1p(-x*x);
  * `choice(key, a, (m,), replace=False)` = `permutation(key, a)[:m]`; `permutation` = `_shuffle`: ceil(3 ln(n) / ln(2^32 - 1)) rounds of
    `key, sub = split(key); a = sort_key_val(bits(sub, n), a)` (stable).
"""
from __future__ import annotations

import numpy as np

U32 = np.uint32
F32 = np.float32
_R = ((13, 15, 26, 6), (17, 29, 16, 24))


def prng_key(seed: int):
    return (0, int(seed) & 0xFFFFFFFF)


def _rotl(x, r):
    return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32(key, x0, x1):
    """vectorised over the counter arrays x0, x1 (uint32)"""
    k0, k1 = U32(key[0]), U32(key[1])
    ks = (k0, k1, U32(k0 ^ k1 ^ U32(0x1BD11BDA)))
    with np.errstate(over="ignore"):
        x0 = (np.asarray(x0, U32) + ks[0]).astype(U32); x1 = (np.asarray(x1, U32) + ks[1]).astype(U32)
        for i in range(5):
            for r in _R[i & 1]:
                x0 = (x0 + x1).astype(U32); x1 = _rotl(x1, r).astype(U32); x1 = x1 ^ x0
            x0 = (x0 + ks[(i + 1) % 3]).astype(U32)
            x1 = (x1 + ks[(i + 2) % 3] + U32(i + 1)).astype(U32)
    return x0, x1


def bits(key, n: int):
    n = int(n)
    half = (n + (n & 1)) // 2
    c0 = np.arange(half, dtype=U32)
    c1 = (np.arange(half, dtype=np.uint64) + half)
    c1 = np.where(c1 < n, c1, 0).astype(U32)               # the pad element is a literal 0
    a, b = threefry2x32(key, c0, c1)
    return np.concatenate([a, b])[:n]


def split(key, m: int = 2):
    return bits(key, 2 * m).reshape(m, 2)


def uniform(key, n: int, minval=0.0, maxval=1.0):
    f = ((bits(key, n) >> U32(9)) | U32(0x3F800000)).view(F32) - F32(1.0)
    lo, hi = F32(minval), F32(maxval)
    return np.maximum(lo, (f * F32(hi - lo)).astype(F32) + lo).astype(F32)


_LT5 = [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503, -0.00417768164, 0.246640727, 1.50140941]
_GE5 = [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613, 0.00943887047, 1.00167406, 2.83297682]


def erfinv32(x):
    x = np.asarray(x, F32)
    w = (-np.log1p((-(x * x)).astype(F32))).astype(F32)
    small = w < F32(5.0)
    ws = np.where(small, w - F32(2.5), np.sqrt(np.maximum(w, F32(0))) - F32(3.0)).astype(F32)
    p = np.where(small, F32(_LT5[0]), F32(_GE5[0])).astype(F32)
    for a, b in zip(_LT5[1:], _GE5[1:]):
        p = (np.where(small, F32(a), F32(b)) + (p * ws).astype(F32)).astype(F32)
    out = (p * x).astype(F32)
    return np.where(np.abs(x) == F32(1.0), x * F32(np.inf), out).astype(F32)


def normal(key, n: int):
    lo = np.nextafter(F32(-1.0), F32(0.0))
    return (F32(1.4142135623730951) * erfinv32(uniform(key, n, lo, 1.0))).astype(F32)


def shuffle(key, x):
    x = np.asarray(x)
    rounds = int(np.ceil(3 * np.log(max(1, x.size)) / np.log(np.iinfo(np.uint32).max)))
    for _ in range(rounds):
        ks = split(key, 2)
        key, sub = (int(ks[0, 0]), int(ks[0, 1])), (int(ks[1, 0]), int(ks[1, 1]))
        x = x[np.argsort(bits(sub, x.size), kind="stable")]
    return x


def choice_no_replace(key, a, m: int):
    return shuffle(key, np.asarray(a))[:m]


def linspace32(start, stop, num):
    """jnp.linspace of 0.3.23 in float32: start*(1-step) + stop*step on iota/div, last point = stop"""
    start, stop = F32(start), F32(stop)
    div = num - 1
    step = (np.arange(div, dtype=F32) / F32(div)).astype(F32)
    out = ((start * (F32(1.0) - step)).astype(F32) + (stop * step).astype(F32)).astype(F32)
    return np.concatenate([out, np.array([stop], F32)])
