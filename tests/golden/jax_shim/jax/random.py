"""jax.random on top of the oracle's restatement of the JAX 0.3.23 PRNG protocol (oracle/oracle_rng.h)."""
import os
import sys

import numpy as np

from ._core import asarr, down, f32

_ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..", ".."))
if _ROOT not in sys.path:
    sys.path.insert(1, _ROOT)
from oracle import oracle as _O  # noqa: E402


def _k(key):
    k = np.asarray(key).astype(np.uint32).reshape(2)
    return int(k[0]), int(k[1])


def PRNGKey(seed):
    s = int(np.asarray(seed)) & 0xFFFFFFFF
    return asarr(np.array([0, s], np.uint32), np.uint32)


def split(key, num=2):
    return asarr(np.asarray(_O.split(_k(key), num), np.uint32), np.uint32)


def normal(key, shape=(), dtype=f32):
    n = int(np.prod(shape)) if len(shape) else 1
    return asarr(np.asarray(_O.normal(_k(key), n), f32).reshape(shape))


def uniform(key, shape=(), dtype=f32, minval=0.0, maxval=1.0):
    n = int(np.prod(shape)) if len(shape) else 1
    return asarr(np.asarray(_O.uniform(_k(key), n, minval, maxval), f32).reshape(shape))


def multivariate_normal(key, mean, cov, shape=None, dtype=f32, method="cholesky"):
    mean = np.asarray(down(np.asarray(mean)), f32); cov = np.asarray(down(np.asarray(cov)), f32)
    shape = tuple(shape) if shape is not None else ()
    d = mean.shape[-1]
    z = np.asarray(normal(key, shape + (d,)))
    L = np.linalg.cholesky(cov).astype(f32)
    return asarr(mean + np.einsum("ij,...j->...i", L, z).astype(f32))


def beta(key, a, b, shape=None, dtype=f32):
    a = np.asarray(down(np.asarray(a)), f32); b = np.asarray(down(np.asarray(b)), f32)
    shape = tuple(shape) if shape is not None else np.broadcast_shapes(a.shape, b.shape)
    aa = np.ascontiguousarray(np.broadcast_to(a, shape), f32).ravel(); bb = np.ascontiguousarray(np.broadcast_to(b, shape), f32).ravel()
    return asarr(np.asarray(_O.beta(_k(key), aa, bb), f32).reshape(shape))


def _random_bits(key, n):
    return np.asarray(_O.bits(_k(key), int(n)), np.uint32)


def _shuffle(key, x):
    """jax/_src/random.py::_shuffle of 0.3.23: ceil(3 ln(n) / ln(2^32 - 1)) rounds of sort-by-random-bits (stable)"""
    x = np.asarray(down(np.asarray(x)))
    rounds = int(np.ceil(3 * np.log(max(1, x.size)) / np.log(np.iinfo(np.uint32).max)))
    for _ in range(rounds):
        ks = np.asarray(split(key, 2))
        key, sub = ks[0], ks[1]
        x = x[np.argsort(_random_bits(sub, x.size), kind="stable")]
    return x


def permutation(key, x):
    if np.ndim(x) == 0:
        x = np.arange(int(x), dtype=np.int32)
    return asarr(_shuffle(key, x))


def choice(key, a, shape=(), replace=True, p=None):
    assert p is None and not replace, "only the form the reference's scene generator uses"
    n = int(np.prod(shape)) if len(shape) else 1
    return asarr(_shuffle(key, a)[:n].reshape(shape))
