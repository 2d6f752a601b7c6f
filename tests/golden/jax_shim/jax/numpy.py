"""jax.numpy for the functions the reference calls: NumPy with float32 defaults (see _core.py)."""
import builtins as _b
import numpy as _np

from ._core import Arr, asarr, down, wrap, f32

newaxis = None
nan = _np.nan
pi = _np.pi
float32 = _np.float32
int32 = _np.int32


def _w(fn):
    def g(*a, **k):
        return wrap(fn(*[down(x) for x in a], **{kk: down(v) for kk, v in k.items()}))
    g.__name__ = getattr(fn, "__name__", "fn")
    return g


for _n in ("dot", "hstack", "vstack", "dstack", "abs", "absolute", "sin", "cos", "tan", "arctan", "arctan2", "sum", "max", "min", "sqrt",
           "where", "maximum", "minimum", "exp", "tile", "argmin", "argmax", "isnan", "nan_to_num", "count_nonzero", "diff", "expand_dims",
           "reshape", "repeat", "outer", "median", "mean", "unwrap", "matmul", "cumsum", "log", "sign", "square", "power", "concatenate", "stack",
           "transpose", "squeeze", "ravel", "einsum", "floor", "ceil", "all", "any", "take", "diag", "trace", "std", "var", "prod", "amax", "amin",
           "logical_and", "logical_or", "logical_not", "isfinite", "zeros_like", "ones_like", "sort", "flip", "roll", "cross"):
    globals()[_n] = _w(getattr(_np, _n))


def asarray(x, dtype=None):
    return asarr(x, dtype)


array = asarray


def shape(x):
    return _np.shape(x)


def zeros(shape, dtype=f32):
    return asarr(_np.zeros(shape, dtype), dtype)


def ones(shape, dtype=f32):
    return asarr(_np.ones(shape, dtype), dtype)


def full(shape, v, dtype=None):
    return asarr(_np.full(shape, down(v)))


def eye(n, m=None, dtype=f32):
    return asarr(_np.eye(n, m, dtype=dtype), dtype)


def identity(n, dtype=f32):
    return asarr(_np.identity(n, dtype), dtype)


def arange(*a, **k):
    return asarr(_np.arange(*[down(x) for x in a], **k))


def linspace(start, stop, num=50, endpoint=True):
    """jnp.linspace (jax/_src/numpy/lax_numpy.py, 0.3.23): float32; start*(1-step) + stop*step on iota/div, last point = stop"""
    start, stop = f32(start), f32(stop)
    if not endpoint or num < 2:
        return asarr(_np.linspace(start, stop, num, endpoint=endpoint).astype(f32))
    div = num - 1
    step = (_np.arange(div, dtype=f32) / f32(div)).astype(f32)
    out = (start * (f32(1.0) - step) + stop * step).astype(f32)
    return asarr(_np.concatenate([out, _np.array([stop], f32)]))


def clip(x, a_min=None, a_max=None):
    x = down(_np.asarray(x))
    if a_min is not None:
        x = _np.maximum(x, down(_np.asarray(a_min)))
    if a_max is not None:
        x = _np.minimum(x, down(_np.asarray(a_max)))
    return wrap(x)


def argsort(x, axis=-1):
    return wrap(_np.argsort(down(_np.asarray(x)), axis=axis, kind="stable").astype(_np.int32))


def quantile(x, q, axis=None):
    """jnp.quantile default: linear interpolation at q*(n-1), computed in float32"""
    a = _np.sort(down(_np.asarray(x)).astype(f32), axis=None if axis is None else axis)
    assert axis is None and a.ndim == 1
    n = a.shape[0]
    qn = f32(q) * f32(n - 1)
    lo, hi = _np.floor(qn), _np.ceil(qn)
    hw = f32(qn - lo); lw = f32(1.0) - hw
    lo_i = int(_b.min(_b.max(lo, 0), n - 1)); hi_i = int(_b.min(_b.max(hi, 0), n - 1))
    return f32(a[lo_i] * lw + a[hi_i] * hw)


def cov(m):
    """jnp.cov(m): rows are variables, ddof = 1, float32"""
    X = down(_np.asarray(m)).astype(f32)
    Xc = (X - X.mean(axis=1, keepdims=True, dtype=f32)).astype(f32)
    return wrap((Xc @ Xc.T / f32(X.shape[1] - 1)).astype(f32))


class linalg:
    @staticmethod
    def solve(a, b):
        return wrap(_np.linalg.solve(down(_np.asarray(a)).astype(f32), down(_np.asarray(b)).astype(f32)).astype(f32))

    @staticmethod
    def norm(x, ord=None, axis=None):
        return wrap(_np.linalg.norm(down(_np.asarray(x)), ord=ord, axis=axis))

    @staticmethod
    def inv(a):
        return wrap(_np.linalg.inv(down(_np.asarray(a)).astype(f32)).astype(f32))

    @staticmethod
    def cholesky(a):
        return wrap(_np.linalg.cholesky(down(_np.asarray(a)).astype(f32)).astype(f32))
