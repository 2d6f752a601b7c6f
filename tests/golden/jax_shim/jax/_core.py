"""NumPy stand-in for the slice of the JAX 0.3.23 API the reference's optimizer uses (TEST INFRASTRUCTURE ONLY).

Purpose: JAX / jaxlib are not installable in this image (no network), so the reference cannot run on its real
runtime.  This package lets the reference's OWN, UNMODIFIED source files (synthetic_static_obs/optimizer/*.py,
compute_beta.py, kernel_computation.py) execute on NumPy float32 so that tests/golden/make_golden_ref.py can record
per-stage input/output vectors that pin oracle/.  What it reproduces from JAX: float32-by-default arithmetic (x64
disabled: every float64 operand is rounded to float32 before an operation), jit = identity, vmap / lax.scan /
lax.cond as Python loops, `.at[idx].set()`, stable argsort, and the PRNG protocol (through oracle/oracle_rng.h --
the Threefry/erfinv/gamma restatement pinned by the known answers in tests/test_cpu_oracle.py).  What it does NOT
reproduce: XLA's exact float32 kernels (matmul association, exp/sin/cos polynomials, LAPACK pivoting) -- results agree
with real JAX only to float32 round-off, which is why the stage tests use the 1e-4 tolerance north_star states.
"""
import numpy as np

f32 = np.float32


def down(x):
    """operand -> plain ndarray / scalar with float64 rounded to float32 (JAX with x64 disabled)"""
    if isinstance(x, np.ndarray):
        x = x.view(np.ndarray)
        if x.dtype == np.float64:
            x = x.astype(f32)
        elif x.dtype == np.int64:
            x = x.astype(np.int32)
        return x
    if isinstance(x, np.float64):
        return f32(x)
    if isinstance(x, (list, tuple)):
        return type(x)(down(v) for v in x)
    return x


def wrap(r):
    if isinstance(r, tuple):
        return tuple(wrap(v) for v in r)
    if isinstance(r, np.ndarray):
        if r.dtype == np.float64:
            r = r.astype(f32)
        return r.view(Arr)
    if isinstance(r, np.float64):
        return f32(r)
    return r


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, v):
        out = np.array(self.arr.view(np.ndarray), copy=True)
        out[down(self.idx) if isinstance(self.idx, np.ndarray) else self.idx] = down(np.asarray(v)) if isinstance(v, np.ndarray) else v
        return wrap(out)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class Arr(np.ndarray):
    """ndarray that behaves like a jax DeviceArray for the operations the reference uses"""
    __array_priority__ = 1000

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kw):
        if out is not None:
            kw["out"] = tuple(o.view(np.ndarray) if isinstance(o, np.ndarray) else o for o in out)
        return wrap(getattr(ufunc, method)(*[down(x) for x in inputs], **kw))

    @property
    def at(self):
        return _At(self)

    def block_until_ready(self):
        return self


def asarr(x, dtype=None):
    a = np.asarray(down(x) if isinstance(x, (np.ndarray, list, tuple)) else x)
    if dtype is not None:
        a = a.astype(dtype)
    elif a.dtype == np.float64:
        a = a.astype(f32)
    elif a.dtype == np.int64:
        a = a.astype(np.int32)
    return a.view(Arr)


def jit(fun=None, static_argnums=None, **kw):
    if fun is None:
        return lambda f: f
    return fun


def _tree_stack(outs):
    o0 = outs[0]
    if isinstance(o0, (tuple, list)):
        return tuple(_tree_stack([o[k] for o in outs]) for k in range(len(o0)))
    return asarr(np.stack([np.asarray(down(o)) if isinstance(o, np.ndarray) else np.asarray(o) for o in outs]))


def vmap(fun, in_axes=0, out_axes=0):
    assert out_axes in (0, (0,)) or out_axes == 0

    def mapped(*args):
        axes = tuple(in_axes) if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        assert len(axes) == len(args)
        n = next(np.shape(a)[ax] for a, ax in zip(args, axes) if ax is not None)
        outs = []
        for i in range(n):
            outs.append(fun(*[a if ax is None else wrap(np.take(np.asarray(a), i, axis=ax)) for a, ax in zip(args, axes)]))
        return _tree_stack(outs)
    return mapped
