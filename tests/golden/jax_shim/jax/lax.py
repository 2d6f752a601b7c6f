"""jax.lax.scan / cond as Python control flow (see _core.py)."""
import numpy as np
from ._core import _tree_stack, wrap


def scan(f, init, xs, length=None):
    carry, ys = init, []
    n = len(xs) if xs is not None else length
    for i in range(n):
        x = wrap(np.asarray(xs)[i]) if xs is not None else None
        carry, y = f(carry, x)
        ys.append(y)
    return carry, _tree_stack(ys)


def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)
