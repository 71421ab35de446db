"""megengine.functional subset used by BaseDet's box-op files; leaf semantics = oracle ASSUMED-1..7."""
import math

import numpy as np

from ..tensor import Tensor, _raw, f32
from . import nn, vision  # noqa: F401


def _t(x):
    return x if isinstance(x, Tensor) else Tensor(x)


def _a(x):
    return _raw(x) if not isinstance(x, Tensor) else x._a


def _as_like(v, ref):
    v = np.asarray(v)
    if v.dtype.kind == "f" or np.asarray(ref).dtype.kind == "f":
        return v.astype(np.float32) if np.asarray(ref).dtype.kind == "f" else v
    return v


def maximum(x, y):  # MegDNN Elemwise MAX: x > y ? x : y (ASSUMED-1)
    x, y = np.broadcast_arrays(_a(_t(x)), _as_like(_a(y), _a(_t(x))))
    return Tensor(np.where(x > y, x, y))


def minimum(x, y):
    x, y = np.broadcast_arrays(_a(_t(x)), _as_like(_a(y), _a(_t(x))))
    return Tensor(np.where(x < y, x, y))


def clip(x, lower=None, upper=None):
    out = _t(x)
    if lower is not None:
        out = maximum(out, lower)
    if upper is not None:
        out = minimum(out, upper)
    return out


def expand_dims(x, axis):
    return Tensor(np.expand_dims(_a(x), axis))


def squeeze(x, axis=None):
    return Tensor(np.squeeze(_a(x), axis))


def concat(inps, axis=0):
    return Tensor(np.concatenate([_a(_t(i)) for i in inps], axis=axis))


def stack(inps, axis=0):
    return Tensor(np.stack([_a(_t(i)) for i in inps], axis=axis))


def _unary(fn):
    def f(x):
        with np.errstate(all="ignore"):
            return Tensor(fn(_a(_t(x))).astype(np.float32))
    return f


log = _unary(np.log)
exp = _unary(np.exp)
sqrt = _unary(np.sqrt)
floor = _unary(np.floor)
abs = _unary(np.abs)


def sigmoid(x):
    a = _a(_t(x))
    with np.errstate(over="ignore"):
        return Tensor((f32(1) / (f32(1) + np.exp(-a).astype(f32)).astype(f32)).astype(f32))


def pow(x, y):
    return Tensor(np.power(_a(_t(x)), _as_like(_a(y), _a(_t(x)))))


def argmax(x, axis=None, keepdims=False):  # first index among equal maxima (ASSUMED-2)
    return Tensor(np.argmax(_a(x), axis=axis).astype(np.int32))


def argmin(x, axis=None, keepdims=False):
    return Tensor(np.argmin(_a(x), axis=axis).astype(np.int32))


def max(x, axis=None, keepdims=False):
    return _t(x).max(axis=axis, keepdims=keepdims)


def min(x, axis=None, keepdims=False):
    return _t(x).min(axis=axis, keepdims=keepdims)


def sum(x, axis=None, keepdims=False):
    return _t(x).sum(axis=axis, keepdims=keepdims)


def full(shape, value, dtype="float32", device=None):
    return Tensor(np.full(shape, _a(value), dtype=np.dtype(dtype)))


def full_like(x, value):
    a = _a(x)
    return Tensor(np.full(a.shape, _a(value), dtype=a.dtype))


def zeros(shape, dtype="float32", device=None):
    return Tensor(np.zeros(shape, dtype=np.dtype(dtype)))


def zeros_like(x):
    return Tensor(np.zeros_like(_a(x)))


def ones(shape, dtype="float32", device=None):
    return Tensor(np.ones(shape, dtype=np.dtype(dtype)))


def arange(start=0, stop=None, step=1, dtype="float32", device=None):  # ASSUMED-7
    if stop is None:
        start, stop = 0, start
    start, stop, step = float(_a(start)), float(_a(stop)), float(_a(step))
    num = int(math.ceil((stop - start) / step))
    num = num if num > 0 else 0
    out = (start + np.arange(num, dtype=np.float64) * step).astype(np.float32)
    return Tensor(out.astype(np.dtype(dtype)))


def broadcast_to(x, shape):
    return Tensor(np.broadcast_to(_a(x), tuple(int(s) for s in shape)).copy())


def repeat(x, repeats, axis=None):
    return Tensor(np.repeat(_a(x), repeats, axis=axis))


def flatten(x, start_axis=0, end_axis=-1):
    a = _a(x)
    end_axis = end_axis % a.ndim if a.ndim else 0
    shape = a.shape[:start_axis] + (-1,) + a.shape[end_axis + 1:]
    return Tensor(a.reshape(shape))


def transpose(x, pattern):
    return Tensor(_a(x).transpose(pattern))


def cond_take(mask, x):  # ASSUMED-4: values + ascending flat int32 indices
    m = _a(mask).reshape(-1).astype(bool)
    idx = np.flatnonzero(m).astype(np.int32)
    return Tensor(_a(x).reshape(-1)[idx]), Tensor(idx)


def argsort(x, descending=False):  # stable (ASSUMED-3)
    a = _a(x)
    key = -a if descending else a
    return Tensor(np.argsort(key, kind="stable").astype(np.int32))


def sort(x, descending=False):
    idx = argsort(x, descending)
    return Tensor(_a(x)[idx._a]), idx


def topk(x, k, descending=False, kth_only=False, no_sort=False):  # ASSUMED-3, k clamped to n
    a = _a(x)
    if int(k) < 0:  # ASSUMED-10: MegDNN TopK takes a signed k, negative = the |k| LARGEST (sampling.py:27 relies on it)
        k, descending = -int(k), not descending
    key = -a if descending else a
    order = np.argsort(key, axis=-1, kind="stable")[..., : builtins_min(int(k), a.shape[-1])].astype(np.int32)
    return Tensor(np.take_along_axis(a, order.astype(np.int64), -1)), Tensor(order)


def gather(x, axis, index):
    return Tensor(np.take_along_axis(_a(x), _a(index).astype(np.int64), axis))


def scatter(x, axis, index, source):
    out = _a(x).copy()
    np.put_along_axis(out, _a(index).astype(np.int64), _a(source).astype(out.dtype), axis)
    return Tensor(out)


def _seq_sum(a, axis, keepdims):
    """fp32 sum accumulated sequentially in index order (ASSUMED-8: the order of MegDNN's reduce is not known)."""
    a = np.asarray(a, np.float32)
    out = np.cumsum(a, axis=axis, dtype=np.float32).take(-1, axis=axis)
    return np.expand_dims(out, axis) if keepdims else out


def mean(x, axis=None, keepdims=False):  # ASSUMED-8: sequential fp32 sum, then one divide by n
    a = _a(_t(x))
    if axis is None:
        a, axis = a.reshape(-1), 0
    return Tensor((_seq_sum(a, axis, keepdims) / f32(a.shape[axis])).astype(f32))


def var(x, axis=None, keepdims=False):  # population variance: mean((x - mean(x)) ** 2)
    a = _a(_t(x))
    if axis is None:
        a, axis = a.reshape(-1), 0
    m = (_seq_sum(a, axis, True) / f32(a.shape[axis])).astype(f32)
    d = (a - m).astype(f32)
    return Tensor((_seq_sum((d * d).astype(f32), axis, keepdims) / f32(a.shape[axis])).astype(f32))


def std(x, axis=None, keepdims=False):
    return Tensor(np.sqrt(var(x, axis, keepdims)._a).astype(f32))


def indexing_one_hot(src, index, axis=1, keepdims=False):
    return Tensor(np.take_along_axis(_a(src), np.expand_dims(_a(index).astype(np.int64), axis), axis).squeeze(axis))


def one_hot(inp, num_classes):
    """F.one_hot: int labels (...,) -> int32 (..., num_classes)."""
    a = _a(inp).astype(np.int64)
    return Tensor((a[..., None] == np.arange(num_classes)).astype(np.int32))


def logsigmoid(x):
    """F.logsigmoid = -softplus(-x), evaluated in the numerically stable form MegEngine uses:
    min(x, 0) - log1p(exp(-|x|)) (ASSUMED-12; fp32 throughout)."""
    a = _a(_t(x)).astype(np.float32)
    with np.errstate(all="ignore"):
        return Tensor((np.minimum(a, f32(0)) - np.log1p(np.exp(-np.abs(a)).astype(f32)).astype(f32)).astype(f32))


def where(mask, x, y):
    return Tensor(np.where(_a(mask), _a(x), _a(y)))


import builtins as _b  # noqa: E402

builtins_min = _b.min
