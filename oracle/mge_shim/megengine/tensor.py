import numpy as np

f32 = np.float32


def _norm(a):
    """MegEngine keeps fp32 / int32 / bool / uint8; numpy's promotions to 64-bit are folded back."""
    a = np.asarray(a)
    if a.dtype == np.float64:
        return a.astype(np.float32)
    if a.dtype == np.int64:
        return a.astype(np.int32)
    return a


def _raw(x):
    if isinstance(x, Tensor):
        return x._a
    if isinstance(x, (list, tuple)):
        if any(isinstance(v, Tensor) for v in x):
            return np.array([_raw(v) for v in x])
    return x


def _idx(i):
    if isinstance(i, Tensor):
        a = i._a
        return a.astype(np.int64) if a.dtype.kind in "iu" else a
    if isinstance(i, tuple):
        return tuple(_idx(v) for v in i)
    if isinstance(i, list):
        return [_idx(v) for v in i]
    return i


class Tensor:
    """Value semantics are numpy's: in-place ops mutate the shared buffer (like mge.Tensor's _reset)."""

    __array_priority__ = 1000

    def __init__(self, data=None, dtype=None, device=None):
        if isinstance(data, Tensor):
            a = data._a
        else:
            a = np.array(_raw(data))
            if dtype is None and a.dtype.kind == "f":
                a = a.astype(np.float32)  # python floats / float64 arrays become fp32 (mge default)
        if dtype is not None:
            a = a.astype(np.dtype(dtype))
        self._a = _norm(a)

    # ---- meta
    @property
    def shape(self):
        return self._a.shape

    @property
    def ndim(self):
        return self._a.ndim

    @property
    def dtype(self):
        return self._a.dtype

    @property
    def size(self):
        return self._a.size

    @property
    def device(self):
        return "cpux"

    def numpy(self):
        return self._a

    def tolist(self):
        return self._a.tolist()

    def item(self):
        return self._a.item()

    def detach(self):
        return self

    def astype(self, dt):
        a = self._a
        dt = np.dtype(dt)
        if dt.kind in "iu" and a.dtype.kind == "f":
            with np.errstate(invalid="ignore"):
                a = np.where(np.isfinite(a), a, -2.0 ** 31)  # x86 cvttss2si on NaN/inf
                a = np.clip(np.trunc(a), -2.0 ** 31, 2.0 ** 31 - 1)
        return Tensor(a.astype(dt))

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return Tensor(self._a.reshape(shape))

    def flatten(self):
        return Tensor(self._a.reshape(-1))

    def transpose(self, *axes):
        if len(axes) == 1 and isinstance(axes[0], (tuple, list)):
            axes = tuple(axes[0])
        return Tensor(self._a.transpose(axes))

    def __len__(self):
        return len(self._a)

    def __bool__(self):
        return bool(self._a)

    def __int__(self):
        return int(self._a)

    def __float__(self):
        return float(self._a)

    def __index__(self):
        return int(self._a)

    def __repr__(self):
        return "Tensor(%r)" % (self._a,)

    def __iter__(self):
        for i in range(len(self._a)):
            yield Tensor(self._a[i])

    # ---- indexing
    def __getitem__(self, i):
        return Tensor(self._a[_idx(i)])

    def __setitem__(self, i, v):
        self._a[_idx(i)] = _raw(v)

    # ---- reductions
    def max(self, axis=None, keepdims=False):
        return Tensor(self._a.max(axis=axis, keepdims=keepdims))

    def min(self, axis=None, keepdims=False):
        return Tensor(self._a.min(axis=axis, keepdims=keepdims))

    def sum(self, axis=None, keepdims=False):
        a = self._a
        if a.dtype == np.bool_:
            return Tensor(a.sum(axis=axis, keepdims=keepdims).astype(np.int32))
        return Tensor(a.sum(axis=axis, keepdims=keepdims, dtype=a.dtype))

    def mean(self, axis=None, keepdims=False):
        return Tensor(self._a.mean(axis=axis, keepdims=keepdims, dtype=np.float32))


def _bin(name, rname=None, inplace=None):
    def op(self, other):
        with np.errstate(all="ignore"):
            return Tensor(getattr(np, name)(self._a, _raw(other)))

    def rop(self, other):
        with np.errstate(all="ignore"):
            return Tensor(getattr(np, name)(_raw(other), self._a))

    def iop(self, other):
        with np.errstate(all="ignore"):
            res = _norm(getattr(np, name)(self._a, _raw(other)))
        if res.shape == self._a.shape and res.dtype == self._a.dtype:
            self._a[...] = res
        else:
            self._a = res
        return self

    return op, rop, iop


for _py, _np in (("add", "add"), ("sub", "subtract"), ("mul", "multiply"), ("truediv", "true_divide"),
                 ("floordiv", "floor_divide"), ("mod", "mod"), ("pow", "power"), ("and", "bitwise_and"),
                 ("or", "bitwise_or")):
    _o, _r, _i = _bin(_np)
    setattr(Tensor, "__%s__" % _py, _o)
    setattr(Tensor, "__r%s__" % _py, _r)
    setattr(Tensor, "__i%s__" % _py, _i)
for _py, _np in (("lt", "less"), ("le", "less_equal"), ("gt", "greater"), ("ge", "greater_equal"),
                 ("eq", "equal"), ("ne", "not_equal")):
    setattr(Tensor, "__%s__" % _py, _bin(_np)[0])
Tensor.__hash__ = lambda self: id(self)
Tensor.__neg__ = lambda self: Tensor(-self._a)
Tensor.__invert__ = lambda self: Tensor(~self._a)


tensor = Tensor  # megengine.tensor is the class itself (the reference does isinstance(x, mge.tensor))
