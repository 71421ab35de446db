"""subgraph(name, dtype, device, nr_inputs, gopt_level): eager interpreter of the builder function.

Every f(op, *args) is ONE fp32 elementwise / indexing step, evaluated in the order the reference source lists
them (structures/op_patch.py:36-76), which is exactly what the un-fused MegEngine CPU path executes."""
import numpy as np

from ...tensor import Tensor
from ..ops import builtin

f32 = np.float32
_STR = {"max": "max", "min": "min", "-": "sub", "+": "add", "*": "mul", "/": "true_div", "pow": "pow"}


def _elem(mode, a, b):
    a, b = np.broadcast_arrays(a, b)
    with np.errstate(all="ignore"):
        if mode == "add":
            return (a + b).astype(a.dtype)
        if mode == "sub":
            return (a - b).astype(a.dtype)
        if mode == "mul":
            return (a * b).astype(a.dtype)
        if mode == "true_div":
            return (a / b).astype(a.dtype)
        if mode == "max":
            return np.where(a > b, a, b)  # MegDNN: x > y ? x : y
        if mode == "min":
            return np.where(a < b, a, b)
        if mode == "pow":
            if np.all(b == 2):
                return (a * a).astype(a.dtype)
            if np.all(b == 0.5):
                return np.sqrt(a).astype(a.dtype)
            return np.power(a, b).astype(a.dtype)
    raise NotImplementedError(mode)


def _f(op, *args):
    arrs = [x._a if isinstance(x, Tensor) else np.asarray(x) for x in args]
    if isinstance(op, str):
        return Tensor(_elem(_STR[op], arrs[0], arrs[1]))
    if isinstance(op, builtin.Elemwise):
        return Tensor(_elem(op.mode.value if hasattr(op.mode, "value") else str(op.mode).lower(), arrs[0], arrs[1]))
    if isinstance(op, builtin.AddAxis):
        out = arrs[0]
        for ax in op.axis:
            out = np.expand_dims(out, ax)
        return Tensor(out)
    if isinstance(op, builtin.RemoveAxis):
        return Tensor(np.squeeze(arrs[0], axis=tuple(op.axis)))
    if isinstance(op, builtin.Reduce):
        assert op.mode == "sum"
        return Tensor(arrs[0].sum(axis=op.axis, keepdims=True, dtype=arrs[0].dtype))
    if isinstance(op, builtin.Subtensor):
        src = arrs[0]
        rest = [int(v) for v in arrs[1:]]
        index = [slice(None)] * src.ndim
        pos = 0
        for axis, has_b, has_e, has_s, has_i in op.items:
            b = e = s = None
            if has_b:
                b = rest[pos]; pos += 1
            if has_e:
                e = rest[pos]; pos += 1
            if has_s:
                s = rest[pos]; pos += 1
            if has_i:
                index[axis] = rest[pos]; pos += 1
            else:
                index[axis] = slice(b, e, s)
        return Tensor(src[tuple(index)])
    raise NotImplementedError(type(op))


def _c(value, dtype="float32", device=None):
    return Tensor(np.array(value, dtype=np.dtype(dtype)))


def subgraph(name, dtype, device, nr_inputs, gopt_level=None):
    def decorator(func):
        def make_op():
            def run(*inputs):
                outputs, _ = func(list(inputs), _f, _c)
                return outputs
            return run
        return make_op
    return decorator
