"""apply(op, *inputs): runs a shim subgraph (megengine.core.tensor.utils.subgraph) eagerly."""


def apply(op, *inputs):
    return op(*inputs)
