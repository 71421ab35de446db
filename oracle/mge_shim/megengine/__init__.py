"""numpy-backed stand-in for megengine (TEST INFRASTRUCTURE ONLY, see ../README.md)."""
import numpy as np

from .tensor import Tensor, tensor  # noqa: F401
from . import functional  # noqa: F401
from . import module  # noqa: F401
from . import random  # noqa: F401


class _Device:
    @staticmethod
    def is_cuda_available():
        return False


device = _Device()


def _full_sync():
    pass
