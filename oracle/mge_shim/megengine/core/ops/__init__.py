from . import builtin  # noqa: F401
