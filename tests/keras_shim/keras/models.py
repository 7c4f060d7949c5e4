from ._core import Model  # noqa: F401
