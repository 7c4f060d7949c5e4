from ._core import Adam  # noqa: F401
