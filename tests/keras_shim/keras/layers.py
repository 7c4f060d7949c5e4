from ._core import Input, Dense, Layer, Lambda, add, Concatenate, activations  # noqa: F401
from ._core import concatenate_layer as concatenate  # noqa: F401
