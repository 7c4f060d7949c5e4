"""Stand-in for Keras 2.2.4 (test infrastructure, see ../README.md)."""
__version__ = "2.2.4-shim"
from . import _core
from . import backend, layers, models, optimizers, activations  # noqa: E402,F401
