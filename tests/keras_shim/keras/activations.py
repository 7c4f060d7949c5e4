from ._core import activations as _a

relu, linear, get = _a.relu, _a.linear, _a.get
