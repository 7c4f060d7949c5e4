"""Stand-in for the three TensorFlow 1.14 symbols the reference touches (test infrastructure, see ../README.md)."""
import torch

__version__ = "1.14.0-shim"


class _Losses:
    @staticmethod
    def huber_loss(labels, predictions, weights=1.0, delta=1.0):
        """tensorflow/python/ops/losses/losses_impl.py huber_loss + compute_weighted_loss with the default reduction
        SUM_BY_NONZERO_WEIGHTS: sum(losses * weights) / (number of elements whose weight is non-zero)."""
        error = predictions - labels
        abs_error = error.abs()
        quadratic = torch.clamp(abs_error, max=delta)
        linear = abs_error - quadratic
        losses = 0.5 * quadratic * quadratic + delta * linear
        w = torch.as_tensor(weights, dtype=losses.dtype).expand_as(losses)
        num_present = (w != 0).to(losses.dtype).sum()
        total = (losses * w).sum()
        return torch.where(num_present > 0, total / torch.clamp(num_present, min=1.0), torch.zeros_like(total))


losses = _Losses()


def set_random_seed(seed):
    from keras import _core
    import numpy as np
    _core._INIT_RNG = np.random.RandomState(seed)


class ConfigProto:
    def __init__(self, *a, **k):
        class _G:
            allow_growth = False
        self.gpu_options = _G()


class Session:
    def __init__(self, *a, **k):
        pass
