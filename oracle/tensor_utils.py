"""Oracle restatement of the numerically relevant helpers of
``pysgmcmc/tensor_utils.py`` (test infrastructure only, see oracle/__init__.py).

All functions are dtype preserving: float32 in -> every intermediate is rounded
to float32 exactly like TensorFlow's op-by-op evaluation (no FMA contraction).
"""
import numpy as np


def _c(x, like):
    """Python scalar -> numpy scalar of `like`'s dtype (TF converts constants
    to the tensor dtype before the op runs)."""
    return np.asarray(like).dtype.type(x)


def safe_divide(x, y, small_constant=1e-16):
    """``x / (y + (2 * sign(y) * c + c))`` -- pysgmcmc/tensor_utils.py:269.

    y > 0 -> y + 3c ; y == 0 -> c ; y < 0 -> y - c.
    """
    y = np.asarray(y)
    x = np.asarray(x, dtype=y.dtype)
    c = _c(small_constant, y)
    two = _c(2.0, y)
    return x / (y + (two * np.sign(y) * c + c))


def safe_sqrt(x, clip_value_min=0.0, clip_value_max=float("inf")):
    """``sqrt(clip(x, 0, inf))`` -- pysgmcmc/tensor_utils.py:319-323."""
    x = np.asarray(x)
    return np.sqrt(np.clip(x, _c(clip_value_min, x), _c(clip_value_max, x)))


def vectorize(array):
    """Row-major flatten to ``(n_elements, 1)`` -- pysgmcmc/tensor_utils.py:87-98."""
    array = np.asarray(array)
    return array.reshape(int(np.prod(array.shape, dtype=np.int64)), 1)


def unvectorize(array, original_shape):
    """Inverse of `vectorize` -- pysgmcmc/tensor_utils.py:153."""
    return np.asarray(array).reshape(original_shape)
