"""Torch counterparts of the numerically relevant helpers of
pysgmcmc/tensor_utils.py (:87-104 vectorize, :153 unvectorize, :269 safe_divide,
:319-323 safe_sqrt).  Host-side conveniences for user cost functions; the CUDA
kernels carry their own copies of the same formulas (csrc/common.cuh).
"""
import torch


def vectorize(tensor):
    """Row-major flatten to ``(n_elements, 1)`` (tensor_utils.py:87-104)."""
    if not isinstance(tensor, torch.Tensor):
        raise ValueError(
            "Unsupported input to tensor_utils.vectorize: "
            "{value} is not a torch.Tensor subclass".format(value=tensor))
    return tensor.reshape(tensor.numel(), 1)


def unvectorize(tensor, original_shape):
    """Inverse of `vectorize` (tensor_utils.py:153)."""
    return tensor.reshape(tuple(original_shape))


def safe_divide(x, y, small_constant=1e-16, name=None):
    """``x / (y + (2 * sign(y) * c + c))`` (tensor_utils.py:269)."""
    y = torch.as_tensor(y)
    x = torch.as_tensor(x, dtype=y.dtype, device=y.device)
    return x / (y + (2.0 * torch.sign(y) * small_constant + small_constant))


def safe_sqrt(x, clip_value_min=0.0, clip_value_max=float("inf"), name=None):
    """``sqrt(clip(x, min, max))`` (tensor_utils.py:319-323)."""
    return torch.sqrt(torch.clamp(torch.as_tensor(x), min=clip_value_min, max=clip_value_max))


# ---- variable names --------------------------------------------------------------------
# tf.Variable objects carry a `name` the reference's diagnostics use as dictionary keys
# (pysgmcmc/diagnostics/sample_chains.py:166-176).  torch tensors have no writable `name`,
# so names live in a side table keyed by tensor identity.
import weakref

_NAMES = {}


def set_name(tensor, name):
    """Give `tensor` a variable name (returned by `get_name`, used by the diagnostics)."""
    key = id(tensor)
    _NAMES[key] = (weakref.ref(tensor, lambda _, key=key: _NAMES.pop(key, None)), str(name))
    return tensor


def get_name(tensor, default=None):
    entry = _NAMES.get(id(tensor))
    if entry is not None and entry[0]() is tensor:
        return entry[1]
    return default
