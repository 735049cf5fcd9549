"""Torch counterparts of the numerically relevant helpers of
pysgmcmc/tensor_utils.py (:87-104 vectorize, :153 unvectorize, :160-208 median,
:269 safe_divide, :319-323 safe_sqrt, :326-419 pdist, :422-577 squareform).
Host-side conveniences for user cost functions; the CUDA kernels carry their own
copies of the same formulas (csrc/common.cuh, csrc/svgd.cu).
"""
import numpy as np
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


def median(tensor):
    """Median of all entries of `tensor` (tensor_utils.py:160-208): the middle value, or the
    mean of the two middle values of an even count.  float32 CUDA tensors go through the
    exact radix select of the SVGD path (K12, ``sgmcmc_median_f32``); anything else is
    sorted with torch."""
    tensor = torch.as_tensor(tensor)
    flat = tensor.reshape(-1)
    if flat.is_cuda and flat.dtype == torch.float32 and flat.numel() > 0:
        from . import _native
        flat = flat.contiguous()
        out = torch.empty(1, dtype=torch.float32, device=flat.device)
        scratch = torch.zeros(512, dtype=torch.int64, device=flat.device)
        with torch.cuda.device(flat.device):
            _native.call("sgmcmc_median_f32", _native.ptr(flat), flat.numel(), _native.ptr(out),
                         _native.ptr(scratch), _native.stream_ptr())
        return out[0]
    values = torch.sort(flat, descending=True).values
    mid_index = flat.numel() // 2
    if flat.numel() % 2 == 1:
        return values[mid_index]
    return (values[mid_index - 1] + values[mid_index]) / 2


def pdist(tensor, metric="euclidean"):
    """Condensed vector of the pairwise euclidean distances of the rows of a 2-d tensor
    (tensor_utils.py:326-419; equals ``scipy.spatial.distance.pdist``)."""
    assert isinstance(tensor, torch.Tensor), "tensor_utils.pdist: Input must be a `torch.Tensor` instance."
    if tensor.dim() != 2:
        raise ValueError('tensor_utils.pdist: A 2-d tensor must be passed.')
    if metric != "euclidean":
        raise NotImplementedError(
            "tensor_utils.pdist: "
            "Metric '{metric}' currently not supported!".format(metric=metric))
    m = tensor.shape[0]
    i, j = torch.triu_indices(m, m, offset=1, device=tensor.device)
    return torch.linalg.vector_norm(tensor[i] - tensor[j], dim=1)


def squareform(tensor):
    """Condensed distance vector -> symmetric distance matrix (tensor_utils.py:422-577;
    vector input only, like the reference)."""
    assert isinstance(tensor, torch.Tensor), "tensor_utils.squareform: Input must be a `torch.Tensor` instance."
    if tensor.dim() != 1:
        raise NotImplementedError("tensor_utils.squareform: Only 1-d (vector) input is supported!")
    n_elements = tensor.shape[0]
    if n_elements == 0:
        return torch.zeros((1, 1), dtype=tensor.dtype, device=tensor.device)
    dimension = int(np.ceil(np.sqrt(n_elements * 2)))
    if dimension * (dimension - 1) != n_elements * 2:
        raise ValueError("Incompatible vector size. It must be a binomial "
                         "coefficient n choose 2 for some integer n >=2.")
    upper_triangular = torch.zeros((dimension, dimension), dtype=tensor.dtype, device=tensor.device)
    i, j = torch.triu_indices(dimension, dimension, offset=1, device=tensor.device)
    upper_triangular[i, j] = tensor
    return upper_triangular + upper_triangular.t()


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
