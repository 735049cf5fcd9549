"""ctypes binding of libsgmcmc_b200.so (the C ABI in include/sgmcmc_b200.h).

There is NO fallback: if the library is missing or a call fails, this raises.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsgmcmc_b200.so")

SAMPLER_SGHMC, SAMPLER_SGLD, SAMPLER_RSGHMC = 0, 1, 2
TARGET_IDS = {"banana": 0, "gmm1": 1, "gmm2": 2, "gmm3": 3}


class NativeError(RuntimeError):
    pass


class Hyper(Structure):
    """sgmcmc_hyper_t"""
    _fields_ = [("epsilon", c_float), ("mdecay", c_float), ("scale_grad", c_float), ("A", c_float),
                ("mass", c_float), ("speed_of_light", c_float), ("D", c_float), ("Bhat", c_float)]


_P = c_void_p  # every device pointer / stream crosses the ABI as an integer address

# name -> argtypes (restype is int unless listed in _RESTYPES); must mirror include/sgmcmc_b200.h
SIGNATURES = {
    "sgmcmc_version": [],
    "sgmcmc_last_error": [],
    "sgmcmc_set_update_tuning": [c_int, c_int],
    "sgmcmc_set_bnn_tuning": [c_int],
    "sgmcmc_set_mlp_tuning": [c_int],
    "sgmcmc_set_persistent_grids": [c_int, c_int],
    "sgmcmc_set_bnn_chunk": [c_int64],
    "sgmcmc_set_bnn_pipeline": [c_int64, c_int],
    "sgmcmc_set_update_reverse": [c_int],
    "sgmcmc_set_bnn_fused": [c_int, c_int],
    "sgmcmc_launch_count": [],
    "sgmcmc_sghmc_step_f32": [_P] * 8 + [c_int64, c_float, c_float, c_float, c_int, c_int,
                                         c_uint64, c_uint64, c_uint64, _P],
    "sgmcmc_sghmc_step_f64": [_P] * 8 + [c_int64, c_double, c_double, c_double, c_int, c_int,
                                         c_uint64, c_uint64, c_uint64, _P],
    "sgmcmc_sgld_step_f32": [_P] * 7 + [c_int64, c_float, c_float, c_float, c_int, c_int,
                                        c_uint64, c_uint64, c_uint64, _P],
    "sgmcmc_sgld_step_f64": [_P] * 7 + [c_int64, c_double, c_double, c_double, c_int, c_int,
                                        c_uint64, c_uint64, c_uint64, _P],
    "sgmcmc_rsghmc_step_f32": [_P] * 4 + [c_int64] + [c_float] * 5 + [c_uint64, c_uint64, c_uint64, _P],
    "sgmcmc_rsghmc_step_f64": [_P] * 4 + [c_int64] + [c_double] * 5 + [c_uint64, c_uint64, c_uint64, _P],
    "sgmcmc_normal_fill_f32": [_P, c_int64, c_uint64, c_uint64, c_uint64, _P],
    "sgmcmc_target_chains_run_f32": [c_int, c_int] + [_P] * 9 + [c_int64, c_int64, c_int64, c_int, c_int64,
                                                                POINTER(Hyper), c_uint64, c_uint64, c_uint64, _P],
    "sgmcmc_mt19937_seed": [_P, _P, c_int64, _P],
    "sgmcmc_mt19937_starts": [_P, _P, c_int64, c_int64, c_uint32, _P],
    "sgmcmc_bnn_nll_grad_f32": [_P] * 7 + [c_int64, c_int, c_int, c_float, c_int64, _P],
    "sgmcmc_mlp_n_params": [POINTER(c_int), c_int],
    "sgmcmc_mlp_workspace_bytes": [POINTER(c_int), c_int, c_int64, c_int],
    "sgmcmc_mlp_nll_grad_f32": [_P] * 8 + [c_int64, c_int64, POINTER(c_int), c_int, c_int, c_float, c_int64, _P],
    "sgmcmc_mlp_predict_f32": [_P] * 4 + [c_int64, c_int64, POINTER(c_int), c_int, c_int64, _P],
    "sgmcmc_bnn_sghmc_run_f32": [_P] * 14 + [c_int64, c_int, c_int, c_float, c_int64, c_int64, c_int64,
                                             c_int, c_int64, c_float, c_float, c_float,
                                             c_uint64, c_uint64, c_uint64, _P],
    "sgmcmc_bnn_sghmc_run_resident_f32": [_P] * 15 + [c_int64, c_int, c_int, c_float, c_int64, c_int64, c_int64,
                                                      c_int, c_int64, c_float, c_float, c_float,
                                                      c_uint64, c_uint64, c_uint64, _P],
    "sgmcmc_bnn_resident_supported": [c_int, c_int],
    "sgmcmc_set_bnn_resident_overlap": [c_int],
    "sgmcmc_bnn_host_pipeline_create": [POINTER(_P), c_int64, c_int, c_int, c_int],
    "sgmcmc_bnn_host_pipeline_destroy": [_P],
    "sgmcmc_bnn_host_pipeline_step": [_P] * 13 + [c_int, c_int, c_float, c_int64, c_int, c_int, c_float, c_float,
                                                  c_float, c_uint64, c_uint64, c_uint64, _P, POINTER(c_int64)],
    "sgmcmc_bnn_host_pipeline_wait": [_P, c_int64],
    "sgmcmc_bnn_predict_f32": [_P, _P, _P, c_int64, c_int, c_int64, _P],
    "sgmcmc_chain_moments_f32": [_P, _P, c_int64, c_int64, c_int64, _P],
    "sgmcmc_variogram_f32": [_P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, _P],
    "sgmcmc_variogram_select_f32": [_P, _P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, _P],
    "sgmcmc_set_svgd_tuning": [c_int],
    "sgmcmc_median_f32": [_P, c_int64, _P, _P, _P],
    "sgmcmc_median_symmetric_f32": [_P, c_int64, _P, _P, _P],
    "sgmcmc_svgd_scratch_bytes": [c_int64, c_int64],
    "sgmcmc_svgd_kernel_matrix_f32": [_P] * 5 + [c_int64, c_int64, c_int64, _P],
    "sgmcmc_svgd_target_run_f32": [c_int] + [_P] * 4 + [c_int64, c_int64, c_int64, c_float, c_float, c_float,
                                                        c_float, _P],
    "sgmcmc_svgd_update_f32": [_P] * 7 + [c_int64, c_int64, c_float, c_float, c_float, c_float, _P],
}
_RESTYPES = {"sgmcmc_last_error": c_char_p, "sgmcmc_launch_count": c_int64, "sgmcmc_svgd_scratch_bytes": c_int64,
             "sgmcmc_mlp_n_params": c_int64, "sgmcmc_mlp_workspace_bytes": c_int64}

_lib = None


def load():
    """Load the shared library (once) and bind every symbol of the header."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            "libsgmcmc_b200.so is not built (%s). Run `python -m pysgmcmc_b200.build` "
            "(needs nvcc); there is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the .so does not export it
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    _lib = lib
    return lib


def call(name, *args):
    """Call an int-returning entry point; raise NativeError with the library's message on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.sgmcmc_last_error()
        raise NativeError("%s failed (%d): %s" % (name, rc, msg.decode() if msg else "?"))
    return rc


def svgd_scratch(n_particles, n_dims, device):
    """Scratch buffer for sgmcmc_svgd_kernel_matrix_f32 (int64 elements so that it is 16-byte aligned)."""
    import torch
    nbytes = int(load().sgmcmc_svgd_scratch_bytes(n_particles, n_dims))
    if nbytes < 0:
        raise NativeError("sgmcmc_svgd_scratch_bytes: unsupported size %d x %d" % (n_particles, n_dims))
    return torch.zeros((nbytes + 7) // 8, dtype=torch.int64, device=device)


def int_array(values):
    """A C `int[]` for the `widths` arguments."""
    values = [int(v) for v in values]
    return (c_int * len(values))(*values), len(values)


def launch_count():
    return int(load().sgmcmc_launch_count())


def ptr(t):
    """Device address of a torch tensor (None -> NULL). The tensor must be contiguous CUDA memory."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NativeError("expected a CUDA tensor; the engine has no CPU path")
    if not t.is_contiguous():
        raise NativeError("expected a contiguous tensor")
    return t.data_ptr()


def stream_ptr(stream=None, device_index=None):
    """cudaStream_t of `stream`, or of torch's current stream (on `device_index` if given)."""
    import torch
    if stream is not None:
        return stream.cuda_stream
    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)     # the handle without building a Stream object
    if raw is not None:
        return raw(torch.cuda.current_device() if device_index is None else device_index)
    return (torch.cuda.current_stream() if device_index is None else torch.cuda.current_stream(device_index)).cuda_stream
