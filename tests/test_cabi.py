"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports
exactly the entry points include/sgmcmc_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

from pysgmcmc_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sgmcmc_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sgmcmc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path():
    names = declared_functions()
    for required in ("sgmcmc_sghmc_step_f32", "sgmcmc_sgld_step_f32", "sgmcmc_rsghmc_step_f32",
                     "sgmcmc_bnn_nll_grad_f32", "sgmcmc_bnn_sghmc_run_f32", "sgmcmc_mt19937_starts",
                     "sgmcmc_target_chains_run_f32", "sgmcmc_chain_moments_f32",
                     "sgmcmc_svgd_kernel_matrix_f32", "sgmcmc_svgd_update_f32", "sgmcmc_median_f32"):
        assert required in names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_native.LIB_PATH), "run `python -m pysgmcmc_b200.build` first"
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), "libsgmcmc_b200.so does not export %s" % name


def test_python_binding_covers_the_header():
    assert sorted(_native.SIGNATURES) == declared_functions()
    _native.load()


def test_version_and_error_reporting_without_gpu():
    lib = _native.load()
    assert lib.sgmcmc_version() >= 100
    assert lib.sgmcmc_set_update_tuning(100, 0) == -1          # SGMCMC_E_INVALID
    assert b"threads" in lib.sgmcmc_last_error()
    assert lib.sgmcmc_set_update_tuning(256, 1) == 0
    with pytest.raises(_native.NativeError):
        _native.call("sgmcmc_set_update_tuning", 0, 3)
    # argument validation happens before any CUDA call
    assert lib.sgmcmc_sghmc_step_f32(None, None, None, None, None, None, None, None, 16,
                                     0.01, 0.05, 1.0, 1, 0, 0, 0, 0, None) == -1
    assert lib.sgmcmc_sghmc_step_f32(None, None, None, None, None, None, None, None, 16,
                                     0.01, 0.05, 1.0, 1, 0, 0, 0, 2, None) == -1   # elem_offset % 4
    assert lib.sgmcmc_mt19937_starts(None, None, 4, 4, 10, None) == -1
    assert lib.sgmcmc_svgd_kernel_matrix_f32(None, None, None, None, None, 4096, 4, 2, None) == -1
    assert lib.sgmcmc_svgd_update_f32(None, None, None, None, None, None, None, 4, 2, 0.1, 0.9, 0.1, 1e-6, None) == -1
    assert lib.sgmcmc_median_f32(None, 0, None, None, None) == -1
    assert lib.sgmcmc_set_svgd_tuning(7) == -1 and lib.sgmcmc_set_svgd_tuning(0) == 0
    # the scratch size is host arithmetic: select state + mean/norms (+ slices of partial Gram tiles)
    assert lib.sgmcmc_svgd_scratch_bytes(10, 2) == 4096 + 4 * 12
    assert lib.sgmcmc_svgd_scratch_bytes(4096, 5252) == 4096 + 4 * (4096 + 5252)
    assert lib.sgmcmc_svgd_scratch_bytes(1024, 5252) > 4096 + 4 * (1024 + 5252)
    assert lib.sgmcmc_svgd_scratch_bytes(50000, 2) == -1


def test_generic_network_layout_is_host_arithmetic():
    """Parameter count and workspace of `get_net` specs (csrc/mlp.cu: make_mlp_layout): the layout of
    tf.trainable_variables() order W_1 b_1 ... W_{L+1} b_{L+1} rho, and the workspace that grows by the
    split canonical operand copies when wide layers go to the tensor cores."""
    lib = _native.load()
    wide, n = _native.int_array([1, 1000, 512, 512, 1])
    assert lib.sgmcmc_mlp_n_params(wide, n) == 777682            # BASELINE.json configs[4]'s network
    default, nd = _native.int_array([1, 50, 50, 50, 1])
    assert lib.sgmcmc_mlp_n_params(default, nd) == 5252          # get_default_net (bayesian_neural_network.py:28-69)
    bad, nb = _native.int_array([1, 50, 2])
    assert lib.sgmcmc_mlp_n_params(bad, nb) == -1                # the output width must be 1
    try:
        assert lib.sgmcmc_set_mlp_tuning(0) == 0
        ffma = lib.sgmcmc_mlp_workspace_bytes(wide, n, 3, 20)
        narrow_ffma = lib.sgmcmc_mlp_workspace_bytes(default, nd, 3, 20)
    finally:
        assert lib.sgmcmc_set_mlp_tuning(1) == 0
    umma = lib.sgmcmc_mlp_workspace_bytes(wide, n, 3, 20)
    # plain activations + gradients: 2 x 20 x (1000 + 512 + 512) floats per chain (+ the partial sums)
    assert 3 * 4 * 2 * 20 * 2024 <= ffma <= 3 * 4 * (2 * 20 * 2024 + 64)
    # + hi / lo planes [round_up(width, 32)][32]: H_1, H_2 for the forward GEMMs of layers 2, 3; dZ_2, dZ_3
    assert umma - ffma == 3 * 4 * 2 * 32 * (1024 + 512 + 512 + 512)
    assert lib.sgmcmc_mlp_workspace_bytes(default, nd, 3, 20) == narrow_ffma     # no layer qualifies
    assert lib.sgmcmc_mlp_workspace_bytes(wide, n, 3, 33) == -1                  # minibatches of up to 32 rows


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "pysgmcmc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_resident_kernel_shape_query_is_host_arithmetic():
    """Which BOHAMIANN shapes the resident kernel (csrc/bnn_resident.cu) can hold on one SM: 7 state arrays of
    ceil(D / 4) * 4 floats plus the activation buffers of the minibatch within 227 KB of shared memory."""
    lib = _native.load()
    assert lib.sgmcmc_bnn_resident_supported(1, 20) == 1          # the reference's configuration (D = 5252)
    assert lib.sgmcmc_bnn_resident_supported(2, 20) == 1          # D % 4 == 2: half-padded last element groups
    assert lib.sgmcmc_bnn_resident_supported(13, 20) == 1
    assert lib.sgmcmc_bnn_resident_supported(13, 32) == 0         # 32-row activation buffers no longer fit beside D = 5852
    assert lib.sgmcmc_bnn_resident_supported(1, 33) == 0          # minibatch rows are held in 32-row buffers
    assert lib.sgmcmc_bnn_resident_supported(64, 20) == 0         # D = 8402: the state alone is 235 KB
    assert lib.sgmcmc_bnn_resident_supported(0, 20) == 0 and lib.sgmcmc_bnn_resident_supported(1, 0) == 0
    # calls with a shape it cannot hold are refused before any device work
    rc = lib.sgmcmc_bnn_sghmc_run_resident_f32(*([None] * 15), 1, 64, 20, 20.0, 100, 1, 1, 0, 1, 0.01, 0.05, 100.0,
                                               0, 0, 0, None)
    assert rc != 0 and lib.sgmcmc_last_error()
