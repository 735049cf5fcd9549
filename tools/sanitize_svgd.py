"""Small SVGD steps through every kernel of csrc/svgd*.cu, for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_svgd.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysgmcmc_b200 import _native  # noqa: E402

dev = torch.device("cuda:0")
for impl in (1, 2):
    for n, D in ((10, 2), (132, 260), (256, 128), (400, 64), (260, 512), (520, 256)):   # incl. sliced K11, two-tile K14
        g = torch.Generator(device=dev).manual_seed(n)
        X = 1.0 + torch.randn((n, D), device=dev, generator=g)
        G = torch.randn((n, D), device=dev, generator=g)
        H = torch.zeros((n, D), device=dev)
        K = torch.empty((n, n), device=dev)
        ksum = torch.empty(n, device=dev)
        bw = torch.zeros(4, device=dev)
        scratch = _native.svgd_scratch(n, D, dev)
        Xs = torch.empty_like(X)
        s = _native.stream_ptr()
        _native.call("sgmcmc_set_svgd_tuning", impl)
        for _ in range(2):
            _native.call("sgmcmc_svgd_kernel_matrix_f32", _native.ptr(X), _native.ptr(K), _native.ptr(ksum),
                         _native.ptr(bw), _native.ptr(scratch), scratch.numel() * 8, n, D, s)
            _native.call("sgmcmc_svgd_update_f32", _native.ptr(X), _native.ptr(G), _native.ptr(H), _native.ptr(K),
                         _native.ptr(ksum), _native.ptr(bw), _native.ptr(Xs), n, D, 0.1, 0.9, 0.1, 1e-6, s)
        torch.cuda.synchronize()
        _native.call("sgmcmc_set_svgd_tuning", 0)
        assert torch.isfinite(X).all()
        print("impl", impl, "n", n, "D", D, "ok", float(bw[1]), flush=True)
# fused one-CTA SVGD on the built-in densities
for target, n, D in ((0, 10, 2), (0, 128, 2), (1, 33, 1)):
    X = torch.randn((n, D), device=dev)
    H = torch.zeros((n, D), device=dev)
    trace = torch.empty((5, n, D), device=dev)
    costs = torch.empty((5, n), device=dev)
    _native.call("sgmcmc_svgd_target_run_f32", target, _native.ptr(X), _native.ptr(H), _native.ptr(trace),
                 _native.ptr(costs), n, 10, 2, 0.1, 0.9, 0.1, 1e-6, _native.stream_ptr())
    torch.cuda.synchronize()
    assert torch.isfinite(trace).all()
    print("fused target", target, "n", n, "ok", flush=True)
big = torch.randn(300000, device=dev)
out = torch.empty(1, device=dev)
scratch = torch.zeros(512, dtype=torch.int64, device=dev)
_native.call("sgmcmc_median_f32", _native.ptr(big), big.numel(), _native.ptr(out), _native.ptr(scratch), _native.stream_ptr())
torch.cuda.synchronize()
print("median", float(out[0]), float(big.median()))
