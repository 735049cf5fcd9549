"""Shared-memory race check of the one-CTA SVGD kernels (fused target run, single-CTA median):
    compute-sanitizer --tool racecheck python tools/racecheck_svgd_small.py
(the tcgen05 kernels are excluded: racecheck does not model the async proxy / mbarrier hand-off)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysgmcmc_b200 import _native  # noqa: E402

dev = torch.device("cuda:0")
for target, n, D in ((0, 10, 2), (0, 37, 2), (2, 16, 1)):
    X = torch.randn((n, D), device=dev)
    H = torch.zeros((n, D), device=dev)
    trace = torch.empty((3, n, D), device=dev)
    costs = torch.empty((3, n), device=dev)
    _native.call("sgmcmc_svgd_target_run_f32", target, _native.ptr(X), _native.ptr(H), _native.ptr(trace),
                 _native.ptr(costs), n, 6, 2, 0.1, 0.9, 0.1, 1e-6, _native.stream_ptr())
    torch.cuda.synchronize()
    print("fused", target, n, bool(torch.isfinite(trace).all()), flush=True)
vals = torch.randn(5000, device=dev)
out = torch.empty(1, device=dev)
scratch = torch.zeros(512, dtype=torch.int64, device=dev)
_native.call("sgmcmc_median_f32", _native.ptr(vals), vals.numel(), _native.ptr(out), _native.ptr(scratch),
             _native.stream_ptr())
torch.cuda.synchronize()
print("median", float(out[0]), float(vals.median()), flush=True)
