"""Sweep of the launch variants of the BNN kernel K4 at the headline shape
(8192 chains, 1-50-50-50-1, batch 20): ms per launch, chain-steps/s, FP32 fraction.
    python tools/bench_k4.py [--chains 8192] [--iters 20]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysgmcmc_b200 import _native  # noqa: E402
from pysgmcmc_b200.models.bnn_cost import default_net_params  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=8192)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--variants", default="0,10")
ap.add_argument("--caps", default="0", help="grid caps (persistent CTAs) to time per variant")
args = ap.parse_args()
dev = torch.device("cuda:0")
C, N, B, D = args.chains, 20000, 20, 5252
g = torch.Generator(device="cpu").manual_seed(0)
X = torch.randn(N, 1, generator=g).to(dev)
y = torch.randn(N, generator=g).to(dev)
theta = torch.cat([p.reshape(C, -1) for p in default_net_params(1, n_chains=C, seed=1, device=dev)], dim=1).contiguous()
theta += 0.05 * torch.randn(theta.shape, device=dev)
starts = torch.randint(0, N - B + 1, (C,), device=dev, dtype=torch.int32)
cost, grad = torch.empty(C, device=dev), torch.empty_like(theta)
p, st = _native.ptr, _native.stream_ptr()


def launch():
    _native.call("sgmcmc_bnn_nll_grad_f32", p(theta), p(X), p(y), p(starts), p(cost), p(grad), None, C, 1, B,
                 float(B), N, st)


ref = None
for v, cap in [(int(x), int(c)) for x in args.variants.split(",") for c in args.caps.split(",")]:
    _native.call("sgmcmc_set_bnn_tuning", v)
    _native.call("sgmcmc_set_persistent_grids", 0, cap)
    grad.fill_(float("nan"))
    try:
        launch()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"variant": v, "error": str(e)[:200]}))
        continue
    if ref is None:
        ref = (cost.clone(), grad.clone())
    ok = bool(torch.allclose(cost, ref[0], rtol=1e-5) and
              ((grad - ref[1]).abs().amax(dim=1) <= 1e-5 * ref[1].abs().amax(dim=1)).all())
    for _ in range(3):
        launch()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.iters):
        launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    tf = 2 * 305000.0 * C / (ms * 1e9)
    print(json.dumps({"variant": v, "max_ctas": cap, "ms": round(ms, 4), "chain_steps_per_s": round(C / ms * 1e3),
                      "fp32_TFLOPs": round(tf, 2), "frac_of_74.4": round(tf / 74.45, 3),
                      "matches_variant0": ok}), flush=True)
_native.call("sgmcmc_set_bnn_tuning", 16)
_native.call("sgmcmc_set_persistent_grids", 0, 0)
