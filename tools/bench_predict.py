"""Throughput of the predictive forward pass K10 (sgmcmc_bnn_predict_f32): n_nets stored
networks x n_points inputs.
    python tools/bench_predict.py
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysgmcmc_b200 import _native  # noqa: E402
from pysgmcmc_b200.models.bnn_cost import default_net_params  # noqa: E402

dev = torch.device("cuda:0")
for n_nets, n_points in ((100, 1000), (100, 100000), (8192, 1000)):
    theta = torch.cat([p.reshape(n_nets, -1) for p in default_net_params(1, n_chains=n_nets, seed=1, device=dev)],
                      dim=1).contiguous()
    X = torch.randn(n_points, 1, device=dev)
    out = torch.empty((n_nets, n_points, 2), device=dev)

    def launch():
        _native.call("sgmcmc_bnn_predict_f32", _native.ptr(theta), _native.ptr(X), _native.ptr(out), n_nets, 1,
                     n_points, _native.stream_ptr())
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    flop = 2.0 * (50 + 2500 + 2500 + 50) * n_nets * n_points
    print(json.dumps({"kernel": "bnn_predict_kernel", "n_nets": n_nets, "n_points": n_points, "ms": round(ms, 4),
                      "evaluations_per_s": round(n_nets * n_points / ms * 1e3), "fp32_TFLOPs": round(flop / ms / 1e9, 2),
                      "out_GBps": round(out.numel() * 4 / ms / 1e6, 1)}), flush=True)
