"""Throughput of the diagnostics reductions K8 on a device trace [n_draws, C, D].
    python tools/bench_diagnostics.py [--chains 8192] [--draws 16]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysgmcmc_b200.diagnostics.sampler_diagnostics import (effective_n_from_trace, gelman_rubin_from_trace,  # noqa: E402
                                                           local_moment_sums, local_variogram_sums)

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=8192)
ap.add_argument("--draws", type=int, default=16)
ap.add_argument("--dims", type=int, default=5252)
args = ap.parse_args()
n, C, D = args.draws, args.chains, args.dims
trace = torch.randn((n, C, D), device="cuda:0")
gb = trace.numel() * 4 / 1e9


def timed(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ms = timed(lambda: local_moment_sums(trace))
print(json.dumps({"kernel": "chain_moments_kernel", "trace_GB": round(gb, 2), "ms": round(ms, 3),
                  "GBps": round(gb / ms * 1e3, 1)}))
lags = 16
ms = timed(lambda: local_variogram_sums(trace, 1, lags))
elems = trace.numel()
print(json.dumps({"kernel": "variogram_window_kernel (lags 1..16 in one pass)", "lags": lags, "ms": round(ms, 3),
                  "ms_per_lag": round(ms / lags, 3), "trace_GBps": round(gb / ms * 1e3, 1),
                  "fp64_TFLOPs": round(3.0 * lags * elems / ms / 1e9, 2)}))
ms = timed(lambda: local_variogram_sums(trace, 17, 4), iters=2)
print(json.dumps({"kernel": "variogram_kernel (general lags, scalar)", "lags": 4, "ms": round(ms, 3),
                  "ms_per_lag": round(ms / 4, 3)}))
t0 = time.perf_counter()
rhat = gelman_rubin_from_trace(trace)
ess = effective_n_from_trace(trace)
torch.cuda.synchronize()
print(json.dumps({"call": "R-hat + ESS of %d chains x %d draws x %d dims" % (C, n, D),
                  "seconds": round(time.perf_counter() - t0, 3), "rhat_mean": float(rhat.mean()),
                  "ess_mean": float(ess.mean())}))
