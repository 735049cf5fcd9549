"""Resident BNN-SGHMC kernel (csrc/bnn_resident.cu) against K4 then K1 per step (sgmcmc_bnn_sghmc_run_f32) at
the benchmarked shapes (N = 20 000, minibatch 20, D = 5252): microseconds per step and chain-steps/s for
1 ... 8192 chains, burn-in and sampling phase, for blocks of steps and for one step per call.

    python tools/bench_resident.py [--chains 1,8,148,296,1184,8192] [--overlap 1,0] [--steps 64]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pysgmcmc_b200 import _native  # noqa: E402
from pysgmcmc_b200.data_batches import DeviceBatchGenerator  # noqa: E402
from pysgmcmc_b200.models.bnn_cost import default_net_params  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chains", default="1,8,148,296,1184,8192")
ap.add_argument("--overlap", default="1,0", help="update beside the gradient (1) / after it (0)")
ap.add_argument("--steps", type=int, default=64)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
dev = torch.device("cuda:0")
N, B, D = 20000, 20, 5252
rng = np.random.RandomState(1)
X = torch.tensor(rng.standard_normal((N, 1)).astype(np.float32), device=dev)
y = torch.tensor(rng.standard_normal(N).astype(np.float32), device=dev)
p = _native.ptr


def timed(fn, reps):
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for C in [int(c) for c in args.chains.split(",")]:
    S = args.steps if C <= 2048 else max(8, args.steps // 4)
    params = default_net_params(1, n_chains=C, seed=1, device=dev)
    theta = torch.cat([q.reshape(C, -1) for q in params], dim=1).contiguous()
    state = [theta, torch.zeros_like(theta)] + [torch.ones_like(theta) for _ in range(4)]
    grad, cost = torch.empty_like(theta), torch.empty(C, device=dev)
    gen = DeviceBatchGenerator(N, B, n_chains=C, seed=1, device=dev)
    starts = gen.next_block(S)
    st = _native.stream_ptr()

    def k4k1(n, burn):
        _native.call("sgmcmc_bnn_sghmc_run_f32", *[p(a) for a in state], p(X), p(y), p(starts), None, None, None,
                     p(grad), p(cost), C, 1, B, float(B), N, n, n if burn else 0, 0, 10 ** 9, 0.01, 0.05, float(N),
                     1, 0, 0, st)

    def resident(n, burn):
        _native.call("sgmcmc_bnn_sghmc_run_resident_f32", *[p(a) for a in state], p(X), p(y), p(starts), None, None,
                     None, None, p(cost), None, C, 1, B, float(B), N, n, n if burn else 0, 0, 10 ** 9, 0.01, 0.05,
                     float(N), 1, 0, 0, st)

    for burn in (True, False):
        ms = timed(lambda: k4k1(S, burn), args.reps)
        line = {"chains": C, "phase": "burn-in" if burn else "sampling", "steps_per_call": S,
                "k4_then_k1_us_per_step": round(1e3 * ms / S, 3), "k4_then_k1_chain_steps_per_s": round(C * S / ms * 1e3)}
        for ov in [int(t) for t in args.overlap.split(",")]:
            _native.call("sgmcmc_set_bnn_resident_overlap", ov)
            name = "resident" if ov else "resident_no_overlap"
            ms = timed(lambda: resident(S, burn), args.reps)
            line[name + "_us_per_step"] = round(1e3 * ms / S, 3)
            line[name + "_chain_steps_per_s"] = round(C * S / ms * 1e3)
            ms1 = timed(lambda: resident(1, burn), args.reps * 4)
            line[name + "_one_step_per_call_us"] = round(1e3 * ms1, 3)
        _native.call("sgmcmc_set_bnn_resident_overlap", 1)
        ms1 = timed(lambda: k4k1(1, burn), args.reps * 4)
        line["k4_then_k1_one_step_per_call_us"] = round(1e3 * ms1, 3)
        assert torch.isfinite(theta).all()
        print(json.dumps(line), flush=True)
    del state, theta, grad, gen, starts
    torch.cuda.empty_cache()
