"""Step time of sgmcmc_bnn_sghmc_run_f32 as one kernel per step (csrc/bnn_fused.cu) against
K4 then K1, burn-in and sampling phase, at the headline shape (8192 chains, D = 5252).
    python tools/bench_fused.py [--chains 8192] [--steps 300] [--caps 0,888,1184]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pysgmcmc_b200 import Session, _native  # noqa: E402
from pysgmcmc_b200.data_batches import DeviceBatchGenerator  # noqa: E402
from pysgmcmc_b200.models.bnn_cost import BayesianNeuralNetworkNLL, default_net_params  # noqa: E402
from pysgmcmc_b200.samplers import SGHMCSampler  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=8192)
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--caps", default="0")
ap.add_argument("--modes", default="1", help="fused modes to time: 1 = fused, 3 = fused without L2 prefetch, 4 = warp-specialised, 6 = the same without prefetch")
args = ap.parse_args()
dev = torch.device("cuda:0")
C, D = args.chains, 5252
X, y = bench.synthetic_sinc()


def timed(fused, cap, burn_in):
    _native.call("sgmcmc_set_bnn_fused", fused, cap)
    gen = DeviceBatchGenerator(20000, 20, n_chains=C, seed=1, device=dev)
    nll = BayesianNeuralNetworkNLL(20000, 20, X=X, y=y, starts_placeholder=gen.starts_placeholder, device=dev)
    s = SGHMCSampler(params=default_net_params(1, n_chains=C, seed=1, device=dev), cost_fun=nll, batch_generator=gen,
                     burn_in_steps=10 ** 9 if burn_in else 20, scale_grad=20000.0, seed=1,
                     session=Session(device=dev, n_chains=C, output="torch"))
    s.run(40, keep_every=40)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s.run(args.steps, keep_every=100)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.steps, s._theta.clone()


for burn_in in (True, False):
    ref = None
    for fused, cap in [(0, 0)] + [(int(m), int(c)) for m in args.modes.split(",") for c in args.caps.split(",")] + [(0, 0)]:
        ms, theta = timed(fused, cap, burn_in)
        if ref is None:
            ref = theta
        bytes_per_elem = (40 if burn_in else 20) if fused else (52 if burn_in else 32)
        print(json.dumps({"phase": "burn-in" if burn_in else "sampling", "fused": fused, "max_ctas": cap,
                          "ms_per_step": round(ms, 4), "chain_steps_per_s": round(C / ms * 1e3),
                          "hbm_GBps_algorithmic": round(bytes_per_elem * C * D / ms / 1e6, 1),
                          "bit_identical_to_unfused": bool(torch.equal(ref, theta))}), flush=True)
_native.call("sgmcmc_set_bnn_fused", 0, 0)
