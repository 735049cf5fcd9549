"""Two-stream pipeline of the BNN-SGHMC step (sgmcmc_set_bnn_pipeline): K4 of chunk j+1 on the
main stream while K1 of chunk j runs on the library's update stream; step time vs chunk size
and ring depth at the headline shape, burn-in and sampling phase.
    python tools/bench_pipeline.py [--chains 8192] [--configs 0:0,1024:2,2048:2]
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
ap.add_argument("--steps", type=int, default=500)
ap.add_argument("--configs", default="0:0,512:2,1024:2,1024:3,2048:2,4096:2,0:0")
ap.add_argument("--phases", default="burn-in,sampling")
ap.add_argument("--k1-threads", type=int, default=256)
args = ap.parse_args()
dev = torch.device("cuda:0")
_native.call("sgmcmc_set_update_tuning", args.k1_threads, 1)
C = args.chains
X, y = bench.synthetic_sinc()
for phase in args.phases.split(","):
    ref = None
    for cfg in args.configs.split(","):
        chunk, ring = [int(v) for v in cfg.split(":")]
        _native.call("sgmcmc_set_bnn_pipeline", chunk, ring)
        gen = DeviceBatchGenerator(20000, 20, n_chains=C, seed=1, device=dev)
        nll = BayesianNeuralNetworkNLL(20000, 20, X=X, y=y, starts_placeholder=gen.starts_placeholder, device=dev)
        s = SGHMCSampler(params=default_net_params(1, n_chains=C, seed=1, device=dev), cost_fun=nll,
                         batch_generator=gen, burn_in_steps=10 ** 9 if phase == "burn-in" else 20,
                         scale_grad=20000.0, seed=1, session=Session(device=dev, n_chains=C, output="torch"))
        s.run(50, keep_every=50)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s.run(args.steps, keep_every=100)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        if ref is None:
            ref = s._theta.clone()
        print(json.dumps({"phase": phase, "k1_threads": args.k1_threads, "chunk_chains": chunk, "ring": ring, "ms_per_step": round(ms, 4),
                          "chain_steps_per_s": round(C / ms * 1e3),
                          "bit_identical": bool(torch.equal(ref, s._theta))}), flush=True)
_native.call("sgmcmc_set_bnn_pipeline", 0, 0)
_native.call("sgmcmc_set_update_tuning", 256, 1)
