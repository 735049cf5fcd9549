"""Step time of sgmcmc_bnn_sghmc_run_f32 (K7-fed K4 + K1) vs the chunk size (L2 reuse of the gradient).
    python tools/bench_chunk.py
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pysgmcmc_b200 import Session, _native  # noqa: E402
from pysgmcmc_b200.data_batches import DeviceBatchGenerator  # noqa: E402
from pysgmcmc_b200.models.bnn_cost import BayesianNeuralNetworkNLL, default_net_params  # noqa: E402
from pysgmcmc_b200.samplers import SGHMCSampler  # noqa: E402

dev = torch.device("cuda:0")
C = 8192
X, y = bench.synthetic_sinc()
ref = None
for chunk in (0, 0, 4096, 2960, 2048, 0, 4096, 5920):
    _native.call("sgmcmc_set_bnn_chunk", chunk)
    gen = DeviceBatchGenerator(20000, 20, n_chains=C, seed=1, device=dev)
    nll = BayesianNeuralNetworkNLL(20000, 20, X=X, y=y, starts_placeholder=gen.starts_placeholder, device=dev)
    s = SGHMCSampler(params=default_net_params(1, n_chains=C, seed=1, device=dev), cost_fun=nll, batch_generator=gen,
                     burn_in_steps=10 ** 9, scale_grad=20000.0, seed=1,
                     session=Session(device=dev, n_chains=C, output="torch"))
    s.run(50, keep_every=50)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s.run(500, keep_every=100)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 500
    if ref is None:
        ref = s._theta.clone()
    same = bool(torch.equal(ref, s._theta))
    print(json.dumps({"chunk": chunk, "ms_per_step": round(ms, 4), "chain_steps_per_s": round(C / ms * 1e3),
                      "bit_identical_to_unchunked": same}), flush=True)
_native.call("sgmcmc_set_bnn_chunk", 0)
