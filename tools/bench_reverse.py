"""L2 reuse between K4 and K1: K1 walking its arrays from the end (sgmcmc_set_update_reverse)
vs from the start, whole BNN-SGHMC step at the headline shape, burn-in and sampling phase.
    python tools/bench_reverse.py
"""
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

dev = torch.device("cuda:0")
C = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
X, y = bench.synthetic_sinc()
for burn_in in (True, False):
    ref = None
    for rev in (1, 3, 1, 3, 0):
        _native.call("sgmcmc_set_update_reverse", rev)
        gen = DeviceBatchGenerator(20000, 20, n_chains=C, seed=1, device=dev)
        nll = BayesianNeuralNetworkNLL(20000, 20, X=X, y=y, starts_placeholder=gen.starts_placeholder, device=dev)
        s = SGHMCSampler(params=default_net_params(1, n_chains=C, seed=1, device=dev), cost_fun=nll,
                         batch_generator=gen, burn_in_steps=10 ** 9 if burn_in else 20, scale_grad=20000.0, seed=1,
                         session=Session(device=dev, n_chains=C, output="torch"))
        s.run(50, keep_every=50)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s.run(1000, keep_every=100)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 1000
        if ref is None:
            ref = s._theta.clone()
        print(json.dumps({"phase": "burn-in" if burn_in else "sampling", "k1_reverse": rev, "ms_per_step": round(ms, 4),
                          "chain_steps_per_s": round(C / ms * 1e3),
                          "bit_identical": bool(torch.equal(ref, s._theta))}), flush=True)
_native.call("sgmcmc_set_update_reverse", 1)
