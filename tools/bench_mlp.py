"""Rates of the layer kernels (csrc/mlp.cu) on the wide 1000-512-512 BNN of BASELINE.json
configs[4] (D = 777 682) and on other architectures: cost + gradient per chain-step, its algorithmic
HBM traffic (theta in + gradient out = 8 B per parameter) against the measured copy peak, the FP32
rate, and the whole BNN-SGHMC step (layer kernels + K1) next to it.

    python tools/bench_mlp.py [--hidden 1000,512,512] [--chains 1,8,64,256] [--iters 10]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pysgmcmc_b200 import Session, _native  # noqa: E402
from pysgmcmc_b200.data_batches import DeviceBatchGenerator  # noqa: E402
from pysgmcmc_b200.models import MLPNet  # noqa: E402
from pysgmcmc_b200.models.bnn_cost import BayesianNeuralNetworkNLL  # noqa: E402
from pysgmcmc_b200.samplers import SGHMCSampler  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--hidden", default="1000,512,512")
ap.add_argument("--chains", default="1,8,64,256")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--n-in", type=int, default=1)
ap.add_argument("--layers", default="tcgen05,ffma", help="implementations of the wide layers to time")
args = ap.parse_args()
dev = torch.device("cuda:0")
hidden = tuple(int(h) for h in args.hidden.split(","))
net = MLPNet(hidden)
N, B = 20000, 20
D = net.n_parameters(args.n_in)
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6650.0
rng = np.random.RandomState(1)
X = rng.standard_normal((N, args.n_in)).astype(np.float32)
y = rng.standard_normal(N).astype(np.float32)
widths = net.widths(args.n_in)
flop = 6.0 * B * sum(widths[l] * widths[l + 1] for l in range(len(widths) - 1))


def timed(fn, iters):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for impl, C in [(m, int(c)) for m in args.layers.split(",") for c in args.chains.split(",")]:
    _native.call("sgmcmc_set_mlp_tuning", 1 if impl == "tcgen05" else 0)
    gen = DeviceBatchGenerator(N, B, n_chains=C, seed=1, device=dev)
    nll = BayesianNeuralNetworkNLL(N, B, X=X, y=y, starts_placeholder=gen.starts_placeholder, device=dev, net=net)
    params = net.init_params(args.n_in, n_chains=C, seed=1, device=dev)
    sampler = SGHMCSampler(params=params, cost_fun=nll, batch_generator=gen, burn_in_steps=10 ** 9,
                           scale_grad=float(N), seed=1, session=Session(device=dev, n_chains=C, output="torch"))
    next(sampler)
    theta, grad = sampler._theta, sampler._grad
    ms_k4 = timed(lambda: nll.native_cost_and_grad(theta, grad), args.iters)
    ms_step = timed(lambda: sampler.run(1, keep_every=10 ** 9), args.iters)
    gbs = 8.0 * C * D / ms_k4 / 1e6
    print(json.dumps({
        "net": "%d-%s-1" % (args.n_in, "-".join(map(str, hidden))), "wide_layers": impl, "params_per_chain": D, "chains": C, "batch": B,
        "k4_layer_kernels_ms": round(ms_k4, 4), "k4_us_per_chain": round(1e3 * ms_k4 / C, 3),
        "k4_algorithmic_GBps": round(gbs, 1), "k4_frac_of_measured_hbm_peak": round(gbs / peak, 4),
        "k4_fp32_TFLOPs": round(flop * C / ms_k4 / 1e9, 2), "k4_frac_of_fp32_peak_74.4": round(flop * C / ms_k4 / 1e9 / 74.45, 3),
        "step_ms": round(ms_step, 4), "chain_steps_per_s": round(C / ms_step * 1e3),
        "step_state_GBps_52B_per_param": round(52.0 * C * D / ms_step / 1e6, 1)}), flush=True)
    del sampler, nll, gen, params
    torch.cuda.empty_cache()
_native.call("sgmcmc_set_mlp_tuning", 1)
