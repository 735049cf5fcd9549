"""Throughput of the non-headline configurations of BASELINE.json (configs 0-2), GPU vs the
NumPy oracle on one host core.  One JSON line per measurement.
    python tools/bench_configs.py
config 0: SGHMC single chain on the 2-D banana, 10 000 samples, through next(sampler)
          (reference semantics: one host round trip per sample) and through sampler.run.
config 1: SGLD on the 1-D Gaussian mixtures, 4096 chains, sampler.run (one K6 launch).
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import samplers as osamplers, targets as otargets  # noqa: E402
from pysgmcmc_b200 import Session  # noqa: E402
from pysgmcmc_b200.diagnostics.objective_functions import (banana_log_likelihood, gmm1_log_likelihood,  # noqa: E402
                                                           to_negative_log_likelihood)
from pysgmcmc_b200.samplers import SGHMCSampler, SGLDSampler  # noqa: E402

DEV = "cuda:0"


def emit(**kw):
    print(json.dumps(kw), flush=True)


def gpu_timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    return time.perf_counter() - t0


# ---- config 0: single banana chain, 10 000 samples -----------------------------------------
n = 10000
banana_nll = to_negative_log_likelihood(banana_log_likelihood)


def single_chain(output, fused=True, prefetch=0):
    params = [torch.tensor(0.0, device=DEV), torch.tensor(6.0, device=DEV)]
    return SGHMCSampler(params=params, cost_fun=banana_nll, burn_in_steps=3000, seed=1,
                        session=Session(device=DEV, output=output, fused=fused, prefetch=prefetch))

s = single_chain("numpy")
for _ in range(100):
    next(s)
dt = gpu_timed(lambda: [next(s) for _ in range(n)])
emit(config="banana SGHMC, 1 chain, next(sampler) -> numpy (reference semantics)", steps=n, seconds=dt,
     chain_steps_per_s=n / dt)
s = single_chain("torch")
for _ in range(100):
    next(s)
dt = gpu_timed(lambda: [next(s) for _ in range(n)])
emit(config="banana SGHMC, 1 chain, next(sampler) -> device tensors (no host sync)", steps=n, seconds=dt,
     chain_steps_per_s=n / dt)
for out in ("torch", "numpy"):
    for S in (64, 1024):
        s = single_chain(out, prefetch=S)
        for _ in range(2 * S):
            next(s)
        dt = gpu_timed(lambda: [next(s) for _ in range(n)])
        emit(config="banana SGHMC, 1 chain, next(sampler) -> %s, Session(prefetch=%d): steps computed %d at a "
                    "time by one K6 launch" % (out, S, S), steps=n, seconds=dt, chain_steps_per_s=n / dt)
s = single_chain("torch", fused=False)
for _ in range(100):
    next(s)
dt = gpu_timed(lambda: [next(s) for _ in range(2000)])
emit(config="banana SGHMC, 1 chain, generic path (torch autograd + K1)", steps=2000, seconds=dt,
     chain_steps_per_s=2000 / dt)
s = single_chain("torch")
s.run(100)
dt = gpu_timed(lambda: s.run(n))
emit(config="banana SGHMC, 1 chain, sampler.run(10000) (one K6 launch)", steps=n, seconds=dt,
     chain_steps_per_s=n / dt)
chain = osamplers.OracleChain("sghmc", np.array([[0.0, 6.0]], dtype=np.float32), otargets.banana_cost_and_grad)
rng = np.random.RandomState(0)
t0 = time.perf_counter()
for _ in range(n):
    chain.next(rng.standard_normal((1, 2)).astype(np.float32))
dt = time.perf_counter() - t0
emit(config="banana SGHMC, 1 chain, NumPy oracle on one host core", steps=n, seconds=dt, chain_steps_per_s=n / dt)

# ---- config 1: 4096 SGLD chains on gmm1 ------------------------------------------------------
C, steps = 4096, 20000
s = SGLDSampler(params=[torch.zeros(C, device=DEV)], cost_fun=to_negative_log_likelihood(gmm1_log_likelihood),
                seed=1, session=Session(device=DEV, n_chains=C, output="torch"))
s.run(100)
dt = gpu_timed(lambda: s.run(steps, keep_every=steps))
emit(config="gmm1 SGLD, 4096 chains, sampler.run (K6)", steps=steps, seconds=dt, chain_steps_per_s=C * steps / dt)
for C2 in (65536, 1 << 20):
    s = SGLDSampler(params=[torch.zeros(C2, device=DEV)], cost_fun=to_negative_log_likelihood(gmm1_log_likelihood),
                    seed=1, session=Session(device=DEV, n_chains=C2, output="torch"))
    s.run(100)
    dt = gpu_timed(lambda: s.run(2000, keep_every=2000))
    emit(config="gmm1 SGLD, %d chains, sampler.run (K6)" % C2, steps=2000, seconds=dt,
         chain_steps_per_s=C2 * 2000 / dt)
chain = osamplers.OracleChain("sgld", np.zeros((C, 1), dtype=np.float32), otargets.cost_and_grad("gmm1"))
t0 = time.perf_counter()
for _ in range(500):
    chain.next(rng.standard_normal((C, 1)).astype(np.float32))
dt = time.perf_counter() - t0
emit(config="gmm1 SGLD, 4096 chains, NumPy oracle on one host core", steps=500, seconds=dt,
     chain_steps_per_s=C * 500 / dt)

# ---- config 2: ONE BOHAMIANN chain (N = 20 000, batch 20) -----------------------------------
from oracle import bnn as obnn  # noqa: E402
from pysgmcmc_b200.data_batches import DeviceBatchGenerator  # noqa: E402
from pysgmcmc_b200.models.bnn_cost import BayesianNeuralNetworkNLL, default_net_params  # noqa: E402

N, B = 20000, 20
rs = np.random.RandomState(1)
Xd = rs.uniform(0, 1, size=(N, 1))
yd = np.sinc(Xd * 10 - 5).sum(axis=1)
Xd = ((Xd - Xd.mean(0)) / Xd.std(0)).astype(np.float32)
yd = ((yd - yd.mean()) / yd.std()).astype(np.float32)


def bnn_chain(C, prefetch=0, output="torch", resident=True, burn_in_steps=1000):
    gen = DeviceBatchGenerator(N, B, n_chains=C, seed=1, device=DEV)
    nll = BayesianNeuralNetworkNLL(N, B, X=Xd, y=yd, starts_placeholder=gen.starts_placeholder, device=DEV)
    s = SGHMCSampler(params=default_net_params(1, n_chains=C, seed=1, device=DEV), cost_fun=nll,
                     batch_generator=gen, burn_in_steps=burn_in_steps, scale_grad=float(N), seed=1,
                     session=Session(device=DEV, n_chains=C, output=output, prefetch=prefetch))
    # the resident kernel (the default up to 2 chains per SM) or K4 then K1 per step, at any chain count
    s.RESIDENT_MAX_CHAINS = 10 ** 9 if resident else 0
    return s


for C in (1, 8, 148, 592):
    for resident in (True, False):
        kernels = "resident kernel (the default up to 2 chains per SM)" if resident else "K4 then K1 per step"
        s = bnn_chain(C, resident=resident)
        s.run(1100, keep_every=10 ** 9)
        dt = gpu_timed(lambda: s.run(5000, keep_every=100))
        emit(config="BNN-SGHMC (1-50-50-50-1, N=20000, batch 20), %d chain(s), sampler.run(5000) after burn-in, %s"
             % (C, kernels), steps=5000, seconds=dt, chain_steps_per_s=C * 5000 / dt, us_per_step=1e6 * dt / 5000)
        s = bnn_chain(C, resident=resident, burn_in_steps=10 ** 9)
        s.run(200, keep_every=10 ** 9)
        dt = gpu_timed(lambda: s.run(3000, keep_every=100))
        emit(config="BNN-SGHMC (1-50-50-50-1, N=20000, batch 20), %d chain(s), sampler.run(3000) in burn-in, %s"
             % (C, kernels), steps=3000, seconds=dt, chain_steps_per_s=C * 3000 / dt, us_per_step=1e6 * dt / 3000)
s = bnn_chain(1)
for _ in range(1100):
    next(s)
dt = gpu_timed(lambda: [next(s) for _ in range(3000)])
emit(config="BNN-SGHMC, 1 chain, next(sampler) -> device tensors, step by step", steps=3000, seconds=dt,
     chain_steps_per_s=3000 / dt, us_per_step=1e6 * dt / 3000)
for out in ("torch", "numpy"):
    s = bnn_chain(1, prefetch=256, output=out)
    for _ in range(1300):
        next(s)
    dt = gpu_timed(lambda: [next(s) for _ in range(5000)])
    emit(config="BNN-SGHMC, 1 chain, next(sampler) -> %s, Session(prefetch=256)" % out, steps=5000, seconds=dt,
         chain_steps_per_s=5000 / dt, us_per_step=1e6 * dt / 5000)
theta = obnn.init_theta(1, seed=1, dtype=np.float32)
holder = {}


def _cg(th):
    Xb, yb = obnn.gather_minibatch(Xd, yd, holder["s"], B)
    c, g, _ = obnn.nll_and_grad(th, Xb, yb, n_examples=N)
    return c, g


chain = osamplers.OracleChain("sghmc", theta, _cg, epsilon=0.01, burn_in_steps=0, scale_grad=float(N))
t0 = time.perf_counter()
for _ in range(2000):
    holder["s"] = rng.randint(0, N - B + 1, size=1)
    chain.next(rng.standard_normal((1, 5252)).astype(np.float32))
dt = time.perf_counter() - t0
emit(config="BNN-SGHMC, 1 chain, NumPy oracle on one host core", steps=2000, seconds=dt, chain_steps_per_s=2000 / dt,
     us_per_step=1e6 * dt / 2000)
