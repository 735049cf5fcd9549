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


def single_chain(output, fused=True):
    params = [torch.tensor(0.0, device=DEV), torch.tensor(6.0, device=DEV)]
    return SGHMCSampler(params=params, cost_fun=banana_nll, burn_in_steps=3000, seed=1,
                        session=Session(device=DEV, output=output, fused=fused))

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
