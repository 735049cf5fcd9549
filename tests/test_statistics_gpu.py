"""North-star correctness part (3): long-run statistics of the CUDA samplers agree with the
reference algorithm's (the CPU oracle run with independent noise), and with the analytic
target where the discretisation bias allows.

Chains are independent, so the states of C chains after burn-in are C i.i.d. draws from
the sampler's stationary law: two-sample KS tests (GPU chains vs oracle chains, same
hyper-parameters, different random streams) at significance 1e-3, plus mean errors
normalised by the standard error.  Sizes keep the oracle to a few seconds.
"""
import numpy as np
import pytest
import torch
from scipy import stats

from oracle import bnn as obnn, samplers as osamplers, targets as otargets
from pysgmcmc_b200 import Session
from pysgmcmc_b200.data_batches import DeviceBatchGenerator
from pysgmcmc_b200.diagnostics.objective_functions import (banana_log_likelihood, gmm1_log_likelihood,
                                                           gmm2_log_likelihood, gmm3_log_likelihood,
                                                           to_negative_log_likelihood)
from pysgmcmc_b200.models.bnn_cost import BayesianNeuralNetworkNLL, parameter_shapes
from pysgmcmc_b200.samplers import RelativisticSGHMCSampler, SGHMCSampler, SGLDSampler
from pysgmcmc_b200.stepsize_schedules import ConstantStepsizeSchedule

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
LOGLIK = {"banana": banana_log_likelihood, "gmm1": gmm1_log_likelihood, "gmm2": gmm2_log_likelihood,
          "gmm3": gmm3_log_likelihood}
CLS = {"sghmc": SGHMCSampler, "sgld": SGLDSampler, "rsghmc": RelativisticSGHMCSampler}


def gpu_final_states(method, target, C, steps, seed, **hyper):
    D = 2 if target == "banana" else 1
    params = [torch.zeros(C, device=DEV) for _ in range(D)]
    if target == "banana":
        params[1] += 6.0                                    # tests/samplers/sampler_testing.py:16-17
    s = CLS[method](params=params, cost_fun=to_negative_log_likelihood(LOGLIK[target]), seed=seed,
                    session=Session(device=DEV, n_chains=C, output="torch"), **hyper)
    assert s._native_target == target
    s.run(steps, keep_every=steps)
    return s._theta.cpu().numpy()


def oracle_final_states(method, target, C, steps, seed, epsilon=None, **hyper):
    rng = np.random.RandomState(seed)
    D = 2 if target == "banana" else 1
    theta0 = np.zeros((C, D), dtype=np.float32)
    if target == "banana":
        theta0[:, 1] = 6.0
    mom = None
    if method == "rsghmc":
        from pysgmcmc_b200.samplers.relativistic_sghmc import _sample_relativistic_momentum
        mom = np.array(_sample_relativistic_momentum(1.0, 1.0, C * D, seed=seed)).reshape(C, D).astype(np.float32)
    chain = osamplers.OracleChain(method, theta0, otargets.cost_and_grad(target), epsilon=epsilon,
                                  momentum=mom, **hyper)
    for _ in range(steps):
        chain.next(rng.standard_normal((C, D)).astype(np.float32))
    return chain.state["theta"]


@pytest.mark.parametrize("method,target,steps,hyper", [
    ("sgld", "gmm1", 6000, dict(burn_in_steps=1000)),                 # config 2: SGLD on the 1-D mixtures
    ("sgld", "gmm2", 6000, dict(burn_in_steps=1000)),
    ("sgld", "gmm3", 6000, dict(burn_in_steps=1000)),
    ("sghmc", "gmm1", 4000, dict(burn_in_steps=1000)),
    ("rsghmc", "gmm2", 3000, dict(stepsize_schedule=ConstantStepsizeSchedule(0.05))),
])
def test_gmm_chains_match_the_oracle_distribution(method, target, steps, hyper):
    C = 4096
    got = gpu_final_states(method, target, C, steps, seed=11, **hyper)[:, 0]
    ohyper = {k: v for k, v in hyper.items() if k != "stepsize_schedule"}
    eps = hyper["stepsize_schedule"].initial_value if "stepsize_schedule" in hyper else None
    want = oracle_final_states(method, target, C, steps, seed=23, epsilon=eps, **ohyper)[:, 0]
    assert np.isfinite(got).all()
    ks = stats.ks_2samp(got, want)
    assert ks.pvalue > 1e-3, "KS p = %.2e" % ks.pvalue
    se = np.sqrt(got.var() / C + want.var() / C)
    assert abs(got.mean() - want.mean()) < 4.5 * se
    assert abs(np.log(got.var() / want.var())) < 0.15


@pytest.mark.parametrize("target", ["gmm1", "gmm2", "gmm3"])
def test_gmm_sgld_long_run_matches_the_analytic_mixture(target):
    """Config 2: 4096 SGLD chains on the 1-D mixtures, 20 000 steps from x0 = 0 (one K6 launch):
    the chain states follow the analytic mixture (KS test) with ~1/3 of the mass per mode.
    (The oracle passes the same test with p > 0.2; the O(eps) bias is below the resolution of
    4096 draws.)"""
    C = 4096
    x = gpu_final_states("sgld", target, C, 20000, seed=3, burn_in_steps=3000)[:, 0]
    var = otargets.GMM_VAR[target]

    def mixture_cdf(v):
        return sum(stats.norm.cdf(v, m, np.sqrt(s2)) for m, s2 in zip(otargets.GMM_MU, var)) / 3.0
    ks = stats.kstest(x, mixture_cdf)
    assert ks.pvalue > 1e-3, "KS D = %.4f, p = %.2e" % (ks.statistic, ks.pvalue)
    weights = np.array([np.mean(x < -2.5), np.mean(np.abs(x) <= 2.5), np.mean(x > 2.5)])
    assert np.abs(weights - 1.0 / 3.0).max() < 0.07, weights


def test_banana_sghmc_moments_match_the_oracle():
    """Config 1 density, 2048 chains: ESS-normalised mean error of both coordinates and of the
    curvature statistic x1 + 0.1 x0^2 (the banana's ridge) against oracle chains."""
    C, steps = 2048, 6000
    got = gpu_final_states("sghmc", "banana", C, steps, seed=5, burn_in_steps=2000)
    want = oracle_final_states("sghmc", "banana", C, steps, seed=6, burn_in_steps=2000)
    for name, f in (("x0", lambda t: t[:, 0]), ("x1", lambda t: t[:, 1]),
                    ("ridge", lambda t: t[:, 1] + 0.1 * t[:, 0] ** 2)):
        a, b = f(got), f(want)
        se = np.sqrt(a.var() / C + b.var() / C)
        assert abs(a.mean() - b.mean()) < 4.5 * se, name
        assert stats.ks_2samp(a, b).pvalue > 1e-3, name
    ridge = got[:, 1] + 0.1 * got[:, 0] ** 2
    assert abs(ridge.mean() - 10.0) < 0.25 and abs(ridge.std() - 1.0) < 0.15    # ~ N(10, 1) under the target


def test_bnn_posterior_predictive_matches_the_oracle():
    """BNN-SGHMC, 48 chains on the GPU vs 48 oracle chains (independent noise and minibatch
    streams), small sinc problem: posterior predictive mean and spread at test points."""
    C, N, B, steps, burn = 48, 60, 20, 700, 300
    rng = np.random.RandomState(1)
    X = np.asarray([rng.uniform(0.0, 1.0, 1) for _ in range(N)])
    y = np.sinc(X * 10 - 5).sum(axis=1)
    X = ((X - X.mean(0)) / X.std(0)).astype(np.float32)
    y = ((y - y.mean()) / y.std()).astype(np.float32)
    theta0 = obnn.init_theta(C, seed=2, dtype=np.float32)
    Xt = np.linspace(X.min(), X.max(), 25, dtype=np.float32)[:, None]

    # GPU
    gen = DeviceBatchGenerator(N, B, n_chains=C, seed=100, device=DEV)
    nll = BayesianNeuralNetworkNLL(N, B, X=X, y=y, starts_placeholder=gen.starts_placeholder, device=DEV)
    params, off = [], 0
    for shp in parameter_shapes(1):
        n = int(np.prod(shp))
        params.append(torch.tensor(theta0[:, off:off + n].reshape((C,) + shp), device=DEV))
        off += n
    s = SGHMCSampler(params=params, cost_fun=nll, batch_generator=gen, burn_in_steps=burn, scale_grad=float(N),
                     seed=9, session=Session(device=DEV, n_chains=C, output="torch"))
    s.run(steps, keep_every=steps)
    f_gpu, _, _ = obnn.forward(s._theta.cpu().numpy().astype(np.float64), np.repeat(Xt[None], C, 0))

    # oracle
    holder = {}
    orng = np.random.RandomState(55)

    def cost_and_grad(theta):
        Xb, yb = obnn.gather_minibatch(X, y, holder["s"], B)
        c, g, _ = obnn.nll_and_grad(theta, Xb, yb, n_examples=N)
        return c, g
    chain = osamplers.OracleChain("sghmc", theta0, cost_and_grad, epsilon=0.01, burn_in_steps=burn,
                                  scale_grad=float(N))
    for _ in range(steps):
        holder["s"] = orng.randint(0, N - B + 1, size=C)
        chain.next(orng.standard_normal((C, 5252)).astype(np.float32))
    f_cpu, _, _ = obnn.forward(chain.state["theta"].astype(np.float64), np.repeat(Xt[None], C, 0))

    # both fit the data comparably and agree on the predictive mean within the sampling error
    mse_gpu = np.mean((obnn.forward(s._theta.cpu().numpy().astype(np.float64), np.repeat(X[None], C, 0))[0] - y) ** 2)
    mse_cpu = np.mean((obnn.forward(chain.state["theta"].astype(np.float64), np.repeat(X[None], C, 0))[0] - y) ** 2)
    assert abs(np.log(mse_gpu / mse_cpu)) < 0.5, (mse_gpu, mse_cpu)
    se = np.sqrt(f_gpu.var(axis=0) / C + f_cpu.var(axis=0) / C) + 1e-3
    z = np.abs(f_gpu.mean(axis=0) - f_cpu.mean(axis=0)) / se
    assert np.mean(z) < 2.0 and np.max(z) < 6.0, z


@pytest.mark.parametrize("target,stepsize", [("gmm2", 1.01), ("gmm2", 3.01), ("gmm3", 2.01), ("banana", 1.01)])
def test_relativistic_ess_reproduces_the_published_table(target, stepsize):
    """The only numbers the reference publishes for this path: mean ESS of Relativistic SGHMC vs
    stepsize (docs/source/notebooks/data/effective_sample_sizes/Relativistic_SGHMC.json; protocol
    of docs/source/experiments/compute_ess.py:177-253: one chain, 20 segments x 10 000 draws kept
    every 10 steps, pymc3 ESS, 5 repeats).  Same protocol on the GPU (K6 + K8): the mean over 5
    repeats lands within 8 % of the published mean (the published repeats spread by 2-4 %).
    This pins the relativistic sampler AND the restated ESS estimator against real reference
    output, statistically."""
    import json
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import ess_vs_stepsize
    published = json.load(open(os.path.join(root, "tests", "golden", "relativistic_ess_published.json")))
    row = [r for r in published[target] if abs(r[0] - stepsize) < 1e-9][0]
    ours = ess_vs_stepsize.mean_ess(target, stepsize, seed=123, dev=torch.device(DEV))
    assert np.isfinite(ours).all()
    ratio = np.mean(ours) / row[1]
    assert 0.92 < ratio < 1.08, (ours, row)
