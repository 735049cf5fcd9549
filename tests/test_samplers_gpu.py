"""GPU parity tests of the sampler classes (the reference-facing API) against the oracle:
generic autograd path (K1-K3), fused target path (K6, per step and multi-step), golden
trajectory files, burn-in bookkeeping and the reference's own seed-determinism test.

Tolerance: north_star part (1): 1e-5 relative in FP32 with injected noise.  The target
gradients are computed on the device (autograd, or by hand in K6) with fp32 rounding
that differs from the oracle's in the last bit, so trajectories are compared with
rtol = 1e-5 (plus atol = 1e-5 * typical scale for coordinates crossing zero); gmm costs
go through exp/log and get 2e-5.
"""
import os
from itertools import islice

import numpy as np
import pytest
import torch

from oracle import philox, samplers as osamplers, targets as otargets
from pysgmcmc_b200 import Session
from pysgmcmc_b200.diagnostics.objective_functions import (
    banana_log_likelihood, gmm1_log_likelihood, gmm2_log_likelihood, gmm3_log_likelihood,
    to_negative_log_likelihood)
from pysgmcmc_b200.sampling import Sampler
from pysgmcmc_b200.samplers import RelativisticSGHMCSampler, SGHMCSampler, SGLDSampler
from pysgmcmc_b200.stepsize_schedules import ConstantStepsizeSchedule

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LOGLIK = {"banana": banana_log_likelihood, "gmm1": gmm1_log_likelihood,
          "gmm2": gmm2_log_likelihood, "gmm3": gmm3_log_likelihood}
CLS = {"sghmc": SGHMCSampler, "sgld": SGLDSampler, "rsghmc": RelativisticSGHMCSampler}
RTOL, ATOL = 1e-5, 2e-5


def as_matrix(sample, C):
    """(list of per-parameter [C] tensors | unwrapped single tensor) -> [C, D] numpy."""
    if isinstance(sample, (list, tuple)):
        return torch.stack([t.reshape(C) for t in sample], dim=1).cpu().numpy()
    return sample.reshape(C, -1).cpu().numpy()


def make_params(theta0):
    """theta0 [C, D] -> list of D tensors of shape [C] (scalar parameters with a chain axis)."""
    return [torch.tensor(theta0[:, d].copy(), device=DEV) for d in range(theta0.shape[1])]


def build(method, target, theta0, fused, seed=None, momentum=None, **hyper):
    C = theta0.shape[0]
    cost = to_negative_log_likelihood(LOGLIK[target])
    if not fused:
        inner = cost
        cost = lambda params: inner(params)       # hides the native tag -> autograd path
    sess = Session(device=DEV, n_chains=C, output="torch")
    eps = hyper.pop("epsilon", None)
    if eps is not None and method != "sgld":
        hyper["stepsize_schedule"] = ConstantStepsizeSchedule(eps)
    sampler = CLS[method](params=make_params(theta0), cost_fun=cost, session=sess, seed=seed, **hyper)
    assert (sampler._native_target is not None) == fused
    if momentum is not None:
        sampler._state_array("p").copy_(torch.as_tensor(momentum, device=DEV))
    return sampler


def golden_cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg


MG = golden_cases()
F32_CASES = [c for c in MG.CASES if c[3] == np.float32]


@pytest.mark.parametrize("fused", [False, True], ids=["autograd+K1", "fused-K6"])
@pytest.mark.parametrize("case", F32_CASES, ids=[c[0] for c in F32_CASES])
def test_golden_trajectories(case, fused):
    """1000 steps, injected noise, 4 chains: theta at the checkpoints and every cost
    against the committed oracle trajectories (tests/golden/trajectories.npz)."""
    name, method, target, dtype, seed, hyper = case
    g = np.load(os.path.join(GOLDEN, "trajectories.npz"))
    theta0 = g[name + "/theta0"]
    momentum = g[name + "/momentum0"] if method == "rsghmc" else None
    hyper = dict(hyper)
    if method == "sgld":
        hyper.pop("epsilon", None)
    sampler = build(method, target, theta0, fused, momentum=momentum, **hyper)
    z_rng = np.random.RandomState(seed + 1)
    k = 0
    for step in range(1, 1001):
        z = z_rng.standard_normal(theta0.shape).astype(np.float32)
        sample, cost = sampler.__next__(feed_dict={sampler.noise: z})
        want_cost = g[name + "/cost"][step - 1]
        np.testing.assert_allclose(cost.cpu().numpy(), want_cost, rtol=2e-5, atol=ATOL,
                                   err_msg="%s cost at step %d" % (name, step))
        if step in MG.CHECKPOINTS:
            got = as_matrix(sample, theta0.shape[0])
            np.testing.assert_allclose(got, g[name + "/theta"][k], rtol=RTOL, atol=ATOL,
                                       err_msg="%s theta at step %d" % (name, step))
            k += 1
    assert sampler.n_iterations == 1000


@pytest.mark.parametrize("method,target", [("sghmc", "banana"), ("sgld", "gmm2"), ("rsghmc", "gmm3")])
def test_multi_step_run_equals_per_step(method, target):
    """sampler.run(n) (one K6 launch, Philox noise) == n calls of next(sampler) == oracle fed
    with the oracle's restatement of the Philox stream."""
    C, steps, burn = 64, 400, 150
    D = 2 if target == "banana" else 1
    rng = np.random.RandomState(5)
    theta0 = rng.uniform(-2, 2, size=(C, D)).astype(np.float32)
    mom = rng.standard_normal((C, D)).astype(np.float32) if method == "rsghmc" else None
    hyper = dict(burn_in_steps=burn) if method != "rsghmc" else dict(epsilon=0.01)
    a = build(method, target, theta0, True, seed=1234, momentum=mom, **dict(hyper))
    b = build(method, target, theta0, True, seed=1234, momentum=mom, **dict(hyper))
    trace, costs = a.run(steps, keep_every=4)
    assert trace.shape == (steps // 4, C, D) and costs.shape == (steps // 4, C)
    assert a.n_iterations == steps
    kept = []
    for s in range(steps):
        sample, cost = next(b)
        if (s + 1) % 4 == 0:
            kept.append((torch.as_tensor(as_matrix(sample, C), device=DEV), cost))
    assert torch.equal(trace, torch.stack([k[0] for k in kept]))
    assert torch.equal(costs, torch.stack([k[1] for k in kept]))
    assert torch.equal(a._theta, b._theta)
    # oracle with the restated Philox normals
    chain = osamplers.OracleChain(method, theta0, otargets.cost_and_grad(target), momentum=mom,
                                  **{k: v for k, v in hyper.items()})
    for s in range(steps):
        z = philox.normals(C * D, seed=1234, step=s).reshape(C, D).astype(np.float32)
        theta, _ = chain.next(z)
    np.testing.assert_allclose(a._theta.cpu().numpy(), theta, rtol=1e-4, atol=1e-4)


def test_reference_mode_single_chain_api():
    """Reference semantics (no chain axis, numpy outputs): list during burn-in, unwrapped
    single parameter afterwards (base_classes.py:302-304 vs :446); cost is pre-update;
    params are live views; minv freezes after burn-in."""
    x = torch.tensor(0.0, device=DEV)
    sampler = SGHMCSampler(params=[x], cost_fun=to_negative_log_likelihood(gmm1_log_likelihood),
                           burn_in_steps=3, seed=1, session=Session(device=DEV))
    assert sampler.is_burning_in
    sample, cost = next(sampler)
    assert isinstance(sample, list) and isinstance(sample[0], np.ndarray) and sample[0].shape == ()
    want_cost, _ = otargets.gmm_cost_and_grad(np.zeros((1, 1), dtype=np.float32))
    assert np.allclose(cost, want_cost[0], rtol=1e-6)
    assert float(x) == float(sample[0])                       # user tensor follows the state
    next(sampler), next(sampler)
    assert not sampler.is_burning_in and sampler.n_iterations == 3
    minv_frozen = [m.copy() for m in sampler.minv]
    sample, cost = next(sampler)
    assert isinstance(sample, np.ndarray)                     # unwrapped
    assert np.array_equal(sampler.minv[0], minv_frozen[0])
    # banana from (0, 0): cost == 50 (notebook known answer, sign convention of a cost)
    p = [torch.tensor(0.0, device=DEV), torch.tensor(0.0, device=DEV)]
    s2 = SGHMCSampler(params=p, cost_fun=to_negative_log_likelihood(banana_log_likelihood),
                      session=Session(device=DEV), seed=3)
    sample, cost = next(s2)
    assert cost == 50.0 and len(sample) == 2
    assert abs(sample[0]) < 0.02 and abs(sample[1] - 1e-3) < 0.02   # -eps^2*grad = +1e-3 on x1, noise 3.2e-3


@pytest.mark.parametrize("cls,kwargs", [(SGHMCSampler, {}), (SGLDSampler, {}),
                                        (RelativisticSGHMCSampler, {})])
@pytest.mark.parametrize("target", ["gmm1", "banana"])
def test_seed_determinism_like_the_reference(cls, kwargs, target):
    """tests/samplers/sampler_testing.py:29-59: two fresh samplers, same seed -> same chain."""
    def fresh_chain(seed, n):
        params = ([torch.tensor(0.0, device=DEV)] if target == "gmm1"
                  else [torch.tensor(0.0, device=DEV), torch.tensor(6.0, device=DEV)])
        loglik = LOGLIK[target]
        sampler = cls(params=params, cost_fun=lambda p: -loglik(p), seed=seed,
                      session=Session(device=DEV), **kwargs)
        return list(islice(sampler, n))
    seed = int(np.random.randint(0, 2 ** 31 - 1))
    n = int(np.random.randint(1, 100))
    for (s1, c1), (s2, c2) in zip(fresh_chain(seed, n), fresh_chain(seed, n)):
        assert np.allclose(c1, c2)
        assert np.allclose(s1, s2)
    (s1, _), (s2, _) = fresh_chain(seed, 1)[0], fresh_chain(seed + 1, 1)[0]
    assert not np.allclose(s1, s2)


def test_sgld_ignores_stepsize_schedule_like_the_reference():
    """sgld.py:96-100 never forwards the schedule: epsilon stays 0.01."""
    x = torch.tensor(0.0, device=DEV)
    s = SGLDSampler(params=[x], cost_fun=to_negative_log_likelihood(gmm1_log_likelihood),
                    stepsize_schedule=ConstantStepsizeSchedule(0.5), session=Session(device=DEV))
    next(s)
    assert float(s.epsilon) == 0.01


def test_factory_builds_samplers():
    x = torch.zeros(4, device=DEV)
    s = Sampler.get_sampler(Sampler.SGHMC, params=[x], cost_fun=lambda p: (p[0] ** 2).sum(),
                            session=Session(device=DEV), dtype=torch.float32)
    assert type(s) is SGHMCSampler and s.dtype == torch.float32
    next(s)
    s = Sampler.get_sampler(Sampler.RelativisticSGHMC, params=[x], cost_fun=lambda p: (p[0] ** 2).sum(),
                            session=Session(device=DEV))
    assert type(s) is RelativisticSGHMCSampler
    with pytest.raises(ValueError, match="does not take any parameter with name 'unknown_argument'"):
        Sampler.get_sampler(Sampler.SGLD, unknown_argument=1, params=[x], cost_fun=lambda p: p[0].sum())


def test_float64_generic_path():
    """dtype=float64 (the reference's default) runs the f64 kernels: compare with the f64 golden."""
    g = np.load(os.path.join(GOLDEN, "trajectories.npz"))
    name = "sghmc_banana_f64"
    theta0 = g[name + "/theta0"]
    C = theta0.shape[0]
    params = [torch.tensor(theta0[:, d].copy(), device=DEV, dtype=torch.float64) for d in range(2)]
    s = SGHMCSampler(params=params, cost_fun=lambda p: -banana_log_likelihood(p), burn_in_steps=300,
                     dtype=torch.float64, session=Session(device=DEV, n_chains=C, output="torch"))
    z_rng = np.random.RandomState(11 + 1)
    for step in range(1, 101):
        z = z_rng.standard_normal(theta0.shape)
        sample, cost = s.__next__(feed_dict={s.noise: z})
    got = as_matrix(sample, C)
    np.testing.assert_allclose(got, g[name + "/theta"][3], rtol=1e-10, atol=1e-12)


def test_invalid_constructor_inputs_assert():
    x = torch.zeros(2, device=DEV)
    good = dict(params=[x], cost_fun=lambda p: p[0].sum(), session=Session(device=DEV))
    for bad in (dict(seed="1"), dict(seed=1.5), dict(batch_generator=[1, 2]), dict(dtype="float32"),
                dict(cost_fun=3), dict(stepsize_schedule=0.01), dict(burn_in_steps=1.5),
                dict(session="session")):
        with pytest.raises(AssertionError):
            SGHMCSampler(**{**good, **bad})


@pytest.mark.parametrize("output", ["torch", "numpy"])
@pytest.mark.parametrize("cls_name", ["sghmc", "sgld", "rsghmc"])
def test_prefetched_next_returns_the_same_pairs(cls_name, output):
    """Session(prefetch=S): next(sampler) hands out steps computed S at a time by one fused launch;
    the (sample, cost) pairs, their container types across the burn-in boundary and n_iterations
    are those of the step-by-step loop, bit for bit."""
    from pysgmcmc_b200.samplers import RelativisticSGHMCSampler, SGHMCSampler, SGLDSampler
    cls = {"sghmc": SGHMCSampler, "sgld": SGLDSampler, "rsghmc": RelativisticSGHMCSampler}[cls_name]
    kw = {} if cls_name == "rsghmc" else {"burn_in_steps": 37}

    def make(prefetch):
        params = [torch.tensor(0.0, device=DEV), torch.tensor(6.0, device=DEV)]
        return cls(params=params, cost_fun=to_negative_log_likelihood(banana_log_likelihood), seed=5,
                   session=Session(device=DEV, output=output, prefetch=prefetch), **kw)
    a, b = make(0), make(16)
    if cls_name == "rsghmc":
        b._state_array("p").copy_(a._state_array("p"))
    for step in range(100):
        (sa, ca), (sb, cb) = next(a), next(b)
        assert type(sa) is type(sb) and a.n_iterations == b.n_iterations == step + 1
        assert getattr(a, "is_burning_in", False) == getattr(b, "is_burning_in", False)
        ha = [np.asarray(x.cpu() if output == "torch" else x) for x in sa]
        hb = [np.asarray(x.cpu() if output == "torch" else x) for x in sb]
        assert all(np.array_equal(x, y) for x, y in zip(ha, hb)), "sample at step %d" % step
        assert np.array_equal(np.asarray(ca.cpu() if output == "torch" else ca),
                              np.asarray(cb.cpu() if output == "torch" else cb)), "cost at step %d" % step
    # leaving the prefetched mode: the steps computed ahead are skipped, the chain goes on from the device state
    ahead = b._pf["n"] - b._pf["pos"]
    b.run(3)
    assert b.n_iterations == 100 + ahead + 3 and b._pf is None
