"""GPU parity tests of the BNN path: K4 (cost + gradient), K5 (BNN-SGHMC steps driven from
C), K10 (predictive forward) and the sampler classes on top, against the oracle.

Tolerances: K4 accumulates 50-term dot products with FMA in fp32 and evaluates tanh on the
SFU (|error| ~ 2e-7 per activation), the oracle is float64.  Cost: rtol 2e-6.  Gradient:
|g - g_ref| <= 2e-5 * max|g_ref| per chain (element-wise rtol is meaningless for the many
entries that are ~1e-8 of the largest one).  Trajectories (injected noise, scale_grad = N as
BayesianNeuralNetwork.train sets it): 1e-5 relative to max|theta| after 200 steps (N = 2000)
and after 1000 steps at the benchmarked shapes (N = 20 000) -- float32 itself separates the
float32 and float64 oracles by 1.4e-5 there; the update alone is bit-exact (teacher-forced test).
"""
import os

import numpy as np
import pytest
import torch

from oracle import bnn as obnn, mt19937 as omt, samplers as osamplers
from pysgmcmc_b200 import Session, _native
from pysgmcmc_b200.data_batches import DeviceBatchGenerator
from pysgmcmc_b200.models.bnn_cost import (BayesianNeuralNetworkNLL, default_net_params,
                                           n_parameters, network_output, parameter_shapes)
from pysgmcmc_b200.samplers import SGHMCSampler, SGLDSampler
from pysgmcmc_b200.stepsize_schedules import ConstantStepsizeSchedule

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
K4_DEFAULT = 16      # tensor-pipe kernel: rounded split, FP32-pipe accumulation, separate cross-term accumulator (csrc/bnn.cu: g_bnn_variant)
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(autouse=True)
def streaming_step_kernels(monkeypatch):
    """This module tests K4 then K1 (all chains streamed through HBM every step).  Samplers with as few chains
    as these tests use would run the resident kernel instead (tests/test_bnn_resident_gpu.py)."""
    monkeypatch.setattr(SGHMCSampler, "RESIDENT_MAX_CHAINS", 0)


def sinc_data(N, n_in=1, seed=1):
    rng = np.random.RandomState(seed)
    X = np.array([rng.uniform(0.0, 1.0, n_in) for _ in range(N)])      # tests/utils.py:24-29
    y = np.sinc(X * 10 - 5).sum(axis=1)                                 # tests/utils.py:32-33
    X = (X - X.mean(axis=0)) / X.std(axis=0)                            # base_model.py:125-133
    y = (y - y.mean()) / y.std()
    return X, y


def k4(theta, X, y, starts, batch, bs_cfg, N, want_grad=True, n_in=1):
    C = theta.shape[0]
    t = torch.as_tensor(theta, dtype=torch.float32, device=DEV).contiguous()
    Xd = torch.as_tensor(X, dtype=torch.float32, device=DEV).contiguous()
    yd = torch.as_tensor(y, dtype=torch.float32, device=DEV).contiguous()
    sd = None if starts is None else torch.as_tensor(starts, dtype=torch.int32, device=DEV)
    cost = torch.empty(C, device=DEV)
    mse = torch.empty(C, device=DEV)
    grad = torch.full_like(t, float("nan")) if want_grad else None
    _native.call("sgmcmc_bnn_nll_grad_f32", _native.ptr(t), _native.ptr(Xd), _native.ptr(yd),
                 _native.ptr(sd), _native.ptr(cost), _native.ptr(grad), _native.ptr(mse), C, n_in, batch,
                 float(bs_cfg), N, _native.stream_ptr())
    torch.cuda.synchronize()
    return cost.cpu().numpy(), None if grad is None else grad.cpu().numpy(), mse.cpu().numpy()


def assert_grad_close(got, want, tol=2e-5):
    scale = np.abs(want).max(axis=1, keepdims=True)
    err = np.abs(got - want) / scale
    assert np.isfinite(got).all()
    assert err.max() <= tol, "max |dg| / max|g| = %.3g" % err.max()


def test_k4_golden_file():
    g = np.load(os.path.join(GOLDEN, "bnn_nll.npz"))
    cost, grad, mse = k4(g["theta"], g["X"], g["y"], g["starts"], 20, 20, g["X"].shape[0])
    np.testing.assert_allclose(cost, g["cost"], rtol=2e-6)
    np.testing.assert_allclose(mse, g["mse"], rtol=2e-5)
    assert_grad_close(grad, g["grad"])


@pytest.mark.parametrize("C", [1, 4, 5, 6, 37, 1000])
@pytest.mark.parametrize("n_in,batch,N", [(1, 20, 20000), (1, 7, 50), (3, 20, 300), (1, 32, 100), (2, 1, 10)])
def test_k4_matches_oracle(C, n_in, batch, N):
    """Ragged cases: chain counts that do not fill a CTA, batches that are not a multiple of
    the 4-row blocking, several input features, batch == 1."""
    if C == 1000 and n_in != 1:
        pytest.skip("large C covered for the headline shape only")
    rng = np.random.RandomState(C * 31 + batch)
    X, y = sinc_data(N, n_in, seed=3)
    theta = obnn.init_theta(C, n_in=n_in, seed=C, dtype=np.float64)
    theta += 0.1 * rng.standard_normal(theta.shape)
    starts = rng.randint(0, N - batch + 1, size=C)
    Xb, yb = obnn.gather_minibatch(X, y, starts, batch)
    wc, wg, wm = obnn.nll_and_grad(theta, Xb, yb, n_examples=N, batch_size=20, n_in=n_in)
    cost, grad, mse = k4(theta, X, y, starts, batch, 20, N, n_in=n_in)
    np.testing.assert_allclose(cost, wc, rtol=3e-6)
    np.testing.assert_allclose(mse, wm, rtol=3e-5, atol=1e-8)   # residuals carry ~1e-7 absolute error
    assert_grad_close(grad, wg)
    # cost-only launch (grad == NULL) gives the same cost
    cost2, _, _ = k4(theta, X, y, starts, batch, 20, N, want_grad=False, n_in=n_in)
    assert np.array_equal(cost, cost2)


def test_k4_reference_prior_golden_inside_the_cost():
    """With zero residual the cost reduces to the two priors the reference pins with golden
    vectors (tests/bayesian_neural_network/test_priors.py): check K4 against them."""
    g = np.load(os.path.join(GOLDEN, "bnn_priors.npz"))
    theta = np.concatenate([g["w%d" % i].ravel() for i in range(9)])[None, :]
    X = np.zeros((20, 1))
    f, rho, _ = obnn.forward(theta, X[None])
    y = f[0]
    N = 100
    cost, _, mse = k4(theta, X, y, None, 20, 20, N)
    lv = obnn.log_variance_prior_log_like(np.full((20, 1), rho[0]))
    wp = float(g["expected_weights"])
    expect = -(-0.5 * rho[0] + lv / N + wp / N)
    assert abs(mse[0]) < 1e-10
    np.testing.assert_allclose(cost[0], expect, rtol=2e-6)


def test_k4_full_size_properties():
    """8192 chains (one GPU's shard of config 4): identical chains give identical results
    whatever CTA slot they land in, and the gradient is linear in the residual scale."""
    C, N = 8192, 20000
    X, y = sinc_data(N)
    theta1 = obnn.init_theta(1, seed=7, dtype=np.float64)
    theta = np.repeat(theta1, C, axis=0)
    starts = np.full(C, 1234)
    cost, grad, _ = k4(theta, X, y, starts, 20, 20, N)
    assert (cost == cost[0]).all() and (grad == grad[0]).all()
    Xb, yb = obnn.gather_minibatch(X, y, starts[:1], 20)
    wc, wg, _ = obnn.nll_and_grad(theta1, Xb, yb, n_examples=N)
    np.testing.assert_allclose(cost[0], wc[0], rtol=3e-6)
    assert_grad_close(grad[:1], wg)


def test_k10_predict_matches_oracle():
    rng = np.random.RandomState(2)
    n_nets, n_points = 7, 77
    theta = obnn.init_theta(n_nets, seed=4, dtype=np.float64) + 0.05 * rng.standard_normal((n_nets, 5252))
    X = rng.uniform(-2, 2, size=(n_points, 1))
    t = torch.as_tensor(theta, dtype=torch.float32, device=DEV)
    out = torch.empty((n_nets, n_points, 2), device=DEV)
    Xd = torch.as_tensor(X, dtype=torch.float32, device=DEV)
    _native.call("sgmcmc_bnn_predict_f32", _native.ptr(t), _native.ptr(Xd), _native.ptr(out), n_nets, 1, n_points,
                 _native.stream_ptr())
    f, rho, _ = obnn.forward(theta, np.repeat(X[None], n_nets, axis=0))
    np.testing.assert_allclose(out[..., 0].cpu().numpy(), f, rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(out[..., 1].cpu().numpy(), np.repeat(rho[:, None], n_points, 1), rtol=1e-6)


def test_torch_cost_equals_oracle_and_k4():
    """The differentiable torch restatement used by the generic path agrees with both."""
    C, N = 3, 200
    X, y = sinc_data(N)
    theta = obnn.init_theta(C, seed=2, dtype=np.float64) + 0.05 * np.random.RandomState(0).standard_normal((C, 5252))
    starts = np.array([0, 17, 180])
    ph = type("P", (), {})
    from pysgmcmc_b200.placeholders import Placeholder
    sp = Placeholder("starts")
    sp.value = torch.as_tensor(starts, dtype=torch.int32, device=DEV)
    nll = BayesianNeuralNetworkNLL(N, 20, X=X, y=y, starts_placeholder=sp, device=DEV, dtype=torch.float64)
    params, off = [], 0
    for shp in parameter_shapes(1):
        n = int(np.prod(shp))
        params.append(torch.tensor(theta[:, off:off + n].reshape((C,) + shp), device=DEV, requires_grad=True))
        off += n
    cost = nll(params)
    Xb, yb = obnn.gather_minibatch(X, y, starts, 20)
    wc, wg, _ = obnn.nll_and_grad(theta, Xb, yb, n_examples=N)
    np.testing.assert_allclose(cost.detach().cpu().numpy(), wc, rtol=1e-12)
    grads = torch.autograd.grad(cost.sum(), params)
    got = np.concatenate([gr.reshape(C, -1).cpu().numpy() for gr in grads], axis=1)
    np.testing.assert_allclose(got, wg, rtol=1e-8, atol=1e-14)


def oracle_bnn_chain(theta0, X, y, seeds, N, batch, steps, burn, z_seed, eps=0.01):
    C, D = theta0.shape
    streams = [omt.MT19937(int(s)) for s in seeds]
    holder = {}

    def cost_and_grad(theta):
        Xb, yb = obnn.gather_minibatch(X, y, holder["starts"], batch)
        c, g, _ = obnn.nll_and_grad(theta, Xb.astype(np.float32), yb.astype(np.float32), n_examples=N)
        return c, g
    chain = osamplers.OracleChain("sghmc", theta0, cost_and_grad, epsilon=eps, burn_in_steps=burn,
                                  scale_grad=float(N))
    zr = np.random.RandomState(z_seed)
    costs = []
    for s in range(steps):
        holder["starts"] = np.array([st.bounded(N - batch) for st in streams])
        z = zr.standard_normal((C, D)).astype(np.float32)
        theta, c = chain.next(z)
        costs.append(c)
    return theta, np.stack(costs), chain


def test_native_cost_refuses_float64_instead_of_casting():
    """The fused cost + gradient kernels are float32; handing them float64 parameters is an error
    (float64 samplers go through the differentiable cost and the float64 update kernels)."""
    X, y = sinc_data(64)
    nll = BayesianNeuralNetworkNLL(64, 20, X=X[:20], y=y[:20], device=DEV)
    theta = torch.zeros((2, 5252), dtype=torch.float64, device=DEV)
    with pytest.raises(TypeError, match="float32"):
        nll.native_cost_and_grad(theta, torch.empty_like(theta))


def test_bnn_sghmc_sampler_trajectory_matches_oracle():
    """next(sampler) on the BNN cost with per-chain on-device minibatch streams and
    injected noise, 200 steps across the burn-in boundary, vs the fp32 oracle."""
    C, N, batch, steps, burn = 6, 2000, 20, 200, 120
    X, y = sinc_data(N)
    theta0 = obnn.init_theta(C, seed=11, dtype=np.float32)
    seeds = np.arange(C) + 40
    want_theta, want_cost, chain = oracle_bnn_chain(theta0, X, y, seeds, N, batch, steps, burn, z_seed=9)

    gen = DeviceBatchGenerator(N, batch, seeds=seeds, device=DEV, block=64)
    nll = BayesianNeuralNetworkNLL(N, batch, X=X, y=y, starts_placeholder=gen.starts_placeholder, device=DEV)
    params, off = [], 0
    for shp in parameter_shapes(1):
        n = int(np.prod(shp))
        params.append(torch.tensor(theta0[:, off:off + n].reshape((C,) + shp), device=DEV))
        off += n
    sampler = SGHMCSampler(params=params, cost_fun=nll, batch_generator=gen, burn_in_steps=burn,
                           scale_grad=float(N), stepsize_schedule=ConstantStepsizeSchedule(0.01),
                           session=Session(device=DEV, n_chains=C, output="torch"))
    zr = np.random.RandomState(9)
    for s in range(steps):
        z = zr.standard_normal((C, 5252)).astype(np.float32)
        sample, cost = sampler.__next__(feed_dict={sampler.noise: z})
        np.testing.assert_allclose(cost.cpu().numpy(), want_cost[s], rtol=1e-4, err_msg="step %d" % s)
    got = sampler._theta.cpu().numpy()
    scale = np.abs(want_theta).max()
    assert np.abs(got - want_theta).max() <= 1e-5 * scale
    # the frozen mass matrix: compared like the parameters, against the scale of the array (its
    # entries are v_hat^-1/2 of a running mean of grad^2; update error and gradient error are
    # separated in the teacher-forced test below)
    minv = sampler._state_array("minv").cpu().numpy()
    assert np.abs(minv - chain.minv).max() <= 5e-4 * np.abs(chain.minv).max()
    assert np.median(np.abs(minv / chain.minv - 1.0)) < 1e-5
    assert not sampler.is_burning_in


@pytest.mark.parametrize("variant", [K4_DEFAULT, 0])
def test_bnn_sghmc_teacher_forced_update_bit_exact_gradient_error_bounded(variant):
    """Separates the two error sources of a BNN-SGHMC trajectory.  Every step the oracle update
    (sghmc.py:165-251) is fed the GPU's OWN gradient and the GPU's state before the step: the
    state after K1 -- theta, V, tau, g, v_hat and the inverse mass matrix -- must then be
    bit-identical (update error = 0, across the burn-in boundary); and the GPU's gradient is
    compared with the float64 oracle gradient at the same point (gradient error
    <= 2e-5 max|g| per chain)."""
    C, N, batch, steps, burn = 6, 20000, 20, 150, 100
    X, y = sinc_data(N)
    theta0 = obnn.init_theta(C, seed=11, dtype=np.float32)
    seeds = np.arange(C) + 40
    _native.call("sgmcmc_set_bnn_tuning", variant)
    try:
        gen = DeviceBatchGenerator(N, batch, seeds=seeds, device=DEV, block=64)
        nll = BayesianNeuralNetworkNLL(N, batch, X=X, y=y, starts_placeholder=gen.starts_placeholder, device=DEV)
        params, off = [], 0
        for shp in parameter_shapes(1):
            n = int(np.prod(shp))
            params.append(torch.tensor(theta0[:, off:off + n].reshape((C,) + shp), device=DEV))
            off += n
        sampler = SGHMCSampler(params=params, cost_fun=nll, batch_generator=gen, burn_in_steps=burn,
                               scale_grad=float(N), stepsize_schedule=ConstantStepsizeSchedule(0.01),
                               session=Session(device=DEV, n_chains=C, output="torch"))
        streams = [omt.MT19937(int(s)) for s in seeds]
        zr = np.random.RandomState(9)
        names = ("v", "tau", "g", "v_hat", "minv")
        frozen = None
        worst_grad = 0.0
        for s in range(steps):
            before = {"theta": sampler._theta.cpu().numpy()}
            before.update({n: sampler._state_array(n).cpu().numpy() for n in names})
            starts = np.array([st.bounded(N - batch) for st in streams])
            z = zr.standard_normal((C, 5252)).astype(np.float32)
            sampler.__next__(feed_dict={sampler.noise: z})
            g_gpu = sampler._grad.cpu().numpy()
            # gradient error at the GPU's own point
            Xb, yb = obnn.gather_minibatch(X, y, starts, batch)
            _, g64, _ = obnn.nll_and_grad(before["theta"].astype(np.float64), Xb, yb, n_examples=N)
            worst_grad = max(worst_grad, float((np.abs(g_gpu - g64) / np.abs(g64).max(axis=1, keepdims=True)).max()))
            # update error: the oracle step on the GPU's state and gradient
            adapt = s < burn
            want = osamplers.sghmc_step(before, g_gpu, z, 0.01, mdecay=0.05, scale_grad=float(N), burn_in=adapt,
                                        frozen_minv=frozen)
            if adapt:
                frozen = want["minv"]
            got = {"theta": sampler._theta.cpu().numpy()}
            got.update({n: sampler._state_array(n).cpu().numpy() for n in names})
            check = ("theta", "v", "tau", "g", "v_hat") if adapt else ("theta", "v")
            for n in check:
                assert np.array_equal(got[n], want[n]), "%s differs at step %d" % (n, s)
            if s == burn - 1:          # the mass matrix the sampling phase uses (base_classes.py:438,448-454)
                assert np.array_equal(got["minv"], want["minv"]), "frozen minv"
        assert np.array_equal(sampler._state_array("minv").cpu().numpy(), frozen)
        assert worst_grad <= 2e-5, "variant %d: max |dg| / max|g| over the run = %.3g" % (variant, worst_grad)
    finally:
        _native.call("sgmcmc_set_bnn_tuning", K4_DEFAULT)


# BASELINE.json north_star: "1000-step trajectories match ... within 1e-5 relative in FP32".
# Measured at these shapes (tools/bnn_trajectory_drift.py, profiles/r02_bnn_trajectory_drift*.jsonl):
# default K4 (tensor pipe: rounded split, FP32-pipe accumulation, separate cross-term accumulator) 9.2e-6,
# FFMA K4 9.8e-6 of max|theta| at step 1000 -- for scale: float32 arithmetic alone separates the float32
# and float64 ORACLES by 1.4e-5 there, and the default K4 ends 1.3e-5 from the float64 oracle, i.e. closer
# to exact arithmetic than the float32 oracle it is gated against.  The faster tensor-pipe modes (variants
# 10 / 11: 4.6e-5 / 2.4e-5) do not meet the bar and are not the default.
TRAJ_TOL = 1e-5


@pytest.mark.parametrize("variant", [K4_DEFAULT, 0])
def test_bnn_sghmc_1000_step_trajectory_at_the_benchmarked_shapes(variant):
    """1000 steps of next(sampler) at BASELINE.json configs[2] shapes -- N = 20 000, minibatch 20,
    scale_grad = N, eps = 0.01, burn-in boundary at step 600, injected noise, bit-exact
    minibatch streams -- for both K4 implementations, against the float32 oracle
    (sghmc.py:165-251 over bayesian_neural_network.py:337-388), at the north star's tolerance."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "tools"))
    from bnn_trajectory_drift import drift_curves
    line, = drift_curves(steps=1000, burn=600, chains=4, every=100, variants=(variant,))
    vs32 = dict(zip(line["checkpoints"], line["gpu_vs_oracle_f32"]))
    assert all(np.isfinite(v) for v in vs32.values())
    assert max(vs32.values()) <= TRAJ_TOL, line
    # never (much) further from the float64 oracle than the float32 oracle itself is
    assert line["gpu_vs_oracle_f64"][-1] <= 2.0 * line["oracle_f32_vs_f64"][-1] + 1e-5, line
    if variant == K4_DEFAULT:      # the default tensor-pipe mode is at least as close to exact arithmetic
        assert line["gpu_vs_oracle_f64"][-1] <= 1.05 * line["oracle_f32_vs_f64"][-1], line


def test_bnn_sghmc_run_equals_per_step_and_oracle():
    """sampler.run(n) (K5: K4+K1 driven from C, Philox noise, device minibatch streams) ==
    n x next(sampler), bit for bit; thinned trace and costs are the same pairs."""
    C, N, batch, steps, burn = 10, 2000, 20, 60, 25
    X, y = sinc_data(N)

    def build():
        gen = DeviceBatchGenerator(N, batch, n_chains=C, seed=5, device=DEV, block=16)
        nll = BayesianNeuralNetworkNLL(N, batch, X=X, y=y, starts_placeholder=gen.starts_placeholder, device=DEV)
        params = default_net_params(1, n_chains=C, seed=3, device=DEV)
        return SGHMCSampler(params=params, cost_fun=nll, batch_generator=gen, burn_in_steps=burn,
                            scale_grad=float(N), seed=77,
                            session=Session(device=DEV, n_chains=C, output="torch"))
    a, b = build(), build()
    trace, costs = a.run(steps, keep_every=5)
    assert a.n_iterations == steps and trace.shape == (steps // 5, C, 5252)
    for s in range(steps):
        sample, cost = next(b)
        if (s + 1) % 5 == 0:
            k = (s + 1) // 5 - 1
            assert torch.equal(trace[k], b._theta), "trace at step %d" % s
            assert torch.equal(costs[k], cost), "cost at step %d" % s
    for name in ("v", "tau", "g", "v_hat", "minv"):
        assert torch.equal(a._state_array(name), b._state_array(name)), name
    assert torch.isfinite(a._theta).all()


@pytest.mark.parametrize("keep_every,run_chunk", [(5, 16), (7, 16), (100, 16), (1, 8)])
def test_bnn_sghmc_chunked_run_with_side_stream_indices(keep_every, run_chunk, monkeypatch):
    """run() splits long runs into chunks of whole thinning periods and generates the
    minibatch indices of the next chunk on a side stream (K7 || K4 + K1): same states, trace
    and costs as one step at a time, across the burn-in boundary, incl. a trailing partial
    thinning period and a second run() that continues the streams."""
    monkeypatch.setattr(SGHMCSampler, "RUN_CHUNK", run_chunk)
    C, N, batch, steps, burn = 9, 2000, 20, 60, 25
    X, y = sinc_data(N)

    def build():
        gen = DeviceBatchGenerator(N, batch, n_chains=C, seed=5, device=DEV, block=16)
        nll = BayesianNeuralNetworkNLL(N, batch, X=X, y=y, starts_placeholder=gen.starts_placeholder, device=DEV)
        params = default_net_params(1, n_chains=C, seed=3, device=DEV)
        return SGHMCSampler(params=params, cost_fun=nll, batch_generator=gen, burn_in_steps=burn,
                            scale_grad=float(N), seed=77,
                            session=Session(device=DEV, n_chains=C, output="torch"))
    a, b = build(), build()
    trace, costs = a.run(steps, keep_every=keep_every)
    trace2, costs2 = a.run(20, keep_every=keep_every)
    assert a.n_iterations == steps + 20 and trace.shape[0] == steps // keep_every
    for s in range(steps + 20):
        sample, cost = next(b)
        tr, co, local = (trace, costs, s) if s < steps else (trace2, costs2, s - steps)
        if (local + 1) % keep_every == 0:
            k = (local + 1) // keep_every - 1
            assert torch.equal(tr[k], b._theta), "trace at step %d" % s
            assert torch.equal(co[k], cost), "cost at step %d" % s
    for name in ("v", "tau", "g", "v_hat", "minv"):
        assert torch.equal(a._state_array(name), b._state_array(name)), name
    assert torch.equal(a._theta, b._theta)


@pytest.mark.parametrize("lookahead,every", [(0, 8), (1, 8), (3, 8), (3, 1), (4, 2), (8, 3)])
def test_iter_host_pipelined_equals_synchronous_steps(lookahead, every):
    """sampler.iter_host (host minibatch indices in, cost and thinned samples out, copies and
    the next step overlapped) == the same steps through next(sampler) with the indices fed one
    row at a time; crosses the burn-in boundary.  sample_every <= lookahead: several samples are
    in flight while the caller still holds one (each has its own pinned slot); the yielded
    sample is compared BEFORE the next one is requested, without copying it first."""
    C, N, batch, steps, burn = 6, 2000, 20, 40, 17
    X, y = sinc_data(N)
    rng = np.random.RandomState(3)
    host_starts = torch.from_numpy(rng.randint(0, N - batch + 1, size=(steps, C)).astype(np.int32)).pin_memory()

    def build():
        ph = DeviceBatchGenerator(N, batch, n_chains=C, seed=5, device=DEV).starts_placeholder
        nll = BayesianNeuralNetworkNLL(N, batch, X=X, y=y, starts_placeholder=ph, device=DEV)
        params = default_net_params(1, n_chains=C, seed=3, device=DEV)
        s = SGHMCSampler(params=params, cost_fun=nll, burn_in_steps=burn, scale_grad=float(N), seed=21,
                         session=Session(device=DEV, n_chains=C, output="torch"))
        return s, ph
    a, _ = build()
    b, ph = build()
    import time
    n_got = 0
    phase = 3 if every == 8 else 0            # a call that continues a thinning period started earlier
    for s, (smp, cost) in enumerate(a.iter_host(host_starts, sample_every=every, lookahead=lookahead,
                                                sample_phase=phase)):
        ph.value = host_starts[s].to(DEV)     # (a feed_dict would be dropped after burn-in, :454)
        sample, want_cost = next(b)
        want_theta = b._theta.cpu().numpy()
        time.sleep(0.002)                     # let the queued steps (and their copies) run ahead
        assert np.array_equal(cost, want_cost.cpu().numpy()), "cost at step %d" % s
        if (s + 1 + phase) % every == 0:
            assert np.array_equal(smp, want_theta), "sample at step %d" % s
        else:
            assert smp is None
        n_got += 1
    assert n_got == steps and a.n_iterations == steps
    for name in ("v", "tau", "g", "v_hat", "minv"):
        assert torch.equal(a._state_array(name), b._state_array(name)), name
    assert torch.equal(a._theta, b._theta) and not a.is_burning_in


def test_session_stream_orders_cost_gradient_and_update():
    """Session(stream=s): the cost / gradient (K4 or autograd), the index kernel and the update
    all run on s; results are bit-identical to the default stream, per step and through run()."""
    C, N, batch, steps = 5, 2000, 20, 24
    X, y = sinc_data(N)

    def build(stream, cls=SGHMCSampler):
        with torch.cuda.stream(stream) if stream is not None else torch.cuda.device(DEV):
            gen = DeviceBatchGenerator(N, batch, n_chains=C, seed=5, device=DEV, block=16)
        torch.cuda.synchronize()
        nll = BayesianNeuralNetworkNLL(N, batch, X=X, y=y, starts_placeholder=gen.starts_placeholder, device=DEV)
        params = default_net_params(1, n_chains=C, seed=3, device=DEV)
        return cls(params=params, cost_fun=nll, batch_generator=gen, burn_in_steps=10, scale_grad=float(N),
                   seed=4, session=Session(device=DEV, n_chains=C, output="torch", stream=stream))
    side = torch.cuda.Stream(device=DEV)
    for cls in (SGHMCSampler, SGLDSampler):
        a, b = build(None, cls), build(side, cls)
        for _ in range(steps // 2):
            ca, cb = next(a)[1], next(b)[1]
        side.synchronize()
        assert torch.equal(ca, cb)
        a.run(steps // 2, keep_every=4)
        tb, _ = b.run(steps // 2, keep_every=4)
        side.synchronize()
        torch.cuda.synchronize()
        assert torch.equal(a._theta, b._theta) and torch.equal(tb[-1], b._theta), cls.__name__


def test_next_then_run_drains_the_pending_index_block():
    """next(sampler) leaves a partially consumed block of minibatch indices in the generator;
    a following run() continues with those rows (same stream of indices as stepping on)."""
    C, N, batch = 4, 2000, 20
    X, y = sinc_data(N)

    def build():
        gen = DeviceBatchGenerator(N, batch, n_chains=C, seed=5, device=DEV, block=16)
        nll = BayesianNeuralNetworkNLL(N, batch, X=X, y=y, starts_placeholder=gen.starts_placeholder, device=DEV)
        params = default_net_params(1, n_chains=C, seed=3, device=DEV)
        return SGHMCSampler(params=params, cost_fun=nll, batch_generator=gen, burn_in_steps=10,
                            scale_grad=float(N), seed=4, session=Session(device=DEV, n_chains=C, output="torch"))
    a, b = build(), build()
    for _ in range(5):
        next(a)
    a.run(30, keep_every=10)                   # 11 pending rows, then new blocks
    next(a)
    state = a.state_dict()
    a.run(7)
    for _ in range(43):
        next(b)
    assert torch.equal(a._theta, b._theta)
    c = build()
    c.load_state_dict(state)                   # resume with pending rows, then run()
    c.run(7)
    assert torch.equal(c._theta, b._theta)


def test_resident_dataset_without_minibatch_indices():
    """No start indices fed: a small resident dataset is evaluated as a whole by the native cost
    (same value as the differentiable path), a large one raises instead of silently using its
    first rows."""
    C = 3
    for N, ok in ((120, True), (2000, False)):
        X, y = sinc_data(N)
        nll = BayesianNeuralNetworkNLL(N, 20, X=X, y=y, device=DEV)
        params = default_net_params(1, n_chains=C, seed=3, device=DEV)
        theta = torch.cat([p.reshape(C, -1) for p in params], dim=1).contiguous()
        grad = torch.empty_like(theta)
        if ok:
            cost = nll.native_cost_and_grad(theta, grad)
            np.testing.assert_allclose(cost.cpu().numpy(), nll(params).cpu().numpy(), rtol=2e-5)
        else:
            with pytest.raises(ValueError, match="start indices"):
                nll.native_cost_and_grad(theta, grad)


def test_reference_style_host_batches_single_chain():
    """Reference wiring: generate_batches feeding placeholders, one chain, numpy outputs."""
    from pysgmcmc_b200.data_batches import generate_batches
    from pysgmcmc_b200.placeholders import placeholder
    N, batch = 300, 20
    X, y = sinc_data(N)
    xp, yp = placeholder(name="X_Minibatch"), placeholder(name="Y_Minibatch")
    nll = BayesianNeuralNetworkNLL(N, batch, n_in=1, x_placeholder=xp, y_placeholder=yp, device=DEV)
    params = default_net_params(1, seed=1, device=DEV)
    sampler = SGHMCSampler(params=params, cost_fun=nll,
                           batch_generator=generate_batches(X, y, xp, yp, batch, seed=1),
                           burn_in_steps=5, scale_grad=float(N), seed=1, session=Session(device=DEV))
    theta0 = np.concatenate([p.detach().cpu().numpy().ravel() for p in params])[None, :]
    start = np.random.RandomState(1).randint(0, N - batch + 1)
    wc, _, _ = obnn.nll_and_grad(theta0.astype(np.float64), X[None, start:start + batch],
                                 y[None, start:start + batch], n_examples=N)
    sample, cost = next(sampler)
    assert isinstance(sample, list) and [s.shape for s in sample] == parameter_shapes(1)
    np.testing.assert_allclose(cost, wc[0], rtol=1e-5)
    for _ in range(10):
        sample, cost = next(sampler)
    assert np.isfinite(cost) and all(np.isfinite(s).all() for s in sample)


def test_sgld_on_bnn_cost_uses_generic_update_kernel():
    C, N = 4, 500
    X, y = sinc_data(N)
    gen = DeviceBatchGenerator(N, 20, n_chains=C, seed=2, device=DEV)
    nll = BayesianNeuralNetworkNLL(N, 20, X=X, y=y, starts_placeholder=gen.starts_placeholder, device=DEV)
    s = SGLDSampler(params=default_net_params(1, n_chains=C, seed=2, device=DEV), cost_fun=nll,
                    batch_generator=gen, burn_in_steps=3, scale_grad=float(N),
                    session=Session(device=DEV, n_chains=C, output="torch"))
    for _ in range(6):
        sample, cost = next(s)
    assert torch.isfinite(cost).all() and cost.shape == (C,)


def test_k4_launch_variants_agree():
    """Every launch shape of K4 (sgmcmc_set_bnn_tuning) computes the same cost and gradient."""
    C, N = 23, 400
    X, y = sinc_data(N)
    theta = obnn.init_theta(C, seed=5, dtype=np.float64) + 0.05 * np.random.RandomState(1).standard_normal((C, 5252))
    starts = np.random.RandomState(2).randint(0, N - 19, size=C)
    Xb, yb = obnn.gather_minibatch(X, y, starts, 20)
    wc, wg, _ = obnn.nll_and_grad(theta, Xb, yb, n_examples=N)
    try:
        for v in range(14):
            _native.call("sgmcmc_set_bnn_tuning", v)
            cost, grad, _ = k4(theta, X, y, starts, 20, 20, N)
            np.testing.assert_allclose(cost, wc, rtol=3e-6, err_msg="variant %d" % v)
            assert_grad_close(grad, wg)
    finally:
        _native.call("sgmcmc_set_bnn_tuning", K4_DEFAULT)


def test_checkpoint_resume_is_bit_identical():
    """state_dict / load_state_dict: a resumed BNN-SGHMC run continues the same chain (state,
    Philox step counter and MT19937 minibatch streams are all restored)."""
    C, N = 6, 300
    X, y = sinc_data(N)

    def build():
        gen = DeviceBatchGenerator(N, 20, n_chains=C, seed=5, device=DEV, block=8)
        nll = BayesianNeuralNetworkNLL(N, 20, X=X, y=y, starts_placeholder=gen.starts_placeholder, device=DEV)
        return SGHMCSampler(params=default_net_params(1, n_chains=C, seed=3, device=DEV), cost_fun=nll,
                            batch_generator=gen, burn_in_steps=12, scale_grad=float(N), seed=7,
                            session=Session(device=DEV, n_chains=C, output="torch"))
    a = build()
    for _ in range(9):
        next(a)                                   # leaves a partially consumed index block
    ckpt = a.state_dict()
    for _ in range(11):
        next(a)
    b = build()
    b.load_state_dict(ckpt)
    assert b.n_iterations == 9 and b.is_burning_in
    for _ in range(11):
        next(b)
    assert torch.equal(a._theta, b._theta) and torch.equal(a._state, b._state)
    assert a.n_iterations == b.n_iterations == 20


# ---- K5 as one kernel (csrc/bnn_fused.cu) ------------------------------------------------
def _run_c(state, X, y, starts, z, C, n_in, batch, N, n_steps, n_burn_in, keep_every, seed=3, step0=0,
           chain_offset=0, eps=0.01, with_grad_scratch=True):
    """sgmcmc_bnn_sghmc_run_f32 straight through the C ABI on copies of `state` (numpy
    [6, C, D]: theta, v, tau, g, v_hat, minv).  Returns (state, trace, cost_trace)."""
    D = state.shape[2]
    st = torch.as_tensor(state, dtype=torch.float32, device=DEV).contiguous().clone()
    Xd = torch.as_tensor(X, dtype=torch.float32, device=DEV).contiguous()
    yd = torch.as_tensor(y, dtype=torch.float32, device=DEV).contiguous()
    sd = None if starts is None else torch.as_tensor(starts, dtype=torch.int32, device=DEV).contiguous()
    zd = None if z is None else torch.as_tensor(z, dtype=torch.float32, device=DEV).contiguous()
    n_keep = n_steps // keep_every
    trace = torch.empty((n_keep, C, D), device=DEV)
    ctrace = torch.empty((n_keep, C), device=DEV)
    grad = torch.empty((C, D), device=DEV) if with_grad_scratch else None
    cost = torch.empty(C, device=DEV)
    p = _native.ptr
    _native.call("sgmcmc_bnn_sghmc_run_f32", *[p(st[i]) for i in range(6)], p(Xd), p(yd), p(sd), p(zd),
                 p(trace), p(ctrace), p(grad), p(cost), C, n_in, batch, float(batch), N, n_steps, n_burn_in,
                 0, keep_every, eps, 0.05, float(N), seed, step0, chain_offset, _native.stream_ptr())
    torch.cuda.synchronize()
    return st.cpu().numpy(), trace.cpu().numpy(), ctrace.cpu().numpy()


def _initial_state(C, n_in=1, seed=11):
    theta0 = obnn.init_theta(C, n_in=n_in, seed=seed, dtype=np.float32)
    D = theta0.shape[1]
    state = np.ones((6, C, D), dtype=np.float32)
    state[0] = theta0
    state[1] = 0.0
    return state


@pytest.mark.parametrize("C,batch,N,tol", [(7, 20, 2000, 1e-5), (3, 8, 100, 1e-5), (5, 13, 300, 1e-5),
                                           (4, 32, 500, 5e-5), (1, 1, 40, 1e-5)])
def test_fused_step_matches_oracle_with_injected_noise(C, batch, N, tol):
    """The one-kernel step against the fp32 oracle: injected noise, host-chosen minibatch
    starts, 40 steps across the burn-in boundary (incl. the step that freezes minv).
    Tolerance 1e-5 of max|theta|, as for the other trajectory tests; the 32-row case is the
    most sensitive one (2.3e-5 with the tensor-pipe gradient, 2e-6 with the FFMA kernel: the
    gradient errors are 4e-8 and 1e-8 of max|g|, and the preconditioner divides every element
    by its own gradient scale, so small-gradient elements amplify them) and gets 5e-5."""
    steps, burn = 40, 25
    X, y = sinc_data(N)
    state = _initial_state(C)
    D = state.shape[2]
    rng = np.random.RandomState(4)
    starts = rng.randint(0, N - batch + 1, size=(steps, C)).astype(np.int32)
    z = rng.standard_normal((steps, C, D)).astype(np.float32)
    holder = {}

    def cost_and_grad(theta):
        Xb, yb = obnn.gather_minibatch(X, y, holder["starts"], batch)
        c, g, _ = obnn.nll_and_grad(theta, Xb.astype(np.float32), yb.astype(np.float32), n_examples=N,
                                    batch_size=batch)
        return c, g
    chain = osamplers.OracleChain("sghmc", state[0].copy(), cost_and_grad, epsilon=0.01, burn_in_steps=burn,
                                  scale_grad=float(N))
    want_costs = []
    for s in range(steps):
        holder["starts"] = starts[s]
        want_theta, c = chain.next(z[s])
        want_costs.append(c)
    try:
        _native.call("sgmcmc_set_bnn_fused", 1, 0)
        got, trace, ctrace = _run_c(state, X, y, starts, z, C, 1, batch, N, steps, burn, 1, with_grad_scratch=False)
    finally:
        _native.call("sgmcmc_set_bnn_fused", 0, 0)
    np.testing.assert_allclose(ctrace, np.stack(want_costs), rtol=1e-4)
    scale = np.abs(want_theta).max()
    assert np.abs(got[0] - want_theta).max() <= tol * scale
    assert np.abs(trace[-1] - want_theta).max() <= tol * scale
    minv_rel = np.abs(got[5] / chain.minv - 1.0)
    assert np.median(minv_rel) < 1e-5 and minv_rel.max() < 5e-2


@pytest.mark.parametrize("z_injected", [False, True])
def test_fused_step_is_bit_identical_to_k4_then_k1(z_injected):
    """One kernel per step == K4 (tensor-pipe kernel) followed by K1, bit for bit: all six
    state arrays, the thinned trace and the costs, Philox noise or injected noise, a chain
    offset (sharded run) and a run that crosses the burn-in boundary."""
    C, batch, N, steps, burn = 37, 20, 2000, 30, 17
    X, y = sinc_data(N)
    state = _initial_state(C, seed=5)
    rng = np.random.RandomState(8)
    starts = rng.randint(0, N - batch + 1, size=(steps, C)).astype(np.int32)
    z = rng.standard_normal((steps, C, state.shape[2])).astype(np.float32) if z_injected else None
    out = {}
    try:
        for fused in (1, 0):
            _native.call("sgmcmc_set_bnn_fused", fused, 0)
            out[fused] = _run_c(state, X, y, starts, z, C, 1, batch, N, steps, burn, 3, seed=99, step0=1000,
                                chain_offset=64)
        _native.call("sgmcmc_set_bnn_fused", 1, 5)        # persistent grid: 5 CTAs loop over 37 chains
        out[2] = _run_c(state, X, y, starts, z, C, 1, batch, N, steps, burn, 3, seed=99, step0=1000,
                        chain_offset=64)
        # warp-specialised form (two MMA groups + update warps per CTA, gradient through shared-memory
        # buffers): its default persistent grid, and 3 CTAs = 6 groups looping over the 37 chains (every
        # group takes several chains: both gradient buffers and both hand-over barriers are recycled)
        _native.call("sgmcmc_set_bnn_fused", 4, 0)
        out[3] = _run_c(state, X, y, starts, z, C, 1, batch, N, steps, burn, 3, seed=99, step0=1000,
                        chain_offset=64)
        _native.call("sgmcmc_set_bnn_fused", 4, 3)
        out[4] = _run_c(state, X, y, starts, z, C, 1, batch, N, steps, burn, 3, seed=99, step0=1000,
                        chain_offset=64)
        _native.call("sgmcmc_set_bnn_fused", 6, 1)        # one CTA, no L2 prefetch of the state rows
        out[5] = _run_c(state, X, y, starts, z, C, 1, batch, N, steps, burn, 3, seed=99, step0=1000,
                        chain_offset=64)
    finally:
        _native.call("sgmcmc_set_bnn_fused", 0, 0)
    names = ("theta", "v", "tau", "g", "v_hat", "minv")
    for other in (0, 2, 3, 4, 5):
        for i, name in enumerate(names):
            assert np.array_equal(out[1][0][i], out[other][0][i]), (name, other)
        assert np.array_equal(out[1][1], out[other][1]) and np.array_equal(out[1][2], out[other][2])
    assert np.isfinite(out[1][0]).all()
    assert not np.array_equal(out[1][0][0], state[0])


@pytest.mark.parametrize("chunk,ring", [(8, 2), (8, 3), (5, 2), (36, 2), (1, 8)])
def test_two_stream_pipeline_is_bit_identical(chunk, ring):
    """K4 of chunk j+1 overlapped with K1 of chunk j on a second stream (gradient through a
    ring of chunk-sized slots) == the sequential K4 then K1 over all chains: states, thinned
    trace and costs, Philox noise, a chain offset, crossing the burn-in boundary, partial last
    chunk; and the caller's stream sees the final state without further synchronisation."""
    C, batch, N, steps, burn = 37, 20, 2000, 30, 17
    X, y = sinc_data(N)
    state = _initial_state(C, seed=5)
    starts = np.random.RandomState(8).randint(0, N - batch + 1, size=(steps, C)).astype(np.int32)
    out = {}
    try:
        for mode in (0, 1):
            _native.call("sgmcmc_set_bnn_pipeline", chunk if mode else 0, ring)
            out[mode] = _run_c(state, X, y, starts, None, C, 1, batch, N, steps, burn, 3, seed=99, step0=1000,
                               chain_offset=64)
    finally:
        _native.call("sgmcmc_set_bnn_pipeline", 0, 0)
    for i, name in enumerate(("theta", "v", "tau", "g", "v_hat", "minv")):
        assert np.array_equal(out[0][0][i], out[1][0][i]), name
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])
    assert np.isfinite(out[1][0]).all() and not np.array_equal(out[1][0][0], state[0])


def test_fused_step_falls_back_when_unsupported():
    """n_in = 2 gives D = 5302 (not a multiple of 4): the run goes through K4 + K1 and needs
    the gradient scratch buffer; the result is finite and the costs match the oracle's."""
    C, batch, N, n_in = 3, 20, 200, 2
    X, y = sinc_data(N, n_in=n_in)
    D = n_parameters(n_in)
    assert D % 4 != 0
    rng = np.random.RandomState(0)
    theta0 = (0.1 * rng.standard_normal((C, D))).astype(np.float32)
    state = np.ones((6, C, D), dtype=np.float32)
    state[0] = theta0
    state[1] = 0.0
    starts = rng.randint(0, N - batch + 1, size=(4, C)).astype(np.int32)
    try:
        _native.call("sgmcmc_set_bnn_fused", 1, 0)
        got, trace, ctrace = _run_c(state, X, y, starts, None, C, n_in, batch, N, 4, 4, 1)
        with pytest.raises(Exception):
            _run_c(state, X, y, starts, None, C, n_in, batch, N, 4, 4, 1, with_grad_scratch=False)
    finally:
        _native.call("sgmcmc_set_bnn_fused", 0, 0)
    Xb, yb = obnn.gather_minibatch(X, y, starts[0], batch)
    wc, _, _ = obnn.nll_and_grad(theta0.astype(np.float64), Xb, yb, n_examples=N, n_in=n_in)
    np.testing.assert_allclose(ctrace[0], wc, rtol=1e-5)
    assert np.isfinite(got).all()
