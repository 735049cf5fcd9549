"""GPU parity tests of the resident BNN-SGHMC kernel (csrc/bnn_resident.cu, K5r): every chain on one SM
for a whole block of steps, through the C ABI (sgmcmc_bnn_sghmc_run_resident_f32).

* gradient of a step == the FFMA K4 (variant 0) bit for bit except d/d rho (same accumulation orders), and
  within the K4 tolerances of the float64 oracle; cost within rtol 3e-6;
* the update == the oracle's sghmc step (sghmc.py:165-251) on the kernel's own gradient, bit for bit, in
  burn-in and in the sampling phase (teacher-forced);
* a block of n steps == n calls of one step (states, thinned trace, costs), across the burn-in boundary;
* 1000-step trajectory at the benchmarked shapes within 1e-5 of the float32 oracle (the north star's bar).
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import bnn as obnn, samplers as osamplers
from pysgmcmc_b200 import _native

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
D = 5252
NAMES = ("theta", "v", "tau", "g", "v_hat", "minv")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))


def sinc_data(N, n_in=1, seed=1):
    rng = np.random.RandomState(seed)
    X = np.array([rng.uniform(0.0, 1.0, n_in) for _ in range(N)])
    y = np.sinc(X * 10 - 5).sum(axis=1)
    X = (X - X.mean(axis=0)) / X.std(axis=0)
    y = (y - y.mean()) / y.std()
    return X.astype(np.float32), y.astype(np.float32)


def fresh_state(C, seed=11, n_in=1):
    theta = obnn.init_theta(C, n_in=n_in, seed=seed, dtype=np.float32)
    st = {"theta": theta, "v": np.zeros_like(theta), "tau": np.ones_like(theta), "g": np.ones_like(theta),
          "v_hat": np.ones_like(theta), "minv": np.ones_like(theta)}
    return {k: torch.tensor(v, device=DEV) for k, v in st.items()}


def run_resident(state, X, y, starts, z, n_steps, n_burn_in, keep_every=1, batch=20, N=None, eps=0.01,
                 want_trace=False, want_grad=False, want_all=False, seed=5, step0=0, chain_offset=0, n_in=None,
                 adapt_forever=0):
    """One call of the C entry point on device copies of `state` (dict of [C, D] tensors, updated in place)."""
    C, Dn = state["theta"].shape
    N = N if N is not None else X.shape[0]
    n_in = X.shape[1] if n_in is None else n_in
    Xd = torch.as_tensor(X, dtype=torch.float32, device=DEV).contiguous()
    yd = torch.as_tensor(y, dtype=torch.float32, device=DEV).contiguous()
    sd = None if starts is None else torch.as_tensor(np.ascontiguousarray(starts), dtype=torch.int32, device=DEV)
    zd = None if z is None else torch.as_tensor(np.ascontiguousarray(z), dtype=torch.float32, device=DEV)
    n_keep = n_steps // keep_every
    trace = torch.full((max(n_keep, 1), C, Dn), float("nan"), device=DEV) if want_trace else None
    cost_trace = torch.full((max(n_keep, 1), C), float("nan"), device=DEV) if want_trace else None
    cost_all = torch.full((max(n_steps, 1), C), float("nan"), device=DEV) if want_all else None
    cost_last = torch.full((C,), float("nan"), device=DEV)
    grad = torch.full((C, Dn), float("nan"), device=DEV) if want_grad else None
    _native.call("sgmcmc_bnn_sghmc_run_resident_f32", *[_native.ptr(state[n]) for n in NAMES],
                 _native.ptr(Xd), _native.ptr(yd), _native.ptr(sd), _native.ptr(zd), _native.ptr(trace),
                 _native.ptr(cost_trace), _native.ptr(cost_all), _native.ptr(cost_last), _native.ptr(grad),
                 C, n_in, batch, float(batch), N, n_steps, n_burn_in, adapt_forever, keep_every,
                 eps, 0.05, float(N), seed, step0, chain_offset, _native.stream_ptr())
    torch.cuda.synchronize()
    return {"trace": trace, "cost_trace": cost_trace, "cost_all": cost_all, "cost_last": cost_last, "grad": grad}


def ffma_k4(theta, X, y, starts, batch, N):
    C, n_in = theta.shape[0], X.shape[1]
    t = theta.contiguous()
    Xd = torch.as_tensor(X, dtype=torch.float32, device=DEV).contiguous()
    yd = torch.as_tensor(y, dtype=torch.float32, device=DEV).contiguous()
    sd = torch.as_tensor(starts, dtype=torch.int32, device=DEV)
    cost, grad = torch.empty(C, device=DEV), torch.empty_like(t)
    _native.call("sgmcmc_set_bnn_tuning", 0)
    try:
        _native.call("sgmcmc_bnn_nll_grad_f32", _native.ptr(t), _native.ptr(Xd), _native.ptr(yd), _native.ptr(sd),
                     _native.ptr(cost), _native.ptr(grad), None, C, n_in, batch, float(batch), N, _native.stream_ptr())
        torch.cuda.synchronize()
    finally:
        _native.call("sgmcmc_set_bnn_tuning", 16)
    return cost.cpu().numpy(), grad.cpu().numpy()


@pytest.fixture(params=[1, 0], ids=["overlap", "no-overlap"])
def overlap(request):
    """The update's gradient-free part beside the gradient (default where it fits shared memory: batch <= 20),
    or the whole update after the gradient; both must give the same bits."""
    _native.call("sgmcmc_set_bnn_resident_overlap", request.param)
    yield request.param
    _native.call("sgmcmc_set_bnn_resident_overlap", 1)


@pytest.mark.parametrize("n_in", [1, 2, 5])       # D % 4 == 0 / 2 / 0: every chain's last element group full or half padding
@pytest.mark.parametrize("batch", [20, 32, 7])
def test_resident_gradient_equals_the_ffma_kernel_and_the_oracle(overlap, batch, n_in):
    C, N = 5, 2000
    X, y = sinc_data(N, n_in=n_in)
    D = 50 * n_in + 5202
    rng = np.random.RandomState(batch)
    starts = rng.randint(0, N - batch + 1, size=(1, C))
    st = fresh_state(C, n_in=n_in)
    theta0 = st["theta"].clone()
    out = run_resident(st, X, y, starts, np.zeros((1, C, D), np.float32), 1, 1, batch=batch, want_grad=True)
    g = out["grad"].cpu().numpy()
    cost_ffma, g_ffma = ffma_k4(theta0, X, y, starts[0], batch, N)
    scalars = [D - 2, D - 1]            # b4 and rho: scalar expressions whose FMA contraction is the compiler's choice
    mask = np.ones(D, bool)
    mask[scalars] = False
    assert np.array_equal(g[:, mask], g_ffma[:, mask]), "gradient differs from the FFMA kernel (same accumulation order)"
    np.testing.assert_allclose(g[:, scalars], g_ffma[:, scalars], rtol=1e-5)
    np.testing.assert_allclose(out["cost_last"].cpu().numpy(), cost_ffma, rtol=3e-6)
    Xb, yb = obnn.gather_minibatch(X.astype(np.float64), y.astype(np.float64), starts[0], batch)
    c64, g64, _ = obnn.nll_and_grad(theta0.cpu().numpy().astype(np.float64), Xb, yb, n_examples=N, batch_size=batch,
                                    n_in=n_in)
    assert (np.abs(g - g64) / np.abs(g64).max(axis=1, keepdims=True)).max() <= 2e-5
    np.testing.assert_allclose(out["cost_last"].cpu().numpy(), c64, rtol=3e-6)


@pytest.mark.parametrize("n_in", [1, 2])
def test_resident_update_is_the_oracle_update_on_its_own_gradient(overlap, n_in):
    """Teacher-forced, step by step, across the burn-in boundary: theta, V, tau, g, v_hat and the frozen
    inverse mass matrix after a one-step call == oracle step on the state before and the kernel's gradient."""
    C, N, batch, steps, burn = 5, 2000, 20, 12, 7
    X, y = sinc_data(N, n_in=n_in)
    D = 50 * n_in + 5202
    rng = np.random.RandomState(2)
    st = fresh_state(C, n_in=n_in)
    frozen = None
    for s in range(steps):
        before = {k: v.cpu().numpy() for k, v in st.items()}
        starts = rng.randint(0, N - batch + 1, size=(1, C))
        z = rng.standard_normal((1, C, D)).astype(np.float32)
        adapt = s < burn
        out = run_resident(st, X, y, starts, z, 1, 1 if adapt else 0, want_grad=True, step0=s)
        want = osamplers.sghmc_step(before, out["grad"].cpu().numpy(), z[0], 0.01, mdecay=0.05, scale_grad=float(N),
                                    burn_in=adapt, frozen_minv=frozen)
        if adapt:
            frozen = want["minv"]
        for n in (("theta", "v", "tau", "g", "v_hat") if adapt else ("theta", "v")):
            assert np.array_equal(st[n].cpu().numpy(), want[n]), "%s differs at step %d" % (n, s)
        if s == burn - 1:
            # n_burn_in == n_steps == 1: this was the last burn-in step of the call, minv is stored
            assert np.array_equal(st["minv"].cpu().numpy(), want["minv"])
        if not adapt:
            for n in ("tau", "g", "v_hat", "minv"):
                assert np.array_equal(st[n].cpu().numpy(), before[n]), "%s must not change after burn-in" % n


@pytest.mark.parametrize("n_in", [1, 2])
@pytest.mark.parametrize("use_z", [True, False])
def test_resident_block_of_steps_equals_single_steps(overlap, use_z, n_in):
    """n steps in one call (state on the SM throughout) == n calls of one step: states, thinned trace, costs;
    Philox noise (counter = element group, step) or injected noise; burn-in ends inside the block."""
    C, N, batch, steps, burn, keep = 7, 2000, 20, 24, 10, 4
    X, y = sinc_data(N, n_in=n_in)
    D = 50 * n_in + 5202
    rng = np.random.RandomState(4)
    starts = rng.randint(0, N - batch + 1, size=(steps, C))
    z = rng.standard_normal((steps, C, D)).astype(np.float32) if use_z else None
    a, b = fresh_state(C, n_in=n_in), fresh_state(C, n_in=n_in)
    out = run_resident(a, X, y, starts, z, steps, burn, keep_every=keep, want_trace=True, want_all=True, step0=100,
                       chain_offset=8)
    for s in range(steps):
        o = run_resident(b, X, y, starts[s:s + 1], None if z is None else z[s:s + 1], 1, 1 if s < burn else 0,
                         step0=100 + s, chain_offset=8)
        assert torch.equal(out["cost_all"][s], o["cost_last"]), "cost of step %d" % s
        if (s + 1) % keep == 0:
            k = (s + 1) // keep - 1
            assert torch.equal(out["trace"][k], b["theta"]), "sample after step %d" % s
            assert torch.equal(out["cost_trace"][k], o["cost_last"])
    assert torch.equal(out["cost_last"], o["cost_last"])
    for n in NAMES:
        assert torch.equal(a[n], b[n]), n
    assert torch.isfinite(a["theta"]).all()


def test_resident_noise_is_k1_noise(overlap):
    """Without injected noise the kernel draws K1's Philox normals: one resident step == K1
    (sgmcmc_sghmc_step_f32) on the resident kernel's gradient, with a chain offset."""
    C, N, batch = 6, 2000, 20
    X, y = sinc_data(N)
    starts = np.random.RandomState(6).randint(0, N - batch + 1, size=(1, C))
    a, b = fresh_state(C), fresh_state(C)
    out = run_resident(a, X, y, starts, None, 1, 1, want_grad=True, seed=31, step0=17, chain_offset=12)
    _native.call("sgmcmc_sghmc_step_f32", *[_native.ptr(b[n]) for n in NAMES], _native.ptr(out["grad"]), None,
                 C * D, 0.01, 0.05, float(N), 1, 1, 31, 17, 12 * D, _native.stream_ptr())
    torch.cuda.synchronize()
    for n in NAMES:
        assert torch.equal(a[n], b[n]), n


def test_resident_1000_step_trajectory_at_the_benchmarked_shapes():
    from bnn_trajectory_drift import drift_curves
    line, = drift_curves(steps=1000, burn=600, chains=4, every=100, variants=("resident",))
    vs32 = dict(zip(line["checkpoints"], line["gpu_vs_oracle_f32"]))
    assert all(np.isfinite(v) for v in vs32.values())
    assert max(vs32.values()) <= 1e-5, line
    assert line["gpu_vs_oracle_f64"][-1] <= 2.0 * line["oracle_f32_vs_f64"][-1] + 1e-5, line


def test_resident_refuses_shapes_it_cannot_hold():
    assert _native.load().sgmcmc_bnn_resident_supported(1, 20) == 1
    assert _native.load().sgmcmc_bnn_resident_supported(3, 32) == 1
    assert _native.load().sgmcmc_bnn_resident_supported(2, 20) == 1      # D % 4 == 2: half-padded last groups
    assert _native.load().sgmcmc_bnn_resident_supported(1, 33) == 0
    st = fresh_state(2)
    X, y = sinc_data(100)
    with pytest.raises(_native.NativeError):
        run_resident(st, X, y, None, None, 1, 1, batch=64)


# ---- the sampler classes on top: with few chains every entry point runs the resident kernel ----------------
def _sampler(C, N, batch, burn, X, y, generator=True, seed=77, limit=None):
    from pysgmcmc_b200 import Session
    from pysgmcmc_b200.data_batches import DeviceBatchGenerator
    from pysgmcmc_b200.models.bnn_cost import BayesianNeuralNetworkNLL, default_net_params
    from pysgmcmc_b200.samplers import SGHMCSampler
    gen = DeviceBatchGenerator(N, batch, n_chains=C, seed=5, device=DEV, block=16)
    nll = BayesianNeuralNetworkNLL(N, batch, X=X, y=y, starts_placeholder=gen.starts_placeholder, device=DEV)
    params = default_net_params(X.shape[1], n_chains=C, seed=3, device=DEV)
    s = SGHMCSampler(params=params, cost_fun=nll, batch_generator=gen if generator else None, burn_in_steps=burn,
                     scale_grad=float(N), seed=seed, session=Session(device=DEV, n_chains=C, output="torch"))
    if limit is not None:
        s.RESIDENT_MAX_CHAINS = limit
    return s, gen.starts_placeholder


@pytest.mark.parametrize("n_in", [1, 2])
@pytest.mark.parametrize("keep_every", [5, 7])
def test_sampler_run_equals_next_with_the_resident_kernel(keep_every, n_in):
    """run(n) (chunks of steps with the chains on their SMs) == n x next() (one launch per step), bit for bit:
    states, thinned trace, costs; the burn-in ends inside; one launch per next() besides the index generator."""
    C, N, batch, steps, burn = 11, 2000, 20, 60, 25
    X, y = sinc_data(N, n_in=n_in)
    a, _ = _sampler(C, N, batch, burn, X, y)
    b, _ = _sampler(C, N, batch, burn, X, y)
    assert a._resident_ok(batch)
    trace, costs = a.run(steps, keep_every=keep_every)
    assert a.n_iterations == steps and trace.shape == (steps // keep_every, C, 50 * n_in + 5202)
    launches0 = _native.load().sgmcmc_launch_count()
    for s in range(steps):
        sample, cost = next(b)
        if (s + 1) % keep_every == 0:
            k = (s + 1) // keep_every - 1
            assert torch.equal(trace[k], b._theta), "trace at step %d" % s
            assert torch.equal(costs[k], cost), "cost at step %d" % s
    n_launches = _native.load().sgmcmc_launch_count() - launches0
    assert n_launches < 1.2 * steps, "%d launches for %d steps: K4 + K1 ran instead of the resident kernel" % (n_launches, steps)
    for name in ("v", "tau", "g", "v_hat", "minv"):
        assert torch.equal(a._state_array(name), b._state_array(name)), name
    assert torch.equal(a._theta, b._theta) and torch.isfinite(a._theta).all() and not a.is_burning_in


@pytest.mark.parametrize("blocks", [True, False], ids=["blocks", "per-step"])
@pytest.mark.parametrize("lookahead,every", [(3, 8), (0, 8), (3, 1), (4, 2), (8, 3), (5, None)])
def test_sampler_iter_host_equals_next_with_the_resident_kernel(lookahead, every, blocks):
    """iter_host on the resident kernel -- blocks of up to lookahead + 1 steps per launch (ending at sample steps),
    or one launch per step -- == next(sampler) fed the same index rows: costs of every step, thinned samples,
    final state; crosses the end of burn-in; a call that continues a thinning period (sample_phase)."""
    C, N, batch, steps, burn = 6, 2000, 20, 43, 17
    X, y = sinc_data(N)
    rng = np.random.RandomState(3)
    host_starts = torch.from_numpy(rng.randint(0, N - batch + 1, size=(steps, C)).astype(np.int32)).pin_memory()
    a, _ = _sampler(C, N, batch, burn, X, y, generator=False, seed=21)
    b, ph = _sampler(C, N, batch, burn, X, y, generator=False, seed=21)
    a.RESIDENT_HOST_BLOCKS = blocks
    phase = 3 if every == 8 else 0
    n_got = 0
    for s, (smp, cost) in enumerate(a.iter_host(host_starts, sample_every=every, lookahead=lookahead,
                                                sample_phase=phase)):
        ph.value = host_starts[s].to(DEV)
        _, want_cost = next(b)
        assert np.array_equal(cost, want_cost.cpu().numpy()), "cost at step %d" % s
        if every and (s + 1 + phase) % every == 0:
            assert np.array_equal(smp, b._theta.cpu().numpy()), "sample at step %d" % s
        else:
            assert smp is None
        n_got += 1
    assert n_got == steps and a.n_iterations == steps
    for name in ("v", "tau", "g", "v_hat", "minv"):
        assert torch.equal(a._state_array(name), b._state_array(name)), name
    assert torch.equal(a._theta, b._theta) and not a.is_burning_in


def test_resident_and_streaming_samplers_agree_to_rounding():
    """The same sampler with the resident kernel and with K4 then K1: different rounding of the gradient's dot
    products, nothing else -- 30 steps apart by ~1e-6 of max|theta|."""
    C, N, batch, steps = 8, 2000, 20, 30
    X, y = sinc_data(N)
    a, _ = _sampler(C, N, batch, 20, X, y)
    b, _ = _sampler(C, N, batch, 20, X, y, limit=0)
    assert a._resident_ok(batch) and not b._resident_ok(batch)
    a.run(steps, keep_every=steps)
    b.run(steps, keep_every=steps)
    ta, tb = a._theta.cpu().numpy(), b._theta.cpu().numpy()
    assert np.abs(ta - tb).max() <= 1e-5 * np.abs(tb).max()
    assert not np.array_equal(ta, tb)          # (they ARE two arithmetics; identical bits would mean one path ran twice)
