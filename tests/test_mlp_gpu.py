"""GPU parity tests of the layer kernels for arbitrary fully connected architectures
(csrc/mlp.cu: K4 / K10 for any `MLPNet`, incl. the wide 1000-512-512 network of BASELINE.json
configs[4]) against the oracle (oracle/bnn.py restates
pysgmcmc/models/bayesian_neural_network.py:28-69,77-141,337-388 for any tuple of hidden widths).

Tolerances as for the specialised K4 (tests/test_bnn_gpu.py): cost rtol 3e-6 against the float64
oracle, gradient |g - g_ref| <= 2e-5 max|g_ref| per chain (FP32 FFMA accumulation over up to
1000-term dot products, tanh on the SFU); trajectories 1e-5 of max|theta|.
"""
import numpy as np
import pytest
import torch

from oracle import bnn as obnn, mt19937 as omt, samplers as osamplers
from pysgmcmc_b200 import Session, _native
from pysgmcmc_b200.data_batches import DeviceBatchGenerator
from pysgmcmc_b200.models import BayesianNeuralNetwork, MLPNet, TorchNet
from pysgmcmc_b200.models.bnn_cost import BayesianNeuralNetworkNLL
from pysgmcmc_b200.samplers import SGHMCSampler
from pysgmcmc_b200.stepsize_schedules import ConstantStepsizeSchedule

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def sinc_data(N, n_in=1, seed=1):
    rng = np.random.RandomState(seed)
    X = np.array([rng.uniform(0.0, 1.0, n_in) for _ in range(N)])
    y = np.sinc(X * 10 - 5).sum(axis=1)
    X = (X - X.mean(axis=0)) / X.std(axis=0)
    y = (y - y.mean()) / y.std()
    return X, y


def mlp_k4(theta, X, y, starts, widths, batch, bs_cfg, N, want_grad=True):
    C = theta.shape[0]
    t = torch.as_tensor(theta, dtype=torch.float32, device=DEV).contiguous()
    Xd = torch.as_tensor(X, dtype=torch.float32, device=DEV).contiguous()
    yd = torch.as_tensor(y, dtype=torch.float32, device=DEV).contiguous()
    sd = None if starts is None else torch.as_tensor(starts, dtype=torch.int32, device=DEV)
    cost, mse = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    grad = torch.full_like(t, float("nan")) if want_grad else None
    w, n_w = _native.int_array(widths)
    nbytes = int(_native.load().sgmcmc_mlp_workspace_bytes(w, n_w, C, batch))
    assert nbytes > 0
    ws = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=DEV)
    _native.call("sgmcmc_mlp_nll_grad_f32", _native.ptr(t), _native.ptr(Xd), _native.ptr(yd), _native.ptr(sd),
                 _native.ptr(cost), _native.ptr(grad), _native.ptr(mse), _native.ptr(ws), ws.numel() * 8, C,
                 w, n_w, batch, float(bs_cfg), N, _native.stream_ptr())
    torch.cuda.synchronize()
    return cost.cpu().numpy(), None if grad is None else grad.cpu().numpy(), mse.cpu().numpy()


ARCHS = [
    # (hidden, n_in, chains, batch, N)
    ((50, 50, 50), 1, 5, 20, 2000),          # the default network through the generic kernels
    ((1000, 512, 512), 1, 3, 20, 20000),     # BASELINE.json configs[4]: D = 777 682 (odd chains 8-byte aligned only)
    ((64,), 3, 4, 8, 500),                   # one hidden layer, several inputs
    ((300, 7, 260, 5), 2, 3, 13, 700),       # ragged widths (scalar paths), odd batch
    ((256, 256), 4, 2, 32, 900),             # the largest minibatch
    ((33,), 1, 1, 1, 40),                    # a single row
]


@pytest.mark.parametrize("hidden,n_in,C,batch,N", ARCHS)
def test_mlp_cost_and_gradient_match_the_oracle(hidden, n_in, C, batch, N):
    X, y = sinc_data(N, n_in)
    theta = obnn.init_theta(C, n_in=n_in, hidden=hidden, seed=3, dtype=np.float64)
    theta += 0.05 * np.random.RandomState(4).standard_normal(theta.shape)      # biases and rho off their defaults
    starts = np.random.RandomState(5).randint(0, N - batch + 1, size=C)
    widths = [n_in] + list(hidden) + [1]
    assert _native.load().sgmcmc_mlp_n_params(*_native.int_array(widths)) == theta.shape[1]
    cost, grad, mse = mlp_k4(theta, X, y, starts, widths, batch, 20, N)
    Xb, yb = obnn.gather_minibatch(X, y, starts, batch)
    wc, wg, wm = obnn.nll_and_grad(theta.astype(np.float32).astype(np.float64), Xb, yb, n_examples=N,
                                   batch_size=20, n_in=n_in, hidden=hidden)
    np.testing.assert_allclose(cost, wc, rtol=3e-6)
    np.testing.assert_allclose(mse, wm, rtol=2e-5)
    assert np.isfinite(grad).all()
    err = np.abs(grad - wg).max(axis=1) / np.abs(wg).max(axis=1)
    assert err.max() <= 2e-5, "max |dg| / max|g| = %.3g" % err.max()
    # cost only (no gradient buffer): same cost
    cost2, _, _ = mlp_k4(theta, X, y, starts, widths, batch, 20, N, want_grad=False)
    np.testing.assert_allclose(cost, cost2, rtol=1e-6)


@pytest.mark.parametrize("hidden,n_in,C,batch,N", [
    ((1000, 512, 512), 1, 3, 20, 20000),     # both wide layers forward and backward on the tensor cores
    ((256, 256), 4, 2, 32, 900),             # the largest minibatch
    ((128, 260, 128), 2, 3, 13, 700),        # partial 128-unit tiles, contraction lengths that are not multiples of 16
    ((512, 7, 256, 256), 1, 2, 20, 500),     # a narrow layer in between: only the last matrix qualifies
    ((256, 256, 64), 1, 2, 8, 300),          # forward on the tensor cores, backward from an FFMA layer above
])
def test_tensor_core_layers_agree_with_the_ffma_layers(hidden, n_in, C, batch, N):
    """csrc/mlp_umma.cu (tcgen05, 3xTF32) against the FFMA layer kernels on the same inputs, and both
    against the float64 oracle."""
    X, y = sinc_data(N, n_in)
    theta = obnn.init_theta(C, n_in=n_in, hidden=hidden, seed=3, dtype=np.float64)
    theta += 0.05 * np.random.RandomState(4).standard_normal(theta.shape)
    starts = np.random.RandomState(5).randint(0, N - batch + 1, size=C)
    widths = [n_in] + list(hidden) + [1]
    try:
        _native.call("sgmcmc_set_mlp_tuning", 0)
        c0, g0, m0 = mlp_k4(theta, X, y, starts, widths, batch, 20, N)
    finally:
        _native.call("sgmcmc_set_mlp_tuning", 1)
    c1, g1, m1 = mlp_k4(theta, X, y, starts, widths, batch, 20, N)
    Xb, yb = obnn.gather_minibatch(X, y, starts, batch)
    wc, wg, wm = obnn.nll_and_grad(theta.astype(np.float32).astype(np.float64), Xb, yb, n_examples=N,
                                   batch_size=20, n_in=n_in, hidden=hidden)
    for cost, grad, mse in ((c0, g0, m0), (c1, g1, m1)):
        np.testing.assert_allclose(cost, wc, rtol=3e-6)
        np.testing.assert_allclose(mse, wm, rtol=2e-5)
        assert np.isfinite(grad).all()
        err = np.abs(grad - wg).max(axis=1) / np.abs(wg).max(axis=1)
        assert err.max() <= 2e-5, "max |dg| / max|g| = %.3g" % err.max()
    assert (np.abs(g1 - g0).max(axis=1) <= 1e-5 * np.abs(g0).max(axis=1)).all()


def test_generic_kernels_agree_with_the_specialised_k4():
    """Same network, same inputs: csrc/mlp.cu against the tensor-pipe K4 (bnn_mma.cuh)."""
    C, N, batch = 16, 2000, 20
    X, y = sinc_data(N)
    theta = obnn.init_theta(C, seed=8, dtype=np.float32)
    starts = np.random.RandomState(5).randint(0, N - batch + 1, size=C)
    cost, grad, _ = mlp_k4(theta, X, y, starts, [1, 50, 50, 50, 1], batch, 20, N)
    t = torch.as_tensor(theta, device=DEV)
    c2, g2 = torch.empty(C, device=DEV), torch.empty_like(t)
    # (named tensors: a temporary would be freed, and its block re-used, before the kernel runs)
    Xd, yd = torch.as_tensor(X, dtype=torch.float32, device=DEV), torch.as_tensor(y, dtype=torch.float32, device=DEV)
    sd = torch.as_tensor(starts, dtype=torch.int32, device=DEV)
    _native.call("sgmcmc_bnn_nll_grad_f32", _native.ptr(t), _native.ptr(Xd), _native.ptr(yd), _native.ptr(sd),
                 _native.ptr(c2), _native.ptr(g2), None, C, 1, batch, 20.0, N, _native.stream_ptr())
    np.testing.assert_allclose(cost, c2.cpu().numpy(), rtol=3e-6)
    g2 = g2.cpu().numpy()
    assert (np.abs(grad - g2).max(axis=1) <= 2e-5 * np.abs(g2).max(axis=1)).all()


@pytest.mark.parametrize("hidden,n_in,n_nets,n_points", [((1000, 512, 512), 1, 3, 77), ((40, 30), 5, 4, 200),
                                                         ((50, 50, 50), 2, 2, 32)])
def test_mlp_predict_matches_the_oracle(hidden, n_in, n_nets, n_points):
    theta = obnn.init_theta(n_nets, n_in=n_in, hidden=hidden, seed=2, dtype=np.float32)
    theta[:, -1] = np.linspace(-3, -1, n_nets)
    Xt = np.random.RandomState(0).uniform(-2, 2, size=(n_points, n_in)).astype(np.float32)
    w, n_w = _native.int_array([n_in] + list(hidden) + [1])
    items = n_nets * ((n_points + 31) // 32)
    ws = torch.empty((int(_native.load().sgmcmc_mlp_workspace_bytes(w, n_w, items, 32)) + 7) // 8, dtype=torch.int64,
                     device=DEV)
    out = torch.full((n_nets, n_points, 2), float("nan"), device=DEV)
    td, Xd = torch.as_tensor(theta, device=DEV), torch.as_tensor(Xt, device=DEV)
    _native.call("sgmcmc_mlp_predict_f32", _native.ptr(td), _native.ptr(Xd), _native.ptr(out), _native.ptr(ws),
                 ws.numel() * 8, n_nets, w, n_w, n_points, _native.stream_ptr())
    f, rho, _ = obnn.forward(theta.astype(np.float64), np.broadcast_to(Xt.astype(np.float64), (n_nets,) + Xt.shape),
                             n_in, hidden)
    got = out.cpu().numpy()
    np.testing.assert_allclose(got[:, :, 0], f, atol=2e-5 * max(1.0, np.abs(f).max()))
    np.testing.assert_allclose(got[:, :, 1], np.broadcast_to(rho[:, None], f.shape), rtol=1e-7)


def test_wide_net_sghmc_trajectory_matches_the_oracle():
    """next(sampler) on the 1000-512-512 network (D = 777 682): 60 steps across the burn-in boundary
    with injected noise and bit-exact minibatch streams against the float32 oracle.  Stepsize 0.002:
    with the default network's 0.01 this much wider net overshoots in its first steps (cost 250 ->
    65 000) and the float32 and float64 ORACLES themselves then differ by 0.9 % in cost at step 28;
    at 0.002 they agree to 1e-6 in cost and 2e-7 in theta over these steps."""
    hidden, C, N, batch, steps, burn, eps = (1000, 512, 512), 2, 20000, 20, 60, 40, 0.002
    X, y = sinc_data(N)
    net = MLPNet(hidden)
    D = net.n_parameters(1)
    assert D == 777682
    theta0 = obnn.init_theta(C, hidden=hidden, seed=11, dtype=np.float32)
    seeds = np.arange(C) + 40
    streams = [omt.MT19937(int(s)) for s in seeds]
    holder = {}

    def cost_and_grad(theta):
        Xb, yb = obnn.gather_minibatch(X, y, holder["starts"], batch)
        c, g, _ = obnn.nll_and_grad(theta, Xb.astype(np.float32), yb.astype(np.float32), n_examples=N, hidden=hidden)
        return c, g
    chain = osamplers.OracleChain("sghmc", theta0, cost_and_grad, epsilon=eps, burn_in_steps=burn,
                                  scale_grad=float(N))
    gen = DeviceBatchGenerator(N, batch, seeds=seeds, device=DEV, block=64)
    nll = BayesianNeuralNetworkNLL(N, batch, X=X, y=y, starts_placeholder=gen.starts_placeholder, device=DEV, net=net)
    assert nll.supports_native and not nll.bnn_native
    params, off = [], 0
    for shp in net.parameter_shapes(1):
        n = int(np.prod(shp))
        params.append(torch.tensor(theta0[:, off:off + n].reshape((C,) + shp), device=DEV))
        off += n
    sampler = SGHMCSampler(params=params, cost_fun=nll, batch_generator=gen, burn_in_steps=burn,
                           scale_grad=float(N), stepsize_schedule=ConstantStepsizeSchedule(eps),
                           session=Session(device=DEV, n_chains=C, output="torch"))
    zr = np.random.RandomState(9)
    for s in range(steps):
        holder["starts"] = np.array([st.bounded(N - batch) for st in streams])
        z = zr.standard_normal((C, D)).astype(np.float32)
        want_theta, want_cost = chain.next(z)
        sample, cost = sampler.__next__(feed_dict={sampler.noise: z})
        np.testing.assert_allclose(cost.cpu().numpy(), want_cost, rtol=1e-4, err_msg="step %d" % s)
    got = sampler._theta.cpu().numpy()
    assert np.abs(got - want_theta).max() <= 1e-5 * np.abs(want_theta).max()
    assert not sampler.is_burning_in
    # run() drives the same kernels from the device loop
    trace, costs = sampler.run(4, keep_every=2)
    assert trace.shape == (2, C, D) and torch.isfinite(trace).all()


@pytest.mark.parametrize("net,dtype", [(MLPNet((64, 32)), torch.float32), (MLPNet((50, 50, 50)), torch.float64),
                                        ("torchnet", torch.float32)])
def test_bayesian_neural_network_with_other_architectures(net, dtype):
    """BayesianNeuralNetwork(get_net=...) with a custom MLP (native layer kernels), a float64
    model (differentiable cost + float64 update kernel) and an arbitrary torch function
    (autograd): the reference's accuracy bar on sinc
    (tests/bayesian_neural_network/test_train_predict.py:12-48: MSE <= 0.1)."""
    if net == "torchnet":
        def forward(x, params):
            W1, b1, W2, b2, rho = params
            chains = W1.dim() == 3
            bias = (lambda b: b[:, None, :]) if chains else (lambda b: b)
            h = torch.nn.functional.softplus(x @ W1 + bias(b1))
            f = h @ W2 + bias(b2)
            return torch.cat([f, torch.ones_like(f) * rho], dim=-1)

        net = TorchNet(forward, lambda n_in: [(n_in, 40), (40,), (40, 1), (1,), (1, 1)])
    rng = np.random.RandomState(1)
    X = rng.uniform(0, 1, size=(100, 1))
    y = np.sinc(X * 10 - 5).sum(axis=1)
    bnn = BayesianNeuralNetwork(session=Session(device=DEV), get_net=net, n_nets=10, n_iters=3000, burn_in_steps=1000,
                                sample_steps=100, seed=1, dtype=dtype)
    bnn.train(X, y)
    Xt = rng.uniform(0, 1, size=(100, 1))
    mean, var = bnn.predict(Xt)
    mse = float(np.mean((mean - np.sinc(Xt * 10 - 5).sum(axis=1)) ** 2))
    assert np.isfinite(mean).all() and (var >= 0).all()
    assert mse <= 0.1, mse
    f_out, noise = bnn.predict(Xt, return_individual_predictions=True)
    assert f_out.shape == (10, 100)


def test_get_net_without_parameters_is_rejected_with_an_explanation():
    with pytest.raises(ValueError, match="parameters are explicit"):
        BayesianNeuralNetwork(session=Session(device=DEV), get_net=lambda inputs, seed=None, dtype=None: inputs)
