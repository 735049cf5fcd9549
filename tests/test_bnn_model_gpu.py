"""GPU tests of BayesianNeuralNetwork.train / predict, modelled on the reference's
tests/bayesian_neural_network/test_train_predict.py and test_seeding.py."""
import numpy as np
import pytest
import torch

from oracle import bnn as obnn
from pysgmcmc_b200 import Session
from pysgmcmc_b200.data_batches import generate_batches
from pysgmcmc_b200.diagnostics.objective_functions import sinc
from pysgmcmc_b200.models.bayesian_neural_network import BayesianNeuralNetwork
from pysgmcmc_b200.models.bnn_cost import default_net_params
from pysgmcmc_b200.sampling import Sampler

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def sinc_problem(seed):
    rng = np.random.RandomState(seed)
    x_train = np.asarray([rng.uniform(0.0, 1.0, 1) for _ in range(100)])
    return x_train, sinc(x_train), np.linspace(0, 1, 100)[:, None]


@pytest.mark.parametrize("normalize", [True, False])
def test_train_predict_performance(normalize):
    """test_train_predict.py:12-48: sinc, 100 points, 1000 burn-in, 10 nets -> test MSE <= 0.1."""
    x_train, y_train, X_test = sinc_problem(1)
    y_test = sinc(X_test)
    bnn = BayesianNeuralNetwork(session=Session(device=DEV), burn_in_steps=1000, n_nets=10, seed=1,
                                normalize_input=normalize, normalize_output=normalize)
    bnn.train(x_train, y_train)
    assert bnn.is_trained and len(bnn.samples) == 10
    mean, var = bnn.predict(X_test)
    assert mean.shape == (100,) and var.shape == (100,) and (var >= 0).all()
    assert np.allclose(np.mean((y_test - mean) ** 2), 0.0, atol=1e-01)
    preds, noise = bnn.predict(X_test, return_individual_predictions=True)
    assert len(preds) == 10 and preds.shape == (10, 100) and noise.shape == (10, 100)
    assert np.allclose(preds.mean(axis=0), mean)
    # predictive forward (K10) against the oracle's forward pass on the stored networks
    theta = torch.stack(list(bnn.samples)).cpu().numpy().astype(np.float64)
    X_ = (X_test - bnn.x_mean) / bnn.x_std if normalize else X_test
    f, _, _ = obnn.forward(theta, np.repeat(X_[None], 10, axis=0))
    if normalize:
        f = f * bnn.y_std + bnn.y_mean
    np.testing.assert_allclose(preds, f, rtol=1e-4, atol=1e-4)


def test_same_seed_same_chain_and_seeded_initial_net():
    """test_seeding.py: same seed -> same initial network; and here also the same posterior samples."""
    a = default_net_params(1, seed=7, device=DEV)
    b = default_net_params(1, seed=7, device=DEV)
    assert all(torch.equal(p, q) for p, q in zip(a, b))
    assert [tuple(p.shape) for p in a] == [(1, 50), (50,), (50, 50), (50,), (50, 50), (50,), (50, 1), (1,), (1, 1)]
    assert float(a[-1]) == pytest.approx(np.log(1e-3)) and float(a[1].abs().sum()) == 0.0
    assert float(a[2].abs().max()) <= 2.0 * np.sqrt(1.3 / 50) + 1e-6      # truncated at 2 std
    x_train, y_train, X_test = sinc_problem(2)
    runs = []
    for _ in range(2):
        bnn = BayesianNeuralNetwork(session=Session(device=DEV), burn_in_steps=50, sample_steps=10,
                                    n_nets=3, seed=5)
        bnn.train(x_train, y_train)
        runs.append(torch.stack(list(bnn.samples)))
    assert torch.equal(runs[0], runs[1])


def test_custom_host_batch_generator_and_sgld():
    """A user generator with generate_batches' signature feeds host minibatches through
    placeholders (reference wiring); SGLD goes through the generic update kernel."""
    x_train, y_train, X_test = sinc_problem(3)
    calls = []

    def my_batches(x, y, x_placeholder, y_placeholder, batch_size, seed):
        calls.append(batch_size)
        return generate_batches(x, y, x_placeholder, y_placeholder, batch_size, seed)
    for method in (Sampler.SGHMC, Sampler.SGLD):
        bnn = BayesianNeuralNetwork(session=Session(device=DEV), sampling_method=method,
                                    batch_generator=my_batches, burn_in_steps=30, sample_steps=5,
                                    n_nets=4, n_iters=200, seed=1)
        bnn.train(x_train, y_train)
        mean, var = bnn.predict(X_test)
        assert len(bnn.samples) == 4 and np.isfinite(mean).all() and np.isfinite(var).all()
    assert calls == [20, 20]


def test_device_and_host_generators_give_the_same_chain():
    """The on-device index stream (K7) reproduces generate_batches(seed): with the same noise
    seed the two wirings produce the same samples up to the per-step kernel differences
    (K5 pipeline vs. per-step calls use identical kernels -> bit-identical)."""
    x_train, y_train, _ = sinc_problem(4)
    kw = dict(burn_in_steps=20, sample_steps=5, n_nets=3, seed=11)
    a = BayesianNeuralNetwork(session=Session(device=DEV), **kw)
    a.train(x_train, y_train)
    b = BayesianNeuralNetwork(session=Session(device=DEV),
                              batch_generator=lambda **k: generate_batches(**k), **kw)
    b.train(x_train, y_train)
    sa, sb = torch.stack(list(a.samples)), torch.stack(list(b.samples))
    assert torch.allclose(sa, sb, rtol=1e-6, atol=1e-7)


def test_multi_chain_training_pools_networks_from_all_chains():
    """Session(n_chains=C): C chains with their own minibatch and noise streams train at once;
    every kept iteration contributes one network per chain (an extension of the reference's
    one-chain model), and the pooled predictive passes the reference's accuracy criterion."""
    x_train, y_train, X_test = sinc_problem(1)
    y_test = sinc(X_test)
    C = 8
    bnn = BayesianNeuralNetwork(session=Session(device=DEV, n_chains=C), burn_in_steps=1000, n_nets=40,
                                sample_steps=100, seed=1)
    bnn.train(x_train, y_train)
    assert bnn.is_trained and len(bnn.samples) == 40
    assert bnn.sampler.n_chains == C and bnn.sampler.n_iterations == 1501       # kept iterations 1100 .. 1500
    nets = torch.stack(list(bnn.samples))
    assert len({float(v) for v in nets[:C, 7]}) == C, "the chains are independent"
    mean, var = bnn.predict(X_test)
    assert np.allclose(np.mean((y_test - mean) ** 2), 0.0, atol=1e-01)
    # same seed -> same pool of networks
    bnn2 = BayesianNeuralNetwork(session=Session(device=DEV, n_chains=C), burn_in_steps=1000, n_nets=40,
                                 sample_steps=100, seed=1)
    bnn2.train(x_train, y_train)
    assert all(torch.equal(a, b) for a, b in zip(bnn.samples, bnn2.samples))
    with pytest.raises(ValueError):
        BayesianNeuralNetwork(session=Session(device=DEV, n_chains=C), batch_generator=lambda **kw: iter(()),
                              n_nets=2).train(x_train, y_train)
