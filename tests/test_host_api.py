"""CPU tests of the host-side mirror of the reference interface: data_batches (like
pysgmcmc/tests/test_data_batches.py), BNN constructor validation (like
tests/bayesian_neural_network/test_invalid_inputs.py), stepsize schedules, the trace adapter,
the relativistic momentum initialiser and the objective functions."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings
from hypothesis.strategies import (complex_numbers, floats, fractions, integers, lists, one_of, sets,
                                   text)

from oracle import targets as otargets
from pysgmcmc_b200.data_batches import generate_batches, generate_shuffled_batches
from pysgmcmc_b200.diagnostics.objective_functions import (banana_log_likelihood, gmm1_log_likelihood,
                                                           gmm2_log_likelihood, gmm3_log_likelihood,
                                                           sinc, to_negative_log_likelihood)
from pysgmcmc_b200.diagnostics.sample_chains import PYSGMCMCTrace
from pysgmcmc_b200.models.bayesian_neural_network import BayesianNeuralNetwork
from pysgmcmc_b200.placeholders import placeholder
from pysgmcmc_b200.samplers.relativistic_sghmc import _sample_relativistic_momentum
from pysgmcmc_b200.sampling import Sampler
from pysgmcmc_b200.stepsize_schedules import ConstantStepsizeSchedule, StepsizeSchedule

NOT_POSITIVE_INT = one_of(floats(), complex_numbers(), lists(integers(), max_size=10),
                          sets(integers(), max_size=10), fractions(), text(), integers(max_value=0))
NOT_NONNEG_INT = one_of(floats(), complex_numbers(), lists(integers(), max_size=10),
                        sets(integers(), max_size=10), fractions(), text(), integers(max_value=-1))


def data(N=100, D=3, seed=0):
    rng = np.random.RandomState(seed)
    return rng.uniform(-10, 10, size=(N, D)), rng.choice([0.0, 1.0], size=N)


# ---- data_batches (tests/test_data_batches.py:79-209) ----
@settings(max_examples=25, deadline=None)
@given(NOT_POSITIVE_INT)
def test_invalid_batch_size(batch_size):
    x, y = data()
    with pytest.raises(AssertionError):
        next(generate_batches(x, y, placeholder(), placeholder(), batch_size=batch_size))


@settings(max_examples=25, deadline=None)
@given(one_of(floats(), text(), integers(max_value=-1), integers(min_value=2 ** 32)))
def test_invalid_seed(seed):
    x, y = data()
    with pytest.raises(AssertionError):
        next(generate_batches(x, y, placeholder(), placeholder(), seed=seed))


def test_label_mismatch_asserts():
    x, y = data()
    with pytest.raises(AssertionError):
        next(generate_batches(x, y[:-1], placeholder(), placeholder()))


@pytest.mark.parametrize("gen", [generate_batches, generate_shuffled_batches])
@pytest.mark.parametrize("batch_size", [1, 7, 20, 100, 250])
def test_batch_shapes_and_whole_set_when_batch_exceeds_n(gen, batch_size):
    x, y = data()
    xp, yp = placeholder(), placeholder()
    batch = next(gen(x.copy(), y.copy(), xp, yp, batch_size=batch_size, seed=1))
    b = min(batch_size, 100)
    assert set(batch.keys()) == {xp, yp}
    assert batch[xp].shape == (b, 3) and batch[yp].shape == (b, 1)
    if batch_size >= 100:
        assert sorted(batch[yp].ravel()) == sorted(y)


@settings(max_examples=10, deadline=None)
@given(integers(min_value=0, max_value=2 ** 32 - 1), integers(min_value=2, max_value=10))
def test_same_seed_same_batches(seed, n_generators):
    x, y = data()
    xp, yp = placeholder(), placeholder()
    gens = [generate_batches(x, y, xp, yp, seed=seed) for _ in range(n_generators)]
    for _ in range(20):
        batches = [next(g) for g in gens]
        for b in batches[1:]:
            assert np.array_equal(b[xp], batches[0][xp]) and np.array_equal(b[yp], batches[0][yp])


def test_batches_are_contiguous_slices_at_the_reference_start_indices():
    x, y = data(N=500)
    xp, yp = placeholder(), placeholder()
    rng = np.random.RandomState()
    rng.seed(42)
    g = generate_batches(x, y, xp, yp, batch_size=20, seed=42)
    for _ in range(50):
        start = rng.randint(0, 500 - 20 + 1)
        b = next(g)
        assert np.array_equal(b[xp], x[start:start + 20])
        assert np.array_equal(b[yp], y[start:start + 20, None])


def test_shuffled_batches_keep_pairs_together():
    x = np.arange(200, dtype=np.float64)[:, None]
    y = np.arange(200, dtype=np.float64)
    xp, yp = placeholder(), placeholder()
    g = generate_shuffled_batches(x, y, xp, yp, batch_size=20, seed=3)
    for _ in range(20):
        b = next(g)
        assert np.array_equal(b[xp].ravel(), b[yp].ravel())
    assert sorted(x.ravel()) == list(range(200))       # rows permuted in place, none lost


# ---- BNN constructor (tests/bayesian_neural_network/test_invalid_inputs.py) ----
@settings(max_examples=20, deadline=None)
@given(NOT_POSITIVE_INT)
def test_bnn_invalid_n_nets(n_nets):
    with pytest.raises(AssertionError):
        BayesianNeuralNetwork(n_nets=n_nets)


@settings(max_examples=20, deadline=None)
@given(NOT_POSITIVE_INT)
def test_bnn_invalid_n_iters(n_iters):
    with pytest.raises(AssertionError):
        BayesianNeuralNetwork(n_iters=n_iters)


@settings(max_examples=20, deadline=None)
@given(NOT_NONNEG_INT)
def test_bnn_invalid_burn_in_steps(burn_in_steps):
    with pytest.raises(AssertionError):
        BayesianNeuralNetwork(burn_in_steps=burn_in_steps)


@settings(max_examples=20, deadline=None)
@given(NOT_POSITIVE_INT)
def test_bnn_invalid_sample_steps(sample_steps):
    with pytest.raises(AssertionError):
        BayesianNeuralNetwork(sample_steps=sample_steps)


@settings(max_examples=20, deadline=None)
@given(NOT_POSITIVE_INT)
def test_bnn_invalid_batch_size(batch_size):
    with pytest.raises(AssertionError):
        BayesianNeuralNetwork(batch_size=batch_size)


@settings(max_examples=20, deadline=None)
@given(one_of(floats(), text(), integers(), lists(integers(), max_size=3)))
def test_bnn_invalid_sampling_method(sampling_method):
    with pytest.raises(ValueError):
        BayesianNeuralNetwork(sampling_method=sampling_method)


def test_bnn_unsupported_sampler_and_predict_before_train():
    with pytest.raises(ValueError):
        BayesianNeuralNetwork(sampling_method=Sampler.RelativisticSGHMC)   # sampling.py:64
    bnn = BayesianNeuralNetwork(burn_in_steps=1000, n_nets=10)
    assert not bnn.is_trained
    with pytest.raises(ValueError, match="untrained"):
        bnn.predict(np.linspace(0, 1, 100)[:, None])                      # test_train_predict.py:51-72


def test_get_net_architectures():
    """get_net: the default callable, MLPNet of any widths, TorchNet; a TensorFlow-style callable
    without parameters is rejected with an explanation (models/networks.py)."""
    import torch
    from pysgmcmc_b200.models import MLPNet, TorchNet
    from pysgmcmc_b200.models.bayesian_neural_network import get_default_net
    from pysgmcmc_b200.models.networks import DEFAULT_NET, as_network
    assert as_network(get_default_net, get_default_net) is DEFAULT_NET
    wide = MLPNet((1000, 512, 512))
    assert wide.n_parameters(1) == 777682 and wide.widths(1) == [1, 1000, 512, 512, 1]
    assert DEFAULT_NET.n_parameters(1) == 5252 and MLPNet((50, 50, 50)) == DEFAULT_NET != wide
    assert DEFAULT_NET.parameter_shapes(1) == [(1, 50), (50,), (50, 50), (50,), (50, 50), (50,), (50, 1), (1,), (1, 1)]
    p = MLPNet((6, 4)).init_params(3, n_chains=2, seed=1, device="cpu")
    assert [tuple(t.shape) for t in p] == [(2, 3, 6), (2, 6), (2, 6, 4), (2, 4), (2, 4, 1), (2, 1), (2, 1, 1)]
    assert float(p[-1][0, 0, 0]) == pytest.approx(np.log(1e-3)) and float(p[1].abs().sum()) == 0.0
    out = MLPNet((6, 4))(torch.zeros(5, 3), p)
    assert tuple(out.shape) == (2, 5, 2)
    q = MLPNet((6, 4)).init_params(3, n_chains=2, seed=1, device="cpu")
    assert all(torch.equal(a, b) for a, b in zip(p, q))                  # same seed, same network
    t = TorchNet(lambda x, prm: x @ prm[0], lambda n_in: [(n_in, 2), (1, 1)])
    assert t.n_parameters(3) == 7 and float(t.init_params(3, device="cpu")[-1]) == pytest.approx(np.log(1e-3))
    with pytest.raises(ValueError, match="parameters are explicit"):
        as_network(lambda inputs, seed=None, dtype=None: inputs, get_default_net)
    with pytest.raises(ValueError):
        MLPNet(())


# ---- schedules, trace adapter, momentum initialiser, objective functions ----
def test_schedules():
    s = ConstantStepsizeSchedule(0.01)
    assert s.initial_value == 0.01 and next(s) == 0.01 and [next(s) for _ in range(4)] == [0.01] * 4
    assert str(ConstantStepsizeSchedule(0.1)) == "ConstantStepsizeSchedule(stepsize=0.1)"
    assert s.update(1, 2, x=3) is None and iter(s) is s
    with pytest.raises(TypeError):
        StepsizeSchedule(0.1)                      # abstract


def test_trace_adapter():
    trace = PYSGMCMCTrace(0, [[0.0, 0.0], [0.2, -0.2], [0.3, -0.5], [0.1, 0.0]], varnames=["x_1:0", "y_1:0"])
    assert trace.varnames == ["x_1:0", "y_1:0"] and len(trace) == 4 and trace.chain == 0
    assert np.array_equal(trace.get_values("x_1:0"), [0.0, 0.2, 0.3, 0.1])
    assert np.array_equal(trace[1], [0.0, -0.2, -0.5, 0.0])
    assert np.array_equal(trace.get_values("y_1:0", burn=1, thin=2), [-0.2, 0.0])
    assert trace.point(1) == {"x_1:0": 0.2, "y_1:0": -0.2}
    with pytest.raises(ValueError, match="FANTASYVARNAME"):
        trace.get_values("FANTASYVARNAME")
    anon = PYSGMCMCTrace(1, [0.5, 0.7])
    assert anon.n_vars == 1 and anon.varnames == ["0"]
    with pytest.raises(AssertionError):
        PYSGMCMCTrace(2, [])


def test_relativistic_momentum_distribution():
    """KS test against the density ~ exp(-m c^2 sqrt(p^2/(m^2 c^2) + 1)) the reference samples
    with arspy (relativistic_sghmc.py:143-223)."""
    from scipy import integrate, stats
    for m, c in ((1.0, 1.0), (1.5, 0.8)):
        p = np.array(_sample_relativistic_momentum(m, c, 20000, seed=3))
        assert len(p) == 20000 and len(_sample_relativistic_momentum(m, c, 10)) == 10
        pdf = lambda x: np.exp(-m * c ** 2 * np.sqrt(x ** 2 / (m ** 2 * c ** 2) + 1.0))
        Z = integrate.quad(pdf, -np.inf, np.inf)[0]
        grid = np.linspace(-40, 40, 8001)
        cdf_grid = np.concatenate([[0.0], np.cumsum((pdf(grid[1:]) + pdf(grid[:-1])) / 2 * np.diff(grid))]) / Z
        cdf = lambda x: np.interp(x, grid, cdf_grid)
        assert stats.kstest(p, cdf).pvalue > 1e-3
    a = _sample_relativistic_momentum(1.0, 1.0, 5, seed=9)
    assert a == _sample_relativistic_momentum(1.0, 1.0, 5, seed=9)
    with pytest.raises(AssertionError):
        _sample_relativistic_momentum(1, 1.0, 5)


def test_objective_functions_match_the_oracle_and_carry_native_tags():
    assert np.allclose(banana_log_likelihood((0, 10)), 0.0)
    assert banana_log_likelihood((0.0, 0.0)) == -50.0
    x = torch.linspace(-8, 8, 33, dtype=torch.float64)
    for fn, name in ((gmm1_log_likelihood, "gmm1"), (gmm2_log_likelihood, "gmm2"), (gmm3_log_likelihood, "gmm3")):
        want = otargets.gmm_log_likelihood(x.numpy()[:, None], var=otargets.GMM_VAR[name])
        assert np.allclose(fn([x]).numpy(), want, rtol=1e-12)
        assert np.allclose([fn([float(v)]) for v in x[:5]], want[:5], rtol=1e-12)
        assert fn.native_target == (name, 1) and to_negative_log_likelihood(fn).native_target == (name, -1)
    nll = to_negative_log_likelihood(banana_log_likelihood)
    assert nll.__name__ == "banana_log_likelihood" and nll((0.0, 0.0)) == 50.0
    assert np.allclose(sinc(np.array([[0.5]])), 1.0)


@settings(max_examples=200, deadline=None)
@given(n_steps=integers(min_value=0, max_value=300), max_steps=integers(min_value=1, max_value=65),
       sample_every=one_of(integers(min_value=1, max_value=120)), sample_phase=integers(min_value=0, max_value=119),
       thinned=integers(min_value=0, max_value=1))
def test_iter_host_blocks_cover_the_steps_and_end_at_sample_steps(n_steps, max_steps, sample_every, sample_phase, thinned):
    """The blocks `iter_host` launches for a sampler on the resident kernel: consecutive, non-empty, at most
    lookahead + 1 steps, covering every step once, and a sample step is always the last step of its block."""
    from pysgmcmc_b200.samplers import SGHMCSampler
    every = sample_every if thinned else None
    blocks = SGHMCSampler._host_blocks(n_steps, max_steps, every, sample_phase)
    assert [b[0] for b in blocks] == [0] + [b[1] for b in blocks[:-1]] if blocks else n_steps == 0
    assert (blocks[-1][1] if blocks else 0) == n_steps
    assert all(0 < s1 - s0 <= max_steps for s0, s1 in blocks)
    if every:
        ends = {s1 - 1 for _, s1 in blocks}
        assert all(s in ends for s in range(n_steps) if (s + 1 + sample_phase) % every == 0)
