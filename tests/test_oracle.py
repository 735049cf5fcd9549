"""CPU tests: pin the oracle (oracle/) against every fixture the reference holds
for the hot path, against NumPy's RandomState, against Random123 known answers,
against torch autograd, and against the committed golden files."""
import os

import numpy as np
import pytest
import torch

from oracle import bnn, diagnostics, mt19937, philox, samplers, targets, tensor_utils

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- reference golden vectors: tests/bayesian_neural_network/test_priors.py:20-81 ----

def test_weight_prior_golden_bit_exact():
    g = np.load(os.path.join(GOLDEN, "bnn_priors.npz"))
    params = [g["w%d" % i] for i in range(9)]
    assert [p.shape for p in params] == [(1, 50), (50,), (50, 50), (50,), (50, 50), (50,),
                                         (50, 1), (1,), (1, 1)]
    result = np.array(bnn.weight_prior_log_like(params))
    assert np.array_equal(result, g["expected_weights"])          # -0.01895130158314839


def test_log_variance_prior_golden_bit_exact():
    g = np.load(os.path.join(GOLDEN, "bnn_priors.npz"))
    result = np.array(bnn.log_variance_prior_log_like(g["f_log_var"]))
    assert np.array_equal(result, g["expected_log_variance"])     # -325.5744411137498


def test_nll_prior_terms_match_standalone_priors():
    """The priors as folded into nll_and_grad equal the standalone (golden-pinned) ones."""
    g = np.load(os.path.join(GOLDEN, "bnn_priors.npz"))
    theta = np.concatenate([g["w%d" % i].ravel() for i in range(9)])[None, :]
    X = np.zeros((1, 20, 1))
    N = 100.0
    f, rho, _ = bnn.forward(theta, X)
    y = f.copy()                                   # zero residual -> data term = -0.5 rho
    cost, _, _ = bnn.nll_and_grad(theta, X, y, n_examples=N, want_grad=False)
    lv = bnn.log_variance_prior_log_like(np.full((20, 1), rho[0]))
    wp = bnn.weight_prior_log_like([g["w%d" % i] for i in range(9)])
    expect = -(-0.5 * rho[0] + lv / N + wp / N)
    assert np.allclose(cost[0], expect, rtol=1e-14)


# ---- tensor_utils doctests (tensor_utils.py:236-265, 300-316) ----

def test_safe_divide_and_sqrt_doctests():
    assert not np.isinf(tensor_utils.safe_divide(np.float32(1.0), np.float32(0.0)))
    assert not np.isinf(tensor_utils.safe_divide(np.float32(1.0), np.float32(-1e-16)))
    assert np.isinf(np.float32(1.0) / (np.float32(-1e-16) + np.float32(1e-16)))
    assert tensor_utils.safe_sqrt(np.float32(-1e-16)) == 0.0
    # the exact offsets: +3c, +c, -c
    c = 1e-16
    assert tensor_utils.safe_divide(1.0, 0.0) == 1.0 / c
    assert tensor_utils.safe_divide(1.0, 1e-16) == 1.0 / (1e-16 + (2 * c + c))
    assert tensor_utils.safe_divide(1.0, -1e-15) == 1.0 / (-1e-15 + (-2 * c + c))
    v = tensor_utils.vectorize(np.arange(6.0).reshape(2, 3))
    assert v.shape == (6, 1)
    assert np.array_equal(tensor_utils.unvectorize(v, (2, 3)), np.arange(6.0).reshape(2, 3))


# ---- objective functions: doctest optimum + notebook known answer ----

def test_banana_known_answers():
    assert np.allclose(targets.banana_log_likelihood(np.array([0.0, 10.0])), 0.0)   # objective_functions.py:54-56
    assert targets.banana_log_likelihood(np.array([0.0, 0.0])) == -50.0             # api_quickstart.ipynb:1104
    cost, _ = targets.banana_cost_and_grad(np.array([0.0, 0.0], dtype=np.float32))
    assert cost == 50.0


@pytest.mark.parametrize("name", ["gmm1", "gmm2", "gmm3"])
def test_gmm_matches_scipy(name):
    from scipy.special import logsumexp
    from scipy.stats import norm
    x = np.linspace(-9, 9, 37)
    var = targets.GMM_VAR[name]
    expect = logsumexp([np.log(1 / 3) + norm.logpdf(x, m, np.sqrt(v))
                        for m, v in zip(targets.GMM_MU, var)], axis=0)
    got = targets.gmm_log_likelihood(x[:, None], var=var)
    assert np.allclose(got, expect, rtol=1e-12)


@pytest.mark.parametrize("name", ["banana", "gmm1", "gmm2", "gmm3"])
def test_target_gradients_match_autograd(name):
    rng = np.random.RandomState(0)
    D = 2 if name == "banana" else 1
    theta = rng.uniform(-6, 6, size=(64, D))
    cost, grad = targets.cost_and_grad(name)(theta)
    t = torch.tensor(theta, requires_grad=True)
    if name == "banana":
        ll = -0.5 * (0.01 * t[:, 0] ** 2 + (t[:, 1] + 0.1 * t[:, 0] ** 2 - 10) ** 2)
    else:
        var = torch.tensor(targets.GMM_VAR[name], dtype=torch.float64)
        mu = torch.tensor(targets.GMM_MU, dtype=torch.float64)
        comps = (np.log(1 / 3) - 0.5 * torch.log(2 * np.pi * var) - 0.5 * (t - mu) ** 2 / var)
        ll = torch.logsumexp(comps, dim=1)
    (-ll).sum().backward()
    assert np.allclose(cost, -ll.detach().numpy(), rtol=1e-12)
    assert np.allclose(grad, t.grad.numpy(), rtol=1e-10, atol=1e-12)


# ---- BNN cost / gradient vs an independent torch restatement ----

def _torch_nll(theta, X, y, N, bs, n_in=1, hidden=(50, 50, 50)):
    lay, D = bnn.layout(n_in, hidden)
    P = {name: theta[:, off:off + int(np.prod(shp))].reshape((-1,) + shp) for name, shp, off in lay}
    h, L = X, len(hidden)
    for l in range(1, L + 1):
        h = torch.tanh(h @ P["W%d" % l] + P["b%d" % l][:, None, :])
    f = (h @ P["W%d" % (L + 1)])[..., 0] + P["b%d" % (L + 1)]
    rho = P["rho"][:, 0, 0]
    fvi = 1.0 / (torch.exp(rho) + 1e-16)
    ll = (-(y - f) ** 2 * (0.5 * fvi[:, None]) - 0.5 * rho[:, None]).sum(1) / bs
    ll = ll + (-(rho - np.log(1e-6)) ** 2 / (0.02 + 3e-16) - 0.5 * np.log(0.01)) / N
    ll = ll + ((-0.5 * theta ** 2).sum(1) / (D + 3e-16)) / N
    return -ll


@pytest.mark.parametrize("n_in,hidden", [(1, (50, 50, 50)), (3, (8, 6, 5)), (2, (12,)), (1, (9, 4, 7, 3, 5)),
                                         (1, (100, 64, 64))])
def test_bnn_nll_grad_matches_autograd(n_in, hidden):
    rng = np.random.RandomState(3)
    C, B, N = 4, 20, 500
    theta = bnn.init_theta(C, n_in, hidden, seed=3, dtype=np.float64)
    theta += 0.1 * rng.standard_normal(theta.shape)
    X = rng.standard_normal((C, B, n_in))
    y = rng.standard_normal((C, B))
    cost, grad, mse = bnn.nll_and_grad(theta, X, y, n_examples=N, n_in=n_in, hidden=hidden)
    t = torch.tensor(theta, requires_grad=True)
    c = _torch_nll(t, torch.tensor(X), torch.tensor(y), N, B, n_in, hidden)
    c.sum().backward()
    assert np.allclose(cost, c.detach().numpy(), rtol=1e-12)
    assert np.allclose(grad, t.grad.numpy(), rtol=1e-9, atol=1e-12)


def test_bnn_layout_default():
    lay, D = bnn.layout()
    assert D == 5252
    assert [off for _, _, off in lay] == [0, 50, 100, 2600, 2650, 5150, 5200, 5250, 5251]


def test_bnn_golden_file():
    g = np.load(os.path.join(GOLDEN, "bnn_nll.npz"))
    Xb, yb = bnn.gather_minibatch(g["X"], g["y"], g["starts"], 20)
    cost, grad, mse = bnn.nll_and_grad(g["theta"], Xb, yb, n_examples=g["X"].shape[0])
    assert np.allclose(cost, g["cost"], rtol=1e-13)
    assert np.allclose(grad, g["grad"], rtol=1e-11, atol=1e-15)


# ---- minibatch starts: bit-exact vs numpy.random.RandomState (data_batches.py:104-120) ----

@pytest.mark.parametrize("seed,N,B", [(1, 20000, 20), (12345, 100, 10), (7, 20, 20), (99, 5, 20),
                                      (2 ** 32 - 1, 1000, 1), (0, 70000, 3)])
def test_minibatch_starts_bit_exact(seed, N, B):
    got = mt19937.minibatch_starts(seed, N, B, 1500)
    rng = np.random.RandomState()
    rng.seed(seed)
    expect = np.array([rng.randint(0, N - min(B, N) + 1) for _ in range(1500)])
    assert np.array_equal(got, expect)


def test_mt19937_raw_stream_bit_exact():
    m = mt19937.MT19937(5489)
    got = np.array([m.next_uint32() for _ in range(2000)], dtype=np.uint64)
    expect = np.random.RandomState(5489).randint(0, 2 ** 32, size=2000, dtype=np.uint64)
    assert np.array_equal(got, expect)


# ---- Philox4x32-10: Random123 known-answer vectors ----

@pytest.mark.parametrize("ctr,key,expect", [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
])
def test_philox_known_answers(ctr, key, expect):
    out = philox.philox4x32_10(np.array([ctr], dtype=np.uint32), np.array([key], dtype=np.uint32))
    assert [int(v) for v in out[0]] == expect


def test_philox_normals_are_standard_normal():
    from scipy import stats
    z = philox.normals(200000, seed=123, step=7)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert stats.kstest(z, "norm").pvalue > 1e-3
    # different steps / offsets give different, reproducible streams
    assert np.array_equal(philox.normals(64, 1, 2, elem_offset=128), philox.normals(192, 1, 2)[128:])
    assert not np.allclose(philox.normals(64, 1, 2), philox.normals(64, 1, 3))


# ---- sampler step semantics ----

def test_sghmc_first_step_by_hand():
    """Step 1 from the initial state (tau=g=v_hat=1, V=0): everything in closed form."""
    theta = np.array([[0.0, 0.0]], dtype=np.float64)
    cost, grad = targets.banana_cost_and_grad(theta)            # grad = (0, -10)
    z = np.array([[0.3, -1.2]])
    st = samplers.sghmc_step(samplers.sghmc_init(theta), grad, z, 0.01)
    r = 0.5
    assert np.allclose(st["tau"], 1 + (-1.0 / (1 + 3e-16) + 1))
    assert np.allclose(st["g"], 1 + (-r + r * grad))
    assert np.allclose(st["v_hat"], 1 + (-r + r * grad ** 2))
    sigma = np.sqrt(2 * 0.01 ** 2 * 0.05 - 0.01 ** 4)           # minv = 1
    assert np.allclose(st["v"], -1e-4 * grad + sigma * z)
    assert np.allclose(st["theta"], st["v"])
    # magnitude of the notebook's first step (api_quickstart.ipynb:1104): |noise| ~ 3.2e-3
    assert np.allclose(sigma, 3.16e-3, rtol=1e-2)


def test_burn_in_freezes_minv_with_one_step_lag():
    """base_classes.py:438-454: minv used after burn-in is the one computed in the LAST
    burn-in step, i.e. from v_hat BEFORE that step's gradient."""
    rng = np.random.RandomState(0)
    chain = samplers.OracleChain("sghmc", np.array([[0.0, 6.0]], dtype=np.float32),
                                 targets.banana_cost_and_grad, burn_in_steps=5)
    v_hat_before_last = None
    for i in range(8):
        if i == 4:
            v_hat_before_last = chain.state["v_hat"].copy()
        chain.next(rng.standard_normal((1, 2)).astype(np.float32))
        if i >= 4:
            assert np.array_equal(chain.minv, tensor_utils.safe_divide(
                np.float32(1), tensor_utils.safe_sqrt(v_hat_before_last)))
    assert not chain.is_burning_in


def test_cost_is_pre_update():
    chain = samplers.OracleChain("sghmc", np.array([[0.0, 0.0]], dtype=np.float32),
                                 targets.banana_cost_and_grad)
    theta, cost = chain.next(np.zeros((1, 2), dtype=np.float32))
    assert cost[0] == 50.0 and not np.array_equal(theta, np.zeros((1, 2)))


def test_trajectory_golden_files():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = np.load(os.path.join(GOLDEN, "trajectories.npz"))
    for name, method, target, dtype, seed, hyper in mg.CASES:
        res = mg.trajectory_case(method, target, dtype, seed, n_steps=100, **hyper)
        k = sum(1 for c in mg.CHECKPOINTS if c <= 100)
        assert np.allclose(res["theta"][:k], g[name + "/theta"][:k], rtol=1e-6, atol=1e-7), name
        assert np.allclose(res["cost"], g[name + "/cost"][:100], rtol=1e-6, atol=1e-7), name


# ---- diagnostics restatement: sanity on known processes ----

def test_diagnostics_iid_and_ar1():
    rng = np.random.RandomState(1)
    m, n = 4, 2000
    iid = rng.standard_normal((m, n, 2))
    assert np.allclose(diagnostics.gelman_rubin(iid), 1.0, atol=0.01)
    ess = diagnostics.effective_n(iid)
    assert (ess > 0.8 * m * n).all() and (ess <= m * n).all()
    phi = 0.9
    ar = np.zeros((m, n, 1))
    e = rng.standard_normal((m, n, 1))
    for i in range(1, n):
        ar[:, i] = phi * ar[:, i - 1] + e[:, i]
    ess_ar = diagnostics.effective_n(ar)[0]
    expect = m * n * (1 - phi) / (1 + phi)
    assert 0.5 * expect < ess_ar < 2.0 * expect
    assert ess_ar == np.floor(ess_ar)
    shifted = iid.copy()
    shifted[0] += 3.0
    assert (diagnostics.gelman_rubin(shifted) > 1.5).all()
