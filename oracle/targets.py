"""Oracle restatement of the test densities of
``pysgmcmc/diagnostics/objective_functions.py:49-98`` with analytic gradients
(test infrastructure only, see oracle/__init__.py).

The reference obtains gradients with ``tf.gradients``; here they are written
out by hand and cross-checked against torch autograd in tests/test_oracle.py.
Pinned by the reference's doctest optima (objective_functions.py:54-56:
banana(0, 10) == 0) and by the notebook known answer banana(0, 0) == -50
(docs/source/notebooks/api_quickstart.ipynb:1104).

Layout: theta is ``[..., D]`` (D=2 for banana: (x0, x1); D=1 for gmm).
Costs are NEGATIVE log likelihoods (tests/samplers/sampler_testing.py:23-26).
"""
import numpy as np

GMM_MU = (-5.0, 0.0, 5.0)
GMM_WEIGHTS = (1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0)
GMM_VAR = {
    "gmm1": (1.0, 1.0, 1.0),                       # objective_functions.py:89-90
    "gmm2": (1.0 / 0.5, 0.5, 1.0 / 0.5),           # :93-94
    "gmm3": (1.0 / 0.3, 0.3, 1.0 / 0.3),           # :97-98
}


def banana_log_likelihood(theta):
    """objective_functions.py:59."""
    theta = np.asarray(theta)
    T = theta.dtype.type if theta.dtype.kind == "f" else np.float64
    x0, x1 = theta[..., 0], theta[..., 1]
    return T(-0.5) * (T(0.01) * x0 ** 2 + (x1 + T(0.1) * x0 ** 2 - T(10)) ** 2)


def banana_cost_and_grad(theta):
    """cost = -loglik ; grad = d cost / d theta."""
    theta = np.asarray(theta)
    T = theta.dtype.type
    x0, x1 = theta[..., 0], theta[..., 1]
    u = x1 + T(0.1) * x0 * x0 - T(10)
    cost = T(0.5) * (T(0.01) * x0 * x0 + u * u)
    g0 = T(0.01) * x0 + T(0.2) * x0 * u
    g1 = u
    return cost, np.stack([g0, g1], axis=-1)


def gmm_log_likelihood(theta, var=GMM_VAR["gmm1"], mu=GMM_MU, weights=GMM_WEIGHTS):
    """objective_functions.py:62-85 (1-D only)."""
    theta = np.asarray(theta)
    T = theta.dtype.type if theta.dtype.kind == "f" else np.float64
    x = theta[..., 0]
    comps = np.stack([
        T(np.log(weights[i])) + (T(-0.5) * T(np.log(2.0 * np.pi * var[i]))
                                 - T(0.5) * ((x - T(mu[i])) ** 2) / T(var[i]))
        for i in range(len(mu))], axis=0)
    m = comps.max(axis=0)
    return m + np.log(np.exp(comps - m).sum(axis=0))


def gmm_cost_and_grad(theta, var=GMM_VAR["gmm1"], mu=GMM_MU, weights=GMM_WEIGHTS):
    theta = np.asarray(theta)
    T = theta.dtype.type
    x = theta[..., 0]
    comps = np.stack([
        T(np.log(weights[i])) + (T(-0.5) * T(np.log(2.0 * np.pi * var[i]))
                                 - T(0.5) * ((x - T(mu[i])) ** 2) / T(var[i]))
        for i in range(len(mu))], axis=0)
    m = comps.max(axis=0)
    e = np.exp(comps - m)
    s = e.sum(axis=0)
    cost = -(m + np.log(s))
    resp = e / s
    grad = sum(resp[i] * ((x - T(mu[i])) / T(var[i])) for i in range(len(mu)))
    return cost, grad[..., None]


def cost_and_grad(name):
    if name == "banana":
        return banana_cost_and_grad
    if name in GMM_VAR:
        return lambda theta: gmm_cost_and_grad(theta, var=GMM_VAR[name])
    raise ValueError(name)
