"""Oracle restatement of the per-step sampler updates of pysgmcmc
(test infrastructure only, see oracle/__init__.py).

Follows, line by line:
  * pysgmcmc/samplers/sghmc.py:109-251            (`sghmc_step`)
  * pysgmcmc/samplers/sgld.py:102-213             (`sgld_step`)
  * pysgmcmc/samplers/relativistic_sghmc.py:100-140 (`rsghmc_step`)
  * pysgmcmc/samplers/base_classes.py:258-310,393-456 (`OracleChain`: burn-in
    bookkeeping, frozen `minv`, cost returned at the PRE-update point)

Everything is element-wise, so state arrays may have any shape (the tests use
``[C chains, D params]``).  All arithmetic is done op by op in the dtype of the
state (float32 or float64) with the reference's parenthesisation; Python
scalars are converted to that dtype first, as TensorFlow does for constants.

Parity: trajectories are **unpinned** by the reference (no golden trajectory,
reference not runnable here); this restatement is the pin.
"""
import numpy as np

from .tensor_utils import safe_divide, safe_sqrt


def _t(dtype):
    return np.dtype(dtype).type


# --------------------------------------------------------------------------- #
# state constructors (initial values: sghmc.py:126-155, sgld.py:117-145)
# --------------------------------------------------------------------------- #

def sghmc_init(theta):
    theta = np.array(theta, copy=True)
    one = np.ones_like(theta)
    return dict(theta=theta, v=np.zeros_like(theta), tau=one.copy(), g=one.copy(),
                v_hat=one.copy(), minv=one / np.sqrt(one))


def sgld_init(theta):
    theta = np.array(theta, copy=True)
    one = np.ones_like(theta)
    return dict(theta=theta, tau=one.copy(), g=one.copy(), v_hat=one.copy(),
                minv=one / np.sqrt(one))


def rsghmc_init(theta, momentum):
    theta = np.array(theta, copy=True)
    return dict(theta=theta, p=np.array(momentum, dtype=theta.dtype, copy=True))


# --------------------------------------------------------------------------- #
# shared burn-in adaptation (sghmc.py:165-196 == sgld.py:153-180)
# --------------------------------------------------------------------------- #

def _adapt(state, grad):
    """Returns (minv_t, updates) using OLD tau / g / v_hat throughout."""
    tau, g, v_hat = state["tau"], state["g"], state["v_hat"]
    one = _t(tau.dtype)(1.0)
    r_t = one / (tau + one)                                     # sghmc.py:168
    tau_t = tau + (safe_divide(-g * g * tau, v_hat) + one)      # sghmc.py:172-176
    minv_t = safe_divide(one, safe_sqrt(v_hat))                 # sghmc.py:179-183
    g_t = g + (-r_t * g + r_t * grad)                           # sghmc.py:186-190
    v_hat_t = v_hat + (-r_t * v_hat + r_t * np.power(grad, _t(grad.dtype)(2.0)))  # :192-196
    return minv_t, dict(tau=tau_t, g=g_t, v_hat=v_hat_t)


# --------------------------------------------------------------------------- #
# SGHMC  (sghmc.py:109-251)
# --------------------------------------------------------------------------- #

def sghmc_scalars(epsilon, mdecay, scale_grad, dtype):
    """Per-step scalar prefixes, computed in `dtype` in the reference's order.

    noise_scale = 2 * eps_s**2 * mdecay * minv - 2 * eps_s**3 * minv**2 * 0 - eps_s**4
    (sghmc.py:211-217) and the drift prefix -(eps**2) (sghmc.py:235).
    """
    T = _t(dtype)
    eps = T(epsilon)
    eps_s = eps / np.sqrt(T(scale_grad))                        # sghmc.py:115
    a = T(2.0) * np.power(eps_s, T(2.0)) * T(mdecay)            # ((2*es^2)*mdecay)
    b = T(2.0) * np.power(eps_s, T(3.0))                        # (2*es^3)  [* minv^2 * noise]
    c = np.power(eps_s, T(4.0))                                 # es^4
    neg_eps2 = -np.power(eps, T(2.0))                           # -(eps**2), UNSCALED eps
    return dict(a=a, b=b, c=c, neg_eps2=neg_eps2, mdecay=T(mdecay))


def sghmc_step(state, grad, z, epsilon, mdecay=0.05, scale_grad=1.0,
               burn_in=True, frozen_minv=None):
    """One SGHMC step. `grad` = d cost / d theta at the OLD theta, `z` ~ N(0,1).

    burn_in=True : adapt tau/g/v_hat, minv from the OLD v_hat (sghmc.py:165-196).
    burn_in=False: use `frozen_minv` (base_classes.py:448-454 feeds the value
                   fetched in the last burn-in step); tau/g/v_hat keep changing in
                   the reference but are unobservable, so they are left alone.
    Returns the new state dict (plus key "minv" = the minv used in this step).
    """
    dt = state["theta"].dtype
    grad = np.asarray(grad, dtype=dt)
    z = np.asarray(z, dtype=dt)
    s = sghmc_scalars(epsilon, mdecay, scale_grad, dt)
    new = dict(state)
    if burn_in:
        minv_t, upd = _adapt(state, grad)
        new.update(upd)
    else:
        minv_t = np.asarray(frozen_minv, dtype=dt)
    noise = _t(dt)(0.0)                                         # sghmc.py:111
    noise_scale = (s["a"] * minv_t - s["b"] * np.square(minv_t) * noise) - s["c"]  # :211-217
    sigma = np.sqrt(np.maximum(noise_scale, _t(dt)(1e-16)))     # sghmc.py:220
    sample = sigma * z                                          # base_classes.py:218
    v_t = state["v"] + ((s["neg_eps2"] * minv_t * grad - s["mdecay"] * state["v"]) + sample)  # :233-238
    new["v"] = v_t
    new["theta"] = state["theta"] + v_t                         # sghmc.py:241-243
    new["minv"] = minv_t
    return new


# --------------------------------------------------------------------------- #
# SGLD  (sgld.py:102-213)
# --------------------------------------------------------------------------- #

def sgld_step(state, grad, z, epsilon, A=1.0, scale_grad=1.0,
              burn_in=True, frozen_minv=None):
    dt = state["theta"].dtype
    T = _t(dt)
    grad = np.asarray(grad, dtype=dt)
    z = np.asarray(z, dtype=dt)
    eps, A_, noise = T(epsilon), T(A), T(0.0)
    new = dict(state)
    if burn_in:
        minv_t, upd = _adapt(state, grad)
        new.update(upd)
    else:
        minv_t = np.asarray(frozen_minv, dtype=dt)
    sigma = safe_sqrt(T(2.0) * eps * safe_divide(minv_t * (A_ - noise), T(scale_grad)))  # sgld.py:186-191
    sample = sigma * z
    new["theta"] = state["theta"] + (-eps * minv_t * A_ * grad + sample)   # sgld.py:201-204
    new["minv"] = minv_t
    return new


# --------------------------------------------------------------------------- #
# Relativistic SGHMC  (relativistic_sghmc.py:100-140)
# --------------------------------------------------------------------------- #

def rsghmc_step(state, grad_cost, z, epsilon, mass=1.0, speed_of_light=1.0,
                D=1.0, Bhat=0.0):
    """`grad_cost` = d cost / d theta; the reference differentiates ``-cost``
    (relativistic_sghmc.py:100-103), so the log-likelihood gradient is -grad_cost."""
    dt = state["theta"].dtype
    T = _t(dt)
    grad = -np.asarray(grad_cost, dtype=dt)
    z = np.asarray(z, dtype=dt)
    eps, m, c, D_, b_hat = T(epsilon), T(mass), T(speed_of_light), T(D), T(Bhat)
    p = state["p"]
    m2c2 = np.square(m) * np.square(c)
    p_grad = eps * p / (m * np.sqrt(p * p / m2c2 + T(1.0)))     # :123
    n = np.sqrt(eps * (T(2.0) * D_ - eps * b_hat)) * z           # :125
    p_t = p + ((eps * grad + n) - D_ * p_grad)                   # :126-129
    p_grad_new = eps * p_t / (m * np.sqrt(p_t * p_t / m2c2 + T(1.0)))   # :131
    new = dict(state)
    new["p"] = p_t
    new["theta"] = state["theta"] + p_grad_new                   # :132-135
    return new


# --------------------------------------------------------------------------- #
# Step driver: base_classes.py:258-310 (MCMCSampler.__next__) and :393-456
# (BurnInMCMCSampler)
# --------------------------------------------------------------------------- #

class OracleChain(object):
    """Drives one of the step functions the way `next(sampler)` does.

    cost_and_grad(theta) -> (cost, grad) evaluated at the CURRENT (pre-update)
    theta; `next()` returns (theta_new, cost_old) like the reference
    (base_classes.py:298-300: the cost graph reads `param` before the assign).
    """

    def __init__(self, method, theta0, cost_and_grad, epsilon=None, burn_in_steps=3000,
                 momentum=None, **hyper):
        self.method = method
        self.cost_and_grad = cost_and_grad
        self.hyper = hyper
        self.n_iterations = 0
        if method == "sghmc":
            self.state = sghmc_init(theta0)
            self.epsilon = 0.01 if epsilon is None else epsilon
            self.burn_in_steps = burn_in_steps
        elif method == "sgld":
            self.state = sgld_init(theta0)
            # sgld.py:96-100 never forwards stepsize_schedule => always 0.01
            self.epsilon = 0.01 if epsilon is None else epsilon
            self.burn_in_steps = burn_in_steps
        elif method == "rsghmc":
            self.state = rsghmc_init(theta0, momentum)
            self.epsilon = 0.001 if epsilon is None else epsilon
            self.burn_in_steps = None
        else:
            raise ValueError(method)
        self.minv = None

    @property
    def is_burning_in(self):
        return self.n_iterations < self.burn_in_steps           # base_classes.py:406

    def next(self, z, **cost_kwargs):
        cost, grad = self.cost_and_grad(self.state["theta"], **cost_kwargs)
        if self.method == "rsghmc":
            self.state = rsghmc_step(self.state, grad, z, self.epsilon, **self.hyper)
        else:
            step = sghmc_step if self.method == "sghmc" else sgld_step
            # base_classes.py:449: burn_in_steps == 0 never freezes minv
            adapt = self.is_burning_in or self.burn_in_steps == 0
            self.state = step(self.state, grad, z, self.epsilon, burn_in=adapt,
                              frozen_minv=None if adapt else self.minv, **self.hyper)
            if adapt:
                self.minv = self.state["minv"]                  # base_classes.py:438
        self.n_iterations += 1
        return self.state["theta"].copy(), cost
