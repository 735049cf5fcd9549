"""Oracle restatement of the BNN cost on the hot path
(test infrastructure only, see oracle/__init__.py).

Follows pysgmcmc/models/bayesian_neural_network.py:
  * get_default_net            :28-69   (3 x tanh dense + linear head + learned
                                         scalar log-variance `output_bias`)
  * log_variance_prior_log_like :77-107
  * weight_prior_log_like       :110-141
  * negative_log_likelihood     :337-388
The reference differentiates with ``tf.gradients``; the hand-written backward
pass below is cross-checked against torch autograd in tests/test_oracle.py.

PINNED (float64, <= 1 ulp) by the reference's golden vectors
tests/data/bayesian_neural_network_priors/{log_variance,weights}.npy
(asserted in tests/bayesian_neural_network/test_priors.py:20-81).

Flat per-chain parameter layout = ``tf.trainable_variables()`` order, each
tensor flattened row-major (kernels are ``[in, out]``):
    W1[n_in,h1] b1[h1] W2[h1,h2] b2[h2] W3[h2,h3] b3[h3] W4[h3,1] b4[1] rho[1,1]
For n_in=1, hidden=(50,50,50): offsets 0,50,100,2600,2650,5150,5200,5250,5251; D=5252.
All functions are vectorised over a leading chain axis: theta ``[C, D]``,
minibatch X ``[C, B, n_in]``, y ``[C, B]``.
"""
import numpy as np

from .tensor_utils import safe_divide


def layout(n_in=1, hidden=(50, 50, 50)):
    """[(name, shape, offset)] and total D, for any number of hidden layers (a user `get_net`
    of the shape of get_default_net with other widths: W_l, b_l per dense layer, then rho)."""
    widths = [n_in] + list(hidden) + [1]
    shapes = []
    for l in range(1, len(widths)):
        shapes.append(("W%d" % l, (widths[l - 1], widths[l])))
        shapes.append(("b%d" % l, (widths[l],)))
    shapes.append(("rho", (1, 1)))
    out, off = [], 0
    for name, shp in shapes:
        out.append((name, shp, off))
        off += int(np.prod(shp))
    return out, off


def unpack(theta, n_in=1, hidden=(50, 50, 50)):
    lay, D = layout(n_in, hidden)
    assert theta.shape[-1] == D, (theta.shape, D)
    lead = theta.shape[:-1]
    return {name: theta[..., off:off + int(np.prod(shp))].reshape(lead + shp)
            for name, shp, off in lay}


def pack(parts, n_in=1, hidden=(50, 50, 50)):
    lay, D = layout(n_in, hidden)
    lead = parts["W1"].shape[:-2]
    return np.concatenate([parts[name].reshape(lead + (-1,)) for name, _, _ in lay], axis=-1)


def _sd_den(y, c=1e-16):
    """Denominator safe_divide actually divides by (tensor_utils.py:269)."""
    T = np.asarray(y).dtype.type
    return y + (T(2.0) * np.sign(y) * T(c) + T(c))


# ------------------------------- priors ---------------------------------- #

def eigen_sum(values, packet=2):
    """Full sum in the order TF 1.x's CPU kernel (Eigen full reducer, SSE2 packets
    of 2 doubles) uses: lane-wise packet accumulation, horizontal add, scalar
    tail.  With this order both float64 goldens of the reference
    (tests/bayesian_neural_network/test_priors.py:20-81) are reproduced BIT-EXACTLY."""
    v = np.asarray(values).ravel()
    T = v.dtype.type
    nv = (len(v) // packet) * packet
    acc = np.zeros(packet, dtype=v.dtype)
    for i in range(0, nv, packet):
        acc = acc + v[i:i + packet]
    s = T(0.0)
    for a in acc:
        s = s + a
    for i in range(nv, len(v)):
        s = s + v[i]
    return s


def log_variance_prior_log_like(log_var, mean=1e-6, var=0.01):
    """bayesian_neural_network.py:102-107.  log_var: ``[B, 1]``."""
    log_var = np.asarray(log_var)
    T = log_var.dtype.type
    mean_, var_ = T(mean), T(var)
    per_row = (safe_divide(-np.square(log_var - np.log(mean_)), T(2.0) * var_)
               - T(0.5) * np.log(var_)).sum(axis=-1)
    return eigen_sum(per_row) / T(per_row.shape[0])


def weight_prior_log_like(parameters, wdecay=1.0, dtype=np.float64):
    """bayesian_neural_network.py:131-141.  parameters: list of arrays (one chain)."""
    T = np.dtype(dtype).type
    log_like, n_params = T(0.0), T(0.0)
    for p in parameters:
        p = np.asarray(p, dtype=dtype)
        log_like = log_like + eigen_sum(-T(wdecay) * T(0.5) * np.square(p))
        n_params = n_params + T(np.float32(np.prod(np.asarray(p.shape, dtype=np.float32))))
    return safe_divide(log_like, n_params)


# --------------------------- network + cost ------------------------------ #

def forward(theta, X, n_in=1, hidden=(50, 50, 50)):
    """get_default_net (:28-69) for any tuple of hidden widths.
    Returns (f_mean [C,B], rho [C], (P, H_0 = X, H_1, ..., H_L))."""
    P = unpack(theta, n_in, hidden)
    L = len(hidden)
    Hs = [X]
    for l in range(1, L + 1):
        Hs.append(np.tanh(Hs[-1] @ P["W%d" % l] + P["b%d" % l][:, None, :]))
    f = (Hs[-1] @ P["W%d" % (L + 1)])[..., 0] + P["b%d" % (L + 1)]
    rho = P["rho"][:, 0, 0]
    return f, rho, (P,) + tuple(Hs)


def nll_and_grad(theta, X, y, n_examples, batch_size=None, n_in=1, hidden=(50, 50, 50),
                 want_grad=True):
    """cost = -log_like of negative_log_likelihood (:337-388) and d cost / d theta.

    theta [C, D], X [C, B, n_in], y [C, B]; `batch_size` is the CONFIGURED
    constant the reference divides by (:377), default B; `n_examples` = N (:380).
    Returns (cost [C], grad [C, D], mse [C]).
    """
    theta = np.asarray(theta)
    T = theta.dtype.type
    X = np.asarray(X, dtype=theta.dtype)
    y = np.asarray(y, dtype=theta.dtype)
    C, B = y.shape
    bs = T(B if batch_size is None else batch_size)
    N = T(n_examples)
    D = theta.shape[-1]
    L = len(hidden)

    f, rho, cache = forward(theta, X, n_in, hidden)
    P, Hs = cache[0], cache[1:]
    f_var_inv = T(1.0) / (np.exp(rho) + T(1e-16))                       # :368
    diff = y - f
    mse = np.square(diff)                                               # :370
    log_like = (-mse * (T(0.5) * f_var_inv[:, None]) - T(0.5) * rho[:, None]).sum(axis=1)  # :372-374
    log_like = log_like / bs                                            # :377
    # prior on the log variance (:383): every row carries the same rho
    lv_den = T(2.0) * T(0.01)
    lv = (safe_divide(-np.square(rho - np.log(T(1e-6))), lv_den) - T(0.5) * np.log(T(0.01)))
    log_like = log_like + lv / N
    # prior on the weights (:386): over ALL trainable variables (rho included)
    wp = safe_divide((-T(1.0) * T(0.5) * np.square(theta)).sum(axis=1), T(D))
    log_like = log_like + wp / N
    cost = -log_like
    mse_mean = mse.mean(axis=1)                                         # :388
    if not want_grad:
        return cost, None, mse_mean

    # ---- backward: d cost ----
    dfm = -(diff * f_var_inv[:, None]) / bs                             # d cost / d f   [C,B]
    drho_data = -((T(0.5) * mse * (np.exp(rho) * f_var_inv * f_var_inv)[:, None] - T(0.5)).sum(axis=1)) / bs
    drho_lv = (T(2.0) * (rho - np.log(T(1e-6))) / _sd_den(lv_den)) / N

    G = {}
    head = L + 1
    G["W%d" % head] = np.einsum("cbh,cb->ch", Hs[L], dfm)[:, :, None]
    G["b%d" % head] = dfm.sum(axis=1)[:, None]
    dH = dfm[:, :, None] * P["W%d" % head][:, None, :, 0]
    for l in range(L, 0, -1):
        dZ = dH * (T(1.0) - Hs[l] * Hs[l])
        G["W%d" % l] = np.einsum("cbi,cbj->cij", Hs[l - 1], dZ)
        G["b%d" % l] = dZ.sum(axis=1)
        if l > 1:
            dH = np.einsum("cbj,cij->cbi", dZ, P["W%d" % l])
    G["rho"] = (drho_data + drho_lv)[:, None, None]
    grad = pack(G, n_in, hidden)
    # weight prior: d/dp of -(1/N) * sum(-0.5 p^2)/(D + 3e-16)
    grad = grad + theta / _sd_den(T(D)) / N
    return cost, grad.astype(theta.dtype), mse_mean


def gather_minibatch(X, y, starts, batch_size):
    """Contiguous slices ``x[start:start+B]`` (data_batches.py:120-123), one per chain."""
    idx = np.asarray(starts, dtype=np.int64)[:, None] + np.arange(batch_size)[None, :]
    return X[idx], y[idx]


def init_theta(n_chains, n_in=1, hidden=(50, 50, 50), seed=1, dtype=np.float32):
    """Synthetic initial weights in the spirit of get_default_net's initialisers
    (:31-57): truncated-normal(0, sqrt(1.3/fan_in)) kernels (tf.contrib
    variance_scaling_initializer(factor=1.0) default: FAN_IN, truncated normal),
    zero biases, rho = log(1e-3) (:59-61).  TF's RNG stream is not reproducible
    (parity unpinned), so this is seeded NumPy."""
    rng = np.random.RandomState(seed)
    lay, D = layout(n_in, hidden)
    theta = np.zeros((n_chains, D), dtype=np.float64)
    for name, shp, off in lay:
        n = int(np.prod(shp))
        if name.startswith("W"):
            std = np.sqrt(1.3 / shp[0])
            w = rng.normal(0.0, 1.0, size=(n_chains, n))
            bad = np.abs(w) > 2.0
            while bad.any():
                w[bad] = rng.normal(0.0, 1.0, size=int(bad.sum()))
                bad = np.abs(w) > 2.0
            theta[:, off:off + n] = w * std
        elif name == "rho":
            theta[:, off] = np.log(1e-3)
    return theta.astype(dtype)
