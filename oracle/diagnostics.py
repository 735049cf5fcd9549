"""Oracle for the convergence diagnostics (test infrastructure only).

The reference delegates to THIRD-PARTY ``pymc3>=3.1`` (requirements.txt:2;
``pymc3.diagnostics.effective_n`` / ``gelman_rubin``, called from
pysgmcmc/diagnostics/sampler_diagnostics.py:110-115,189-194), which is absent
from /root/reference and not installable here: **parity unpinned** bit-wise (statistically
pinned: the published ESS table of the reference is reproduced within ~1 % with this
estimator).  This file
restates pymc3 3.1's published algorithm and the formulas documented in the
reference's own docstrings (sampler_diagnostics.py:76-82,153-161):

  R_hat   = sqrt(V_hat / W),  W = mean_j s_j^2,  B = n * var_j(mean_j) (ddof=1),
            V_hat = W (n-1)/n + B/n
  n_eff   = m n / (1 + 2 sum_{t=1}^{T} rho_t),  rho_t = 1 - V_t / (2 V_hat),
            V_t = mean over chains and draws of (x_{i} - x_{i-t})^2,
            stop at the first t with rho_{t-1} + rho_t < 0 (made even), floored and
            capped at m n (pymc3 3.1 behaviour; matches the integer-valued entries
            of docs/source/notebooks/data/effective_sample_sizes/*.json).

x: ``[m chains, n draws, D]``.
"""
import numpy as np


def chain_moments(x):
    x = np.asarray(x, dtype=np.float64)
    return x.mean(axis=1), x.var(axis=1, ddof=1)


def v_hat_and_w(x):
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[1]
    means, variances = chain_moments(x)
    B = n * means.var(axis=0, ddof=1)
    W = variances.mean(axis=0)
    return W * (n - 1) / n + B / n, W


def gelman_rubin(x):
    v_hat, W = v_hat_and_w(x)
    return np.sqrt(v_hat / W)


def variogram(x, t):
    x = np.asarray(x, dtype=np.float64)
    return np.mean((x[:, t:, :] - x[:, :-t, :]) ** 2, axis=(0, 1))


def effective_n(x):
    x = np.asarray(x, dtype=np.float64)
    m, n, D = x.shape
    v_hat, _ = v_hat_and_w(x)
    out = np.empty(D)
    for d in range(D):
        rho = np.ones(n)
        negative_autocorr = False
        t = 1
        while not negative_autocorr and t < n:
            vg = np.mean((x[:, t:, d] - x[:, :-t, d]) ** 2)
            rho[t] = 1.0 - vg / (2.0 * v_hat[d])
            negative_autocorr = (rho[t - 1] + rho[t]) < 0
            t += 1
        if t % 2:
            t -= 1
        neff = m * n / (1.0 + 2.0 * rho[1:t - 1].sum())
        out[d] = min(m * n, np.floor(neff))
    return out
