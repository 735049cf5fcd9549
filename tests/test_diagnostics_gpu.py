"""GPU tests of the diagnostics reductions (K8) and the reference-facing entry points,
against oracle/diagnostics.py (float64 NumPy).  Tolerance: the kernels accumulate in fp64
on fp32 data, so sums agree to ~1e-12 relative; R_hat rtol 1e-9; ESS is a floored integer
and must match exactly except when the unfloored value sits within 1e-9 of an integer."""
import numpy as np
import pytest
import torch

from oracle import diagnostics as odiag
from pysgmcmc_b200 import Session
from pysgmcmc_b200.diagnostics import effective_sample_sizes, gelman_rubin
from pysgmcmc_b200.diagnostics.objective_functions import gmm1_log_likelihood, to_negative_log_likelihood
from pysgmcmc_b200.diagnostics.sampler_diagnostics import (effective_n_from_trace, gelman_rubin_from_trace,
                                                           local_moment_sums, local_variogram_sums)
from pysgmcmc_b200.samplers import SGHMCSampler
from pysgmcmc_b200.tensor_utils import set_name

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def ar1(m, n, D, phi, seed):
    rng = np.random.RandomState(seed)
    x = np.zeros((m, n, D), dtype=np.float32)
    e = rng.standard_normal((m, n, D)).astype(np.float32)
    x[:, 0] = e[:, 0]
    for i in range(1, n):
        x[:, i] = phi * x[:, i - 1] + e[:, i]
    return x + rng.standard_normal((1, 1, D)).astype(np.float32) * 3


@pytest.mark.parametrize("m,n,D", [(2, 100, 2), (7, 50, 33), (300, 40, 70), (1000, 10, 5),
                                   # vectorised kernels: D % 4 == 0 / D % 2 == 0, ragged chain groups,
                                   # draws not a multiple of the window, fewer draws than lags
                                   (70, 37, 8), (33, 100, 5252), (130, 17, 6), (5, 3, 4), (65, 16, 12), (3, 33, 1028)])
def test_moment_and_variogram_sums(m, n, D):
    x = ar1(m, n, D, 0.7, seed=m)
    trace = torch.as_tensor(np.ascontiguousarray(x.transpose(1, 0, 2)), device=DEV)   # [n, m, D]
    sums = local_moment_sums(trace).cpu().numpy()
    means, variances = odiag.chain_moments(x)
    np.testing.assert_allclose(sums[0], means.sum(0), rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(sums[1], (means ** 2).sum(0), rtol=1e-11)
    np.testing.assert_allclose(sums[2], variances.sum(0), rtol=1e-10)
    vg = local_variogram_sums(trace, 1, min(8 if D % 2 else 16, n - 1)).cpu().numpy()
    for b in range(vg.shape[0]):
        t = 1 + b
        want = odiag.variogram(x, t) * (m * (n - t))
        np.testing.assert_allclose(vg[b], want, rtol=1e-10)


def test_variogram_of_selected_dimensions():
    """sgmcmc_variogram_select_f32 (a few dimensions, any lags) == the columns of the full sums."""
    from pysgmcmc_b200.diagnostics.sampler_diagnostics import local_variogram_select_sums
    m, n, D = 300, 45, 70
    x = ar1(m, n, D, 0.8, seed=4)
    trace = torch.as_tensor(np.ascontiguousarray(x.transpose(1, 0, 2)), device=DEV)
    dims = np.array([69, 0, 13, 14])
    got = local_variogram_select_sums(trace, dims, 17, 40).cpu().numpy()      # lags 17 .. 56 (>= n: zero)
    for b in range(40):
        t = 17 + b
        want = odiag.variogram(x, t)[dims] * (m * (n - t)) if t < n else np.zeros(4)
        np.testing.assert_allclose(got[b], want, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("phi", [0.0, 0.5, 0.95])
def test_rhat_and_ess_match_oracle(phi):
    m, n, D = 16, 400, 6
    x = ar1(m, n, D, phi, seed=3)
    x[0, :, 0] += 2.0                              # one dimension with a stray chain
    trace = torch.as_tensor(np.ascontiguousarray(x.transpose(1, 0, 2)), device=DEV)
    np.testing.assert_allclose(gelman_rubin_from_trace(trace).cpu().numpy(), odiag.gelman_rubin(x), rtol=1e-9)
    np.testing.assert_array_equal(effective_n_from_trace(trace), odiag.effective_n(x))


def test_reference_entry_points():
    """sampler_diagnostics.py:89-107,169-187: dict keyed by variable name, one value per dimension."""
    def get_sampler(session):
        x = set_name(torch.tensor([1.0, 2.0], device=DEV), "x:0")
        return SGHMCSampler(params=[x], cost_fun=lambda params: (params[0] ** 2).sum(), session=session,
                            burn_in_steps=10)
    ess = effective_sample_sizes(get_sampler)
    assert isinstance(ess, dict) and list(ess)[0].startswith("x") and len(ess["x:0"]) == 2
    factors = gelman_rubin(get_sampler)
    assert isinstance(factors, dict) and len(factors["x:0"]) == 2 and np.isfinite(factors["x:0"]).all()


def test_many_chains_from_one_sampler_gmm():
    """4096 SGLD-style chains in one sampler (config 2 shape) feed the diagnostics directly."""
    C = 4096
    s = SGHMCSampler(params=[torch.zeros(C, device=DEV)], burn_in_steps=200,
                     cost_fun=to_negative_log_likelihood(gmm1_log_likelihood), seed=1,
                     session=Session(device=DEV, n_chains=C, output="torch"))
    s.run(200)
    trace, _ = s.run(400, keep_every=4)
    rhat = gelman_rubin_from_trace(trace)
    ess = effective_n_from_trace(trace)
    assert rhat.shape == (1,) and torch.isfinite(rhat).all() and 1 <= ess[0] <= C * 100
