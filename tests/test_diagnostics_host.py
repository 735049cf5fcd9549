"""CPU tests of the host side of the diagnostics: finalisation from chain sums against the
oracle, and the N>1 path (all-reduce of per-rank sums over gloo, world_size 2)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import diagnostics as odiag
from pysgmcmc_b200.diagnostics.sampler_diagnostics import ChainSums, _all_reduce_sum, effective_n_from_variograms


def ar1(m, n, D, phi, seed):
    rng = np.random.RandomState(seed)
    x = np.zeros((m, n, D))
    e = rng.standard_normal((m, n, D))
    for i in range(1, n):
        x[:, i] = phi * x[:, i - 1] + e[:, i]
    return x


def host_sums(x):
    """What K8 computes on the device, in NumPy (test-side stand-in: no GPU here)."""
    means, variances = odiag.chain_moments(x)
    return torch.tensor(np.stack([means.sum(0), (means ** 2).sum(0), variances.sum(0)]))


def host_variogram_block(x, group=None):
    m, n, _ = x.shape

    def block(lag0, k):
        v = np.stack([((x[:, t:, :] - x[:, :-t, :]) ** 2).sum(axis=(0, 1)) for t in range(lag0, lag0 + k)])
        return _all_reduce_sum(torch.tensor(v), group)
    return block


def test_finalisation_matches_oracle_single_rank():
    for phi in (0.0, 0.6, 0.97):
        x = ar1(6, 300, 4, phi, seed=1)
        cs = ChainSums(host_sums(x), x.shape[0], x.shape[1])
        np.testing.assert_allclose(cs.gelman_rubin().numpy(), odiag.gelman_rubin(x), rtol=1e-10)
        v_hat, _ = cs.v_hat_and_w()
        ess = effective_n_from_variograms(v_hat, cs.m, cs.n, host_variogram_block(x))
        np.testing.assert_array_equal(ess, odiag.effective_n(x))


def _worker(rank, world, port, x, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = np.array_split(np.arange(x.shape[0]), world)[rank]
    xs = x[shard]
    cs = ChainSums(host_sums(xs), len(shard), x.shape[1])
    v_hat, _ = cs.v_hat_and_w()
    ess = effective_n_from_variograms(v_hat, cs.m, cs.n, host_variogram_block(xs))
    out[rank] = (cs.m, cs.gelman_rubin().numpy(), ess)
    dist.destroy_process_group()


def test_two_ranks_gloo_equal_the_unsharded_result():
    x = ar1(10, 200, 3, 0.8, seed=4)           # 10 chains: ranks own 5 + 5
    x[3] += 1.5
    manager = mp.Manager()
    out = manager.dict()
    port = 29500 + int(np.random.randint(0, 2000))
    mp.spawn(_worker, args=(2, port, x, out), nprocs=2, join=True)
    for rank in range(2):
        m, rhat, ess = out[rank]
        assert m == 10
        np.testing.assert_allclose(rhat, odiag.gelman_rubin(x), rtol=1e-10)
        np.testing.assert_array_equal(ess, odiag.effective_n(x))
