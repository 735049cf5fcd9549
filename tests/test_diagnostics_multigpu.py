"""The one NCCL exchange of the path (K9): per-rank chain sums from K8 all-reduced over
NVLink, R-hat / ESS finalised on every rank.  Needs >= 2 GPUs (skipped otherwise; run with
`gpurun --gpus 2 -- python -m pytest tests/test_diagnostics_multigpu.py -m gpu`)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import diagnostics as odiag

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, x, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from pysgmcmc_b200.diagnostics.sampler_diagnostics import effective_n_from_trace, gelman_rubin_from_trace
    shard = np.array_split(np.arange(x.shape[0]), world)[rank]
    trace = torch.as_tensor(np.ascontiguousarray(x[shard].transpose(1, 0, 2)), device="cuda:%d" % rank)
    rhat = gelman_rubin_from_trace(trace).cpu().numpy()
    ess = effective_n_from_trace(trace)
    out[rank] = (rhat, ess)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_chains_nccl_all_reduce_equals_unsharded():
    rng = np.random.RandomState(0)
    m, n, D = 64, 200, 37
    x = np.zeros((m, n, D), dtype=np.float32)
    e = rng.standard_normal((m, n, D)).astype(np.float32)
    for i in range(1, n):
        x[:, i] = 0.85 * x[:, i - 1] + e[:, i]
    x[5] += 1.0
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(2, 29500 + int(rng.randint(0, 2000)), x, out), nprocs=2, join=True)
    for rank in range(2):
        rhat, ess = out[rank]
        np.testing.assert_allclose(rhat, odiag.gelman_rubin(x), rtol=1e-9)
        np.testing.assert_array_equal(ess, odiag.effective_n(x))
