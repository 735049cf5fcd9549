"""Convergence diagnostics -- same entry points as
pysgmcmc/diagnostics/sampler_diagnostics.py:47-194 (`effective_sample_sizes`,
`gelman_rubin`), computed by GPU reductions (K8, csrc/moments.cu) plus one all-reduce of
per-dimension chain sums when chains are sharded over several GPUs (K9).

The reference delegates to pymc3 >= 3.1 (third-party, not vendored); the formulas are the
ones its docstrings state (:76-82, :153-161) and pymc3 3.1 implements:

    W = mean_j s_j^2,  B = n var_j(mean_j),  V_hat = W (n-1)/n + B/n,  R_hat = sqrt(V_hat / W)
    n_eff = m n / (1 + 2 sum_{t=1}^{T} rho_t),  rho_t = 1 - V_t / (2 V_hat),
    V_t = mean over chains and draws of (x_i - x_{i-t})^2, stop at the first t with
    rho_{t-1} + rho_t < 0 (made even); floored and capped at m n.

Everything that crosses GPUs is a SUM over chains, so ranks combine with
``all_reduce(SUM)`` of ``[3, D]`` (+ ``[n_lags, D]``) float64 values and finalise
redundantly.  Parity unpinned by the reference (see oracle/diagnostics.py).
"""
import numpy as np
import torch

from .. import _native
from ..tensor_utils import get_name

__all__ = ("effective_sample_sizes", "gelman_rubin", "gelman_rubin_from_trace",
           "effective_n_from_trace", "diagnose_trace", "ChainSums")


# ---------------------------------------------------------------------------------------
# rank-local GPU reductions (K8)
# ---------------------------------------------------------------------------------------
def local_moment_sums(trace):
    """trace ``[n, C, D]`` float32 CUDA -> float64 ``[3, D]``: sum_j mean_j, sum_j mean_j^2,
    sum_j var_j(ddof=1) over the local chains."""
    n, C, D = trace.shape
    sums = torch.zeros((3, D), dtype=torch.float64, device=trace.device)
    with torch.cuda.device(trace.device):
        _native.call("sgmcmc_chain_moments_f32", _native.ptr(trace), _native.ptr(sums), n, C, D,
                     _native.stream_ptr())
    return sums


def local_variogram_sums(trace, lag0, n_lags):
    """float64 ``[n_lags, D]``: sum over local chains and draws of (x_i - x_{i-t})^2, t = lag0.."""
    n, C, D = trace.shape
    out = torch.zeros((n_lags, D), dtype=torch.float64, device=trace.device)
    with torch.cuda.device(trace.device):
        _native.call("sgmcmc_variogram_f32", _native.ptr(trace), _native.ptr(out), n, C, D, lag0, n_lags,
                     _native.stream_ptr())
    return out


def local_variogram_select_sums(trace, dims, lag0, n_lags):
    """float64 ``[n_lags, len(dims)]``: the same sums for the listed dimensions only."""
    n, C, D = trace.shape
    idx = torch.as_tensor(np.asarray(dims, dtype=np.int64), device=trace.device)
    out = torch.zeros((n_lags, idx.numel()), dtype=torch.float64, device=trace.device)
    with torch.cuda.device(trace.device):
        _native.call("sgmcmc_variogram_select_f32", _native.ptr(trace), _native.ptr(idx), _native.ptr(out),
                     n, C, D, idx.numel(), lag0, n_lags, _native.stream_ptr())
    return out


# ---------------------------------------------------------------------------------------
# cross-rank combination (K9) and finalisation -- pure torch, runs on any device
# ---------------------------------------------------------------------------------------
def _all_reduce_sum(t, group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class ChainSums(object):
    """Per-dimension sums over ALL chains of the job (after the all-reduce)."""

    def __init__(self, moment_sums, n_chains_local, n_draws, group=None):
        packed = torch.cat([moment_sums.reshape(-1),
                            torch.tensor([float(n_chains_local)], dtype=torch.float64,
                                         device=moment_sums.device)])
        packed = _all_reduce_sum(packed, group)
        self.sums = packed[:-1].reshape(3, -1)
        self.m = int(round(float(packed[-1])))
        self.n = int(n_draws)
        self.group = group

    def v_hat_and_w(self):
        m, n = self.m, self.n
        mean_of_means = self.sums[0] / m
        var_of_means = (self.sums[1] - m * mean_of_means ** 2) / (m - 1) if m > 1 else torch.zeros_like(self.sums[0])
        B = n * var_of_means
        W = self.sums[2] / m
        return W * (n - 1) / n + B / n, W

    def gelman_rubin(self):
        v_hat, W = self.v_hat_and_w()
        return torch.sqrt(v_hat / W)


def effective_n_from_variograms(v_hat, m, n, variogram_block, select_block=None):
    """pymc3-3.1 style ESS from V_hat and a callable ``variogram_block(lag0, n_lags) ->
    [n_lags, D]`` of globally summed squared lag differences.  Lags are requested in blocks
    until every dimension has hit its stopping rule; once fewer than a quarter of the
    dimensions are still open, ``select_block(lag0, n_lags, dims) -> [n_lags, len(dims)]``
    (if given) is asked for those dimensions only."""
    D = v_hat.shape[0]
    v_hat = v_hat.detach().cpu().numpy()
    rho_prev = np.ones(D)
    rho_sum = np.zeros(D)                  # sum of rho[1 : t_stop - 1]
    done = np.zeros(D, dtype=bool)
    history = [np.ones(D)]                 # rho[0] = 1
    t_stop = np.full(D, n)                 # value of t when the loop ends
    t, block = 1, 16
    while t < n and not done.all():
        k = min(block, n - t)
        active = np.flatnonzero(~done)
        if select_block is not None and t > 1 and 4 * active.size <= D:
            vg = np.full((k, D), np.nan)
            vg[:, active] = select_block(t, k, active).detach().cpu().numpy()
        else:
            vg = variogram_block(t, k).detach().cpu().numpy()
        for b in range(k):
            lag = t + b
            rho = 1.0 - (vg[b] / (m * (n - lag))) / (2.0 * v_hat)
            history.append(np.where(done, np.nan, rho))
            newly = (~done) & ((rho_prev + rho) < 0)
            t_stop[newly] = lag + 1        # loop exits with t = lag + 1
            done |= newly
            rho_prev = np.where(done, rho_prev, rho)
        t += k
        block = min(2 * block, 1024)       # slowly mixing chains need thousands of lags
    rho_all = np.stack(history)            # [t_max, D]
    t_end = np.where(t_stop % 2 == 1, t_stop - 1, t_stop)
    # sum of rho[1 : t_end - 1] per dimension, all dimensions at once
    lag = np.arange(rho_all.shape[0])[:, None]
    inside = (lag >= 1) & (lag < np.maximum(1, t_end - 1)[None, :])
    s = np.nansum(np.where(inside, rho_all, 0.0), axis=0)
    return np.minimum(m * n, np.floor(m * n / (1.0 + 2.0 * s)))


def gelman_rubin_from_trace(trace, group=None):
    """R_hat per dimension for a device trace ``[n_draws, C_local, D]`` (all ranks together)."""
    n, C, _ = trace.shape
    return ChainSums(local_moment_sums(trace), C, n, group).gelman_rubin()


def effective_n_from_trace(trace, group=None):
    """ESS per dimension for a device trace ``[n_draws, C_local, D]`` (all ranks together)."""
    n, C, _ = trace.shape
    cs = ChainSums(local_moment_sums(trace), C, n, group)
    v_hat, _ = cs.v_hat_and_w()
    return effective_n_from_variograms(
        v_hat, cs.m, n, lambda lag0, k: _all_reduce_sum(local_variogram_sums(trace, lag0, k), group),
        lambda lag0, k, dims: _all_reduce_sum(local_variogram_select_sums(trace, dims, lag0, k), group))


class _Stopwatch(object):
    """Accumulates device time (CUDA events on the current stream) and host time per label;
    a no-op when no `timings` dict was asked for."""

    def __init__(self, timings, device):
        self.t, self.device, self.pending = timings, device, []

    def device_section(self, label, fn):
        if self.t is None:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream(self.device))
        out = fn()
        e1.record(torch.cuda.current_stream(self.device))
        self.pending.append((label, e0, e1))
        return out

    def host_section(self, label, fn):
        if self.t is None:
            return fn()
        import time
        torch.cuda.synchronize(self.device)
        t0 = time.perf_counter()
        out = fn()
        self.t[label] = self.t.get(label, 0.0) + (time.perf_counter() - t0) * 1e3
        return out

    def close(self):
        if self.t is None:
            return
        torch.cuda.synchronize(self.device)
        for label, e0, e1 in self.pending:
            self.t[label] = self.t.get(label, 0.0) + e0.elapsed_time(e1)
        self.pending = []


def diagnose_trace(trace, group=None, timings=None):
    """R-hat and ESS per dimension of a device trace ``[n_draws, C_local, D]``, all ranks of
    `group` together, in one pass: K8 moment sums -> all-reduce -> V_hat, W, R-hat; K8 variogram
    sums per lag block -> all-reduce -> ESS (sampler_diagnostics.py:76-82,153-161).  Returns
    ``(r_hat [D] float64 tensor, ess [D] float64 ndarray)``.

    `timings` (a dict) receives, in milliseconds: ``k8_ms`` (the reduction kernels),
    ``allreduce_ms`` (the collectives, device time on the current stream), ``finalize_ms``
    (host arithmetic of the stopping rule), plus ``allreduce_calls`` and ``allreduce_bytes``."""
    import time
    n, C, D = trace.shape
    sw = _Stopwatch(timings, trace.device)
    if timings is not None:
        timings.update({"allreduce_calls": 0, "allreduce_bytes": 0})
        torch.cuda.synchronize(trace.device)
    t_begin = time.perf_counter()

    def reduce(t):
        if timings is not None:
            timings["allreduce_calls"] += 1
            timings["allreduce_bytes"] += t.numel() * t.element_size()
        return sw.device_section("allreduce_ms", lambda: _all_reduce_sum(t, group))

    sums = sw.device_section("k8_ms", lambda: local_moment_sums(trace))
    packed = torch.cat([sums.reshape(-1), torch.tensor([float(C)], dtype=torch.float64, device=trace.device)])
    packed = reduce(packed)
    cs = ChainSums.__new__(ChainSums)
    cs.sums, cs.m, cs.n, cs.group = packed[:-1].reshape(3, -1), int(round(float(packed[-1]))), int(n), group
    v_hat, W = cs.v_hat_and_w()
    r_hat = torch.sqrt(v_hat / W)

    def block(lag0, k):
        return reduce(sw.device_section("k8_ms", lambda: local_variogram_sums(trace, lag0, k)))

    def select(lag0, k, dims):
        return reduce(sw.device_section("k8_ms", lambda: local_variogram_select_sums(trace, dims, lag0, k)))

    ess = effective_n_from_variograms(v_hat, cs.m, n, block, select)
    sw.close()
    if timings is not None:
        # the stopping rule runs on the host between the lag blocks: what is left of the wall
        # time after the kernels and collectives it waited for
        timings["wall_ms"] = (time.perf_counter() - t_begin) * 1e3
        timings["finalize_ms"] = max(0.0, timings["wall_ms"] - timings.get("k8_ms", 0.0)
                                     - timings.get("allreduce_ms", 0.0))
    return r_hat, ess


# ---------------------------------------------------------------------------------------
# the reference's entry points
# ---------------------------------------------------------------------------------------
def _chains_from_get_sampler(get_sampler, n_chains, samples_per_chain):
    """`get_sampler(session)` is called once per chain like
    pysgmcmc/diagnostics/sample_chains.py:367-382 does; each sampler advances on the device
    and its `samples_per_chain` draws are stacked into one ``[n, m, D]`` trace.  A sampler
    built with ``Session(n_chains=C)`` contributes C chains at once."""
    from ..session import Session
    traces, names, sizes = [], None, None
    for _ in range(n_chains):
        session = Session(output="torch")
        sampler = get_sampler(session)
        trace, _ = sampler.run(samples_per_chain)
        traces.append(trace.to(torch.float32))
        if names is None:
            names = [get_name(p, "param_%d" % i) for i, p in enumerate(sampler.params)]
            sizes = list(sampler._sizes)
    return torch.cat(traces, dim=1).contiguous(), names, sizes


def _per_variable(values, names, sizes):
    out, off = {}, 0
    values = np.asarray(values)
    for name, n in zip(names, sizes):
        out[name] = values[off:off + n]
        off += n
    return out


def effective_sample_sizes(get_sampler, n_chains=2, samples_per_chain=100):
    """ESS of the sampler returned by `get_sampler` (sampler_diagnostics.py:47-115):
    dict ``variable name -> array with one value per dimension``."""
    trace, names, sizes = _chains_from_get_sampler(get_sampler, n_chains, samples_per_chain)
    return _per_variable(effective_n_from_trace(trace), names, sizes)


def gelman_rubin(get_sampler, n_chains=2, samples_per_chain=100):
    """Potential scale reduction factors (sampler_diagnostics.py:118-194)."""
    trace, names, sizes = _chains_from_get_sampler(get_sampler, n_chains, samples_per_chain)
    if trace.shape[1] < 2:
        raise ValueError("Gelman-Rubin diagnostic requires multiple chains of the same length.")
    return _per_variable(gelman_rubin_from_trace(trace).cpu().numpy(), names, sizes)
