"""Execution context of the engine: what the reference's ``session`` argument becomes.

The reference passes a ``tf.Session`` to every sampler
(pysgmcmc/samplers/base_classes.py:23-25,76); here the same argument carries the
device, the CUDA stream and the engine options, so sampler signatures stay
identical to the reference's.
"""
import torch


class Session(object):
    """
    device : CUDA device the chain state lives on.
    n_chains : None -> reference semantics, one chain, `params` have the reference's
        shapes.  C -> every tensor in `params` carries a leading chain axis of size C
        and all C chains advance in one kernel launch.
    output : "numpy" -> ``next(sampler)`` returns host ndarrays like the reference
        (one device->host copy per step); "torch" -> device tensors, no sync.
    chain_offset : global index of this process' first chain (multi-GPU sharding);
        only shifts the Philox substreams so a shard reproduces the un-sharded run.
    fused : use the fused whole-step kernels (K5/K6) where the cost function is native.
    prefetch : S > 1 -> ``next(sampler)`` hands out the results of steps that were computed S
        at a time by ONE launch of the fused kernels (built-in target densities, the native BNN
        cost), so a Python loop over ``next()`` costs a few microseconds per step instead of a
        kernel launch (+ device-to-host copy) per step.  The (sample, cost) pairs are exactly the
        ones the step-by-step loop returns; what differs is that the sampler's live state
        (`sampler.params`) runs up to S - 1 steps AHEAD of the sample last returned, so it is
        opt-in.  Ignored (step-by-step) whenever something is fed, the stepsize schedule is not
        constant or the cost function is not native.
    """

    def __init__(self, device="cuda:0", n_chains=None, output="numpy", stream=None,
                 chain_offset=0, fused=True, prefetch=0):
        assert output in ("numpy", "torch")
        assert n_chains is None or (isinstance(n_chains, int) and n_chains > 0)
        assert isinstance(prefetch, int) and prefetch >= 0
        self.prefetch = prefetch
        self.device = torch.device(device)
        self.n_chains = n_chains
        self.output = output
        self.stream = stream
        self.chain_offset = int(chain_offset)
        self.fused = fused

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False
