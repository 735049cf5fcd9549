"""The BNN cost function of the hot path as a sampler `cost_fun`:
negative log likelihood of a network with a learned log-variance output, i.e.
pysgmcmc/models/bayesian_neural_network.py:77-141 (priors) and :337-388
(`negative_log_likelihood`) over a `get_net` architecture (:28-69; models/networks.py).

* ``native_cost_and_grad`` runs kernel K4 for all chains at once: the specialised 50-50-50
  kernels (csrc/bnn_mma.cuh, csrc/bnn.cu) for the BOHAMIANN default -- `SGHMCSampler` recognises
  that case and drives K4 + K1 from C (K5) without returning to Python between steps -- and the
  layer kernels of csrc/mlp.cu for every other `MLPNet` (any widths / depth, minibatches of up
  to 32 rows).
* ``__call__(params)`` is the same cost written in differentiable torch ops: it serves
  float64 samplers, SGLD / relativistic samplers on wide minibatches, arbitrary `TorchNet`
  architectures (generic autograd path) and the full-dataset logging of
  `BayesianNeuralNetwork.train`.

Flat per-chain layout = ``tf.trainable_variables()`` order of the reference; for the default
network: W1[n_in,50] b1[50] W2[50,50] b2[50] W3[50,50] b3[50] W4[50,1] b4[1] rho[1,1].
"""
import math

import numpy as np
import torch

from .. import _native
from ..placeholders import Placeholder
from ..tensor_utils import safe_divide
from .networks import DEFAULT_NET, MLPNet

HIDDEN = 50
MAX_NATIVE_BATCH = 256          # rows of one minibatch the BNN kernels accept (csrc/bnn.cu)


def parameter_shapes(n_in):
    return [(n_in, HIDDEN), (HIDDEN,), (HIDDEN, HIDDEN), (HIDDEN,), (HIDDEN, HIDDEN), (HIDDEN,),
            (HIDDEN, 1), (1,), (1, 1)]


def n_parameters(n_in):
    return sum(int(np.prod(s)) for s in parameter_shapes(n_in))


def default_net_params(n_in, n_chains=None, seed=None, dtype=torch.float32, device="cuda:0"):
    """Initial parameters of `get_default_net` (bayesian_neural_network.py:28-61); see
    `MLPNet.init_params`.  Returns a list of 9 tensors (with a leading chain axis when
    `n_chains` is given)."""
    return DEFAULT_NET.init_params(n_in, n_chains=n_chains, seed=seed, dtype=dtype, device=device)


def network_output(params, X):
    """`get_default_net` in torch ops: returns ``[..., n_points, 2]`` = (mean, log variance).
    `params` may carry a leading chain axis (then X is ``[C, B, n_in]`` or ``[B, n_in]``)."""
    return DEFAULT_NET(X, params)


def log_variance_prior_log_like(log_var, mean=1e-6, var=0.01, dtype=None):
    """bayesian_neural_network.py:77-107 (log_var: ``[..., B, 1]``)."""
    return (safe_divide(-torch.square(log_var - math.log(mean)),
                        torch.as_tensor(2.0 * var, dtype=log_var.dtype, device=log_var.device))
            - 0.5 * math.log(var)).sum(dim=-1).mean(dim=-1)


def weight_prior_log_like(parameters, wdecay=1.0, dtype=None, chains=False):
    """bayesian_neural_network.py:110-141."""
    log_like, n_params = 0.0, 0.0
    for p in parameters:
        sq = -wdecay * 0.5 * torch.square(p)
        log_like = log_like + (sq.reshape(p.shape[0], -1).sum(dim=1) if chains else sq.sum())
        n_params += float(np.prod(p.shape[1:] if chains else p.shape))
    return safe_divide(log_like, torch.as_tensor(n_params, dtype=log_like.dtype, device=log_like.device))


class BayesianNeuralNetworkNLL(object):
    """Cost function object: ``cost = -log_like`` of :337-388.

    Two ways to supply the minibatch, mirroring the reference's feed mechanism:

    * device-resident data + on-device start indices (`DeviceBatchGenerator`): pass the
      whole (normalised) dataset `X`, `y` and the generator's ``starts_placeholder``;
      chain j uses rows ``X[starts[j] : starts[j] + batch_size]``;
    * host minibatches (`generate_batches`): pass ``x_placeholder`` / ``y_placeholder``;
      every step the fed ``(B, n_in)`` / ``(B, 1)`` arrays are uploaded and used by all
      chains (the reference's single-chain behaviour).
    """

    def __init__(self, n_examples, batch_size=20, n_in=None, X=None, y=None, starts_placeholder=None,
                 x_placeholder=None, y_placeholder=None, n_chains=None, device="cuda:0",
                 dtype=torch.float32, net=None):
        self.net = DEFAULT_NET if net is None else net
        self.device = torch.device(device)
        self.dtype = dtype
        self.n_examples = int(n_examples)
        self.batch_size = int(batch_size)               # the CONFIGURED constant of :377
        self.n_chains = n_chains
        self.x_placeholder, self.y_placeholder = x_placeholder, y_placeholder
        self.starts_placeholder = starts_placeholder
        if X is not None:
            self.X = torch.as_tensor(np.asarray(X) if not isinstance(X, torch.Tensor) else X
                                     ).to(self.device, dtype).reshape(len(X), -1).contiguous()
            self.y = torch.as_tensor(np.asarray(y) if not isinstance(y, torch.Tensor) else y
                                     ).to(self.device, dtype).reshape(-1).contiguous()
            self.n_in = self.X.shape[1]
        else:
            assert x_placeholder is not None and y_placeholder is not None and n_in is not None
            self.X = self.y = None
            self.n_in = int(n_in)
        self.actual_batch = min(self.batch_size, self.n_examples)    # data_batches.py:111
        self.n_params = self.net.n_parameters(self.n_in)
        #: the default architecture: K4 + K1 can be driven from C (K5, `SGHMCSampler.run` / `iter_host`)
        self.bnn_native = self.net == DEFAULT_NET
        #: any MLPNet: cost + gradient by the layer kernels of csrc/mlp.cu
        self.mlp_native = isinstance(self.net, MLPNet) and not self.bnn_native
        self._cost = None
        self._mlp_ws = None
        self.last_mse = None

    @property
    def supports_native(self):
        """Whether `native_cost_and_grad` exists for this architecture and minibatch (else the
        sampler differentiates `__call__` with autograd)."""
        if self.bnn_native:
            return True
        return self.mlp_native and self.actual_batch <= 32 and (self.X is None or self.starts_placeholder is not None
                                                                 or self.X.shape[0] <= 32)

    # ---- where the current minibatch comes from ---------------------------------------
    def full_dataset_batch(self):
        """Rows per chain when NO minibatch indices are fed: the cost is then over the WHOLE
        resident dataset, like the differentiable path (`_torch_batch`).  The kernels hold one
        minibatch per chain in shared memory, so this only exists for small datasets; a larger
        one raises instead of silently using its first rows."""
        n = int(self.X.shape[0])
        if n > MAX_NATIVE_BATCH:
            raise ValueError(
                "BayesianNeuralNetworkNLL: the dataset is resident on the device (%d rows) but no "
                "minibatch start indices were fed through `starts_placeholder`; the native cost "
                "evaluates at most %d rows per chain. Pass a DeviceBatchGenerator as the sampler's "
                "batch_generator (or feed starts)." % (n, MAX_NATIVE_BATCH))
        return n

    def _device_batch(self):
        """(X, y, starts or None, batch) for the native kernels."""
        if self.X is not None:
            starts = None
            if self.starts_placeholder is not None and self.starts_placeholder.value is not None:
                starts = self.starts_placeholder.tensor(self.device, torch.int32).contiguous()
            if starts is not None:
                return self.X, self.y, starts, self.actual_batch
            return self.X, self.y, None, self.full_dataset_batch()
        xb = self.x_placeholder.tensor(self.device, torch.float32).reshape(-1, self.n_in).contiguous()
        yb = self.y_placeholder.tensor(self.device, torch.float32).reshape(-1).contiguous()
        return xb, yb, None, xb.shape[0]

    def _torch_batch(self):
        if self.X is not None:
            if self.starts_placeholder is not None and self.starts_placeholder.value is not None:
                starts = self.starts_placeholder.tensor(self.device, torch.int64)
                idx = starts[:, None] + torch.arange(self.actual_batch, device=self.device)[None, :]
                return self.X[idx].to(self.dtype), self.y[idx].to(self.dtype)
            return self.X.to(self.dtype), self.y.to(self.dtype)
        xb = self.x_placeholder.tensor(self.device, self.dtype).reshape(-1, self.n_in)
        yb = self.y_placeholder.tensor(self.device, self.dtype).reshape(-1)
        return xb, yb

    # ---- native path (K4) ---------------------------------------------------------------
    def native_cost_and_grad(self, theta, grad_out, want_mse=False):
        C, D = theta.shape
        assert D == self.n_params, "parameter layout does not match the %d-input network" % self.n_in
        if theta.dtype != torch.float32 or grad_out.dtype != torch.float32:
            raise TypeError("the native BNN cost + gradient kernels are float32 (got %s); float64 samplers use the "
                            "differentiable cost (call the object) with the float64 update kernels"
                            % str(theta.dtype))
        X, y, starts, batch = self._device_batch()
        assert starts is None or starts.shape[0] == C
        if self._cost is None or self._cost.shape[0] != C:
            self._cost = torch.empty(C, dtype=torch.float32, device=self.device)
        mse = torch.empty(C, dtype=torch.float32, device=self.device) if want_mse else None
        with torch.cuda.device(self.device):
            if self.bnn_native:
                _native.call("sgmcmc_bnn_nll_grad_f32", _native.ptr(theta), _native.ptr(X), _native.ptr(y),
                             _native.ptr(starts), _native.ptr(self._cost), _native.ptr(grad_out),
                             _native.ptr(mse), C, self.n_in, batch, float(self.batch_size),
                             self.n_examples, _native.stream_ptr())
            else:
                if batch > 32:
                    raise ValueError("the layer kernels (csrc/mlp.cu) take minibatches of up to 32 rows (got %d)"
                                     % batch)
                widths, n_w = _native.int_array(self.net.widths(self.n_in))
                ws = self._mlp_workspace(widths, n_w, C, batch)
                _native.call("sgmcmc_mlp_nll_grad_f32", _native.ptr(theta), _native.ptr(X), _native.ptr(y),
                             _native.ptr(starts), _native.ptr(self._cost), _native.ptr(grad_out), _native.ptr(mse),
                             _native.ptr(ws), ws.numel() * 8, C, widths, n_w, batch, float(self.batch_size),
                             self.n_examples, _native.stream_ptr())
        self.last_mse = mse
        return self._cost

    def _mlp_workspace(self, widths, n_w, n_items, batch):
        """Activation workspace of the layer kernels (int64 elements: 16-byte aligned), kept."""
        nbytes = int(_native.load().sgmcmc_mlp_workspace_bytes(widths, n_w, n_items, batch))
        if nbytes < 0:
            raise _native.NativeError("sgmcmc_mlp_workspace_bytes: unsupported architecture %r" % (self.net,))
        if self._mlp_ws is None or self._mlp_ws.numel() * 8 < nbytes:
            self._mlp_ws = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=self.device)
        return self._mlp_ws

    # ---- generic differentiable path ------------------------------------------------------
    def __call__(self, params, *_):
        params = list(params)
        chains = params[0].dim() == 3
        X, Y = self._torch_batch()
        out = self.net(X, params)
        f_mean, f_log_var = out[..., 0:1], out[..., 1:2]
        # minibatches per chain: Y [C, B]; one batch for all chains: Y [B] broadcasts over them
        Y = Y.reshape(Y.shape + (1,)) if Y.dim() == f_mean.dim() - 1 else Y.reshape(f_mean.shape[-2:])
        f_var_inv = 1.0 / (torch.exp(f_log_var) + 1e-16)
        mse = torch.square(Y - f_mean)
        log_like = (-mse * (0.5 * f_var_inv) - 0.5 * f_log_var).sum(dim=(-1, -2))
        log_like = log_like / self.batch_size
        log_like = log_like + log_variance_prior_log_like(f_log_var) / self.n_examples
        log_like = log_like + weight_prior_log_like(params, chains=chains) / self.n_examples
        self.last_mse = mse.mean(dim=(-1, -2)).detach()
        return -log_like
