"""BayesianNeuralNetwork -- the caller of the sampler hot path, same interface as
pysgmcmc/models/bayesian_neural_network.py:146-630 (`train` / `predict`), running on the
CUDA kernels: minibatch indices K7, cost + gradient K4, SGHMC update K1 (driven from C,
K5) and the predictive forward pass K10.

What is the same: constructor arguments and their validation, normalisation, the sampling
schedule (burn-in, one network kept every `sample_steps` iterations after burn-in, stop at
`n_nets`), the minibatch index stream (bit-exact with the reference's
``generate_batches(seed=seed)``), the predictive mean / variance formulas.
What differs: `session` is a :class:`pysgmcmc_b200.Session` (with ``Session(n_chains=C)`` C
independent chains sample in parallel and each kept iteration yields C networks); `dtype` is a
torch dtype and defaults to float32 (float64 models sample through the differentiable torch cost
and the float64 update kernels -- the fused BNN kernels are float32 only); `get_net` is
`get_default_net`, an ``MLPNet(hidden=...)`` of any widths (native kernels: csrc/mlp.cu) or a
``TorchNet`` (autograd), see models/networks.py; weight initialisation uses a torch generator
(TensorFlow's stream is not reproducible).
"""
import logging
from collections import deque
from time import time

import numpy as np
import torch

from .. import _native
from ..data_batches import DeviceBatchGenerator, generate_batches
from ..placeholders import placeholder
from ..sampling import Sampler
from ..session import Session
from ..stepsize_schedules import ConstantStepsizeSchedule
from .base_model import (BaseModel, zero_mean_unit_var_normalization,
                         zero_mean_unit_var_unnormalization)
from .bnn_cost import (BayesianNeuralNetworkNLL, default_net_params, log_variance_prior_log_like,  # noqa: F401
                       network_output, weight_prior_log_like)
from .networks import DEFAULT_NET, MLPNet, TorchNet, as_network  # noqa: F401


def get_default_net(inputs, params):
    """The default architecture (bayesian_neural_network.py:28-69): three 50-unit tanh layers,
    a linear head and a learned log-variance, as a function of explicit parameters."""
    return network_output(params, inputs)


class BayesianNeuralNetwork(object):
    def __init__(self, session=None, sampling_method=Sampler.SGHMC,
                 get_net=get_default_net,
                 batch_generator=generate_batches,
                 batch_size=20,
                 stepsize_schedule=ConstantStepsizeSchedule(np.sqrt(1e-4)),
                 n_nets=100, n_iters=50000,
                 burn_in_steps=1000, sample_steps=100,
                 normalize_input=True, normalize_output=True,
                 seed=None, dtype=torch.float32, **sampler_kwargs):
        # input checks of bayesian_neural_network.py:238-268: wrong types / values raise
        # AssertionError (pinned by tests/bayesian_neural_network/test_invalid_inputs.py:17-100)
        counts = {"n_nets": (n_nets, 1), "n_iters": (n_iters, 1), "burn_in_steps": (burn_in_steps, 0),
                  "sample_steps": (sample_steps, 1), "batch_size": (batch_size, 1)}
        for name, (value, lowest) in counts.items():
            assert isinstance(value, int) and value >= lowest, "%s must be an integer >= %d" % (name, lowest)
        assert isinstance(dtype, torch.dtype)
        assert callable(get_net) and callable(batch_generator)
        assert all(hasattr(stepsize_schedule, a) for a in ("update", "__next__"))

        if not Sampler.is_supported(sampling_method):
            raise ValueError(
                "'BayesianNeuralNetwork.__init__' received unsupported input "
                "for parameter 'sampling_method'. Input was: {input}.\n"
                "Supported sampling methods are enumerated in "
                "'Sampler' enum type.".format(input=sampling_method)
            )
        #: the architecture behind `get_net` (raises ValueError for callables without parameters)
        self.net = as_network(get_net, default_callable=get_default_net)

        self.sampling_method = sampling_method
        self.stepsize_schedule = stepsize_schedule
        self.get_net = get_net
        self.batch_generator = batch_generator
        self.normalize_input = normalize_input
        self.normalize_output = normalize_output
        self.n_nets = n_nets
        self.n_iters = n_iters
        self.batch_size = batch_size
        self.sampler_kwargs = sampler_kwargs
        self.burn_in_steps = burn_in_steps
        self.sample_steps = sample_steps
        self.samples = deque(maxlen=n_nets)
        self.seed = seed
        self.dtype = dtype
        self.session = Session() if session is None else session
        self.is_trained = False

    # ------------------------------------------------------------------ train
    @BaseModel._check_shapes_train
    def train(self, X, y, *args, **kwargs):
        """Sample `n_nets` networks from the posterior given X ``(N, D)``, y ``(N,)``
        (bayesian_neural_network.py:391-533)."""
        start_time = time()

        self.X, self.y = X, y
        if self.normalize_input:
            self.X, self.x_mean, self.x_std = zero_mean_unit_var_normalization(self.X)
        if self.normalize_output:
            self.y, self.y_mean, self.y_std = zero_mean_unit_var_normalization(self.y)

        n_datapoints, n_inputs = X.shape
        device = self.session.device
        # Session(n_chains=C): C independent chains sample the same posterior in one set of
        # kernel launches and every kept iteration contributes C networks (an extension; the
        # reference is one chain per model, which is what n_chains=None gives)
        n_chains = self.session.n_chains
        chain_session = Session(device=device, n_chains=n_chains, output="torch", stream=self.session.stream,
                                chain_offset=self.session.chain_offset)

        # minibatches: the reference's default generator is replaced by its on-device twin
        # (same RandomState stream, data resident in HBM); any other generator is called like
        # the reference calls it and feeds host minibatches through placeholders
        if self.batch_generator is generate_batches:
            seed = int(np.random.randint(1, 100000)) if self.seed is None else self.seed    # data_batches.py:101-102
            seeds = [seed] if n_chains is None else [(seed + j) % 2 ** 32 for j in range(n_chains)]
            batches = DeviceBatchGenerator(n_datapoints, self.batch_size, seeds=seeds, device=device)
            self.nll = BayesianNeuralNetworkNLL(n_datapoints, self.batch_size, X=self.X, y=self.y,
                                                starts_placeholder=batches.starts_placeholder,
                                                device=device, dtype=self.dtype, net=self.net)
        else:
            if n_chains is not None:
                raise ValueError("Session(n_chains=C) needs the default `generate_batches` (per-chain "
                                 "minibatch streams are generated on the device)")
            self.X_Minibatch = placeholder(name="X_Minibatch")
            self.Y_Minibatch = placeholder(name="Y_Minibatch")
            batches = self.batch_generator(x=self.X, x_placeholder=self.X_Minibatch,
                                           y=self.y, y_placeholder=self.Y_Minibatch,
                                           batch_size=self.batch_size, seed=self.seed)
            self.nll = BayesianNeuralNetworkNLL(n_datapoints, self.batch_size, n_in=n_inputs,
                                                x_placeholder=self.X_Minibatch,
                                                y_placeholder=self.Y_Minibatch,
                                                device=device, dtype=self.dtype, net=self.net)

        self.network_params = self.net.init_params(n_inputs, n_chains=n_chains, seed=self.seed, dtype=self.dtype,
                                                   device=device)
        self._param_shapes = self.net.parameter_shapes(n_inputs)
        self.samples.clear()

        self.sampler_kwargs.update({
            "params": self.network_params,
            "cost_fun": self.nll,
            "batch_generator": batches,
            "session": chain_session,
            "seed": self.seed,
            "dtype": self.dtype,
            "stepsize_schedule": self.stepsize_schedule,
        })
        if Sampler.is_burn_in_mcmc(self.sampling_method):
            self.sampler_kwargs.update({
                "scale_grad": n_datapoints,
                "burn_in_steps": self.burn_in_steps,
            })

        self.sampler = Sampler.get_sampler(self.sampling_method, **self.sampler_kwargs)

        logging.info("Starting sampling")

        def log_full_training_error(iteration_index, is_sampling):
            if not logging.getLogger().isEnabledFor(logging.INFO):
                return
            total_nll, total_mse = self._full_data_nll_mse()
            seconds_elapsed = time() - start_time
            if is_sampling:
                logging.info("Iter {:8d} : NLL = {:.4e} MSE = {:.4e} "
                             "Time = {:5.2f}".format(iteration_index, total_nll, total_mse, seconds_elapsed))
            else:
                logging.info("Iter {:8d} : NLL = {:.4e} MSE = {:.4e} "
                             "Samples = {} Time = {:5.2f}".format(iteration_index, total_nll, total_mse,
                                                                  len(self.samples), seconds_elapsed))

        logging_intervals = {"burn-in": 512, "sampling": self.sample_steps}

        # The reference inspects every iteration of islice(sampler, n_iters); only the
        # iterations where it logs or keeps a sample matter, so the chain advances on the
        # device from one such iteration to the next (bayesian_neural_network.py:510-531).
        steps_done = 0
        for iteration_index in range(self.n_iters):
            burning_in = iteration_index <= self.burn_in_steps
            log_burn = burning_in and iteration_index % logging_intervals["burn-in"] == 0
            keep = not burning_in and iteration_index % logging_intervals["sampling"] == 0
            if not (log_burn or keep):
                continue
            self.sampler.run(iteration_index + 1 - steps_done, keep_every=10 ** 9)
            steps_done = iteration_index + 1
            if log_burn:
                log_full_training_error(iteration_index, is_sampling=False)
            if keep:
                log_full_training_error(iteration_index, is_sampling=True)
                for row in self.sampler._theta.clone():           # one network per chain
                    self.samples.append(row)
                if len(self.samples) >= self.n_nets:
                    break
        if steps_done < self.n_iters and len(self.samples) < self.n_nets:
            self.sampler.run(self.n_iters - steps_done, keep_every=10 ** 9)

        self.is_trained = True

    def _full_data_nll_mse(self):
        """NLL and MSE of the current parameters on the full training set (the reference's
        `log_full_training_error`, :470-493; differentiable torch path, logging only)."""
        saved = None
        ph = getattr(self.nll, "starts_placeholder", None)
        if ph is not None:
            saved, ph.value = ph.value, None
        try:
            if self.nll.X is not None:
                cost = self.nll([p.detach() for p in self.network_params])
            else:
                self.X_Minibatch.value, self.Y_Minibatch.value = self.X, self.y.reshape(-1, 1)
                cost = self.nll([p.detach() for p in self.network_params])
            # (averaged over the chains when there are several)
            return float(torch.as_tensor(cost).float().mean()), float(torch.as_tensor(self.nll.last_mse).float().mean())
        finally:
            if ph is not None:
                ph.value = saved

    # ------------------------------------------------------------------ predict
    def compute_network_output(self, params, input_data):
        """Network output ``(N, 2)`` = (mean, log variance) for one parameter sample
        (bayesian_neural_network.py:535-557); `params` is a flat ``[D]`` tensor or the list of
        parameter arrays."""
        device = self.session.device
        if isinstance(params, (list, tuple)):
            params = torch.cat([torch.as_tensor(np.asarray(p) if not isinstance(p, torch.Tensor) else p
                                                ).reshape(-1) for p in params])
        theta = params.to(device=device).reshape(1, -1).contiguous()
        return self._forward(theta, input_data)[0]

    def _forward(self, theta, input_data):
        """(mean, log variance) of every stored network at every input: ``[n_nets, N, 2]``.
        float32 models: K10 (default architecture) or the layer kernels (any MLPNet); float64
        models and TorchNet architectures: the differentiable torch function, in the model's dtype."""
        device = self.session.device
        native = self.dtype == torch.float32 and isinstance(self.net, MLPNet)
        X = torch.as_tensor(np.asarray(input_data), dtype=torch.float32 if native else self.dtype,
                            device=device).contiguous()
        n_nets, n_points = theta.shape[0], X.shape[0]
        if not native:
            theta = theta.to(self.dtype)
            params, off = [], 0
            for shp in self.net.parameter_shapes(X.shape[1]):
                n = int(np.prod(shp))
                params.append(theta[:, off:off + n].reshape((n_nets,) + tuple(shp)))
                off += n
            with torch.no_grad():
                out = self.net(X, params)
            return out.cpu().numpy().astype(np.float64)
        theta = theta.to(torch.float32).contiguous()
        out = torch.empty((n_nets, n_points, 2), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            stream = _native.stream_ptr(self.session.stream)
            if self.net == DEFAULT_NET:
                _native.call("sgmcmc_bnn_predict_f32", _native.ptr(theta), _native.ptr(X), _native.ptr(out),
                             n_nets, X.shape[1], n_points, stream)
            else:
                widths, n_w = _native.int_array(self.net.widths(X.shape[1]))
                items = n_nets * ((n_points + 31) // 32)
                nbytes = int(_native.load().sgmcmc_mlp_workspace_bytes(widths, n_w, items, 32))
                ws = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=device)
                _native.call("sgmcmc_mlp_predict_f32", _native.ptr(theta), _native.ptr(X), _native.ptr(out),
                             _native.ptr(ws), ws.numel() * 8, n_nets, widths, n_w, n_points, stream)
        return out.cpu().numpy().astype(np.float64)

    @BaseModel._check_shapes_predict
    def predict(self, X_test, return_individual_predictions=False, *args, **kwargs):
        """Predictive mean and variance at X_test ``(N, D)`` (bayesian_neural_network.py:560-630)."""
        if not self.is_trained:
            raise ValueError(
                "Calling `bnn.predict()` on an untrained "
                "Bayesian Neural Network 'bnn' is not supported! "
                "Please call `bnn.train()` before calling `bnn.predict()`"
            )

        if self.normalize_input:
            X_, _, _ = zero_mean_unit_var_normalization(X_test, self.x_mean, self.x_std)
        else:
            X_ = X_test

        theta = torch.stack(list(self.samples)).contiguous()
        out = self._forward(theta, X_)                      # [n_nets, N, 2]
        f_out = out[:, :, 0]
        theta_noise = np.exp(out[:, :, 1])

        if return_individual_predictions:
            if self.normalize_output:
                f_out = zero_mean_unit_var_unnormalization(f_out, self.y_mean, self.y_std)
                theta_noise *= self.y_std ** 2
            return f_out, theta_noise

        mean_prediction = np.mean(f_out, axis=0)
        variance_prediction = np.mean((f_out - mean_prediction) ** 2, axis=0)

        if self.normalize_output:
            mean_prediction = zero_mean_unit_var_unnormalization(mean_prediction, self.y_mean, self.y_std)
            variance_prediction *= self.y_std ** 2

        return mean_prediction, variance_prediction
