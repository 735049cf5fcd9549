"""Network architectures for `BayesianNeuralNetwork(get_net=...)`.

The reference's `get_net(inputs, seed, dtype)` builds a TensorFlow graph whose variables are
collected implicitly (``tf.trainable_variables()``, pysgmcmc/models/bayesian_neural_network.py:28-69,
:337-366, :432).  Here parameters are explicit, so an architecture is an object with

* ``parameter_shapes(n_inputs) -> [shape, ...]`` -- the parameter tensors of ONE chain in sampler
  order (the flat per-chain layout is their row-major concatenation),
* ``init_params(n_inputs, n_chains, seed, dtype, device) -> [tensor, ...]`` -- initial values
  (with a leading chain axis when `n_chains` is given),
* ``__call__(inputs, params) -> [..., n_points, 2]`` -- (mean, log variance) in differentiable
  torch ops; `params` may carry a leading chain axis.

`MLPNet` (fully connected tanh layers + linear head + learned log variance, any widths / depth)
is what the CUDA kernels implement natively: ``MLPNet((50, 50, 50))`` is `get_default_net` (K4 /
K10), any other widths run the layer kernels of csrc/mlp.cu -- e.g. ``MLPNet((1000, 512, 512))``,
the wide network of BASELINE.json configs[4].  `TorchNet` wraps an arbitrary torch function: cost
and gradient then go through autograd, the sampler update stays on the engine's kernels.
"""
import math

import numpy as np
import torch


class MLPNet(object):
    """n_inputs -> hidden[0] -> ... -> hidden[-1] -> 1 with tanh hidden layers, a linear head and
    a learned scalar log variance concatenated as the second output column."""

    native = True

    def __init__(self, hidden=(50, 50, 50)):
        hidden = tuple(int(h) for h in hidden)
        if not 1 <= len(hidden) <= 7 or any(h < 1 for h in hidden):
            raise ValueError("MLPNet: 1 to 7 hidden layers of positive width (got %r)" % (hidden,))
        self.hidden = hidden

    def widths(self, n_inputs):
        return [int(n_inputs)] + list(self.hidden) + [1]

    def parameter_shapes(self, n_inputs):
        w = self.widths(n_inputs)
        shapes = []
        for l in range(1, len(w)):
            shapes += [(w[l - 1], w[l]), (w[l],)]
        return shapes + [(1, 1)]

    def n_parameters(self, n_inputs):
        return sum(int(np.prod(s)) for s in self.parameter_shapes(n_inputs))

    def init_params(self, n_inputs, n_chains=None, seed=None, dtype=torch.float32, device="cuda:0"):
        """Initialisers of `get_default_net` (bayesian_neural_network.py:31-61): kernels ~ truncated
        normal(0, sqrt(1.3 / fan_in)) (tf.contrib `variance_scaling_initializer(factor=1.0)`: FAN_IN,
        truncated at two standard deviations and rescaled), zero biases, log variance log(1e-3).
        TensorFlow's random stream cannot be reproduced (and the reference pins none,
        tests/bayesian_neural_network/test_seeding.py:14-46 only asks for same seed -> same net), so
        the draws come from a seeded torch generator."""
        gen = torch.Generator(device="cpu")
        gen.manual_seed(int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed))
        lead = () if n_chains is None else (n_chains,)
        out = []
        shapes = self.parameter_shapes(n_inputs)
        for k, shp in enumerate(shapes):
            if k == len(shapes) - 1:
                out.append(torch.full(lead + shp, math.log(1e-3), dtype=dtype))
            elif len(shp) == 2:
                std = math.sqrt(1.3 / shp[0])
                w = torch.empty(lead + shp, dtype=torch.float64)
                torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=gen)
                out.append((w * std).to(dtype))
            else:
                out.append(torch.zeros(lead + shp, dtype=dtype))
        return [p.to(device) for p in out]

    def __call__(self, inputs, params):
        params = list(params)
        rho = params[-1]
        chains = params[0].dim() == 3
        bias = (lambda b: b[:, None, :]) if chains else (lambda b: b)
        h = inputs
        n_dense = (len(params) - 1) // 2
        for l in range(n_dense):
            z = h @ params[2 * l] + bias(params[2 * l + 1])
            h = torch.tanh(z) if l < n_dense - 1 else z
        return torch.cat([h, torch.ones_like(h) * rho], dim=-1)

    def __eq__(self, other):
        return isinstance(other, MLPNet) and other.hidden == self.hidden

    def __hash__(self):
        return hash(("MLPNet", self.hidden))

    def __repr__(self):
        return "MLPNet(hidden=%r)" % (self.hidden,)


class TorchNet(object):
    """An arbitrary architecture written in torch ops.

    forward(inputs, params) -> ``[..., n_points, 2]`` (mean, log variance); it must broadcast over
    a leading chain axis of `params` if the model is used with ``Session(n_chains=C)``.
    shapes(n_inputs) -> list of parameter shapes; init(index, shape, generator) -> CPU tensor of
    that shape for parameter `index` (default: truncated normal scaled by sqrt(1.3 / fan_in) for
    matrices, zeros for vectors, log(1e-3) for a trailing (1, 1) parameter -- the log variance of
    `get_default_net`)."""

    native = False

    def __init__(self, forward, shapes, init=None):
        assert callable(forward) and callable(shapes)
        self.forward, self.shapes, self.init = forward, shapes, init

    def parameter_shapes(self, n_inputs):
        return [tuple(s) for s in self.shapes(n_inputs)]

    def n_parameters(self, n_inputs):
        return sum(int(np.prod(s)) for s in self.parameter_shapes(n_inputs))

    def init_params(self, n_inputs, n_chains=None, seed=None, dtype=torch.float32, device="cuda:0"):
        gen = torch.Generator(device="cpu")
        gen.manual_seed(int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed))
        lead = () if n_chains is None else (n_chains,)
        out = []
        shapes = self.parameter_shapes(n_inputs)
        for k, shp in enumerate(shapes):
            if self.init is not None:
                w = torch.as_tensor(self.init(k, lead + shp, gen))
            elif k == len(shapes) - 1 and shp == (1, 1):
                w = torch.full(lead + shp, math.log(1e-3), dtype=torch.float64)
            elif len(shp) == 2:
                w = torch.empty(lead + shp, dtype=torch.float64)
                torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=gen)
                w = w * math.sqrt(1.3 / shp[0])
            else:
                w = torch.zeros(lead + shp, dtype=torch.float64)
            out.append(w.to(dtype).to(device))
        return out

    def __call__(self, inputs, params):
        return self.forward(inputs, list(params))


DEFAULT_NET = MLPNet((50, 50, 50))


def as_network(get_net, default_callable=None):
    """What `BayesianNeuralNetwork(get_net=...)` accepts: the default callable (-> the
    50-50-50 MLP), an `MLPNet` / `TorchNet`, or any object with the three methods above."""
    if get_net is default_callable or get_net is None:
        return DEFAULT_NET
    if all(hasattr(get_net, a) for a in ("parameter_shapes", "init_params")) and callable(get_net):
        return get_net
    raise ValueError(
        "get_net must be `get_default_net`, an MLPNet(hidden=...) / TorchNet(...) or an object with "
        "`parameter_shapes(n_inputs)`, `init_params(n_inputs, n_chains, seed, dtype, device)` and "
        "`__call__(inputs, params)`: parameters are explicit in this engine, a bare TensorFlow-style "
        "`get_net(inputs, seed, dtype)` has no variables to sample (see models/networks.py)")
