"""Oracle restatement of Stein variational gradient descent as pysgmcmc implements it
(test infrastructure only, see oracle/__init__.py).

Follows, line by line:
  * pysgmcmc/tensor_utils.py:160-208   (`median`: full descending sort, middle value or
                                        the mean of the two middle values)
  * pysgmcmc/tensor_utils.py:326-419   (`pdist`: condensed vector of ``norm(x_i - x_j)``, i < j)
  * pysgmcmc/tensor_utils.py:422-577   (`squareform`: condensed vector -> symmetric matrix)
  * pysgmcmc/samplers/svgd.py:150-182  (`svgd_kernel`: RBF kernel, median bandwidth)
  * pysgmcmc/samplers/svgd.py:125-148  (`svgd_step`: Stein direction, AdaGrad history, update)

All arithmetic is done op by op in the dtype of the particles with the reference's
parenthesisation (Python scalars converted to that dtype first, as TensorFlow does
for constants).

Pinned by the reference's own tests: `pdist` and `squareform` are asserted equal to
``scipy.spatial.distance.pdist/squareform`` (pysgmcmc/tests/test_tensor_utils.py:46-86
and the doctests :356-366, :447-455) and scipy is importable here; `median` by its
doctests (:180-192).  SVGD *trajectories*: **parity unpinned** (the reference has no
test or golden for svgd.py, only docs/source/notebooks/SVGD.ipynb's plot); this
restatement is the pin.

Quirks kept on purpose: `lnpgrad` is the gradient of the COST (= minus log density)
and the particles move by ``-epsilon * adj_grad`` (svgd.py:125,143-148), so the
kernel-gradient term enters with the sign the reference gives it; the median runs over
all n*n entries of the squared-distance matrix including the n zeros of the diagonal;
the AdaGrad history starts at zero (no first-step special case).
"""
import numpy as np


def _t(dtype):
    return np.dtype(dtype).type


def median(tensor):
    """tensor_utils.py:194-208."""
    flat = np.asarray(tensor).reshape(-1)
    n_elements = flat.shape[0]
    values = np.sort(flat)[::-1]                 # tf.nn.top_k(..., sorted=True): descending
    mid_index = n_elements // 2
    if n_elements % 2 == 1:
        return values[mid_index]
    return (values[mid_index - 1] + values[mid_index]) / _t(flat.dtype)(2)


def pdist(tensor, metric="euclidean"):
    """tensor_utils.py:393-419."""
    tensor = np.asarray(tensor)
    if tensor.ndim != 2:
        raise ValueError('tensor_utils.pdist: A 2-d tensor must be passed.')
    if metric != "euclidean":
        raise NotImplementedError(
            "tensor_utils.pdist: Metric '{metric}' currently not supported!".format(metric=metric))
    m = tensor.shape[0]
    distances = []
    for i in range(m):
        for j in range(i + 1, m):
            diff = tensor[i] - tensor[j]
            distances.append(np.sqrt(np.sum(diff * diff)))        # tf.norm
    return np.asarray(distances, dtype=tensor.dtype)


def squareform(tensor):
    """tensor_utils.py:459-577 (vector -> matrix only)."""
    tensor = np.asarray(tensor)
    if tensor.ndim != 1:
        raise NotImplementedError("tensor_utils.squareform: Only 1-d (vector) input is supported!")
    n_elements = tensor.shape[0]
    if n_elements == 0:
        return np.zeros((1, 1), dtype=tensor.dtype)
    dimension = int(np.ceil(np.sqrt(n_elements * 2)))
    if dimension * (dimension - 1) != n_elements * 2:
        raise ValueError("Incompatible vector size. It must be a binomial "
                         "coefficient n choose 2 for some integer n >=2.")
    upper = np.zeros((dimension, dimension), dtype=tensor.dtype)
    upper[np.triu_indices(dimension, k=1)] = tensor               # row-major i < j order
    return upper + upper.T


def svgd_kernel(particles):
    """svgd.py:150-182.  Returns (kernel_matrix, kernel_gradients, h)."""
    particles = np.asarray(particles)
    T = _t(particles.dtype)
    n_particles = T(particles.shape[0])
    pairwise_distances = squareform(pdist(particles)) ** 2
    h = np.sqrt(T(0.5) * median(pairwise_distances) / np.log(n_particles + T(1.0)))
    kernel_matrix = np.exp(-pairwise_distances / (h * h) / T(2))
    kernel_sum = np.sum(kernel_matrix, axis=1)
    kernel_gradients = (-(kernel_matrix @ particles)) + particles * kernel_sum[:, None]
    return kernel_matrix, kernel_gradients / (h * h), h


def svgd_init(particles):
    particles = np.array(particles, copy=True)
    return dict(theta=particles, historical_grad=np.zeros_like(particles))


def svgd_step(state, grad, epsilon, alpha=0.9, fudge_factor=1e-6):
    """One update of all particles (svgd.py:125-148).  `grad` = d cost_i / d particle_i,
    shape ``[n_particles, D]``.  Updates `state` in place."""
    theta, hist = state["theta"], state["historical_grad"]
    T = _t(theta.dtype)
    n_particles = T(theta.shape[0])
    kernel_matrix, kernel_gradients, _ = svgd_kernel(theta)
    grad_theta = (kernel_matrix @ np.asarray(grad, dtype=theta.dtype) + kernel_gradients) / n_particles
    hist_t = T(alpha) * hist + T(1. - alpha) * (grad_theta * grad_theta)
    adj_grad = grad_theta / (T(fudge_factor) + np.sqrt(hist_t))
    state["historical_grad"] = hist_t
    state["theta"] = theta - T(epsilon) * adj_grad
    return state


class OracleSVGD(object):
    """`sample, cost = next(sampler)` for SVGD with a cost/gradient callable
    ``cost_and_grad(particles[n, D]) -> (cost[n], grad[n, D])`` (svgd.py:81-148 on top of
    base_classes.py:258-310: the cost returned belongs to the PRE-update particles)."""

    def __init__(self, particles, cost_and_grad, epsilon=0.1, alpha=0.9, fudge_factor=1e-6):
        self.state = svgd_init(particles)
        self.cost_and_grad = cost_and_grad
        self.epsilon, self.alpha, self.fudge_factor = epsilon, alpha, fudge_factor

    def __iter__(self):
        return self

    def __next__(self):
        cost, grad = self.cost_and_grad(self.state["theta"])
        svgd_step(self.state, grad, self.epsilon, self.alpha, self.fudge_factor)
        return self.state["theta"].copy(), np.asarray(cost)
