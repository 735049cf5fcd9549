"""Regenerates the fixtures under tests/golden/.

Run in the BUILD container (needs /root/reference for part 1):
    python tests/golden/make_golden.py

0. relativistic_ess_published.json -- REFERENCE DATA: the published ESS-vs-stepsize table
   (see make_published_ess).
1. bnn_priors.npz  -- REFERENCE DATA: the reference's own golden vectors
   (pysgmcmc/tests/data/bayesian_neural_network_priors/{weights_inputs,weights,
   log_variance}.npy, asserted bit-exactly in
   pysgmcmc/tests/bayesian_neural_network/test_priors.py:20-81), repacked from a
   pickled object array into a plain npz so the GPU box (no /root/reference) can
   read them.
2. trajectories.npz / bnn_nll.npz -- ORACLE DATA: the reference cannot run here
   (TensorFlow 1.x), and it holds no golden sampler trajectory, so these are
   produced by oracle/ (the line-by-line restatement) on seeded inputs.  They
   pin the oracle against silent edits and give the CUDA tests a second,
   file-based comparison point.  Noise is NOT stored: it is
   ``RandomState(seed).standard_normal`` (a frozen legacy stream).
3. svgd.npz -- ORACLE DATA (float64): the reference has no SVGD test or golden either
   (only docs/source/notebooks/SVGD.ipynb's plot); the notebook's configuration
   (banana, 10 particles, N(0, 1) starts, stepsize 0.1) and a 64-particle one, first
   50 steps (later steps are chaotic, see tests/test_svgd.py), plus kernel matrix and
   bandwidth of the start.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle import bnn, samplers, svgd, targets  # noqa: E402

REF = "/root/reference/pysgmcmc/tests/data/bayesian_neural_network_priors"
CHECKPOINTS = (1, 2, 10, 100, 500, 1000)


def make_priors():
    wi = np.load(os.path.join(REF, "weights_inputs.npy"), allow_pickle=True)
    out = {"w%d" % i: np.asarray(w, dtype=np.float64) for i, w in enumerate(wi)}
    out["expected_weights"] = np.load(os.path.join(REF, "weights.npy"))
    out["expected_log_variance"] = np.load(os.path.join(REF, "log_variance.npy"))
    out["f_log_var"] = np.full((20, 1), -11.25474104)   # test_priors.py:23-44
    np.savez(os.path.join(HERE, "bnn_priors.npz"), **out)


def trajectory_case(method, target, dtype, seed, n_chains=4, n_steps=1000, **hyper):
    """theta at CHECKPOINTS and all costs for `n_chains` chains started at
    spread-out points (chain 0 = the reference's test start:
    tests/samplers/sampler_testing.py:15-18)."""
    rng = np.random.RandomState(seed)
    D = 2 if target == "banana" else 1
    theta0 = np.zeros((n_chains, D))
    theta0[0] = (0.0, 6.0) if target == "banana" else (0.0,)
    theta0[1:] = rng.uniform(-3, 3, size=(n_chains - 1, D))
    theta0 = theta0.astype(dtype)
    momentum = rng.standard_normal((n_chains, D)).astype(dtype) if method == "rsghmc" else None
    chain = samplers.OracleChain(method, theta0, targets.cost_and_grad(target),
                                 momentum=momentum, **hyper)
    z_rng = np.random.RandomState(seed + 1)
    thetas, costs = [], []
    for step in range(1, n_steps + 1):
        z = z_rng.standard_normal((n_chains, D)).astype(dtype)
        theta, cost = chain.next(z)
        costs.append(cost)
        if step in CHECKPOINTS:
            thetas.append(theta)
    out = dict(theta0=theta0, theta=np.stack(thetas), cost=np.stack(costs))
    if momentum is not None:
        out["momentum0"] = momentum
    return out


CASES = [
    # name, method, target, dtype, seed, hyper
    ("sghmc_banana_f32", "sghmc", "banana", np.float32, 11, dict(burn_in_steps=300)),
    ("sghmc_banana_f64", "sghmc", "banana", np.float64, 11, dict(burn_in_steps=300)),
    ("sghmc_gmm1_f32", "sghmc", "gmm1", np.float32, 12, dict(burn_in_steps=300, scale_grad=4.0, mdecay=0.1)),
    ("sghmc_banana_noburn_f32", "sghmc", "banana", np.float32, 13, dict(burn_in_steps=0)),
    ("sgld_gmm1_f32", "sgld", "gmm1", np.float32, 21, dict(burn_in_steps=300)),
    ("sgld_gmm3_f32", "sgld", "gmm3", np.float32, 22, dict(burn_in_steps=300, A=2.0, scale_grad=3.0)),
    ("sgld_banana_f32", "sgld", "banana", np.float32, 23, dict(burn_in_steps=300)),
    ("rsghmc_banana_f32", "rsghmc", "banana", np.float32, 31, dict(epsilon=0.01)),
    ("rsghmc_gmm2_f32", "rsghmc", "gmm2", np.float32, 32,
     dict(epsilon=0.05, mass=1.5, speed_of_light=0.8, D=1.2, Bhat=0.1)),
]


def make_trajectories():
    out = {}
    for name, method, target, dtype, seed, hyper in CASES:
        res = trajectory_case(method, target, dtype, seed, **hyper)
        for k, v in res.items():
            out["%s/%s" % (name, k)] = v
    np.savez_compressed(os.path.join(HERE, "trajectories.npz"), **out)


def bnn_case(seed=5, n_chains=3, N=200, B=20):
    rng = np.random.RandomState(seed)
    X = rng.uniform(0, 1, size=(N, 1))
    y = np.sinc(X * 10 - 5).sum(axis=1)
    X = (X - X.mean(0)) / X.std(0)
    y = (y - y.mean()) / y.std()
    theta = bnn.init_theta(n_chains, seed=seed, dtype=np.float64)
    theta += 0.05 * rng.standard_normal(theta.shape)            # non-zero biases
    starts = rng.randint(0, N - B + 1, size=n_chains)
    return X, y, theta, starts


def make_bnn():
    X, y, theta, starts = bnn_case()
    Xb, yb = bnn.gather_minibatch(X, y, starts, 20)
    cost, grad, mse = bnn.nll_and_grad(theta, Xb, yb, n_examples=X.shape[0])
    np.savez_compressed(os.path.join(HERE, "bnn_nll.npz"), X=X, y=y, theta=theta, starts=starts,
                        cost=cost, grad=grad, mse=mse)


SVGD_CASES = (("banana_10", 10, 4), ("banana_64", 64, 5))
SVGD_CHECKPOINTS = (1, 2, 5, 10, 20, 50)


def svgd_case(n_particles, seed, n_steps=50):
    X0 = np.random.RandomState(seed).standard_normal((n_particles, 2))
    K, kgrad, h = svgd.svgd_kernel(X0)
    ref = svgd.OracleSVGD(X0, targets.banana_cost_and_grad)
    thetas, costs = [], []
    for step in range(1, n_steps + 1):
        theta, cost = next(ref)
        costs.append(cost)
        if step in SVGD_CHECKPOINTS:
            thetas.append(theta)
    return dict(theta0=X0, kernel_matrix=K, kernel_gradients=kgrad, bandwidth=np.float64(h),
                theta=np.stack(thetas), cost=np.stack(costs),
                historical_grad=ref.state["historical_grad"])


def make_svgd():
    out = {}
    for name, n, seed in SVGD_CASES:
        for k, v in svgd_case(n, seed).items():
            out["%s/%s" % (name, k)] = v
    np.savez_compressed(os.path.join(HERE, "svgd.npz"), **out)


def make_published_ess():
    """REFERENCE DATA: the only numbers the reference publishes for this path -- mean ESS of
    Relativistic SGHMC vs stepsize (docs/source/notebooks/data/effective_sample_sizes/
    Relativistic_SGHMC.json, produced by docs/source/experiments/compute_ess.py:177-253:
    20 segments x 10 000 kept draws, keep_every = 10, 5 repeats).  Repacked as
    {target: [[stepsize, mean over repeats, min, max], ...]} sorted by stepsize."""
    import json
    src = "/root/reference/docs/source/notebooks/data/effective_sample_sizes/Relativistic_SGHMC.json"
    data = json.load(open(src))
    out = {}
    for target, table in data.items():
        rows = []
        for eps, reps in table.items():
            vals = np.array(reps, dtype=np.float64).ravel()
            rows.append([float(eps), float(vals.mean()), float(vals.min()), float(vals.max())])
        out[target] = sorted(rows)
    with open(os.path.join(HERE, "relativistic_ess_published.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    make_priors()
    make_published_ess()
    make_trajectories()
    make_bnn()
    make_svgd()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
