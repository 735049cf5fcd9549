"""Drift of a BNN-SGHMC trajectory on the GPU against the float32 AND the float64 oracle at the
shapes the benchmark is quoted on (BASELINE.json configs[2]: N = 20 000, minibatch 20,
scale_grad = N, eps = 0.01, burn-in boundary inside the run), with injected N(0,1) draws and the
bit-exact minibatch index stream, for both implementations of K4.

    python tools/bnn_trajectory_drift.py [--steps 1000] [--burn 600] [--chains 4] [--every 100]

Prints one JSON line per K4 implementation: max |theta_gpu - theta_oracle| / max |theta| at
every checkpoint for the float32 oracle and for the float64 oracle, the drift of the float32
ORACLE against the float64 one (what float32 arithmetic alone costs), and the first checkpoint
at which 1e-5 is crossed (None if never).  The test in tests/test_bnn_gpu.py asserts on the same
function.  (Test infrastructure: imports oracle/.)
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N, BATCH, D = 20000, 20, 5252


def sinc_data(n=N, seed=1):
    rng = np.random.RandomState(seed)
    X = np.array([rng.uniform(0.0, 1.0, 1) for _ in range(n)])
    y = np.sinc(X * 10 - 5).sum(axis=1)
    X = (X - X.mean(axis=0)) / X.std(axis=0)
    y = (y - y.mean()) / y.std()
    return X, y


def oracle_run(theta0, X, y, seeds, steps, burn, z_seed, dtype, every):
    """Checkpoints [(step, theta)] of the oracle chain in `dtype` (same z and minibatches)."""
    from oracle import bnn as obnn, mt19937 as omt, samplers as osamplers
    C = theta0.shape[0]
    streams = [omt.MT19937(int(s)) for s in seeds]
    holder = {}
    Xd, yd = X.astype(dtype), y.astype(dtype)

    def cost_and_grad(theta):
        Xb, yb = obnn.gather_minibatch(Xd, yd, holder["starts"], BATCH)
        c, g, _ = obnn.nll_and_grad(theta, Xb, yb, n_examples=N)
        return c, g
    chain = osamplers.OracleChain("sghmc", theta0.astype(dtype), cost_and_grad, epsilon=0.01,
                                  burn_in_steps=burn, scale_grad=float(N))
    zr = np.random.RandomState(z_seed)
    out = []
    for s in range(steps):
        holder["starts"] = np.array([st.bounded(N - BATCH) for st in streams])
        z = zr.standard_normal((C, D)).astype(np.float32)
        theta, _ = chain.next(z.astype(dtype))
        if (s + 1) % every == 0:
            out.append((s + 1, theta.copy()))
    return out, chain


def gpu_run(theta0, X, y, seeds, steps, burn, z_seed, every, variant, dev="cuda:0"):
    import torch
    from pysgmcmc_b200 import Session, _native
    from pysgmcmc_b200.data_batches import DeviceBatchGenerator
    from pysgmcmc_b200.models.bnn_cost import BayesianNeuralNetworkNLL, parameter_shapes
    from pysgmcmc_b200.samplers import SGHMCSampler
    from pysgmcmc_b200.stepsize_schedules import ConstantStepsizeSchedule
    C = theta0.shape[0]
    if variant == "resident":
        return gpu_run_resident(theta0, X, y, seeds, steps, burn, z_seed, every, dev), None
    _native.call("sgmcmc_set_bnn_tuning", variant)
    try:
        gen = DeviceBatchGenerator(N, BATCH, seeds=seeds, device=dev, block=256)
        nll = BayesianNeuralNetworkNLL(N, BATCH, X=X, y=y, starts_placeholder=gen.starts_placeholder, device=dev)
        params, off = [], 0
        for shp in parameter_shapes(1):
            n = int(np.prod(shp))
            params.append(torch.tensor(theta0[:, off:off + n].reshape((C,) + shp), device=dev))
            off += n
        sampler = SGHMCSampler(params=params, cost_fun=nll, batch_generator=gen, burn_in_steps=burn,
                               scale_grad=float(N), stepsize_schedule=ConstantStepsizeSchedule(0.01),
                               session=Session(device=dev, n_chains=C, output="torch"))
        sampler.RESIDENT_MAX_CHAINS = 0          # K4 (this variant) then K1, not the resident kernel
        zr = np.random.RandomState(z_seed)
        out = []
        for s in range(steps):
            z = zr.standard_normal((C, D)).astype(np.float32)
            sampler.__next__(feed_dict={sampler.noise: z})
            if (s + 1) % every == 0:
                out.append((s + 1, sampler._theta.cpu().numpy()))
        return out, sampler
    finally:
        _native.call("sgmcmc_set_bnn_tuning", DEFAULT_VARIANT)


def gpu_run_resident(theta0, X, y, seeds, steps, burn, z_seed, every, dev="cuda:0"):
    """The same chain through the resident kernel (csrc/bnn_resident.cu): one call per checkpoint interval,
    the chains on their SMs in between; minibatch indices from the device generator (K7), z injected."""
    import torch
    from pysgmcmc_b200 import _native
    from pysgmcmc_b200.data_batches import DeviceBatchGenerator
    C = theta0.shape[0]
    gen = DeviceBatchGenerator(N, BATCH, seeds=seeds, device=dev, block=256)
    Xd, yd = torch.tensor(X, dtype=torch.float32, device=dev), torch.tensor(y, dtype=torch.float32, device=dev)
    theta = torch.tensor(theta0, device=dev)
    state = [theta, torch.zeros_like(theta)] + [torch.ones_like(theta) for _ in range(4)]   # theta, V, tau, g, v_hat, minv
    cost = torch.empty(C, device=dev)
    zr = np.random.RandomState(z_seed)
    out, done = [], 0
    while done < steps:
        n = min(every, steps - done)
        starts = gen.next_block(n)
        z = torch.tensor(np.stack([zr.standard_normal((C, D)).astype(np.float32) for _ in range(n)]), device=dev)
        _native.call("sgmcmc_bnn_sghmc_run_resident_f32", *[_native.ptr(a) for a in state], _native.ptr(Xd),
                     _native.ptr(yd), _native.ptr(starts), _native.ptr(z), None, None, None, _native.ptr(cost), None,
                     C, 1, BATCH, float(BATCH), N, n, min(n, max(0, burn - done)), 0, 1, 0.01, 0.05, float(N),
                     0, done, 0, _native.stream_ptr())
        done += n
        out.append((done, theta.cpu().numpy()))
    return out


DEFAULT_VARIANT = 16
K4_NAMES = {0: "FFMA (variant 0)", 10: "tensor-pipe 3xTF32 (variant 10: truncating split, chained accumulation)",
            11: "tensor-pipe 3xTF32 (variant 11: rounded split)",
            12: "tensor-pipe 3xTF32 (variant 12: FP32-pipe accumulation across k-steps)",
            13: "tensor-pipe 3xTF32 (variant 13: rounded split + FP32-pipe accumulation)",
            14: "tensor-pipe 3xTF32 (variant 14: 13 with packed FP32 splitting of the weight fragments)",
            15: "tensor-pipe 3xTF32 (variant 15: 13 with the cross terms in their own accumulator)",
            16: "tensor-pipe 3xTF32 (variant 16: 15 with packed FP32 splitting of the weight fragments)",
            "resident": "resident kernel (csrc/bnn_resident.cu: FFMA accumulation order, chain on one SM)"}


def drift_curves(steps=1000, burn=600, chains=4, every=100, variants=(16, 0), z_seed=9, theta_seed=11):
    from oracle import bnn as obnn
    X, y = sinc_data()
    theta0 = obnn.init_theta(chains, seed=theta_seed, dtype=np.float32)
    seeds = np.arange(chains) + 40
    o32, _ = oracle_run(theta0, X, y, seeds, steps, burn, z_seed, np.float32, every)
    o64, _ = oracle_run(theta0, X, y, seeds, steps, burn, z_seed, np.float64, every)
    rel = lambda a, b: float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / np.abs(b).max())
    lines = []
    for variant in variants:
        got, _ = gpu_run(theta0, X.astype(np.float32), y.astype(np.float32), seeds, steps, burn, z_seed, every,
                         variant)
        vs32 = [rel(g[1], o[1]) for g, o in zip(got, o32)]
        vs64 = [rel(g[1], o[1]) for g, o in zip(got, o64)]
        o32_vs64 = [rel(a[1], b[1]) for a, b in zip(o32, o64)]
        cross = next((g[0] for g, d in zip(got, vs32) if d > 1e-5), None)
        lines.append({"k4": K4_NAMES.get(variant, "variant %s" % variant),
                      "config": {"N": N, "batch": BATCH, "scale_grad": N, "eps": 0.01, "burn_in_steps": burn,
                                 "steps": steps, "chains": chains},
                      "checkpoints": [g[0] for g in got],
                      "gpu_vs_oracle_f32": vs32, "gpu_vs_oracle_f64": vs64, "oracle_f32_vs_f64": o32_vs64,
                      "first_checkpoint_above_1e-5_vs_f32": cross})
    return lines


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--burn", type=int, default=600)
    ap.add_argument("--chains", type=int, default=4)
    ap.add_argument("--every", type=int, default=100)
    ap.add_argument("--variants", default="16,13,0")
    a = ap.parse_args()
    for line in drift_curves(a.steps, a.burn, a.chains, a.every, tuple(v if v == "resident" else int(v) for v in a.variants.split(","))):
        print(json.dumps(line), flush=True)
