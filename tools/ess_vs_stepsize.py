"""The reference's ESS-vs-stepsize experiment (docs/source/experiments/compute_ess.py:177-253)
re-run on the GPU for Relativistic SGHMC and compared with the numbers the reference
publishes (tests/golden/relativistic_ess_published.json, the only published data for this
path).

Protocol of the reference, per stepsize: ONE chain from theta0 = (0, 6) [banana] / 0 [gmm],
20 consecutive segments of 10 000 draws each, one draw kept every 10 steps (2e6 steps), the 20
segments treated as chains, ESS per variable by the pymc3 estimator, averaged over variables;
5 repeats.  Here the 5 repeats are 5 independent chains of one K6 launch and the estimator is
K8 + the host finalisation (diagnostics/sampler_diagnostics.py).

    python tools/ess_vs_stepsize.py [--targets banana,gmm2,gmm3] [--every 8] [--out file.json]

Sampling-quality numbers are stochastic and depend on details the reference does not pin (the
arspy momentum draw, TensorFlow's noise stream), so the comparison is of curve shape, order of
magnitude and location of the optimum, not of individual values.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pysgmcmc_b200 import Session  # noqa: E402
from pysgmcmc_b200.diagnostics.objective_functions import (banana_log_likelihood, gmm2_log_likelihood,  # noqa: E402
                                                           gmm3_log_likelihood, to_negative_log_likelihood)
from pysgmcmc_b200.diagnostics.sampler_diagnostics import effective_n_from_trace  # noqa: E402
from pysgmcmc_b200.samplers import RelativisticSGHMCSampler  # noqa: E402
from pysgmcmc_b200.stepsize_schedules import ConstantStepsizeSchedule  # noqa: E402

LOGLIK = {"banana": banana_log_likelihood, "gmm2": gmm2_log_likelihood, "gmm3": gmm3_log_likelihood}
N_SEGMENTS, DRAWS, KEEP_EVERY, REPEATS = 20, 10000, 10, 5


def mean_ess(target, stepsize, seed, dev):
    D = 2 if target == "banana" else 1
    params = [torch.zeros(REPEATS, device=dev) for _ in range(D)]
    if target == "banana":
        params[1] += 6.0
    s = RelativisticSGHMCSampler(params=params, cost_fun=to_negative_log_likelihood(LOGLIK[target]),
                                 stepsize_schedule=ConstantStepsizeSchedule(stepsize), seed=seed,
                                 session=Session(device=dev, n_chains=REPEATS, output="torch"))
    trace, _ = s.run(N_SEGMENTS * DRAWS * KEEP_EVERY, keep_every=KEEP_EVERY)      # [200000, R, D]
    out = []
    for r in range(REPEATS):
        x = trace[:, r, :]
        if not bool(torch.isfinite(x).all()):
            out.append(float("nan"))
            continue
        seg = x.reshape(N_SEGMENTS, DRAWS, D).permute(1, 0, 2).contiguous()       # [draws, segments, D]
        out.append(float(np.mean(effective_n_from_trace(seg))))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--targets", default="banana,gmm2,gmm3")
    ap.add_argument("--every", type=int, default=8, help="use every n-th published stepsize")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    published = json.load(open(os.path.join(ROOT, "tests", "golden", "relativistic_ess_published.json")))
    dev = torch.device("cuda:0")
    result = {}
    for target in args.targets.split(","):
        rows = published[target][::args.every]
        table = []
        for eps, ref_mean, ref_min, ref_max in rows:
            ours = mean_ess(target, eps, seed=int(eps * 1000) + 7, dev=dev)
            finite = [v for v in ours if np.isfinite(v)]
            row = {"stepsize": eps, "published_mean": ref_mean, "published_min": ref_min, "published_max": ref_max,
                   "ours_mean": float(np.mean(finite)) if finite else None, "ours": ours}
            table.append(row)
            print(json.dumps(dict(target=target, **row)), flush=True)
        result[target] = table
        good = [r for r in table if r["ours_mean"]]
        if good:
            lr = np.array([np.log10(r["ours_mean"] / r["published_mean"]) for r in good])
            best_ours = max(good, key=lambda r: r["ours_mean"])["stepsize"]
            best_pub = max(table, key=lambda r: r["published_mean"])["stepsize"]
            summary = {"target": target, "n_stepsizes": len(good), "median_log10_ratio": float(np.median(lr)),
                       "max_abs_log10_ratio": float(np.abs(lr).max()), "argmax_stepsize_ours": best_ours,
                       "argmax_stepsize_published": best_pub,
                       "spearman": float(__import__("scipy.stats").stats.spearmanr(
                           [r["ours_mean"] for r in good], [r["published_mean"] for r in good]).correlation)}
            result[target + "_summary"] = summary
            print(json.dumps(summary), flush=True)
    if args.out:
        json.dump(result, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
