"""Bandwidth sweep of the streaming update kernels K1-K3 (config 5 of BASELINE.json:
update-kernel bandwidth vs D).  Prints one JSON line per (kernel, n, tuning).

    python tools/bench_update.py [--sizes 1e5,1e6,...] [--sweep-tuning] [--iters 20]

Timing: CUDA events on the launching stream around `iters` back-to-back launches after 3
warm-up launches; the state (>= 5 arrays of n floats) is larger than the 126 MB L2 for
n >= 1e7, and for smaller n an L2 flush (256 MB memset) runs between launches (timed
separately and excluded).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysgmcmc_b200 import _native  # noqa: E402

BYTES = {"sghmc_burnin": 44, "sghmc_sampling": 24, "sgld_burnin": 36, "sgld_sampling": 16, "rsghmc": 20}


def peak_gbs():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        return json.load(open(path))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def launcher(kind, n, dev):
    st = _native.stream_ptr()
    mk = lambda v: torch.full((n,), v, device=dev, dtype=torch.float32)
    theta, v, tau, g, v_hat, minv = mk(0.1), mk(0.0), mk(1.0), mk(1.0), mk(1.0), mk(1.0)
    grad = torch.randn(n, device=dev)
    p = _native.ptr
    step = [0]

    def sghmc(burn):
        def f():
            _native.call("sgmcmc_sghmc_step_f32", p(theta), p(v), p(tau), p(g), p(v_hat), p(minv), p(grad),
                         None, n, 0.01, 0.05, 20000.0, burn, 0, 1, step[0], 0, st)
            step[0] += 1
        return f

    def sgld(burn):
        def f():
            _native.call("sgmcmc_sgld_step_f32", p(theta), p(tau), p(g), p(v_hat), p(minv), p(grad), None, n,
                         0.01, 1.0, 20000.0, burn, 0, 1, step[0], 0, st)
            step[0] += 1
        return f

    def rsghmc():
        _native.call("sgmcmc_rsghmc_step_f32", p(theta), p(v), p(grad), None, n, 0.001, 1.0, 1.0, 1.0, 0.0,
                     1, step[0], 0, st)
        step[0] += 1

    keep = (theta, v, tau, g, v_hat, minv, grad)
    return {"sghmc_burnin": sghmc(1), "sghmc_sampling": sghmc(0), "sgld_burnin": sgld(1),
            "sgld_sampling": sgld(0), "rsghmc": rsghmc}[kind], keep


def time_kernel(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    total = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    return total / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1e5,1e6,1e7,43024384,1e8,1e9")
    ap.add_argument("--kinds", default=",".join(BYTES))
    ap.add_argument("--sweep-tuning", action="store_true")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    peak, which = peak_gbs()
    tunings = [(t, u) for t in (128, 256, 512) for u in (1, 2)] if args.sweep_tuning else [(0, 0)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for size in args.sizes.split(","):
        n = int(float(size))
        for kind in args.kinds.split(","):
            try:
                fn, keep = launcher(kind, n, dev)
            except torch.OutOfMemoryError:
                print(json.dumps({"kernel": kind, "n": n, "error": "oom"}))
                continue
            for threads, unroll in tunings:
                _native.call("sgmcmc_set_update_tuning", threads, unroll)
                small = n * BYTES[kind] < (256 << 20)
                ms = time_kernel(fn, args.iters, flush if small else None)
                gbs = BYTES[kind] * n / ms / 1e6
                print(json.dumps({"kernel": kind, "n": n, "threads": threads, "unroll": unroll,
                                  "ms": round(ms, 5), "GBps": round(gbs, 1),
                                  "frac_of_%s_peak" % which: round(gbs / peak, 4),
                                  "frac_of_8TBps": round(gbs / 8000, 4),
                                  "Gelem_steps_per_s": round(n / ms / 1e6, 2)}), flush=True)
            del fn, keep
            torch.cuda.empty_cache()
    _native.call("sgmcmc_set_update_tuning", 256, 1)


if __name__ == "__main__":
    main()
