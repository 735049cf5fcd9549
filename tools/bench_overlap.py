"""Experiment: overlap the HBM-bound update (K1) of one half of the chains with the
FP32-bound BNN gradient (K4) of the other half, both as persistent (capped-grid) kernels on
two streams so that they are co-resident on every SM.
    python tools/bench_overlap.py
Prints, per configuration, the time of one full step (both halves) sequentially and pipelined.

Round-1 result (NEGATIVE, kept as a record): without further measures the two kernels never
overlap, because K4 asks for a large shared-memory carveout and K1 for none, and an SM cannot
host two carveouts at once.  Forcing both to the maximum-shared carveout does make them
overlap (pipelined 0.675 ms vs 0.99 ms for the same capped kernels back to back), but the
smaller L1 slows K4 by 28 % (0.295 -> 0.377 ms) and the capped grids slow both kernels, so
the best pipelined step (0.675 ms) loses to the plain sequential default (0.58-0.61 ms).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysgmcmc_b200 import _native  # noqa: E402
from pysgmcmc_b200.models.bnn_cost import default_net_params  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=8192)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
dev = torch.device("cuda:0")
C, N, B, D = args.chains, 20000, 20, 5252
H = C // 2
g = torch.Generator(device="cpu").manual_seed(0)
X = torch.randn(N, 1, generator=g).to(dev)
y = torch.randn(N, generator=g).to(dev)
theta = torch.cat([p.reshape(C, -1) for p in default_net_params(1, n_chains=C, seed=1, device=dev)], dim=1).contiguous()
starts = torch.randint(0, N - B + 1, (C,), device=dev, dtype=torch.int32)
cost, grad = torch.empty(C, device=dev), torch.zeros_like(theta)
state = [torch.zeros_like(theta)] + [torch.ones_like(theta) for _ in range(4)]   # v, tau, g, v_hat, minv
p = _native.ptr
s_compute, s_update = torch.cuda.Stream(), torch.cuda.Stream()


def k4(lo, hi, stream):
    _native.call("sgmcmc_bnn_nll_grad_f32", p(theta[lo:hi]), p(X), p(y), p(starts[lo:hi]), p(cost[lo:hi]),
                 p(grad[lo:hi]), None, hi - lo, 1, B, float(B), N, stream.cuda_stream)


def k1(lo, hi, stream):
    _native.call("sgmcmc_sghmc_step_f32", p(theta[lo:hi]), *[p(a[lo:hi]) for a in state], p(grad[lo:hi]), None,
                 (hi - lo) * D, 0.01, 0.05, float(N), 1, 0, 1, 0, lo * D, stream.cuda_stream)


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def pipelined(n_steps):
    """All K4s in order on one stream, all K1s in order on the other; K1_X(s) waits for
    K4_X(s), K4_X(s+1) waits for K1_X(s).  Steady state: K4_A || K1_B, then K4_B || K1_A."""
    halves = ((0, H), (H, C))
    done_k1 = [None, None]
    for s in range(n_steps):
        for h, (lo, hi) in enumerate(halves):
            if done_k1[h] is not None:
                s_compute.wait_event(done_k1[h])
            k4(lo, hi, s_compute)
            ev = torch.cuda.Event()
            ev.record(s_compute)
            s_update.wait_event(ev)
            k1(lo, hi, s_update)
            done_k1[h] = torch.cuda.Event()
            done_k1[h].record(s_update)
    torch.cuda.current_stream().wait_stream(s_compute)
    torch.cuda.current_stream().wait_stream(s_update)


cur = torch.cuda.current_stream()
STEPS = 10
for variant, bnn_per_sm, upd_per_sm, upd_threads in [
        (0, 0, 0, 256), (0, 2, 2, 256), (0, 1, 2, 256), (0, 1, 4, 256),
        (1, 2, 1, 256), (1, 2, 2, 128), (1, 3, 1, 128), (9, 3, 1, 256), (9, 2, 2, 256), (9, 3, 2, 128),
        (4, 2, 1, 256), (7, 2, 1, 256)]:
    _native.call("sgmcmc_set_bnn_tuning", variant)
    _native.call("sgmcmc_set_update_tuning", upd_threads, 1)
    _native.call("sgmcmc_set_persistent_grids", 0, 0)
    seq = timed(lambda: (k4(0, C, cur), k1(0, C, cur)), args.iters)
    _native.call("sgmcmc_set_persistent_grids", 148 * upd_per_sm, 148 * bnn_per_sm)
    only4 = timed(lambda: k4(0, C, cur), args.iters)
    only1 = timed(lambda: k1(0, C, cur), args.iters)
    pipe = timed(lambda: pipelined(STEPS), 3) / STEPS
    print(json.dumps({"k4_variant": variant, "k4_ctas_per_sm": bnn_per_sm, "k1_ctas_per_sm": upd_per_sm,
                      "k1_threads": upd_threads, "sequential_full_grid_ms": round(seq, 4),
                      "k4_capped_ms": round(only4, 4), "k1_capped_ms": round(only1, 4),
                      "pipelined_step_ms": round(pipe, 4)}), flush=True)
_native.call("sgmcmc_set_bnn_tuning", 0)
_native.call("sgmcmc_set_update_tuning", 256, 1)
_native.call("sgmcmc_set_persistent_grids", 0, 0)
