"""Experiment: can the HBM-bound update (K1) of one half of the chains overlap the
FP32-bound BNN gradient (K4) of the other half when they run on two streams?
    python tools/bench_overlap.py [--variant 10]
Prints sequential vs concurrent time for one full step of 8192 chains.
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
ap.add_argument("--variants", default="0")
ap.add_argument("--k1-threads", default="256,128")
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
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def k4(lo, hi, stream):
    _native.call("sgmcmc_bnn_nll_grad_f32", p(theta[lo:hi]), p(X), p(y), p(starts[lo:hi]), p(cost[lo:hi]),
                 p(grad[lo:hi]), None, hi - lo, 1, B, float(B), N, stream.cuda_stream)


def k1(lo, hi, stream):
    _native.call("sgmcmc_sghmc_step_f32", p(theta[lo:hi]), *[p(a[lo:hi]) for a in state], p(grad[lo:hi]), None,
                 (hi - lo) * D, 0.01, 0.05, float(N), 1, 0, 1, 0, lo * D, stream.cuda_stream)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        fn()
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.iters


cur = torch.cuda.current_stream()
for v in [int(x) for x in args.variants.split(",")]:
    for kt in [int(x) for x in args.k1_threads.split(",")]:
        _native.call("sgmcmc_set_bnn_tuning", v)
        _native.call("sgmcmc_set_update_tuning", kt, 1)
        seq = timed(lambda: (k4(0, C, cur), k1(0, C, cur)))
        only4 = timed(lambda: k4(0, C, cur))
        only1 = timed(lambda: k1(0, C, cur))

        def pipelined():
            # one step of both halves, half B one phase behind half A
            k4(0, H, s1); k1(H, C, s2)
            torch.cuda.current_stream().wait_stream(s1)
            ev1, ev2 = torch.cuda.Event(), torch.cuda.Event()
            ev1.record(s1); ev2.record(s2)
            s1.wait_event(ev2); s2.wait_event(ev1)
            k1(0, H, s1); k4(H, C, s2)
            ev3, ev4 = torch.cuda.Event(), torch.cuda.Event()
            ev3.record(s1); ev4.record(s2)
            s1.wait_event(ev4); s2.wait_event(ev3)
        conc = timed(pipelined)
        print(json.dumps({"k4_variant": v, "k1_threads": kt, "k4_ms": round(only4, 4), "k1_ms": round(only1, 4),
                          "sequential_step_ms": round(seq, 4), "two_stream_step_ms": round(conc, 4)}), flush=True)
_native.call("sgmcmc_set_bnn_tuning", 0)
_native.call("sgmcmc_set_update_tuning", 256, 1)
