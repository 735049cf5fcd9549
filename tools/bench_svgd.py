"""Per-kernel times of one SVGD step (K11-K14, csrc/svgd*.cu), FFMA and tcgen05 implementations,
and of the whole `next(sampler)` with the BNN cost kernel K4 supplying the gradients.
    python tools/bench_svgd.py [--quick | --k14-sweep [n D]]
Rooflines: K14 is 4 n^2 D flop (two products sharing the K operand): the FFMA kernel against the
FP32 peak of 74.4 TFLOP/s (148 SM x 128 lanes x 2 x 1.965 GHz), the tcgen05 kernel (3 TF32 products
per fp32 product) against the dense TF32 tensor peak; K11 is n^2 D / 2 MAC on the upper triangle;
K12/K13 are latency / L2 bound.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysgmcmc_b200 import _native  # noqa: E402

dev = torch.device("cuda:0")
FP32_PEAK = 148 * 128 * 2 * 1.965e9 / 1e12


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def k14_sweep():
    """K14 only: FFMA, tcgen05 with 2/3/4 producer register buffers.  `--k14-sweep [n D]`"""
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    shapes = [(int(args[0]), int(args[1]))] if len(args) >= 2 else [(1024, 5252), (4096, 5252), (8192, 1024)]
    for n, D in shapes:
        g = torch.Generator(device=dev).manual_seed(1)
        X = torch.randn((n, D), device=dev, generator=g)
        G = torch.randn((n, D), device=dev, generator=g)
        H = torch.zeros((n, D), device=dev)
        K = torch.rand((n, n), device=dev, generator=g)
        K = ((K + K.t()) * 0.5).contiguous()
        ksum = K.sum(1).contiguous()
        bw = torch.tensor([1.0, 1.0, 1.0, 0.0], device=dev)
        Xs = torch.empty_like(X)
        s = _native.stream_ptr()

        def update():
            _native.call("sgmcmc_svgd_update_f32", _native.ptr(X), _native.ptr(G), _native.ptr(H), _native.ptr(K),
                         _native.ptr(ksum), _native.ptr(bw), _native.ptr(Xs), n, D, 0.0, 0.9, 0.1, 1e-6, s)
        row = {"n_particles": n, "n_dims": D}
        for impl in (1, 2, 23, 24):
            _native.call("sgmcmc_set_svgd_tuning", impl)
            ms = timed(update, 5 if n * n * D > 5e10 else 20)
            _native.call("sgmcmc_set_svgd_tuning", 0)
            row["impl_%d_ms" % impl] = round(ms, 4)
            row["impl_%d_TFLOPs" % impl] = round(4.0 * n * n * D / ms / 1e9, 1)
        print(json.dumps(row), flush=True)


def main():
    if "--k14-sweep" in sys.argv:
        return k14_sweep()
    quick = "--quick" in sys.argv
    shapes = [(10, 2), (1024, 2), (128, 128), (256, 5252), (1024, 5252), (2048, 5252), (4096, 5252), (4096, 64),
              (8192, 1024)]
    if quick:
        shapes = [(10, 2), (1024, 5252)]
    for n, D in shapes:
        g = torch.Generator(device=dev).manual_seed(1)
        X = torch.randn((n, D), device=dev, generator=g)
        G = torch.randn((n, D), device=dev, generator=g)
        H = torch.zeros((n, D), device=dev)
        K = torch.empty((n, n), device=dev)
        ksum = torch.empty(n, device=dev)
        bw = torch.zeros(4, device=dev)
        scratch = _native.svgd_scratch(n, D, dev)
        Xs = torch.empty_like(X)
        X0 = X.clone()
        s = _native.stream_ptr()

        def kernel_matrix():
            _native.call("sgmcmc_svgd_kernel_matrix_f32", _native.ptr(X), _native.ptr(K), _native.ptr(ksum),
                         _native.ptr(bw), _native.ptr(scratch), scratch.numel() * 8, n, D, s)

        def median_only():
            _native.call("sgmcmc_median_f32", _native.ptr(K), n * n, _native.ptr(bw), _native.ptr(scratch), s)

        def update():
            # eps = 0 keeps the particles where they are, so every repetition does the same work
            _native.call("sgmcmc_svgd_update_f32", _native.ptr(X), _native.ptr(G), _native.ptr(H), _native.ptr(K),
                         _native.ptr(ksum), _native.ptr(bw), _native.ptr(Xs), n, D, 0.0, 0.9, 0.1, 1e-6, s)

        reps = 5 if n * n * D > 5e10 else 20
        _native.call("sgmcmc_set_svgd_tuning", 1)
        ms_km_ffma = timed(kernel_matrix, reps)
        _native.call("sgmcmc_set_svgd_tuning", 0)
        ms_km = timed(kernel_matrix, reps)
        ms_med = timed(median_only, reps)
        kernel_matrix()
        flop_up = 4.0 * n * n * D
        flop_sq = 1.5 * n * n * D
        row = {"n_particles": n, "n_dims": D,
               "k11_k12_k13_kernel_matrix_ms": round(ms_km, 4), "k11_k12_k13_ffma_ms": round(ms_km_ffma, 4),
               "k12_median_ms": round(ms_med, 4),
               "k11_sqdist_TFLOPs_equiv": round(flop_sq / max(ms_km - ms_med, 1e-6) / 1e9, 2)}
        best = None
        for impl, tag in ((1, "ffma"), (2, "tcgen05")):
            if impl == 2 and (n % 4 or D % 4):
                continue
            _native.call("sgmcmc_set_svgd_tuning", impl)
            ms_up = timed(update, reps)
            _native.call("sgmcmc_set_svgd_tuning", 0)
            assert torch.equal(X, X0)
            row["k14_%s_ms" % tag] = round(ms_up, 4)
            # algorithmic fp32 flop (the tcgen05 kernel executes 3 TF32 products per fp32 product)
            row["k14_%s_TFLOPs" % tag] = round(flop_up / ms_up / 1e9, 2)
            best = ms_up if best is None else min(best, ms_up)
        row["k14_ffma_frac_of_fp32_peak"] = round(row["k14_ffma_TFLOPs"] / FP32_PEAK, 3)
        if "k14_tcgen05_TFLOPs" in row:
            # dense TF32 tensor peak taken as half the measured bf16 peak (MEASURED_PEAKS.json)
            row["k14_tcgen05_executed_tf32_TFLOPs"] = round(3 * row["k14_tcgen05_TFLOPs"], 1)
        ms_auto = timed(update, reps)
        row["k14_auto_ms"] = round(ms_auto, 4)
        row["svgd_step_ms"] = round(ms_km + ms_auto, 4)
        row["particle_updates_per_s"] = round(n / (ms_km + ms_auto) * 1e3)
        print(json.dumps(row), flush=True)

    # whole next(sampler) on the BNN posterior: one particle = one 1-50-50-50-1 network
    if not quick:
        from pysgmcmc_b200 import Session
        from pysgmcmc_b200.models.bnn_cost import BayesianNeuralNetworkNLL, default_net_params
        from pysgmcmc_b200.samplers import SVGDSampler
        rng = np.random.RandomState(1)
        Xd = rng.rand(20, 1).astype(np.float32)
        yd = np.sinc(Xd * 10 - 5).sum(1).astype(np.float32)
        for n in (256, 1024, 2048):
            flat = torch.cat([p.reshape(n, -1) for p in default_net_params(1, n_chains=n, seed=1, device=dev)], dim=1)
            nll = BayesianNeuralNetworkNLL(20, batch_size=20, X=Xd, y=yd, device=dev)
            sampler = SVGDSampler([flat[i].clone() for i in range(n)], nll, session=Session(device=dev, output="torch"))
            ms = timed(sampler._step_on_device, 10)
            print(json.dumps({"workload": "BNN-SVGD next(sampler) on device (K4 + K11-K14)", "n_particles": n,
                              "n_dims": flat.shape[1], "ms_per_step": round(ms, 4),
                              "particle_steps_per_s": round(n / ms * 1e3)}), flush=True)


if __name__ == "__main__":
    main()
