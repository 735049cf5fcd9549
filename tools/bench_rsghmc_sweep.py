"""BASELINE.json configs[4]: bandwidth sweep of the relativistic SGHMC update kernel (K3,
pysgmcmc/samplers/relativistic_sghmc.py:120-140 generalised element-wise) over
D = 1e5 ... 1e9 parameters, including the wide 1000-512-512 BNN's parameter count, on 1/2/4/8 GPUs.

    python tools/bench_rsghmc_sweep.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/bench_rsghmc_sweep.py                  # N GPUs

The update is element-wise, so it shards with no collective.  Two ways, both reported:
  weak   : every GPU advances its own chain of D parameters (aggregate = N x D elements per step);
  strong : ONE chain of D parameters split into N contiguous shards (`elem_offset` shifts the
           Philox counter so the shards reproduce the un-sharded noise stream).
Per point: CUDA events on the launching stream around `iters` launches after 3 warm-ups, barrier
before and after, MAX over ranks; state smaller than 256 MB gets an L2 flush (a 256 MB memset,
outside the timed events) before every launch, larger state is its own flush.  One JSON line per
point (rank 0): GB/s = 20 B x elements / time, against MEASURED_PEAKS.json and against 8 TB/s.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pysgmcmc_b200 import _native  # noqa: E402

BYTES_PER_ELEM = 20          # theta, grad, p in; theta, p out
WIDE_BNN_D = 1 * 1000 + 1000 + 1000 * 512 + 512 + 512 * 512 + 512 + 512 + 1 + 1   # 1-1000-512-512-1 + rho


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1e5,%d,1e6,1e7,1e8,1e9" % WIDE_BNN_D)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peak = peak_gbs()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st, p = _native.stream_ptr(), _native.ptr

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for size in args.sizes.split(","):
        D = int(float(size))
        for mode in ("weak", "strong"):
            if mode == "strong" and world == 1:
                continue
            # strong: shard boundaries on multiples of 4 elements (one Philox counter per group)
            per = D if mode == "weak" else (((D + world - 1) // world + 3) & ~3)
            lo = 0 if mode == "weak" else min(D, rank * per)
            n = D if mode == "weak" else max(0, min(D, lo + per) - lo)
            total = D * world if mode == "weak" else D
            g = torch.Generator(device=dev).manual_seed(1 + rank)
            theta = torch.randn(max(n, 1), device=dev, generator=g)
            mom = torch.randn(max(n, 1), device=dev, generator=g)
            grad = torch.randn(max(n, 1), device=dev, generator=g)
            step = [0]

            def launch():
                if n > 0:
                    _native.call("sgmcmc_rsghmc_step_f32", p(theta), p(mom), p(grad), None, n, 0.001, 1.0, 1.0,
                                 1.0, 0.0, 1, step[0], lo, st)
                step[0] += 1
            for _ in range(3):
                launch()
            small = n * BYTES_PER_ELEM < (256 << 20)
            barrier()
            if small:
                ms = 0.0
                for _ in range(args.iters):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    launch()
                    e1.record()
                    torch.cuda.synchronize()
                    ms += e0.elapsed_time(e1)
                ms /= args.iters
            else:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.iters):
                    launch()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.iters
            barrier()
            ms = max_over_ranks(ms)
            assert bool(torch.isfinite(theta).all())
            if rank == 0:
                gbs = BYTES_PER_ELEM * total / ms / 1e6
                print(json.dumps({
                    "kernel": "rsghmc_update_kernel (K3)", "scaling": mode, "n_gpus": world, "D": D,
                    "elements_total": total, "elements_per_gpu": per if mode == "strong" else D,
                    "wide_bnn_1000_512_512": D == WIDE_BNN_D, "ms": round(ms, 5), "GBps_aggregate": round(gbs, 1),
                    "GBps_per_gpu": round(gbs / world, 1),
                    "frac_of_measured_peak_per_gpu": round(gbs / world / peak, 4),
                    "frac_of_8TBps_per_gpu": round(gbs / world / 8000.0, 4),
                    "l2": "flushed before every launch" if small else "state larger than L2"}), flush=True)
            del theta, mom, grad
            torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
