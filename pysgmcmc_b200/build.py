"""Builds pysgmcmc_b200/libsgmcmc_b200.so in-tree with nvcc for sm_100a.

    python -m pysgmcmc_b200.build [--force] [--verbose]

The library is plain C ABI (include/sgmcmc_b200.h); it links the CUDA runtime
statically so that it loads with ctypes on a box without a GPU (symbol checks)
and next to torch's own runtime on the B200 box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libsgmcmc_b200.so")
OBJ_DIR = os.path.join(HERE, "csrc", "_obj")

SOURCES = ["capi.cu", "update_kernels.cu", "target_chains.cu", "mt19937.cu", "bnn.cu", "mlp.cu", "mlp_umma.cu", "bnn_fused.cu", "bnn_resident.cu", "host_pipeline.cu", "moments.cu", "svgd.cu", "svgd_umma.cu", "svgd_sqdist_umma.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "sgmcmc_b200.h"))
    return hdrs


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    deps = _deps()
    objs, procs = [], []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [path] + deps):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", path, "-o", obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                                text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write("---- nvcc %s ----\n%s\n" % (src, out))
        log = os.path.join(OBJ_DIR, src.replace(".cu", ".ptxas.log"))
        with open(log, "w") as f:
            f.write(out)
        failed = failed or p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed (see output above)")
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
               "-o", LIB] + objs
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
