"""GPU parity tests (through the C ABI) of the streaming update kernels K1-K3, the noise
stream and the minibatch index kernel K7 against the CPU oracle.

Tolerances (north_star part 1: 1e-5 relative in FP32 with injected noise):
  * single step, injected grad + noise: BIT-EXACT (same IEEE ops in the same order);
  * 1000-step trajectories with injected noise and synthetic gradients: BIT-EXACT;
  * in-kernel Philox noise: uniforms are bit-exact by construction, normals go through the
    SFU approximations (lg2/sqrt/sin/cos.approx) -> atol 2e-5 / rtol 1e-5 (worst seen 8e-6).
"""
import numpy as np
import pytest
import torch

from oracle import mt19937, philox, samplers
from pysgmcmc_b200 import _native

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
NP = {torch.float32: np.float32, torch.float64: np.float64}
SUF = {torch.float32: "f32", torch.float64: "f64"}


def dev(a, dtype=None):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(DEV)


def random_state(rng, n, dtype):
    """A mid-burn-in looking state (positive v_hat, tau >= 1) plus edge values."""
    st = dict(theta=rng.standard_normal(n), v=0.01 * rng.standard_normal(n),
              tau=1.0 + 5.0 * rng.uniform(size=n), g=rng.standard_normal(n),
              v_hat=rng.uniform(0.05, 4.0, size=n), minv=rng.uniform(0.2, 3.0, size=n))
    k = min(3, n)
    st["v_hat"][:k] = (0.0, 1e-30, 1.0)[:k]   # safe_divide / safe_sqrt edge cases
    st["g"][:k] = (0.0, -2.0, 1.0)[:k]
    return {k: v.astype(dtype) for k, v in st.items()}


def stream():
    return _native.stream_ptr()


def call_sghmc(t, grad, z, eps, mdecay, scale_grad, burn_in, store_minv, dtype, seed=0, step=0, off=0):
    _native.call("sgmcmc_sghmc_step_" + SUF[dtype], *[_native.ptr(t[k]) for k in
                 ("theta", "v", "tau", "g", "v_hat", "minv")], _native.ptr(grad), _native.ptr(z),
                 t["theta"].numel(), eps, mdecay, scale_grad, int(burn_in), int(store_minv),
                 seed, step, off, stream())


def call_sgld(t, grad, z, eps, A, scale_grad, burn_in, store_minv, dtype, seed=0, step=0, off=0):
    _native.call("sgmcmc_sgld_step_" + SUF[dtype], *[_native.ptr(t[k]) for k in
                 ("theta", "tau", "g", "v_hat", "minv")], _native.ptr(grad), _native.ptr(z),
                 t["theta"].numel(), eps, A, scale_grad, int(burn_in), int(store_minv),
                 seed, step, off, stream())


def call_rsghmc(t, grad, z, eps, m, c, D, Bhat, dtype, seed=0, step=0, off=0):
    _native.call("sgmcmc_rsghmc_step_" + SUF[dtype], _native.ptr(t["theta"]), _native.ptr(t["p"]),
                 _native.ptr(grad), _native.ptr(z), t["theta"].numel(), eps, m, c, D, Bhat,
                 seed, step, off, stream())


def assert_same(got, want, name, exact=True):
    got = got.cpu().numpy()
    if exact:
        bad = ~((got == want) | (np.isnan(got) & np.isnan(want)))
        assert not bad.any(), "%s: %d / %d elements differ, max rel %.3g" % (
            name, bad.sum(), bad.size, np.max(np.abs(got - want)[bad] / (np.abs(want[bad]) + 1e-30)))
    else:
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-6, err_msg=name)


# ------------------------------------------------------------------------------------
# single step, every mode, odd sizes (tail group), both dtypes
# ------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("n", [1, 3, 4, 5, 1023, 4097, 100003])
@pytest.mark.parametrize("burn_in", [True, False])
def test_sghmc_single_step_bit_exact(dtype, n, burn_in):
    rng = np.random.RandomState(n)
    npdt = NP[dtype]
    st = random_state(rng, n, npdt)
    grad = (3 * rng.standard_normal(n)).astype(npdt)
    z = rng.standard_normal(n).astype(npdt)
    eps, mdecay, sg = 0.01, 0.05, 20000.0
    want = samplers.sghmc_step(st, grad, z, eps, mdecay, sg, burn_in=burn_in, frozen_minv=st["minv"])
    t = {k: dev(v) for k, v in st.items()}
    call_sghmc(t, dev(grad), dev(z), eps, mdecay, sg, burn_in, True, dtype)
    for k in ("theta", "v", "tau", "g", "v_hat", "minv"):
        assert_same(t[k], want[k], "sghmc %s" % k)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("n", [1, 6, 4097, 100003])
@pytest.mark.parametrize("burn_in", [True, False])
def test_sgld_single_step_bit_exact(dtype, n, burn_in):
    rng = np.random.RandomState(100 + n)
    npdt = NP[dtype]
    st = random_state(rng, n, npdt)
    st.pop("v")
    grad = (3 * rng.standard_normal(n)).astype(npdt)
    z = rng.standard_normal(n).astype(npdt)
    eps, A, sg = 0.01, 1.5, 3.0
    want = samplers.sgld_step(st, grad, z, eps, A, sg, burn_in=burn_in, frozen_minv=st["minv"])
    t = {k: dev(v) for k, v in st.items()}
    call_sgld(t, dev(grad), dev(z), eps, A, sg, burn_in, True, dtype)
    for k in ("theta", "tau", "g", "v_hat", "minv"):
        assert_same(t[k], want[k], "sgld %s" % k)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("n", [1, 7, 4097, 100003])
def test_rsghmc_single_step_bit_exact(dtype, n):
    rng = np.random.RandomState(200 + n)
    npdt = NP[dtype]
    st = dict(theta=rng.standard_normal(n).astype(npdt), p=(2 * rng.standard_normal(n)).astype(npdt))
    grad = (3 * rng.standard_normal(n)).astype(npdt)
    z = rng.standard_normal(n).astype(npdt)
    hyper = dict(mass=1.5, speed_of_light=0.8, D=1.2, Bhat=0.1)
    want = samplers.rsghmc_step(st, grad, z, 0.05, **hyper)
    t = {k: dev(v) for k, v in st.items()}
    call_rsghmc(t, dev(grad), dev(z), 0.05, 1.5, 0.8, 1.2, 0.1, dtype)
    for k in ("theta", "p"):
        assert_same(t[k], want[k], "rsghmc %s" % k)


def test_misaligned_pointers_take_the_scalar_path():
    """Views that start 4 bytes into an allocation are not 16-byte aligned."""
    n = 1001
    rng = np.random.RandomState(9)
    st = random_state(rng, n, np.float32)
    grad = rng.standard_normal(n).astype(np.float32)
    z = rng.standard_normal(n).astype(np.float32)
    want = samplers.sghmc_step(st, grad, z, 0.01, 0.05, 1.0, burn_in=True)
    t = {}
    for k, v in st.items():
        buf = torch.zeros(n + 1, dtype=torch.float32, device=DEV)
        buf[1:] = dev(v)
        t[k] = buf[1:]
        assert t[k].data_ptr() % 16 != 0
    call_sghmc(t, dev(grad), dev(z), 0.01, 0.05, 1.0, True, True, torch.float32)
    for k in ("theta", "v", "tau", "g", "v_hat", "minv"):
        assert_same(t[k], want[k], "misaligned %s" % k)


def test_store_minv_flag_and_untouched_arrays():
    n = 4096
    rng = np.random.RandomState(3)
    st = random_state(rng, n, np.float32)
    grad = rng.standard_normal(n).astype(np.float32)
    z = rng.standard_normal(n).astype(np.float32)
    t = {k: dev(v) for k, v in st.items()}
    call_sghmc(t, dev(grad), dev(z), 0.01, 0.05, 1.0, True, False, torch.float32)
    assert_same(t["minv"], st["minv"], "minv must not be written when store_minv == 0")
    t = {k: dev(v) for k, v in st.items()}
    call_sghmc(t, dev(grad), dev(z), 0.01, 0.05, 1.0, False, False, torch.float32)
    for k in ("tau", "g", "v_hat", "minv"):
        assert_same(t[k], st[k], "%s must not be written after burn-in" % k)


def test_invalid_arguments_raise():
    t = torch.zeros(8, device=DEV)
    with pytest.raises(_native.NativeError, match="NULL"):
        _native.call("sgmcmc_sghmc_step_f32", _native.ptr(t), None, None, None, None, None, None, None,
                     8, 0.01, 0.05, 1.0, 1, 0, 0, 0, 0, stream())
    with pytest.raises(_native.NativeError, match="multiple of 4"):
        _native.call("sgmcmc_normal_fill_f32", _native.ptr(t), 8, 0, 0, 6, stream())
    # n == 0 is a no-op, not an error
    _native.call("sgmcmc_normal_fill_f32", _native.ptr(t), 0, 0, 0, 0, stream())


# ------------------------------------------------------------------------------------
# 1000-step trajectories, injected noise, gradient of a synthetic quadratic computed
# identically on both sides (numpy fp32 vs torch fp32 element-wise mul) -> bit-exact
# ------------------------------------------------------------------------------------
@pytest.mark.parametrize("method", ["sghmc", "sgld", "rsghmc"])
def test_thousand_step_trajectory_bit_exact(method):
    C, D, steps, burn = 8, 37, 1000, 300
    n = C * D
    rng = np.random.RandomState(42)
    curv = rng.uniform(0.5, 3.0, size=n).astype(np.float32)     # grad = curv * theta
    theta0 = rng.standard_normal(n).astype(np.float32)
    if method == "sghmc":
        st = samplers.sghmc_init(theta0)
    elif method == "sgld":
        st = samplers.sgld_init(theta0)
    else:
        st = samplers.rsghmc_init(theta0, rng.standard_normal(n).astype(np.float32))
    t = {k: dev(v) for k, v in st.items()}
    curv_d = dev(curv)
    frozen = None
    for s in range(steps):
        z = rng.standard_normal(n).astype(np.float32)
        grad = curv * st["theta"]
        grad_d = curv_d * t["theta"]
        burn_in = s < burn
        if method == "sghmc":
            st = samplers.sghmc_step(st, grad, z, 0.01, 0.05, 4.0, burn_in=burn_in, frozen_minv=frozen)
            call_sghmc(t, grad_d, dev(z), 0.01, 0.05, 4.0, burn_in, s == burn - 1, torch.float32)
        elif method == "sgld":
            st = samplers.sgld_step(st, grad, z, 0.01, 1.0, 4.0, burn_in=burn_in, frozen_minv=frozen)
            call_sgld(t, grad_d, dev(z), 0.01, 1.0, 4.0, burn_in, s == burn - 1, torch.float32)
        else:
            st = samplers.rsghmc_step(st, grad, z, 0.01)
            call_rsghmc(t, grad_d, dev(z), 0.01, 1.0, 1.0, 1.0, 0.0, torch.float32)
        if burn_in and method != "rsghmc":
            frozen = st["minv"]
    for k in t:
        if k == "minv":
            continue
        assert_same(t[k], st[k], "%s %s after %d steps" % (method, k, steps))
    if method != "rsghmc":
        assert_same(t["minv"], frozen, "frozen minv")


# ------------------------------------------------------------------------------------
# the engine's noise stream
# ------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,off", [(1, 0), (5, 0), (4096, 0), (100003, 8), (1000, 2 ** 34)])
def test_normal_fill_matches_oracle_philox(n, off):
    out = torch.empty(n, device=DEV)
    _native.call("sgmcmc_normal_fill_f32", _native.ptr(out), n, 0x1234567 + (5 << 32), 77, off, stream())
    want = philox.normals(n, seed=0x1234567 + (5 << 32), step=77, elem_offset=off)
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=1e-5, atol=2e-5)


def test_normal_fill_statistics_and_determinism():
    from scipy import stats
    n = 1 << 22
    a, b = torch.empty(n, device=DEV), torch.empty(n, device=DEV)
    _native.call("sgmcmc_normal_fill_f32", _native.ptr(a), n, 1, 0, 0, stream())
    _native.call("sgmcmc_normal_fill_f32", _native.ptr(b), n, 1, 0, 0, stream())
    assert torch.equal(a, b)
    _native.call("sgmcmc_normal_fill_f32", _native.ptr(b), n, 1, 1, 0, stream())
    assert not torch.equal(a, b)
    x = a.cpu().numpy().astype(np.float64)
    assert abs(x.mean()) < 3e-3 and abs(x.std() - 1) < 3e-3
    assert stats.kstest(x[:200000], "norm").pvalue > 1e-3
    assert abs(np.corrcoef(x[:-1], x[1:])[0, 1]) < 3e-3


def test_in_kernel_noise_equals_filled_noise_and_shards():
    """z == NULL uses Philox(elem_offset + e, step): (1) same as feeding normal_fill's
    output, (2) a shard with elem_offset reproduces the slice of the full run."""
    n, seed, step = 40000, 99, 5
    rng = np.random.RandomState(1)
    st = random_state(rng, n, np.float32)
    grad = rng.standard_normal(n).astype(np.float32)
    z = torch.empty(n, device=DEV)
    _native.call("sgmcmc_normal_fill_f32", _native.ptr(z), n, seed, step, 0, stream())
    a = {k: dev(v) for k, v in st.items()}
    b = {k: dev(v) for k, v in st.items()}
    call_sghmc(a, dev(grad), z, 0.01, 0.05, 1.0, True, True, torch.float32)
    call_sghmc(b, dev(grad), None, 0.01, 0.05, 1.0, True, True, torch.float32, seed=seed, step=step)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    lo = 12000
    c = {k: dev(v[lo:]) for k, v in st.items()}
    call_sghmc(c, dev(grad[lo:]), None, 0.01, 0.05, 1.0, True, True, torch.float32, seed=seed,
               step=step, off=lo)
    for k in a:
        assert torch.equal(a[k][lo:], c[k]), k


# ------------------------------------------------------------------------------------
# launch tuning must not change results
# ------------------------------------------------------------------------------------
def test_tuning_knobs_do_not_change_results():
    n = 300007
    rng = np.random.RandomState(5)
    st = random_state(rng, n, np.float32)
    grad = rng.standard_normal(n).astype(np.float32)
    ref = None
    try:
        for threads, unroll, reverse, cap in [(th, un, rv, 0) for th in (128, 256, 512) for un in (1, 2)
                                              for rv in (1, 0)] + [(256, 1, 1, 37), (256, 1, 0, 37)]:
            if True:
                _native.call("sgmcmc_set_update_tuning", threads, unroll)
                _native.call("sgmcmc_set_update_reverse", reverse)     # walk order of K1 (L2 reuse)
                _native.call("sgmcmc_set_persistent_grids", cap, 0)    # capped grid looping over the array
                t = {k: dev(v) for k, v in st.items()}
                call_sghmc(t, dev(grad), None, 0.01, 0.05, 1.0, True, True, torch.float32, seed=3, step=1)
                torch.cuda.synchronize()
                if ref is None:
                    ref = t
                else:
                    for k in t:
                        assert torch.equal(t[k], ref[k]), (threads, unroll, reverse, cap, k)
    finally:
        _native.call("sgmcmc_set_update_tuning", 256, 1)
        _native.call("sgmcmc_set_update_reverse", 1)
        _native.call("sgmcmc_set_persistent_grids", 0, 0)


# ------------------------------------------------------------------------------------
# full-size properties (config 4 shard: 8192 chains x 5252 params = 43 M elements)
# ------------------------------------------------------------------------------------
def test_full_size_properties():
    C, D = 8192, 5252
    n = C * D
    g = torch.Generator(device=DEV).manual_seed(0)
    theta = torch.randn(n, device=DEV, generator=g)
    t = dict(theta=theta.clone(), v=torch.zeros(n, device=DEV), tau=torch.ones(n, device=DEV),
             g=torch.ones(n, device=DEV), v_hat=torch.ones(n, device=DEV), minv=torch.ones(n, device=DEV))
    grad = torch.randn(n, device=DEV, generator=g)
    # (1) from the initial state r = 1/2, minv = 1: closed form of the first step
    call_sghmc(t, grad, None, 0.01, 0.05, 20000.0, True, True, torch.float32, seed=7, step=0)
    z = torch.empty(n, device=DEV)
    _native.call("sgmcmc_normal_fill_f32", _native.ptr(z), n, 7, 0, 0, stream())
    eps_s = np.float32(0.01) / np.sqrt(np.float32(20000.0))
    sigma = float(np.sqrt(np.float32(2) * eps_s ** 2 * np.float32(0.05) - eps_s ** 4))
    v_want = -1e-4 * grad + sigma * z
    assert torch.allclose(t["v"], v_want, rtol=1e-5, atol=1e-9)
    assert torch.equal(t["theta"], theta + t["v"])
    assert torch.allclose(t["g"], 0.5 + 0.5 * grad, rtol=1e-6, atol=1e-7)
    assert torch.allclose(t["v_hat"], 0.5 + 0.5 * grad * grad, rtol=1e-6, atol=1e-7)
    assert torch.equal(t["minv"], torch.ones_like(theta))
    # (2) sharding invariance at scale: second half with elem_offset == slice of the full run
    h = n // 2
    s = dict(theta=theta[h:].clone(), v=torch.zeros(n - h, device=DEV), tau=torch.ones(n - h, device=DEV),
             g=torch.ones(n - h, device=DEV), v_hat=torch.ones(n - h, device=DEV),
             minv=torch.ones(n - h, device=DEV))
    call_sghmc(s, grad[h:].contiguous(), None, 0.01, 0.05, 20000.0, True, True, torch.float32,
               seed=7, step=0, off=h)
    for k in s:
        assert torch.equal(s[k], t[k][h:]), k


# ------------------------------------------------------------------------------------
# K7: minibatch start indices, bit-exact vs numpy RandomState
# ------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,B", [(20000, 20), (100, 10), (20, 20), (5, 20), (70000, 3)])
def test_minibatch_starts_bit_exact(N, B):
    seeds = np.array([1, 2, 12345, 2 ** 32 - 1, 0, 77], dtype=np.uint64)
    C, steps = len(seeds), 1500
    state = torch.empty((625, C), dtype=torch.int32, device=DEV)
    seeds_d = torch.as_tensor(seeds.astype(np.int64)).to(torch.int32).to(DEV)
    _native.call("sgmcmc_mt19937_seed", _native.ptr(state), _native.ptr(seeds_d), C, stream())
    out = torch.empty((steps, C), dtype=torch.int32, device=DEV)
    Beff = min(B, N)
    # two consecutive calls must continue the streams (crosses several twists)
    _native.call("sgmcmc_mt19937_starts", _native.ptr(state), _native.ptr(out), C, 700, N - Beff, stream())
    _native.call("sgmcmc_mt19937_starts", _native.ptr(state), _native.ptr(out[700:]), C, steps - 700,
                 N - Beff, stream())
    got = out.cpu().numpy()
    for j, seed in enumerate(seeds):
        rng = np.random.RandomState()
        rng.seed(int(seed))
        want = np.array([rng.randint(0, N - Beff + 1) for _ in range(steps)])
        assert np.array_equal(got[:, j], want), "seed %d" % seed
        assert np.array_equal(want, mt19937.minibatch_starts(int(seed), N, B, steps))


def test_device_batch_generator_matches_host_generator():
    from pysgmcmc_b200.data_batches import DeviceBatchGenerator, generate_batches
    from pysgmcmc_b200.placeholders import placeholder
    N, B = 500, 20
    x = np.arange(N, dtype=np.float64)[:, None]
    y = np.arange(N, dtype=np.float64)
    gen = DeviceBatchGenerator(N, B, seeds=[5, 6, 7], device=DEV, block=64)
    xp, yp = placeholder(), placeholder()
    hosts = [generate_batches(x, y, xp, yp, B, seed=s) for s in (5, 6, 7)]
    for _ in range(200):
        starts = next(gen)[gen.starts_placeholder].cpu().numpy()
        for j, h in enumerate(hosts):
            assert next(h)[xp][0, 0] == starts[j]
