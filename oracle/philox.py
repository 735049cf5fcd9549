"""Oracle for the in-kernel noise generator (test infrastructure only).

The reference draws noise with ``tf.random_normal`` (base_classes.py:218-220);
TF's stream is third-party and not reproducible, so the engine defines its own
counter-based stream and parity on trajectories is established with INJECTED
noise.  This file restates the engine's generator so the in-kernel path can be
checked too:

  Philox4x32-10 (Salmon et al., SC'11; Random123), PINNED by the Random123
  known-answer vectors in tests/test_oracle.py.
  counter = (group_lo, group_hi, step_lo, step_hi), key = (seed_lo, seed_hi),
  group = global flat element index // 4; the 4 outputs serve elements
  4*group .. 4*group+3 via two Box-Muller pairs:
      u = float32(x) * 2^-32 + 2^-33   (round-to-nearest, in (0, 1])
      z0 = sqrt(-2 ln u0) * cos(2 pi u1),  z1 = sqrt(-2 ln u0) * sin(2 pi u1)
      z2, z3 likewise from (u2, u3).
"""
import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(ctr, key):
    """ctr: uint32 [..., 4], key: uint32 [..., 2] (broadcastable) -> uint32 [..., 4]."""
    c = [np.asarray(ctr[..., i], dtype=np.uint64) for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint64)
    k1 = np.asarray(key[..., 1], dtype=np.uint64)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = [(hi1 ^ c[1] ^ k0) & mask, lo1, (hi0 ^ c[3] ^ k1) & mask, lo0]
        k0 = (k0 + np.uint64(W0)) & mask
        k1 = (k1 + np.uint64(W1)) & mask
    return np.stack([x.astype(np.uint32) for x in np.broadcast_arrays(*c)], axis=-1)


def uniform_open(x):
    """uint32 -> float32 in (0, 1]: float32(x) * 2^-32 + 2^-33, each op rounded to fp32."""
    xf = np.asarray(x, dtype=np.uint32).astype(np.float32)
    return xf * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)


def normals(n_elems, seed, step, elem_offset=0):
    """The engine's N(0,1) stream for global elements [elem_offset, elem_offset+n).

    Returned in float64 (exact transform of the fp32 uniforms); the kernel
    evaluates log/sincos in fp32, so compare with rtol~1e-5 / atol~1e-6.
    """
    assert elem_offset % 4 == 0
    n_groups = (n_elems + 3) // 4
    group = np.arange(n_groups, dtype=np.uint64) + np.uint64(elem_offset // 4)
    ctr = np.stack([group & np.uint64(0xFFFFFFFF), group >> np.uint64(32),
                    np.full_like(group, step & 0xFFFFFFFF),
                    np.full_like(group, (step >> 32) & 0xFFFFFFFF)], axis=-1).astype(np.uint32)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    x = philox4x32_10(ctr, key)
    u = uniform_open(x).astype(np.float64)
    r0 = np.sqrt(-2.0 * np.log(u[:, 0]))
    r1 = np.sqrt(-2.0 * np.log(u[:, 2]))
    z = np.stack([r0 * np.cos(2 * np.pi * u[:, 1]), r0 * np.sin(2 * np.pi * u[:, 1]),
                  r1 * np.cos(2 * np.pi * u[:, 3]), r1 * np.sin(2 * np.pi * u[:, 3])], axis=-1)
    return z.reshape(-1)[:n_elems]
