"""Oracle: Philox4x32-10 counter RNG, numpy restatement of the device generator.

TEST INFRASTRUCTURE -- see oracle/__init__.py.

The reference never seeds TensorFlow (locator.py:170-171 seeds numpy only), so
weight init / dropout masks / shuffle order are free choices of the new build.
The CUDA path draws them from Philox4x32-10 (Salmon et al. 2011, the published
algorithm; constants below) keyed by (seed, stream) with the element index as
counter; this file mirrors that bit-for-bit so tests can start the oracle from
the same weights and masks as the device.

Layout contract shared with locator_b200/csrc/philox.cuh:
  counter = (idx_lo, idx_hi, stream, 0), key = (seed_lo, seed_hi)
  element e of a tensor uses block idx = e // 4, lane = e % 4
  uniform u = (x >> 8) * 2**-24   in [0, 1)
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0 = np.asarray(c0, dtype=np.uint32).copy()
    c1 = np.asarray(c1, dtype=np.uint32).copy()
    c2 = np.asarray(c2, dtype=np.uint32).copy()
    c3 = np.asarray(c3, dtype=np.uint32).copy()
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & MASK).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & MASK).astype(np.uint32)
            n0 = hi1 ^ c1 ^ k0
            n1 = lo1
            n2 = hi0 ^ c3 ^ k1
            n3 = lo0
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def random_u32(n, seed, stream):
    """n uint32 values: element e = lane (e%4) of block (e//4)."""
    nblk = (n + 3) // 4
    idx = np.arange(nblk, dtype=np.uint64)
    c0 = (idx & MASK).astype(np.uint32)
    c1 = (idx >> np.uint64(32)).astype(np.uint32)
    c2 = np.full(nblk, np.uint32(stream & 0xFFFFFFFF), dtype=np.uint32)
    c3 = np.zeros(nblk, dtype=np.uint32)
    r = philox4x32_10(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(r, axis=1).reshape(-1)[:n]


def uniform01(n, seed, stream):
    return ((random_u32(n, seed, stream) >> np.uint32(8)).astype(np.float32)) * np.float32(2.0**-24)


def glorot_uniform(fan_in, fan_out, seed, stream):
    """Keras glorot_uniform: U(-l, l), l = sqrt(6/(fan_in+fan_out)); [fan_in, fan_out] fp32."""
    limit = np.float32(np.sqrt(6.0 / (fan_in + fan_out)))
    u = uniform01(fan_in * fan_out, seed, stream)
    w = (np.float32(2.0) * u - np.float32(1.0)) * limit
    return w.reshape(fan_in, fan_out).astype(np.float32)


def dropout_keep(batch, width, p, seed, step):
    """keep mask [batch, width] for optimizer step `step` (0-based): keep iff u >= p.

    stream = 0x40000000 + step; element index = b * width + j.
    """
    u = uniform01(batch * width, seed, 0x40000000 + step)
    return (u >= np.float32(p)).reshape(batch, width)
