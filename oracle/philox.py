"""Philox4x32-10 counter-based RNG, numpy restatement of the generator the CUDA driver uses
(learn_region_grow_b200/csrc/lrg_rng.cuh).  TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference consumes one global MT19937 stream (test_region_grow.py:21,238,250,266-267); that cannot be
sharded over rooms/GPUs, so the engine's throughput mode defines its own stream: every draw is
``philox(key=seed, counter=(element, (seed point << 8) | stream, step within the region, room))[0]`` -- a pure function of
its coordinates (oracle/lrg_driver.py PhiloxRng).
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)

STREAM_INLIER_KEY = 0      # sort keys for choice(n, 512, replace=False) over the inlier list
STREAM_NEIGHBOR_KEY = 1
STREAM_ADD_UNIFORM = 2     # uniforms compared with the add confidences
STREAM_REMOVE_UNIFORM = 3
STREAM_INLIER_PAD = 4      # with-replacement padding draws when the list is shorter than 512
STREAM_NEIGHBOR_PAD = 5


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds; all args broadcastable uint32 arrays; returns word 0..3."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & _MASK for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & _MASK, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & _MASK, lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def draw_u32(seed, room, step, stream, n):
    """n raw 32-bit draws for elements 0..n-1 of (room, step, stream)."""
    e = np.arange(n, dtype=np.uint64)
    return philox4x32_10(e, np.uint64(stream), np.uint64(step), np.uint64(room),
                         seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)[0]


def u32_to_unit_float(x):
    """Top 24 bits -> float32 in [0,1) (what the device compares against the float32 confidence)."""
    return (x >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
