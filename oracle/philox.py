"""Philox4x32-10 counter RNG in NumPy (oracle side of the sampler / dropout streams).

TEST INFRASTRUCTURE.  The reference draws its samples with an unseeded
``tf.multinomial`` (reinforcement_multisampling_tf_s2vt.py:335) and its dropout
masks with an unseeded DropoutWrapper (:85-87), so it is itself not
reproducible.  Both sides of the parity tests therefore use the
same counter-based stream, defined here and in csrc/philox.cuh:

    counter = (c0, c1, c2, c3), key = (seed_lo, seed_hi)
    uniform = ((x >> 9) + 0.5) * 2**-23          (strictly inside (0, 1), exact in fp32)

    sampler : c0 = vocab index // 4, c1 = decode step, c2 = global row, c3 = STREAM_SAMPLE
              lane (vocab index % 4) selects one of the four outputs
    dropout : c0 = unit // 4, c1 = time step (0 .. T_v+T_c-1), c2 = global row,
              c3 = STREAM_DROP1 / STREAM_DROP2, lane = unit % 4
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85

STREAM_SAMPLE = 0x53414D50  # 'SAMP'
STREAM_DROP1 = 0x44525031   # 'DRP1'
STREAM_DROP2 = 0x44525032   # 'DRP2'


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds.  All inputs broadcastable uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    sh = np.uint64(32)
    for r in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> sh, p0 & mask
        hi1, lo1 = p1 >> sh, p1 & mask
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n1 = lo1
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        n3 = lo0
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def u32_to_uniform(x):
    """uint32 -> float32 uniform strictly in (0,1): ((x>>9)+0.5)*2^-23."""
    return ((x >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23)


def uniforms(seed, stream, rows, step, n):
    """float32 [len(rows), n] uniforms for element index 0..n-1 of each global row at `step`."""
    rows = np.asarray(rows, dtype=np.uint32).reshape(-1, 1)
    ngrp = (n + 3) // 4
    grp = np.arange(ngrp, dtype=np.uint32).reshape(1, -1)
    o = philox4x32_10(grp, np.uint32(step), rows, np.uint32(stream), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = np.stack([u32_to_uniform(x) for x in o], axis=-1).reshape(rows.shape[0], ngrp * 4)
    return u[:, :n]


def gumbel(seed, rows, step, n):
    """float32 Gumbel(0,1) noise -log(-log(u)) for the categorical sampler (fp32 arithmetic)."""
    u = uniforms(seed, STREAM_SAMPLE, rows, step, n)
    return -np.log(-np.log(u, dtype=np.float32), dtype=np.float32)


def dropout_mask(seed, stream, rows, step, n, keep):
    """float32 [len(rows), n]: 1/keep where kept (u < keep), else 0 (DropoutWrapper output_keep_prob)."""
    u = uniforms(seed, stream, rows, step, n)
    return np.where(u < np.float32(keep), np.float32(1.0) / np.float32(keep), np.float32(0.0)).astype(np.float32)
