"""Deterministic synthetic layers for the parity tests.

``W ~ N(0, 0.02^2)`` fp32; calibration activations per sample ``i``:
``X_i = Z_i + lam * F_i @ B / sqrt(r)`` with ``Z_i ~ N(0,1)``, ``F_i ~ N(0,1)^{L x r}``,
``B ~ N(0,1)^{r x m}`` fixed per layer, rounded ONCE to fp16 and handed to both sides
(the oracle sees ``X.astype(float32)``), so tensor-core products are exact and only the
accumulation order differs.  No per-channel scale heterogeneity.

numpy ``default_rng`` streams: the golden generator and the tests call the same code on
the same numpy, and each fixture stores a checksum of its inputs to catch RNG drift.
"""

import numpy as np


def make_weight(n, m, seed, scale=0.02, row_offset=0.0):
    rng = np.random.default_rng(seed)
    W = rng.standard_normal((n, m), dtype=np.float32) * np.float32(scale)
    if row_offset:
        W += (rng.standard_normal((n, 1), dtype=np.float32) * np.float32(row_offset))
    return np.ascontiguousarray(W, dtype=np.float32)


def make_activations(num_samples, seq_len, m, seed, lam=0.5, r=64):
    """Returns fp16 array (num_samples, seq_len, m)."""
    rng_b = np.random.default_rng(seed)
    B = rng_b.standard_normal((r, m), dtype=np.float32)
    out = np.empty((num_samples, seq_len, m), dtype=np.float16)
    for i in range(num_samples):
        rng = np.random.default_rng(seed + 1 + i)
        Z = rng.standard_normal((seq_len, m), dtype=np.float32)
        if lam:
            F = rng.standard_normal((seq_len, r), dtype=np.float32)
            Z += np.float32(lam / np.sqrt(r)) * (F @ B)
        out[i] = Z.astype(np.float16)
    return out


def checksum(*arrays):
    """Order-sensitive fp64 checksum used to detect RNG drift between generator and test."""
    acc = 0.0
    for a in arrays:
        a64 = np.asarray(a, dtype=np.float64).reshape(-1)
        w = np.arange(1, a64.shape[0] + 1, dtype=np.float64)
        acc += float((a64 * np.cos(w)).sum())
    return acc
