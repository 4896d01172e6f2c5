"""Shared helpers for the parity tests."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# north_star tolerances: fp32 paths 1e-5 relative, bf16 paths 1e-2 relative.
REL_F32 = 1e-5
REL_BF16 = 1e-2


def rel_err(a, ref):
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert a.shape == ref.shape, (a.shape, ref.shape)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - ref)) / (np.max(np.abs(ref)) + 1e-30))


def assert_close(a, ref, rel, what="", atol=0.0):
    """max|a-ref| <= rel * max|ref| + atol (norm-wise relative error; atol only for
    quantities that are analytically zero, e.g. the softmax-shift gradient db2 of DIN-B)."""
    a64 = np.asarray(a, dtype=np.float64)
    r64 = np.asarray(ref, dtype=np.float64)
    assert a64.shape == r64.shape, (what, a64.shape, r64.shape)
    if a64.size == 0:
        return
    err = float(np.max(np.abs(a64 - r64)))
    bound = rel * float(np.max(np.abs(r64))) + atol
    assert err <= bound, f"{what}: max abs error {err:.3e} > {bound:.3e} (rel {rel:.1e}, max|ref| {np.max(np.abs(r64)):.3e})"


def glorot_uniform(rng, fan_in, fan_out, shape=None):
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape or (fan_in, fan_out)).astype(np.float32)


def interacting_params(rng, D, U, random_affine=True):
    W = np.concatenate([glorot_uniform(rng, D, U) for _ in range(4)], axis=1)
    b = (rng.standard_normal(4 * U) * 0.1).astype(np.float32)
    if random_affine:
        gamma = (1.0 + 0.1 * rng.standard_normal(U)).astype(np.float32)
        beta = (0.1 * rng.standard_normal(U)).astype(np.float32)
    else:
        gamma, beta = np.ones(U, np.float32), np.zeros(U, np.float32)
    return W, b, gamma, beta
