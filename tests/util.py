"""Shared helpers for the parity tests."""
import numpy as np


def relmax(a, b):
    """max-norm error of `a` against `b`, relative to the max-norm of `b` (the parity metric of DESIGN.md §3)."""
    a, b = np.asarray(a), np.asarray(b)
    scale = np.abs(b).max()
    return float(np.abs(a - b).max() / (scale if scale > 0 else 1.0))


def mixed_field(n, seed=0, sigma_jitter=True):
    """Random particle cloud with overlap ~ FLOWUnsteady's (sigma ~ 2 x spacing), a few static particles and a few
    zero-strength probes; deterministic."""
    rng = np.random.default_rng(seed)
    x = rng.random((n, 3))
    g = rng.standard_normal((n, 3)) / n
    s = 2.125 * n ** (-1.0 / 3.0) * (0.7 + 0.6 * rng.random(n) if sigma_jitter else np.ones(n))
    static = np.zeros(n)
    if n >= 16:
        static[rng.choice(n, n // 16, replace=False)] = 1.0
        g[rng.choice(n, n // 16, replace=False)] = 0.0
    return x, g, s, static
