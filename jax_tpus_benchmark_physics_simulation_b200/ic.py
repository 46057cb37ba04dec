"""Initial conditions for the 2-D LJ fluid (host side, numpy).

The reference draws R ~ U[0,1)*box and V ~ N(0,1)*sqrt(kT) from jax.random (MD:133-135).  That
placement is unphysical (closest pairs ~0.05 sigma, fp32 overflow by step 2), so
parity and throughput work uses the lattice-plus-jitter state named by BASELINE.json.north_star
and specified in SURVEY.md §8d.  ``reference_style_uniform`` keeps the reference's distribution
(statistically, not bit-for-bit: threefry parity with jax.random is version dependent).
"""
from __future__ import annotations

import numpy as np


def box_size(N: int, rho: float) -> np.float32:
    """MD:30 — ``jnp.sqrt(N / rho)``: python-float quotient, fp32 square root."""
    return np.sqrt(np.float32(N / rho), dtype=np.float32)


def lattice_jitter(N: int, rho: float = 0.8, kT: float = 1.0, seed: int = 0,
                   jitter: float = 0.05):
    """sqrt(N) x sqrt(N) square lattice, spacing a = box/sqrt(N), positions
    (i+1/2, j+1/2)*a + U(-jitter, jitter)*a; velocities N(0,1)*sqrt(kT), COM not removed
    (matches MD:135).  Generated in float64 from ``seed`` and cast once to float32.
    Returns (R (N,2) f32, V (N,2) f32, box f32)."""
    n = int(round(np.sqrt(N)))
    if n * n != N:
        raise ValueError(f"lattice_jitter needs a perfect-square N, got {N}")
    box = box_size(N, rho)
    a = float(box) / n
    rng = np.random.default_rng(seed)
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    R = np.stack([(ii.ravel() + 0.5) * a, (jj.ravel() + 0.5) * a], axis=1)
    R = R + rng.uniform(-jitter, jitter, size=(N, 2)) * a
    V = rng.standard_normal((N, 2)) * np.sqrt(kT)
    R32 = R.astype(np.float32)
    # keep inside the closed interval [0, box] the reference's jnp.mod produces (MD:72)
    np.clip(R32, np.float32(0.0), box, out=R32)
    return R32, V.astype(np.float32), box


def reference_style_uniform(N: int, rho: float = 0.8, kT: float = 1.0, seed: int = 42):
    """MD:133-135 distribution (uniform box placement, Maxwell velocities).  Timing only."""
    box = box_size(N, rho)
    rng = np.random.default_rng(seed)
    R = (rng.uniform(0.0, 1.0, size=(N, 2)).astype(np.float32) * box).astype(np.float32)
    V = (rng.standard_normal((N, 2)).astype(np.float32) * np.float32(np.sqrt(kT))).astype(np.float32)
    return R, V, box
