"""Seeded synthetic clouds of the benchmark configs.  numpy only."""
from __future__ import annotations

import numpy as np


def uniform(seed: int, n: int, d: int = 3) -> np.ndarray:
    """U[0,1) cloud generated directly in float32."""
    return np.random.default_rng(seed).random((n, d), dtype=np.float32)


def lidar(seed: int, n: int) -> np.ndarray:
    """Ring-structured 64-beam spinning-lidar-like cloud (density ~ 1/r, negative coordinates).

    Trigonometry is evaluated in float64 and rounded once to float32, so the array is the same on every
    host CPU (float32 SIMD sin/cos differ by an ulp between AVX2 and AVX-512 builds of numpy).
    """
    g = np.random.default_rng(seed)
    az = g.random(n, dtype=np.float32).astype(np.float64) * (2 * np.pi)
    beam = g.integers(0, 64, n)
    el = np.deg2rad(-25.0 + beam * (28.0 / 63.0))
    r = 2.0 + 78.0 * g.random(n, dtype=np.float32).astype(np.float64) * g.random(n, dtype=np.float32).astype(np.float64)
    xyz = np.stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el)], axis=1)
    return np.ascontiguousarray(xyz, dtype=np.float32)


def grid_ties(seed: int, n: int, d: int = 3, levels: int = 6) -> np.ndarray:
    """Integer-lattice cloud: full of exact distance ties and duplicates (tie-rule stress)."""
    g = np.random.default_rng(seed)
    return g.integers(0, levels, (n, d)).astype(np.float32)


def uniform_batch(base_seed: int, b: int, n: int, d: int = 3) -> np.ndarray:
    """[B,N,D] batch, cloud i seeded base_seed+i (identical to uniform(base_seed+i, n, d))."""
    out = np.empty((b, n, d), dtype=np.float32)
    for i in range(b):
        out[i] = uniform(base_seed + i, n, d)
    return out
