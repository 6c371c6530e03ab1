"""Shared helpers for the parity tests (seeded synthetic clouds)."""
import numpy as np


def rand_cloud(seed, B, N, scale=1.0, offset=0.0):
    rng = np.random.default_rng(seed)
    return (rng.random((B, N, 3), dtype=np.float32) * np.float32(scale) + np.float32(offset)).astype(np.float32)


def lattice_cloud(seed, B, N, side=6):
    """Integer-lattice points (many exact distance ties) -> stresses the lowest-index tie rule."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, side, size=(B, N, 3)).astype(np.float32)


def shape_cloud(seed, B, N):
    """Points on a noisy superquadric-ish surface in [-0.5, 0.5]^3 (stand-in for a generated shape)."""
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((B, N, 3))
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    r = 0.35 + 0.1 * np.sin(3 * v[..., :1]) * np.cos(2 * v[..., 1:2])
    return (v * r * np.array([1.0, 0.7, 0.5])).astype(np.float32)
