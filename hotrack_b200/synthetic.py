"""Seeded synthetic point clouds shared by bench.py, the tests and the golden generator.

Shapes follow SURVEY.md section 8(d): hand-like clouds span roughly +-0.5 after
canonicalisation, so radii 0.1 / 0.2 capture tens of neighbours at N >= 1024.
"""
import numpy as np


def ball(B, N, seed=0, radius=0.5):
    """Uniform in a ball."""
    rng = np.random.RandomState(seed)
    v = rng.randn(B, N, 3)
    v /= np.linalg.norm(v, axis=-1, keepdims=True) + 1e-12
    r = radius * rng.rand(B, N, 1) ** (1.0 / 3.0)
    return (v * r).astype(np.float32)


def shell(B, N, seed=0, radius=0.4, noise=0.01):
    """Noisy sphere surface (depth-camera-like)."""
    rng = np.random.RandomState(seed)
    v = rng.randn(B, N, 3)
    v /= np.linalg.norm(v, axis=-1, keepdims=True) + 1e-12
    return (v * radius + noise * rng.randn(B, N, 3)).astype(np.float32)


def lattice(B, N, seed=0, step=1.0 / 16.0, extent=8):
    """Points on a coarse lattice: masses of EXACT distance ties and exact duplicates."""
    rng = np.random.RandomState(seed)
    return (rng.randint(-extent, extent + 1, size=(B, N, 3)) * step).astype(np.float32)


def duplicates(B, N, seed=0, frac=0.25):
    """Ball cloud where a fraction of the points are exact copies of other points."""
    rng = np.random.RandomState(seed)
    pts = ball(B, N, seed)
    ndup = int(N * frac)
    for b in range(B):
        src = rng.randint(0, N, size=ndup)
        dst = rng.randint(0, N, size=ndup)
        pts[b, dst] = pts[b, src]
    return pts


def keypoints(B, n=21, seed=0, sigma=0.3):
    rng = np.random.RandomState(seed + 7919)
    return (sigma * rng.randn(B, n, 3)).astype(np.float32)


KINDS = {"ball": ball, "shell": shell, "lattice": lattice, "duplicates": duplicates}


def make(kind, B, N, seed=0):
    return KINDS[kind](B, N, seed)
