"""Start structures (compute_init_struct, initial_structure_tools.py:256-289).

The default, ``hilbert``, is generated on the device by the engine (mmm_hilbert_points /
mmm_hilbert_init); ``hilbert_points_host`` is the same integer decode in numpy for callers that
have no engine yet.  The other eight curves are host numpy and draw from ``np.random`` (seeded
earlier by the loaders, utils.py:233/439) in the reference's order.  All curves are written to
the init CIF unscaled, i.e. as Angstrom, and read back as nm / 10 (model.py:753).
"""
from __future__ import annotations

import numpy as np


def hilbert_points_host(n: int, p: int = 8) -> np.ndarray:
    """First n lattice points of the order-p 3-D Hilbert curve (Skilling transpose -> axes, as
    hilbertcurve 2.0.5 ``points_from_distances``), vectorised over all points."""
    h = np.arange(n, dtype=np.uint64)
    X = [np.zeros(n, dtype=np.uint32) for _ in range(3)]
    for b in range(3 * p):  # bit string MSB first; axis a takes characters a, a+3, ...
        bit = ((h >> np.uint64(3 * p - 1 - b)) & np.uint64(1)).astype(np.uint32)
        X[b % 3] = (X[b % 3] << np.uint32(1)) | bit
    t = X[2] >> np.uint32(1)
    X[2] = X[2] ^ X[1]
    X[1] = X[1] ^ X[0]
    X[0] = X[0] ^ t
    q = 2
    while q != (2 << (p - 1)):
        pm = np.uint32(q - 1)
        for i in (2, 1, 0):
            hit = (X[i] & np.uint32(q)) != 0
            t = (X[0] ^ X[i]) & pm
            x0_inv = X[0] ^ pm
            x0_exc = X[0] ^ t
            xi_exc = X[i] ^ t
            if i == 0:
                X[0] = np.where(hit, x0_inv, X[0])  # exchange with itself is the identity
            else:
                X[i] = np.where(hit, X[i], xi_exc)
                X[0] = np.where(hit, x0_inv, x0_exc)
        q <<= 1
    return np.stack(X, axis=1).astype(np.int32)


def polymer_circle(n, z_stretch=1.0, radius=5.0):
    ang = 360.0 / float(n)
    i = np.arange(n)
    z = np.cumsum(np.full(n, z_stretch / n)) if z_stretch != 0 else np.zeros(n)
    return np.column_stack((radius * np.cos(ang * i * np.pi / 180), radius * np.sin(ang * i * np.pi / 180), z))


def helix_structure(n, radius=1, pitch=2):
    th = np.linspace(0, 4 * np.pi, n)
    return np.column_stack((radius * np.cos(th), radius * np.sin(th), np.linspace(0, pitch * n, n)))


def spiral_structure(n, initial_radius=1, pitch=1, growth_factor=0.05):
    th = np.linspace(0, 4 * np.pi, n)
    r = initial_radius + growth_factor * np.arange(n)
    return np.column_stack((r * np.cos(th), r * np.sin(th), np.linspace(0, pitch * n, n)))


def sphere_surface_structure(n, radius=1):
    phi = np.random.uniform(0, 2 * np.pi, n)
    costheta = np.random.uniform(-1, 1, n)
    u = np.random.uniform(0, 1, n)
    th = np.arccos(costheta)
    r = radius * u ** (1 / 3)
    return np.column_stack((r * np.sin(th) * np.cos(phi), r * np.sin(th) * np.sin(phi), r * np.cos(th)))


def confined_random_walk(n, box_size=5):
    v = np.zeros((n, 3))
    steps = np.random.choice([-1, 1], size=(max(n - 1, 0), 3))
    for i in range(1, n):  # clipping makes this recurrence sequential
        v[i] = np.clip(v[i - 1] + steps[i - 1], -box_size, box_size)
    return v


def trefoil_knot_structure(n, scale=5):
    t = np.linspace(0, 2 * np.pi, n)
    return np.column_stack((scale * (np.sin(t) + 2 * np.sin(2 * t)), scale * (np.cos(t) - 2 * np.cos(2 * t)),
                            -scale * np.sin(3 * t)))


def random_walk_structure(n, step_size=1):
    d = np.random.normal(size=(max(n - 1, 0), 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([np.zeros((1, 3)), np.cumsum(step_size * d, axis=0)])[:n]


def self_avoiding_random_walk(n, step=1.0, bead_radius=0.5, epsilon=0.001):
    pts = [np.zeros(3)]
    for _ in range(n - 1):
        trials, cand = 0, pts[-1]
        while trials < 1000:
            v = np.random.normal(0, 1, 3)
            nv = np.linalg.norm(v)
            cand = pts[-1] + step * (v / nv if nv > 0 else np.array([1.0, 0.0, 0.0]))
            if np.min(np.linalg.norm(np.asarray(pts) - cand, axis=1)) < 2 * bead_radius - epsilon:
                trials += 1
            else:
                break
        pts.append(cand)
    return np.asarray(pts)


def compute_init_struct(n_beads: int, mode="hilbert", engine=None) -> np.ndarray:
    """(N,3) start coordinates in the units the init CIF is written in."""
    mode = getattr(mode, "value", mode)
    if mode == "hilbert":
        pts = engine.hilbert_points(8) if engine is not None else hilbert_points_host(n_beads, 8)
        return pts.astype(np.float64)
    table = {
        "rw": random_walk_structure, "confined_rw": confined_random_walk, "knot": trefoil_knot_structure,
        "self_avoiding_rw": self_avoiding_random_walk, "circle": lambda n: polymer_circle(n, 50, 5),
        "helix": helix_structure, "spiral": spiral_structure, "sphere": sphere_surface_structure,
    }
    if mode not in table:
        raise ValueError(f"Invalid option for initial structure: {mode!r}. Choose one of: rw, confined_rw, knot, "
                         "self_avoiding_rw, circle, helix, spiral, sphere, hilbert.")
    return np.asarray(table[mode](n_beads), dtype=np.float64)
