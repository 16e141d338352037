"""Structure report — the numeric part of the reference's ``analyze_structure``
(src/multimm/plots.py:630-829), re-hosted without matplotlib: same quantities, same report text,
the curves the reference only plots are saved as arrays.

The one O(N^2) quantity (mean of the full distance matrix, plots.py:663-664) comes from the engine
when one is passed (a tiled kernel that never materialises the matrix: the reference needs N^2 x 8
bytes, 320 GB at N = 2e5) and from a blocked numpy loop otherwise.  The O(N x window) local-Rg loop
(plots.py:712-720) is replaced by prefix sums, O(N).
"""
from __future__ import annotations

import os

import numpy as np


def mean_pair_distance_host(V: np.ndarray, block: int = 256) -> float:
    n = len(V)
    total = 0.0
    for a in range(0, n, block):
        d = V[a:a + block, None, :] - V[None, :, :]
        total += float(np.sqrt((d * d).sum(axis=2)).sum())
    return total / (float(n) * float(n))


def local_rg(V: np.ndarray, window: int) -> np.ndarray:
    """sqrt(mean |x - cm|^2) over each window [i, i + window), i in [0, N - window): prefix sums."""
    n = len(V)
    if n - window <= 0:
        return np.zeros(0)
    c1 = np.vstack([np.zeros((1, 3)), np.cumsum(V, axis=0)])
    c2 = np.concatenate([[0.0], np.cumsum((V * V).sum(axis=1))])
    i = np.arange(n - window)
    s1 = c1[i + window] - c1[i]
    s2 = c2[i + window] - c2[i]
    var = s2 / window - (s1 * s1).sum(axis=1) / (window * window)
    return np.sqrt(np.maximum(var, 0.0))


CONTACT_BINS = 1000  # block-mean contact map saved with every report (the N x N matrix never exists)


def device_pair_stats(V, bins: int, device: int = 0, log_scale: bool = True):
    """(B x B block-mean contact map, np.mean(cdist(V, V))) of ANY (n, 3) array on the GPU
    (csrc/mmm_analysis.cu: mmm_contact_map), one pass over the N^2 pairs in FP64."""
    import ctypes as C

    from . import _lib

    V = np.ascontiguousarray(V, dtype=np.float64)
    n = len(V)
    bins = max(1, min(int(bins), n))
    out = np.empty((bins, bins), dtype=np.float64)
    mean = C.c_double()
    lib = _lib.load()
    rc = lib.mmm_contact_map(int(device), V.ctypes.data_as(C.c_void_p), n, bins, int(bool(log_scale)),
                             out.ctypes.data_as(C.c_void_p), C.byref(mean))
    if rc != 0:
        raise _lib.Error(rc, lib.mmm_last_error(None).decode() or "CUDA error: contact-map pass failed")
    return out, float(mean.value)


def analyze_structure(V, save_path, name="structure", engine=None, device=None) -> dict:
    """`device` (a CUDA device index): the O(N^2) quantities — the mean pair distance of the report and
    the block-mean contact map (saved as <name>_contact_map.npy; the heat-map the reference draws only
    below 5e4 beads, model.py:1095-1104) — come from one pass on the GPU, for any number of rows of V."""
    V = np.asarray(V, dtype=np.float64)
    V = V[np.isfinite(V).all(axis=1)]
    n = len(V)
    base = os.path.join(save_path, "analysis")
    os.makedirs(base, exist_ok=True)

    r_cm = np.mean(V, axis=0)
    Vc = V - r_cm
    rg = np.sqrt(np.mean(np.sum(Vc ** 2, axis=1)))
    ree = np.linalg.norm(V[-1] - V[0])
    if device is not None:
        cmap, mean_dist = device_pair_stats(V, CONTACT_BINS, device=device)
        np.save(os.path.join(base, f"{name}_contact_map.npy"), cmap)
    elif engine is not None and engine.n == n:
        engine.set_positions(V)
        mean_dist = engine.mean_pair_distance()
    else:
        mean_dist = mean_pair_distance_host(V)
    try:
        from scipy.spatial import ConvexHull

        volume = ConvexHull(V).volume
    except Exception:
        volume = np.nan
    density = n / volume if volume > 0 else np.nan
    G = np.dot(Vc.T, Vc) / n
    eigvals = np.sort(np.linalg.eigvalsh(G))
    l1, l2, l3 = eigvals
    asphericity = l3 - 0.5 * (l1 + l2)
    acylindricity = l2 - l1
    bonds = np.linalg.norm(np.diff(V, axis=0), axis=1)
    v1, v2 = V[1:-1] - V[:-2], V[2:] - V[1:-1]
    cos_angles = np.sum(v1 * v2, axis=1) / (np.linalg.norm(v1, axis=1) * np.linalg.norm(v2, axis=1) + 1e-8)
    angles = np.arccos(np.clip(cos_angles, -1, 1))
    separations = np.arange(1, min(500, n // 2))
    spatial = np.array([np.mean(np.linalg.norm(V[s:] - V[:n - s], axis=1)) for s in separations])
    window = max(10, n // 100)
    lrg = local_rg(V, window)

    with open(os.path.join(base, f"{name}_report.txt"), "w") as f:
        f.write("===== STRUCTURE ANALYSIS =====\n\n")
        f.write(f"N beads: {n}\n\n")
        f.write("---- Global ----\n")
        f.write(f"Rg: {rg:.4f}\n")
        f.write(f"Ree: {ree:.4f}\n")
        f.write(f"Mean distance: {mean_dist:.4f}\n\n")
        f.write("---- Volume ----\n")
        f.write(f"Volume: {volume:.4f}\n")
        f.write(f"Density: {density:.6f}\n\n")
        f.write("---- Shape ----\n")
        f.write(f"Eigenvalues: {eigvals}\n")
        f.write(f"Asphericity: {asphericity:.6f}\n")
        f.write(f"Acylindricity: {acylindricity:.6f}\n\n")
        f.write("---- Local properties ----\n")
        f.write(f"Mean bond length: {np.mean(bonds):.4f}\n")
        f.write(f"Mean angle (rad): {np.mean(angles):.4f}\n\n")
        f.write("Interpretation:\n")
        f.write("Rg ~ size of polymer\n")
        f.write("Distance vs separation → scaling law\n")
        f.write("Angles → stiffness\n")
        f.write("Local Rg → domain compaction\n")
    # what the reference draws (plots.py:765-829) saved as data
    np.savez_compressed(os.path.join(base, f"{name}_curves.npz"), bonds=bonds, angles=angles, separations=separations,
                        spatial_dists=spatial, local_rg=lrg)
    return dict(n=n, rg=rg, ree=ree, mean_dist=mean_dist, volume=volume, density=density, eigvals=eigvals,
                asphericity=asphericity, acylindricity=acylindricity, mean_bond=float(np.mean(bonds)),
                mean_angle=float(np.mean(angles)), separations=separations, spatial_dists=spatial, local_rg=lrg)


def contact_map(V, log_scale: bool = True, reorder_by_diagonal: bool = False, bins: int | None = None) -> np.ndarray:
    """The matrix the reference's ``get_heatmap`` draws (plots.py:540-561): contact strength
    1 / (d + 1)^(2/3) of every bead pair, log1p-transformed, optionally reordered by distance from
    the centroid.  ``bins=None`` is that N x N matrix (the reference refuses it at N >= 5e4,
    model.py:1095); ``bins=B`` returns the B x B matrix of block means over consecutive beads
    instead, built block by block so that the N x N matrix never exists — the form in which a
    genome-wide structure can still be looked at."""
    V = np.asarray(V, dtype=np.float64)
    n = len(V)
    if reorder_by_diagonal:
        V = V[np.argsort(np.linalg.norm(V - np.mean(V, axis=0), axis=1))]

    def strength(a, b):
        d = a[:, None, :] - b[None, :, :]
        m = 1.0 / (np.sqrt((d * d).sum(axis=2)) + 1.0) ** (2.0 / 3.0)
        return np.log1p(m) if log_scale else m

    if bins is None:
        out = np.empty((n, n))
        for a in range(0, n, 1024):
            out[a:a + 1024] = strength(V[a:a + 1024], V)
        return out
    bins = max(1, min(int(bins), n))  # no empty blocks
    edges = np.linspace(0, n, bins + 1).astype(np.int64)
    width = np.diff(edges)
    out = np.zeros((bins, bins))
    for p in range(bins):
        a0, a1 = edges[p], edges[p + 1]
        cols = np.zeros(n)
        for a in range(a0, a1, 512):  # column sums of this block row, 512 beads at a time
            cols += strength(V[a:min(a + 512, a1)], V).sum(axis=0)
        out[p] = np.add.reduceat(cols, edges[:-1]) / ((a1 - a0) * width)
    return out
