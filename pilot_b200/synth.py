"""Synthetic inputs of the shapes BASELINE.json names and a duck-typed
AnnData stand-in (anndata itself is not installed in this image).

Host-side data generation only; used by tests/, bench.py and __graft_entry__.smoke().
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import pandas as pd
import scipy.spatial.distance as ssd

# name -> (cells, dim, types, samples, seed)
CONFIGS: Dict[str, tuple] = {
    "c1": (200_000, 30, 10, 20, 1),      # cosine, exact EMD, CPU-runnable
    "c2": (1_000_000, 50, 30, 100, 2),   # Sinkhorn reg 0.1 on 1 x B200
    "c3": (5_000_000, 50, 40, 600, 3),   # EMD + Sinkhorn on 8 x B200
    "c4": (2_000_000, 50, 64, 2_000, 4), # log-domain Sinkhorn reg 0.01
}


class _View:
    def __init__(self, X):
        self.X = X


class FakeAnnData:
    """The subset of the AnnData surface the hot path touches: .obs, .obsm, .uns,
    .var_names and adata[:, names].X (Trajectory.py:255-263, 291-296)."""

    def __init__(self, obs: pd.DataFrame, obsm: Optional[dict] = None, X: Optional[np.ndarray] = None,
                 var_names=None):
        self.obs = obs
        self.obsm = obsm or {}
        self.uns = {}
        self.X = X
        self.var_names = pd.Index(var_names if var_names is not None else [])

    def __getitem__(self, key):
        rows, names = key
        assert rows == slice(None)
        idx = [self.var_names.get_loc(n) for n in names]
        return _View(self.X[:, idx])


def make_cells(n_cells: int, dim: int, n_types: int, n_samples: int, seed: int, dtype=np.float32,
               labels: str = "str", type_prefix: str = "ct", sample_prefix: str = "s"):
    """Cells path (C1-C4): type centroids mu_k ~ N(0, I); per-sample mixtures
    theta_s ~ Dirichlet(0.5); cell -> sample uniform (shuffled so first-appearance order
    differs from sorted order); type ~ Cat(theta_s); x = mu_type + 0.5 N(0, I).
    Returns (X [n, dim], obs DataFrame with cell_types / sampleID / status)."""
    rng = np.random.default_rng(seed)
    mu = rng.normal(size=(n_types, dim))
    theta = rng.dirichlet(0.5 * np.ones(n_types), size=n_samples)
    smp = rng.integers(0, n_samples, size=n_cells)
    cdf = np.cumsum(theta, axis=1)
    u = rng.random(n_cells)
    typ = (u[:, None] > cdf[smp]).sum(axis=1).clip(0, n_types - 1)
    X = (mu[typ] + 0.5 * rng.normal(size=(n_cells, dim))).astype(dtype)
    # label strings in a scrambled order so code order != appearance order != sorted order
    tnames = np.array([f"{type_prefix}{(7 * k + 3) % n_types:03d}" for k in range(n_types)], dtype=object)
    snames = np.array([f"{sample_prefix}{(11 * s + 5) % n_samples:05d}" for s in range(n_samples)], dtype=object)
    status = np.where(smp % 2 == 1, "case", "ctrl").astype(object)
    obs = pd.DataFrame({"cell_types": tnames[typ], "sampleID": snames[smp], "status": status})
    if labels == "categorical":
        for c in obs.columns:
            obs[c] = obs[c].astype("category")
    elif labels == "int":
        obs["cell_types"] = typ.astype(np.int64)
        obs["sampleID"] = (smp.astype(np.int64) * 3 + 100)
    return np.ascontiguousarray(X), obs


def make_adata(config: str = "c1", labels: str = "str", scale: float = 1.0, emb_key: str = "X_PCA",
               dtype=np.float32) -> FakeAnnData:
    n, d, k, s, seed = CONFIGS[config]
    n = max(int(n * scale), k * s)
    X, obs = make_cells(n, d, k, s, seed, dtype=dtype, labels=labels)
    return FakeAnnData(obs, obsm={emb_key: X})


def smooth_props(counts: np.ndarray, regulizer: float = 0.2) -> np.ndarray:
    """PILOT smoothing of a count table (Trajectory.py:405-430) -- vectorised, host side,
    for synthetic pairs-only inputs."""
    counts = counts.astype(np.float64)
    n_k = counts.sum(axis=0)
    prior = (n_k / (counts.sum() - 1)) * regulizer
    sp = 0.0
    for p in prior:
        sp = sp + p
    return (counts + prior) / (counts.sum(axis=1, keepdims=True) + sp)


def make_pairs(n_samples: int, n_types: int, seed: int = 5, dim: int = 50, metric: str = "cosine",
               cells_per_sample: int = 5000):
    """Pairs-only path (C5 and kernel micro-benchmarks): counts ~ Multinomial(5000,
    Dirichlet(1)), PILOT smoothing, cost = pdist(N(0, I)[K x dim]) / max.
    Returns (props float64 [S, K], cost_norm float64 [K, K])."""
    rng = np.random.default_rng(seed)
    theta = rng.dirichlet(np.ones(n_types), size=n_samples)
    counts = np.stack([rng.multinomial(cells_per_sample, t) for t in theta])
    props = smooth_props(counts)
    cent = rng.normal(size=(n_types, dim))
    cost = ssd.squareform(ssd.pdist(cent, metric))
    return np.ascontiguousarray(props), np.ascontiguousarray(cost / cost.max())
