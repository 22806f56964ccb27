"""The first consumer of the distance matrix, GPU side: the diffusion-map embedding of
``pilotpy.pl.trajectory`` (/root/reference/pilotpy/plot/ploting.py:38-143).

The reference computes ``EMD / EMD.max()`` (:95) and hands it to pydiffmap,
``DiffusionMap.from_sklearn(n_evecs=2, epsilon=1, alpha=0.5, k=64).fit_transform(EMD)`` (:109-110), which treats
the ROWS of the S x S matrix as S-dimensional points.  At S = 20 000 the dominant cost is the k-nearest-
neighbour search over those rows (an S x S x S contraction); the matrix already sits in HBM after
``wasserstein_distance``.  Here:

* the Gram matrix is one library DGEMM (``torch.mm``), selection and distances run in the CUDA kernel behind
  ``pilot_knn_rows`` (include/pilot_b200.h) -- checked against scikit-learn's NearestNeighbors, the routine
  pydiffmap itself calls;
* the rest is S x k sparse algebra restated from pydiffmap's published algorithm (Gaussian kernel
  exp(-d^2 / (4 eps)) on the kNN graph, right-normalisation by q^-alpha, row-normalisation, generator
  (P - I) / eps, leading eigenpairs with ARPACK, coordinates psi / sqrt(-lambda)) on the host with SciPy --
  the same calls pydiffmap makes.  pydiffmap is not installable in this image, so parity of this second half
  against pydiffmap itself is UNPINNED (tests compare with oracle/diffmap_oracle.py, a scikit-learn + SciPy
  restatement); eigenvectors are defined up to sign.

``trajectory(adata, ...)`` writes ``adata.uns['embedding']`` like the reference (:143); the scatter plot is
drawn only when matplotlib is importable (it is not in this image).
"""
from __future__ import annotations

import os

import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spsl
import torch

from . import ops


def knn_graph(X, k: int):
    """(idx [S,k] int32, dist [S,k] float64): the k nearest rows of every row of X, self included, sorted by
    distance -- what NearestNeighbors(n_neighbors=k).fit(X).kneighbors(X) returns.  X: ndarray or CUDA tensor."""
    if not torch.cuda.is_available():
        raise ops._lib.PilotLibraryError("pilot_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    Xd = X if isinstance(X, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).cuda()
    idx, dist = ops.knn_rows(Xd.contiguous(), int(k))
    return idx.cpu().numpy(), dist.cpu().numpy()


def diffusion_embedding(EMD, n_evecs: int = 2, epsilon=1, alpha: float = 0.5, knn: int = 64) -> np.ndarray:
    """Diffusion-map coordinates [S, n_evecs] of the rows of ``EMD`` (already divided by its maximum)."""
    if isinstance(epsilon, str):
        raise NotImplementedError("automatic bandwidth selection (epsilon='bgh') is not part of PILOT's call "
                                  "(ploting.py:109 passes epsilon=1)")
    S = EMD.shape[0]
    k0 = min(int(knn), S)
    idx, dist = knn_graph(EMD, k0)
    # Gaussian kernel on the kNN graph (not symmetrised, as pydiffmap uses kneighbors_graph's matrix as is)
    vals = np.exp(-dist.ravel() ** 2 / (4.0 * float(epsilon)))
    indptr = np.arange(0, S * k0 + 1, k0)
    K = sps.csr_matrix((vals, idx.ravel().astype(np.int64), indptr), shape=(S, S))
    q = np.asarray(K.sum(axis=1)).ravel()
    K = K @ sps.diags(np.power(q, -alpha))                  # right normalisation by the density estimate
    row = np.asarray(K.sum(axis=1)).ravel()
    P = sps.diags(1.0 / row) @ K                            # Markov matrix
    L = (P - sps.eye(S)) / float(epsilon)                   # generator
    n_eig = n_evecs + 1
    if S <= n_eig + 1:
        w, v = np.linalg.eig(L.toarray())
    else:
        w, v = spsl.eigs(L.tocsr(), k=n_eig, which="LR", v0=np.ones(S))
    ix = np.argsort(w.real)[::-1][1:n_eig]                   # drop the trivial eigenpair (lambda = 0)
    evals, evecs = w.real[ix], v.real[:, ix]
    return evecs @ np.diag(np.sqrt(-1.0 / evals))


def trajectory(adata, n_evecs=2, epsilon=1, alpha=0.5, knn=64, sample_col=1, clusters="status", label_act=False,
               colors=("#377eb8", "#ff7f00", "#e41a1c"), location_labels="center", figsize=(12, 12), font_size=24,
               axes_line_width=1, axes_color="black", facecolor="white", point_size=100, cmap="viridis",
               fontsize_legend=24, alpha_trans=1, plot_titel="Trajectory of the disease progression"):
    """``pilotpy.pl.trajectory`` (ploting.py:38-143): diffusion-map embedding of the samples from
    ``adata.uns['EMD']`` into ``adata.uns['embedding']``; the figure is drawn if matplotlib is available."""
    EMD = adata.uns["EMD"] / adata.uns["EMD"].max()
    path = "Results_PILOT/plots"
    if not os.path.exists(path):
        os.makedirs(path)
    embedding = diffusion_embedding(EMD, n_evecs=n_evecs, epsilon=epsilon, alpha=alpha, knn=knn)
    try:
        import matplotlib.pyplot as plt
    except Exception:
        plt = None
    if plt is not None:
        df = adata.uns["annot"]
        df = df.drop_duplicates(subset=[df.columns[sample_col]])
        plt.rcParams.update({"font.size": font_size})
        fig = plt.figure(figsize=figsize)
        ax = plt.gca()
        ax.set(facecolor=facecolor)
        for category in df[clusters].unique():
            aux = np.array(df[clusters] == category)
            g = embedding[aux]
            ax.scatter(g[:, 0], g[:, 1], alpha=alpha_trans, label=category, s=point_size)
            if label_act:
                for txt, (x0, y0) in zip(np.array(df[df[clusters] == category].sampleID), g[:, :2]):
                    ax.annotate(txt, (x0, y0), fontsize=font_size)
        ax.legend(loc=location_labels, fontsize=fontsize_legend)
        plt.title(plot_titel)
        plt.savefig(path + "/" + plot_titel + ".pdf")
        plt.close(fig)
    adata.uns["embedding"] = embedding
