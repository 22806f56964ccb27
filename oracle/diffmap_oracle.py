"""oracle/diffmap_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of what pilotpy.pl.trajectory computes before it plots
(/root/reference/pilotpy/plot/ploting.py:95-110):
    EMD = EMD / EMD.max()
    DiffusionMap.from_sklearn(n_evecs=2, epsilon=1, alpha=0.5, k=64).fit_transform(EMD)
The arithmetic lives in pydiffmap (PyPI ``pydiffmap``, a dependency of pilotpy, not vendored under
/root/reference and not installable here) and, below it, in scikit-learn's NearestNeighbors and SciPy's ARPACK
wrapper, which ARE present.  The kNN step is therefore the real thing (``knn`` below calls scikit-learn exactly as
pydiffmap's Kernel.fit/compute do); the steps after it restate pydiffmap 0.2.0.1's published algorithm:
Gaussian kernel exp(-d^2/(4 eps)) on the kneighbors_graph distances, q = row sums, K <- K diag(q^-alpha),
P = diag(1/rowsum) K, L = (P - I)/eps, eigs(L, k=n_evecs+1, which='LR'), drop the trivial pair,
coordinates = evecs * sqrt(-1/evals).  **Parity of that second half against pydiffmap is UNPINNED.**

Only tests/ may import this module.
"""
import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spsl
from sklearn.neighbors import NearestNeighbors


def knn(X, k):
    nn = NearestNeighbors(n_neighbors=k, metric="euclidean").fit(X)
    dist, idx = nn.kneighbors(X)
    return idx, dist


def diffusion_embedding(EMD, n_evecs=2, epsilon=1.0, alpha=0.5, k=64):
    S = EMD.shape[0]
    k0 = min(k, S)
    nn = NearestNeighbors(n_neighbors=k0, metric="euclidean").fit(EMD)
    A = nn.kneighbors_graph(EMD, mode="distance").tocsr()
    K = A.copy()
    K.data = np.exp(-K.data ** 2 / (4.0 * epsilon))
    # kneighbors_graph stores the self-distance 0 explicitly only for the query == training case: make sure the
    # diagonal carries exp(0) = 1 like every other neighbour
    K = K.tolil()
    K.setdiag(1.0)
    K = K.tocsr()
    q = np.asarray(K.sum(axis=1)).ravel()
    K = K @ sps.diags(np.power(q, -alpha))
    P = sps.diags(1.0 / np.asarray(K.sum(axis=1)).ravel()) @ K
    L = (P - sps.eye(S)) / epsilon
    w, v = spsl.eigs(L.tocsr(), k=n_evecs + 1, which="LR", v0=np.ones(S))
    ix = np.argsort(w.real)[::-1][1:n_evecs + 1]
    return v.real[:, ix] @ np.diag(np.sqrt(-1.0 / w.real[ix]))
