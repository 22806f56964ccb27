"""(f)-1: row kNN on the distance matrix and the diffusion-map embedding of pilotpy.pl.trajectory
(ploting.py:95-110).  The kNN kernel is checked against scikit-learn's NearestNeighbors -- the routine pydiffmap
calls; the embedding against the SciPy restatement in oracle/diffmap_oracle.py (up to the sign of each eigenvector;
parity against pydiffmap itself is unpinned: it is not installable here)."""
import numpy as np
import pytest
import torch

from oracle import diffmap_oracle as do
from pilot_b200 import ops, pairs, pl, synth

pytestmark = pytest.mark.gpu


def emd_matrix(S, K, seed):
    P, M = synth.make_pairs(S, K, seed=seed)
    E = pairs.all_pairs(torch.from_numpy(P).cuda(), torch.from_numpy(M).cuda(), "unreg").cpu().numpy()
    return E / E.max()


@pytest.mark.parametrize("S,k", [(50, 50), (300, 64), (1000, 64), (777, 10), (64, 1), (2500, 128)])
def test_knn_rows_match_sklearn(S, k):
    X = emd_matrix(S, 12, 100 + S)
    idx, dist = pl.knn_graph(X, k)
    widx, wdist = do.knn(X, k)
    assert idx.shape == (S, k) and dist.shape == (S, k)
    assert (idx[:, 0] == np.arange(S)).all() and (dist[:, 0] == 0).all()          # the point itself
    # the Gram-matrix form loses ~1e-16 |x|^2 absolutely in d^2, i.e. ~1e-6 in d near 0 (scikit-learn uses the same
    # form and, not knowing that query == training set, reports self-distances of that size instead of 0)
    np.testing.assert_allclose(dist, wdist, rtol=1e-9, atol=3e-6)
    same = (idx == widx).mean()
    assert same > 0.999, same                                                     # ties / 1-ulp swaps only
    assert (np.diff(dist, axis=1) >= 0).all()
    for i in range(0, S, max(1, S // 20)):                                        # same neighbour SETS
        assert len(set(idx[i]) ^ set(widx[i])) <= 2


def test_knn_large_unstaged_rows():
    """S > 24 576 rows do not fit the shared-memory staging: the kernel re-reads the Gram row."""
    rng = np.random.default_rng(0)
    S = 26_000
    X = rng.random((S, 24))
    G = torch.from_numpy(X).cuda()
    # knn_rows takes an S x S' point matrix in general: here 24-dimensional points
    idx, dist = ops.knn_rows(G, 16)
    widx, wdist = do.knn(X, 16)
    np.testing.assert_allclose(dist.cpu().numpy(), wdist, rtol=1e-9, atol=3e-6)
    assert (idx.cpu().numpy() == widx).mean() > 0.999


@pytest.mark.parametrize("S", [120, 900])
def test_diffusion_embedding_matches_restatement(S):
    X = emd_matrix(S, 10, 7 + S)
    got = pl.diffusion_embedding(X, n_evecs=2, epsilon=1, alpha=0.5, knn=64)
    want = do.diffusion_embedding(X, n_evecs=2, epsilon=1.0, alpha=0.5, k=64)
    assert got.shape == (S, 2)
    for c in range(2):
        s = np.sign(np.dot(got[:, c], want[:, c]))
        np.testing.assert_allclose(got[:, c], s * want[:, c], rtol=1e-6, atol=1e-8 * np.abs(want[:, c]).max())


def test_trajectory_writes_embedding(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    from pilot_b200 import tl
    adata = synth.make_adata("c1", scale=0.2)
    tl.wasserstein_distance(adata)
    pl.trajectory(adata)
    emb = adata.uns["embedding"]
    assert emb.shape == (20, 2) and np.isfinite(emb).all()
    assert (tmp_path / "Results_PILOT" / "plots").exists()
