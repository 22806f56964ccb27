"""(f)-2: silhouette score of the samples on the distance matrix -- Sil_computing (Trajectory.py:593-612) =
sklearn.metrics.silhouette_score(EMD, labels, metric='cosine').  scikit-learn is importable here, so the kernel is
pinned against the very routine the reference calls."""
import numpy as np
import pytest
import torch
from sklearn import metrics

from pilot_b200 import ops, pairs, synth, tl

pytestmark = pytest.mark.gpu


def emd_matrix(S, K, seed):
    P, M = synth.make_pairs(S, K, seed=seed)
    E = pairs.all_pairs(torch.from_numpy(P).cuda(), torch.from_numpy(M).cuda(), "unreg").cpu().numpy()
    return E / E.max()


@pytest.mark.parametrize("metric", ["cosine", "euclidean", "precomputed"])
@pytest.mark.parametrize("S,L", [(40, 2), (300, 3), (1000, 7), (777, 120)])
def test_silhouette_matches_sklearn(S, L, metric):
    E = emd_matrix(S, 12, 7 + S)
    rng = np.random.default_rng(S + L)
    labels = rng.integers(0, L, size=S)
    labels[:L] = np.arange(L)
    if L == 7:
        labels[labels == 3] = 2          # a label value that never occurs (LabelEncoder compacts)
        labels[5] = 3                    # ... and a singleton cluster: its sample scores 0
    names = np.array(["case", "ctrl", "x"] + [f"c{i}" for i in range(200)])[labels]   # string labels, as real_labels
    want = metrics.silhouette_samples(E, names, metric=metric)
    got = ops.silhouette_rows(torch.from_numpy(E).cuda(), names, metric).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-11)
    score = tl.Sil_computing(E, list(names), metric=metric)
    assert abs(score - metrics.silhouette_score(E, names, metric=metric)) <= 1e-11


def test_silhouette_large_unstaged_and_errors():
    """S > 24 576 rows do not fit the shared-memory staging; label-count and diagonal checks are scikit-learn's."""
    rng = np.random.default_rng(1)
    S = 25_100
    D = rng.random((S, 24))
    X = torch.from_numpy(D).cuda()
    G = torch.mm(X, X.t())                                   # a valid "matrix whose rows are points"
    labels = rng.integers(0, 5, size=S)
    got = ops.silhouette_rows(G, labels, "precomputed")      # G is read as distances: any non-negative matrix will do
    Gh = G.cpu().numpy()
    pick = rng.integers(0, S, size=40)
    for i in pick:
        sums = np.bincount(labels, weights=Gh[i], minlength=5)
        cnt = np.bincount(labels, minlength=5).astype(float)
        a = sums[labels[i]] / (cnt[labels[i]] - 1)
        b = np.min(np.delete(sums / cnt, labels[i]))
        assert abs(got[i].item() - (b - a) / max(a, b)) <= 1e-10
    E = emd_matrix(30, 6, 3)
    with pytest.raises(ValueError, match="Number of labels"):
        tl.Sil_computing(E, np.zeros(30, dtype=int))
    with pytest.raises(ValueError, match="Number of labels"):
        tl.Sil_computing(E, np.arange(30))
    with pytest.raises(ValueError, match="non-zero elements on the diagonal"):
        tl.Sil_computing(E + np.eye(30), np.arange(30) % 3, metric="precomputed")
    with pytest.raises(ValueError):
        tl.Sil_computing(E, np.arange(30) % 3, metric="manhattan")
