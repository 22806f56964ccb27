"""Pin the OT part of the oracle: exact EMD against SciPy-HiGHS and closed forms, the C Sinkhorn
against its NumPy twin and the algorithm's invariants."""
import numpy as np
import pytest
import scipy.spatial.distance as ssd
from scipy.optimize import linprog

from oracle import pilot_oracle as po
from pilot_b200 import synth


def highs_emd(a, b, M):
    n, m = M.shape
    A = np.zeros((n + m, n * m))
    for i in range(n):
        A[i, i * m:(i + 1) * m] = 1
    for j in range(m):
        A[n + j, j::m] = 1
    r = linprog(M.ravel(), A_eq=A[:-1], b_eq=np.concatenate([a, b])[:-1], bounds=(0, None), method="highs",
                options=dict(primal_feasibility_tolerance=1e-10, dual_feasibility_tolerance=1e-10))
    assert r.status == 0
    return r.fun


@pytest.mark.parametrize("K", [2, 3, 10, 30, 64])
def test_emd_matches_highs(K):
    P, M = synth.make_pairs(8, K, seed=100 + K)
    for i in range(0, 8, 2):
        a, b = P[i], P[i + 1]
        v, code, st = po.emd2(a, b, M, return_stats=True)
        assert code == 1
        ref = highs_emd(a, b * a.sum() / b.sum(), M)
        assert abs(v - ref) <= 1e-11 * abs(ref) + 1e-15
        assert st["pivots"] > 0 and st["arcs_priced"] >= st["pivots"]


def test_emd_known_answers():
    rng = np.random.default_rng(0)
    P, M = synth.make_pairs(4, 12, seed=7)
    # identical histograms, zero-diagonal cost -> exactly 0
    assert po.emd2(P[0], P[0], M) == 0.0
    # K = 2 closed form
    M2 = np.array([[0.0, 0.7], [0.7, 0.0]])
    a, b = np.array([0.3, 0.7]), np.array([0.55, 0.45])
    assert abs(po.emd2(a, b, M2) - abs(a[0] - b[0]) * 0.7) < 1e-15
    # 1-D chain cost |i-j|: EMD = sum |CDF_a - CDF_b|
    K = 17
    Mc = np.abs(np.subtract.outer(np.arange(K), np.arange(K))).astype(float)
    a = rng.dirichlet(np.ones(K)); b = rng.dirichlet(np.ones(K))
    b = b * a.sum() / b.sum()
    assert abs(po.emd2(a, b, Mc) - np.abs(np.cumsum(a) - np.cumsum(b))[:-1].sum()) < 1e-12
    # zero-mass entries are dropped, not fatal
    a0 = a.copy(); a0[3] = 0; a0 /= a0.sum()
    assert np.isfinite(po.emd2(a0, b / b.sum(), Mc))
    # mass mismatch raises like ot.emd2 (check_marginals)
    with pytest.raises(AssertionError):
        po.emd2(a, 2 * b, Mc)
    # symmetric cost -> symmetric distance
    assert abs(po.emd2(P[1], P[2], M) - po.emd2(P[2], P[1], M)) < 1e-15


@pytest.mark.parametrize("K,reg", [(10, 0.1), (30, 0.1), (64, 0.1), (40, 0.01)])
def test_sinkhorn_c_matches_numpy(K, reg):
    P, M = synth.make_pairs(6, K, seed=200 + K)
    for i, j in ((0, 1), (2, 3), (4, 4), (5, 0)):
        G, info = po.sinkhorn_stabilized_np(P[i], P[j], M, reg, return_info=True)
        c, info_c = po.sinkhorn2(P[i], P[j], M, reg, return_info=True)
        ref = float((M * G).sum())
        assert abs(c - ref) <= 1e-12 * abs(ref) + 1e-300
        assert info["iters"] == info_c["iters"] and info["absorptions"] == info_c["absorptions"]
        assert info["status"] == info_c["status"]
        # u is updated last -> row marginals exact; converged -> column marginals within stopThr
        np.testing.assert_allclose(G.sum(axis=1), P[i], rtol=1e-12)
        if info["status"] == 0:
            assert info["iters"] % 20 == 1
            assert np.linalg.norm(G.sum(axis=0) - P[j]) <= 1e-9
        else:
            assert info["iters"] == 1000


def test_sinkhorn_tends_to_emd_and_is_not_symmetric():
    P, M = synth.make_pairs(3, 10, seed=3)
    e = po.emd2(P[0], P[1], M)
    s1, s2 = po.sinkhorn2(P[0], P[1], M, 0.1), po.sinkhorn2(P[0], P[1], M, 0.02)
    assert s1 >= s2 >= e * (1 - 1e-6)
    assert po.sinkhorn2(P[0], P[0], M, 0.1) > 0  # S(a, a) != 0


def test_wasserstein_d_containers():
    P, M = synth.make_pairs(5, 6, seed=9)
    rep = {f"s{i}": P[i] for i in range(5)}
    EMD, df = po.wasserstein_d(rep, M, "reg", 0.1)
    assert EMD.shape == (5, 5) and df.index.name == "sampleID"
    assert np.array_equal(df.to_numpy(), EMD.T)
    assert list(df.columns) == list(rep.keys())
