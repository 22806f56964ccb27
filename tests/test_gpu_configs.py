"""Parity at the BASELINE configuration shapes: C2 end to end, C4- and C5-shaped
pair problems on row slices of the full pair space, entry-wise against the oracle."""
import numpy as np
import pytest
import torch

from oracle import pilot_oracle as po
from pilot_b200 import _lib, ops, pairs, synth, tl

pytestmark = pytest.mark.gpu


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def test_config_c2_end_to_end():
    """configs[1]: 1M cells, 50-dim f32, 30 types, 100 samples, Sinkhorn reg 0.1 -- full drop-in call;
    proportions bit-exact, cost <= 1e-12, 16 rows of the ordered pair matrix <= 1e-9 vs the oracle."""
    adata = synth.make_adata("c2", labels="categorical")
    tl.wasserstein_distance(adata, regularized="reg", reg=0.1)
    annot, data = adata.uns["annot"], adata.uns["data"]
    props = po.cluster_representations(annot)
    assert list(props.keys()) == list(adata.uns["proportions"].keys())
    for k in props:
        assert np.array_equal(props[k], adata.uns["proportions"][k])
    dis, _ = po.cost_matrix(annot, data, "cosine")
    np.testing.assert_allclose(adata.uns["cost"].to_numpy(), dis, rtol=1e-12, atol=1e-15)
    P = np.stack(list(props.values()))
    want, _, _ = po.sinkhorn_rows(P, dis / dis.max(), 0.1, 0, 16)
    np.testing.assert_allclose(adata.uns["EMD"][:16], want, rtol=1e-9)
    assert adata.uns["EMD"].shape == (100, 100) and len(adata.uns["real_labels"]) == 100
    assert adata.uns["real_labels"] == po.return_real_labels(annot)


def _full_path_check(config, regularized, reg, n_entries, rng_seed):
    """wasserstein_distance on a whole BASELINE cohort: proportions bit-exact, cost <= 1e-12, random entries of
    the first 16 rows of the S x S matrix <= 1e-9 against the oracle; Sinkhorn also iteration/absorption counts."""
    adata = synth.make_adata(config, labels="categorical")
    tl.wasserstein_distance(adata, regularized=regularized, reg=reg)
    annot, data = adata.uns["annot"], adata.uns["data"]
    props = po.cluster_representations(annot)
    assert list(props.keys()) == list(adata.uns["proportions"].keys())
    for k in props:
        assert np.array_equal(props[k], adata.uns["proportions"][k]), "proportions must be bit-exact"
    dis, _ = po.cost_matrix(annot, data, "cosine")
    np.testing.assert_allclose(adata.uns["cost"].to_numpy(), dis, rtol=1e-12, atol=1e-15)
    P = np.stack(list(props.values()))
    M = dis / dis.max()
    S = P.shape[0]
    EMD = adata.uns["EMD"]
    assert EMD.shape == (S, S) and EMD.flags.c_contiguous and len(adata.uns["real_labels"]) == S
    np.testing.assert_array_equal(adata.uns["EMD_df"].to_numpy(), EMD.T)
    r = np.random.default_rng(rng_seed)
    rows = 16
    picks = r.choice(rows * S, n_entries, replace=False)
    if regularized == "unreg":
        for g in picks:
            i, j = divmod(int(g), S)
            w = po.emd2(P[i], P[j], M)
            assert abs(EMD[i, j] - w) <= 1e-9 * max(w, 1e-300) + 1e-15, (i, j)
        assert np.array_equal(EMD, EMD.T) and np.abs(np.diag(EMD)).max() == 0.0
    else:
        out, iters, absn, status = ops.sinkhorn_pairs(dev(P), dev(M), reg, ops.make_range(rows * S, _lib.PAIRS_FULL),
                                                      want_info=True)
        out, iters, absn = out.cpu().numpy(), iters.cpu().numpy(), absn.cpu().numpy()
        for g in picks:
            i, j = divmod(int(g), S)
            c, info = po.sinkhorn2(P[i], P[j], M, reg, return_info=True)
            assert iters[g] == info["iters"] and absn[g] == info["absorptions"], (i, j)
            assert abs(EMD[i, j] - c) <= 1e-9 * abs(c), (i, j)
            assert abs(out[g] - EMD[i, j]) <= 1e-12 * abs(c)
    assert adata.uns["real_labels"] == po.return_real_labels(annot)


@pytest.mark.parametrize("regularized,reg", [("unreg", 0.1), ("reg", 0.1)])
def test_config_c3_full_path(regularized, reg):
    """configs[2], paper upper scale: 5M cells x 50 dims, 40 cell types, 600 samples, exact EMD and Sinkhorn
    (the KP = 48 / KC = 40 panel tiles)."""
    _full_path_check("c3", regularized, reg, 1500, 3)


def test_config_c4_full_path():
    """configs[3], pathomics shape: 2M structures, 64 clusters, 2 000 samples, stabilised Sinkhorn reg 0.01
    (4e6 problems running to the iteration cap with ~30 absorptions each)."""
    _full_path_check("c4", "reg", 0.01, 120, 4)


def test_config_c4_shape_pairs():
    """C4 shape: 2 000 samples x 64 clusters, stabilised Sinkhorn reg 0.01 (iteration cap, ~30 absorptions):
    the first 6 rows of the ordered pair space (12 000 problems), 96 of them checked against the oracle
    including iteration and absorption counts."""
    S, K = 2000, 64
    P, M = synth.make_pairs(S, K, seed=4)
    rows = 6
    rng = ops.make_range(rows * S, _lib.PAIRS_FULL)
    out, iters, absn, status = ops.sinkhorn_pairs(dev(P), dev(M), 0.01, rng, want_info=True)
    out, iters, absn = out.cpu().numpy(), iters.cpu().numpy(), absn.cpu().numpy()
    r = np.random.default_rng(0)
    for g in r.choice(rows * S, 96, replace=False):
        i, j = divmod(int(g), S)
        c, info = po.sinkhorn2(P[i], P[j], M, 0.01, return_info=True)
        assert iters[g] == info["iters"] and absn[g] == info["absorptions"], (i, j)
        assert abs(out[g] - c) <= 1e-9 * abs(c)


def test_config_c5_shape_pairs():
    """C5 shape: 20 000 samples x 64 types.  A 48-row slice of the exact-EMD upper triangle and a 24-row
    slice of the ordered Sinkhorn matrix; 2 000 / 300 random entries against the oracle; EMD <= Sinkhorn."""
    S, K = 20000, 64
    P, M = synth.make_pairs(S, K, seed=5)
    Pd, Md = dev(P), dev(M)
    n_emd = 48 * S - 48 * 49 // 2
    emd = ops.emd_pairs(Pd, Md, ops.make_range(n_emd, _lib.PAIRS_UPPER)).cpu().numpy()
    r = np.random.default_rng(1)
    for g in r.choice(n_emd, 2000, replace=False):
        i, j = pairs.global_to_ij(int(g), S, _lib.PAIRS_UPPER)
        w = po.emd2(P[i], P[j], M)
        assert abs(emd[g] - w) <= 1e-9 * w, (i, j)
    n_sk = 24 * S
    sk, iters, _, _ = ops.sinkhorn_pairs(Pd, Md, 0.1, ops.make_range(n_sk, _lib.PAIRS_FULL), want_info=True)
    sk, iters = sk.cpu().numpy(), iters.cpu().numpy()
    for g in r.choice(n_sk, 300, replace=False):
        i, j = divmod(int(g), S)
        c, info = po.sinkhorn2(P[i], P[j], M, 0.1, return_info=True)
        assert iters[g] == info["iters"]
        assert abs(sk[g] - c) <= 1e-9 * c
    # size-independent property on the shared part of both slices: entropic cost >= exact cost
    for i in range(24):
        row_emd = np.array([emd[pairs_index(i, j, S)] for j in range(i + 1, i + 200)])
        row_sk = sk[i * S + i + 1: i * S + i + 200]
        assert (row_sk >= row_emd * (1 - 1e-9)).all()


def pairs_index(i, j, S):
    return i * (2 * S - i - 1) // 2 + (j - i - 1)
