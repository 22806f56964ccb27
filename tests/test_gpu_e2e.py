"""End-to-end drop-in parity of pilot_b200.tl.wasserstein_distance on a duck-typed AnnData:
the committed golden fixtures (reference outputs) and the oracle on BASELINE config C1."""
import os

import numpy as np
import pandas as pd
import pytest

from conftest import build_golden_adata, golden_kwargs, golden_names, load_golden
from oracle import pilot_oracle as po
from pilot_b200 import synth, tl

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names())
def test_golden_fixture(name):
    g = load_golden(name)
    kw = golden_kwargs(g["case"])
    adata = build_golden_adata(g["case"])
    tl.wasserstein_distance(adata, **kw)
    u = adata.uns
    assert set(u) >= {"data", "annot", "proportions", "cost", "EMD_df", "EMD", "real_labels"}
    assert isinstance(u["proportions"], dict)
    assert [str(k) for k in u["proportions"].keys()] == list(g["samples"])
    P = np.stack(list(u["proportions"].values()))
    assert np.array_equal(P, g["props"]), "proportions must be bit-exact"
    assert [str(c) for c in u["cost"].columns] == list(g["cells"])
    assert u["cost"].index.name == "cell_types"
    np.testing.assert_allclose(u["cost"].to_numpy(), g["cost"], rtol=1e-12, atol=1e-15)
    assert isinstance(u["EMD"], np.ndarray) and u["EMD"].dtype == np.float64 and u["EMD"].flags.c_contiguous
    np.testing.assert_allclose(u["EMD"], g["EMD"], rtol=1e-9, atol=1e-14)
    assert u["EMD_df"].index.name == "sampleID"
    np.testing.assert_allclose(u["EMD_df"].to_numpy(), g["EMD_df"], rtol=1e-9, atol=1e-14)
    assert [str(x) for x in u["real_labels"]] == list(g["real_labels"])
    assert list(u["annot"].columns) == ["cell_type", "sampleID", "status"]
    assert len(u["data"]) == len(u["annot"])


def test_kidney_igan_real_data_golden():
    """The reference's own test case (test/test_pilot.py:7-15) on REAL data: Kidney_IgAN_G.h5ad, 24 227 glomeruli x
    14 features, 634 biopsies, data_type='Pathomics'.  Inputs and the reference's outputs were stored by
    tests/golden/make_kidney_golden.py (reference functions exec'd verbatim on the file read by pilot_b200.h5ad)."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "g6_kidney_igan_G.npz"),
                allow_pickle=True)
    obs = pd.DataFrame({
        "Cell_type": g["cell_type"],
        "sampleID": pd.Categorical.from_codes(g["sample_codes"], categories=list(g["sample_categories"])),
        "status": pd.Categorical.from_codes(g["status_codes"], categories=list(g["status_categories"])),
    })
    adata = synth.FakeAnnData(obs, X=g["X"], var_names=list(g["var_names"]))
    tl.wasserstein_distance(adata, clusters_col="Cell_type", sample_col="sampleID", status="status",
                            data_type="Pathomics")
    u = adata.uns
    assert [str(k) for k in u["proportions"].keys()] == list(g["samples"])
    P = np.stack(list(u["proportions"].values()))
    assert np.array_equal(P, g["props"]), "proportions must be bit-exact"
    assert list(u["cost"].columns) == list(g["cells"])
    np.testing.assert_allclose(u["cost"].to_numpy(), g["cost"], rtol=1e-12, atol=1e-15)
    E = u["EMD"]
    assert E.shape == (634, 634)
    np.testing.assert_allclose(E[g["EMD_rows"]], g["EMD_sub"], rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(E.sum(axis=0), g["EMD_colsum"], rtol=1e-10)
    np.testing.assert_allclose(np.diag(E), g["EMD_diag"], atol=1e-14)
    np.testing.assert_allclose(u["EMD_df"].to_numpy(), E.T, rtol=0, atol=0)
    assert [str(x) for x in u["real_labels"]] == list(g["real_labels"])


def test_categorical_columns_with_unused_categories():
    """A cohort subset in scanpy keeps the categories of the full one: samples and cell types that no longer occur
    must vanish from every output exactly as `.unique()` makes them vanish in the reference (Trajectory.py:402-425)."""
    X, obs = synth.make_cells(60_000, 12, 9, 14, seed=8, labels="categorical")
    keep = ~obs["sampleID"].isin(obs["sampleID"].cat.categories[[2, 9]]) & (obs["cell_types"] != obs["cell_types"].cat.categories[4])
    obs2 = obs[keep.to_numpy()].copy()                    # categories untouched: 14 samples / 9 types declared
    assert len(obs2["sampleID"].cat.categories) == 14 and obs2["sampleID"].nunique() == 12
    adata = synth.FakeAnnData(obs2, obsm={"X_PCA": np.ascontiguousarray(X[keep.to_numpy()])})
    tl.wasserstein_distance(adata, emb_matrix="X_PCA", clusters_col="cell_types", sample_col="sampleID", status="status")
    annot, data = adata.uns["annot"], adata.uns["data"]
    wprops = po.cluster_representations(annot)
    assert len(wprops) == 12 and list(wprops.keys()) == list(adata.uns["proportions"].keys())
    for k in wprops:
        assert wprops[k].shape == (8,) and np.array_equal(wprops[k], adata.uns["proportions"][k])
    wdis, wcost = po.cost_matrix(annot, data, "cosine")
    assert adata.uns["cost"].shape == (8, 8)
    np.testing.assert_allclose(adata.uns["cost"].to_numpy(), wdis, rtol=1e-12, atol=1e-15)
    wEMD, _ = po.wasserstein_d(wprops, wdis / wdis.max())
    np.testing.assert_allclose(adata.uns["EMD"], wEMD, rtol=1e-9, atol=1e-15)
    assert adata.uns["real_labels"] == po.return_real_labels(annot)


@pytest.mark.parametrize("labels", ["str", "categorical"])
def test_config_c1_full(labels):
    """BASELINE configs[0]: 200K cells, 30-dim, 10 types, 20 samples, cosine, exact EMD."""
    adata = synth.make_adata("c1", labels=labels)
    tl.wasserstein_distance(adata, emb_matrix="X_PCA", clusters_col="cell_types", sample_col="sampleID",
                            status="status")
    annot, data = adata.uns["annot"], adata.uns["data"]
    wprops = po.cluster_representations(annot)
    assert list(wprops.keys()) == list(adata.uns["proportions"].keys())
    for k in wprops:
        assert np.array_equal(wprops[k], adata.uns["proportions"][k])
    wdis, wcost = po.cost_matrix(annot, data, "cosine")
    np.testing.assert_allclose(adata.uns["cost"].to_numpy(), wdis, rtol=1e-12, atol=1e-15)
    pd.testing.assert_index_equal(adata.uns["cost"].index, wcost.index)
    wEMD, wdf = po.wasserstein_d(wprops, wdis / wdis.max())
    np.testing.assert_allclose(adata.uns["EMD"], wEMD, rtol=1e-9, atol=1e-15)
    pd.testing.assert_index_equal(adata.uns["EMD_df"].index, wdf.index)
    assert adata.uns["real_labels"] == po.return_real_labels(annot)


def test_sinkhorn_e2e_and_regularized_string_semantics():
    adata = synth.make_adata("c1", scale=0.1)
    # anything but the string "unreg" selects Sinkhorn (Trajectory.py:507)
    tl.wasserstein_distance(adata, regularized=True, reg=0.1)
    annot, data = adata.uns["annot"], adata.uns["data"]
    props = po.cluster_representations(annot)
    dis, _ = po.cost_matrix(annot, data, "cosine")
    want, wdf = po.wasserstein_d(props, dis / dis.max(), regularized="reg", reg=0.1)
    np.testing.assert_allclose(adata.uns["EMD"], want, rtol=1e-9)
    np.testing.assert_allclose(adata.uns["EMD_df"].to_numpy(), want.T, rtol=1e-9)
    assert np.abs(np.diag(adata.uns["EMD"])).min() > 0


def test_unnormalised_counts_raise_like_pot():
    adata = synth.make_adata("c1", scale=0.05)
    with pytest.raises(AssertionError):
        tl.wasserstein_distance(adata, normalization=False)


def test_wasserstein_d_public_function():
    P, M = synth.make_pairs(15, 10, seed=8)
    rep = {f"s{i}": P[i] for i in range(15)}
    EMD, df = tl.wasserstein_d(rep, M)
    want, wdf = po.wasserstein_d(rep, M)
    np.testing.assert_allclose(EMD, want, rtol=1e-9, atol=1e-15)
    pd.testing.assert_frame_equal(df, wdf, rtol=1e-9, atol=1e-15)
