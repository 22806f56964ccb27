"""The oracle restatement vs the committed golden fixtures (generated from the reference's own
code by tests/golden/make_golden.py) and, when /root/reference is mounted, vs the reference live."""
import numpy as np
import pandas as pd
import pytest

from conftest import build_golden_adata, golden_kwargs, golden_names, load_golden
from oracle import pilot_oracle as po
from oracle import ref_exec
from pilot_b200 import synth


def _annot_data(adata, data_type):
    if data_type == "scRNA":
        data = pd.DataFrame(adata.obsm["X_PCA"])
    else:
        data = pd.DataFrame(adata.X)
    annot = adata.obs[["cell_types", "sampleID", "status"]].copy()
    annot.columns = ["cell_type", "sampleID", "status"]
    return data.reset_index(drop=True), annot.reset_index(drop=True)


@pytest.mark.parametrize("name", golden_names())
def test_restatement_matches_golden(name):
    g = load_golden(name)
    case = g["case"]
    kw = golden_kwargs(case)
    adata = build_golden_adata(case)
    data, annot = _annot_data(adata, kw["data_type"])
    props = po.cluster_representations(annot, regulizer=kw["regulizer"])
    assert [str(k) for k in props.keys()] == list(g["samples"])
    P = np.stack(list(props.values()))
    assert P.dtype == np.float64
    assert np.array_equal(P, g["props"]), "proportions must be bit-exact"
    dis, cost_df = po.cost_matrix(annot, data, metric=kw["metric"])
    assert [str(c) for c in cost_df.columns] == list(g["cells"])
    np.testing.assert_allclose(dis, g["cost"], rtol=1e-12, atol=1e-15)
    EMD, emd_df = po.wasserstein_d(props, dis / dis.max(), regularized=kw["regularized"], reg=kw["reg"])
    np.testing.assert_allclose(EMD, g["EMD"], rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(emd_df.to_numpy(), g["EMD_df"], rtol=1e-9, atol=1e-14)
    assert np.array_equal(g["EMD_df"], g["EMD"].T)
    assert [str(x) for x in po.return_real_labels(annot)] == list(g["real_labels"])


@pytest.mark.skipif(not ref_exec.available(), reason="/root/reference not mounted (GPU box)")
@pytest.mark.parametrize("labels", ["str", "categorical", "int"])
@pytest.mark.parametrize("seed", [21, 22, 23])
def test_restatement_matches_reference_live(labels, seed):
    ref = ref_exec.load()
    X, obs = synth.make_cells(1500 + 7 * seed, 6, 5 + seed % 4, 6 + seed % 5, seed, labels=labels)
    annot = obs.copy()
    annot.columns = ["cell_type", "sampleID", "status"]
    data = pd.DataFrame(X)
    for regulizer, norm in ((0.2, True), (0.7, True), (0.2, False)):
        a = ref.Cluster_Representations(annot, regulizer=regulizer, normalization=norm)
        b = po.cluster_representations(annot, regulizer=regulizer, normalization=norm)
        assert list(a.keys()) == list(b.keys())
        for k in a:
            assert np.array_equal(np.asarray(a[k], dtype=np.float64), b[k])
    for metric in ("cosine", "euclidean"):
        d1, c1 = ref.cost_matrix(annot, data, metric=metric)
        d2, c2 = po.cost_matrix(annot, data, metric=metric)
        np.testing.assert_allclose(d2, d1, rtol=1e-12, atol=1e-15)
        assert list(c1.columns) == list(c2.columns)
    assert ref.return_real_labels(annot) == po.return_real_labels(annot)


def test_median_is_in_input_dtype():
    X, obs = synth.make_cells(2001, 4, 3, 4, 5, dtype=np.float32)
    annot = obs.rename(columns={"cell_types": "cell_type"})
    cent = po.centroid_medians(annot, X)
    assert cent.dtype == np.float32
