"""Host-side glue of pilot_b200.tl (no GPU needed): factorisation helpers and the pandas
containers of the adata.uns contract."""
import numpy as np
import pandas as pd
import pytest

from oracle import pilot_oracle as po
from pilot_b200 import synth, tl


def first_appearance_perm(codes, n_codes):
    first = np.full(n_codes, len(codes), dtype=np.int64)
    np.minimum.at(first, codes, np.arange(len(codes)))
    p = np.argsort(first, kind="stable")
    return p[first[p] < len(codes)].astype(np.int32)


@pytest.mark.parametrize("labels", ["str", "categorical", "int"])
def test_unique_in_order_equals_pandas_unique(labels):
    X, obs = synth.make_cells(5000, 3, 9, 14, 3, labels=labels)
    if labels == "categorical":
        obs["cell_types"] = obs["cell_types"].cat.add_categories(["never_seen"])
    for name in ("cell_types", "sampleID"):
        col = obs[name]
        codes, lab = tl._raw_codes(col)
        assert codes.dtype in (np.int8, np.int16, np.int32) and codes.min() >= 0
        perm = first_appearance_perm(codes, len(lab))
        mine = tl._unique_in_order(col, lab, perm)
        ref = col.unique()
        assert type(mine) is type(ref)
        assert list(mine) == list(ref)
        f1 = pd.DataFrame(np.zeros((len(perm), len(perm))))
        f1.columns = mine
        f2 = pd.DataFrame(np.zeros((len(perm), len(perm))))
        f2.columns = ref
        pd.testing.assert_index_equal(f1.columns, f2.columns)


def test_missing_labels_rejected():
    with pytest.raises(ValueError):
        tl._raw_codes(pd.Series(["a", None, "b"], name="cell_type"))


def test_frames_match_reference_assembly():
    P, M = synth.make_pairs(5, 4, seed=1)
    ids = [f"p{i}" for i in range(5)]
    EMD = np.arange(25, dtype=float).reshape(5, 5)
    mine = tl._emd_frame(EMD, ids)
    ref = pd.DataFrame.from_dict(EMD).T
    ref.columns = ids
    ref["sampleID"] = ids
    ref = ref.set_index("sampleID")
    pd.testing.assert_frame_equal(mine, ref)
    cells = np.array(["a", "b", "c", "d"], dtype=object)
    c1 = tl._cost_frame(M, cells)
    assert c1.index.name == "cell_types" and list(c1.columns) == list(cells)
    np.testing.assert_array_equal(c1.to_numpy(), M)


def test_return_real_labels_matches_oracle():
    X, obs = synth.make_cells(3000, 3, 4, 11, 8)
    annot = obs.rename(columns={"cell_types": "cell_type"})
    assert tl.return_real_labels(annot) == po.return_real_labels(annot)


def test_emd_mass_check_like_pot():
    P = np.array([[0.5, 0.5], [0.6, 0.4 + 1e-5]])
    with pytest.raises(AssertionError):
        tl._check_emd_inputs(P, np.zeros((2, 2)))
    tl._check_emd_inputs(np.array([[0.5, 0.5], [0.25, 0.75]]), np.zeros((2, 2)))


def test_extract_mirrors_reference_columns(tmp_path):
    adata = synth.make_adata("c1", scale=0.001)
    data, annot = tl.extract_data_anno_scRNA_from_h5ad(adata, emb_matrix="X_PCA", clusters_col="cell_types",
                                                       sample_col="sampleID", status="status")
    assert list(annot.columns) == ["cell_type", "sampleID", "status"]
    assert list(data.columns)[:2] == ["PCA_1", "PCA_2"] and data.shape[1] == 30
    assert tl.path_to_results == "Results_PILOT/plots"
    import os
    assert os.path.isdir("Results_PILOT/plots")
    with pytest.raises(KeyError):
        tl.extract_data_anno_scRNA_from_h5ad(adata, emb_matrix="missing")


@pytest.mark.parametrize("kind", ["categorical", "object", "int", "str", "mixed"])
def test_labelled_frames_equal_the_reference_statements(kind):
    """_cost_frame / _emd_frame build the frame in one step; it must be what the reference's
    from_dict / columns / new column / set_index statements (Trajectory.py:470-473, 518-521) build."""
    import pandas as pd
    from pilot_b200 import tl
    K = 7
    rng = np.random.default_rng(0)
    mat = rng.random((K, K))
    names = [f"ct{i}" for i in range(K)]
    labels = {"categorical": pd.Categorical.from_codes(rng.permutation(K), categories=names),
              "object": np.array(names, dtype=object), "int": np.arange(K) * 3,
              "str": pd.array(names, dtype="str") if hasattr(pd, "StringDtype") else np.array(names, dtype=object),
              "mixed": np.array([1, "a", 2.5, "b", 3, "c", None], dtype=object)}[kind]
    for name, build in (("cell_types", tl._cost_frame), ("sampleID", tl._emd_frame)):
        lab = list(labels) if name == "sampleID" else labels
        want = tl._labelled_square_reference(mat, lab, name)
        got = build(mat, lab)
        pd.testing.assert_frame_equal(got, want, check_exact=True, check_index_type=True, check_column_type=True)
        assert type(got.index) is type(want.index) and type(got.columns) is type(want.columns)
        assert got.index.name == want.index.name and got.columns.name == want.columns.name
        np.testing.assert_array_equal(got.to_numpy(), mat.T)


def test_band_rows_cover_the_pair_space():
    from pilot_b200 import _lib, pairs
    for S in (1, 2, 3, 17, 100, 20000):
        for mode in (_lib.PAIRS_FULL, _lib.PAIRS_UPPER):
            total = S * S if mode == _lib.PAIRS_FULL else S * (S - 1) // 2
            assert pairs.row_start(S, S, mode) == total
            for nb in (1, 2, 7, 64, 1000):
                e = pairs.band_rows(S, mode, nb)
                assert e[0] == 0 and e[-1] == S and all(a < b for a, b in zip(e[:-1], e[1:]))
                sizes = [pairs.row_start(b, S, mode) - pairs.row_start(a, S, mode) for a, b in zip(e[:-1], e[1:])]
                assert sum(sizes) == total
                if S == 20000 and nb == 64:
                    assert max(sizes) <= 1.2 * total / nb  # balanced
    # a row's first problem maps back to (row, first column)
    for S in (5, 33):
        for i in range(S - 1):
            assert pairs.global_to_ij(pairs.row_start(i, S, _lib.PAIRS_UPPER), S, _lib.PAIRS_UPPER) == (i, i + 1)
            assert pairs.global_to_ij(pairs.row_start(i, S, _lib.PAIRS_FULL), S, _lib.PAIRS_FULL) == (i, 0)


def test_precomputed_distance_writes_the_uns_contract(tmp_path, monkeypatch):
    """Precomputed_distance (Trajectory.py:1687-1727) is host-only; data_type is a keyword here (the reference
    reads an undefined name, :1716)."""
    import pandas as pd
    from pilot_b200 import synth, tl
    monkeypatch.chdir(tmp_path)
    X, obs = synth.make_cells(500, 4, 3, 5, seed=1)
    adata = synth.FakeAnnData(obs, obsm={"X_PCA": X})
    D = np.arange(25.0).reshape(5, 5)
    cost = pd.DataFrame(np.eye(3))
    feats = {"a": np.ones(3)}
    tl.Precomputed_distance(adata, D, cost, feats)
    assert adata.uns["EMD"] is D and adata.uns["cost"] is cost and adata.uns["proportions"] is feats
    assert list(adata.uns["annot"].columns) == ["cell_type", "sampleID", "status"]
    assert len(adata.uns["real_labels"]) == 5 and len(adata.uns["data"]) == 500


def test_object_column_factorize_fast_path_equals_pandas():
    """The pointer-first factorisation of object label columns returns exactly pd.factorize(sort=False)."""
    import pandas as pd
    from pilot_b200 import tl
    rng = np.random.default_rng(0)
    n = 120_000
    names = np.array([f"type{k}" for k in range(17)], dtype=object)
    shared = pd.Series(names[rng.integers(0, 17, n)], dtype=object)                                  # 17 shared objects
    dup = pd.Series(np.array([f"s{v}" for v in rng.integers(0, 40, n)], dtype=object), dtype=object)  # every cell its own object
    mixed = shared.copy()
    mixed.iloc[::40] = [f"type{k}" for k in rng.integers(0, 17, len(mixed.iloc[::40]))]   # equal values, other objects
    ints = pd.Series(rng.integers(0, 9, n).astype(object), dtype=object)
    for col in (shared, mixed, ints):
        got = tl._factorize_object_column(col)
        assert got is not None
        want_codes, want_labels = pd.factorize(col, sort=False)
        assert np.array_equal(got[0], want_codes) and list(got[1]) == list(want_labels)
    got = tl._factorize_object_column(dup)                     # every cell its own object: the C helper, or None
    if tl._host_helper() is None:
        assert got is None                                     # (pointer-first has no gain: pd.factorize is used)
    else:
        want_codes, want_labels = pd.factorize(dup, sort=False)
        assert np.array_equal(got[0], want_codes) and list(got[1]) == list(want_labels)
    withnan = shared.copy()
    withnan.iloc[5] = np.nan
    got = tl._factorize_object_column(withnan)
    want_codes, want_labels = pd.factorize(withnan, sort=False)
    assert np.array_equal(got[0], want_codes) and list(got[1]) == list(want_labels) and got[0][5] == -1
    codes, labels = tl._raw_codes(shared.rename("cell_type"))
    assert codes.dtype == np.int32 and list(labels) == list(pd.factorize(shared, sort=False)[1])


def test_factorisation_cache_hits_only_on_identical_object_columns():
    """Object label columns are factorised once; a later call with the same objects in every cell reuses the
    result, any changed cell misses (tl._factor_cache_get validates every pointer)."""
    rng = np.random.default_rng(5)
    n = 60_000
    base = pd.Series(rng.integers(0, 25, n)).astype(str).astype(object)     # every cell its own str object
    df = pd.DataFrame({"a": base})
    tl._factor_cache.clear()

    def codes_of(frame):
        annot = frame[["a"]].reset_index(drop=True)
        return tl._raw_codes(annot["a"])

    c0, l0 = codes_of(df)
    assert len(tl._factor_cache) == 1
    c1, l1 = codes_of(df)                                                      # same objects: served from the cache
    assert c1 is c0 and list(l1) == list(l0)
    wc, wl = pd.factorize(df["a"], sort=False)
    assert np.array_equal(c0, wc) and list(l0) == list(wl)
    df.loc[n // 3, "a"] = "other"                                              # one cell changes: exact validation misses
    c2, l2 = codes_of(df)
    wc, wl = pd.factorize(df["a"], sort=False)
    assert np.array_equal(c2, wc) and list(l2) == list(wl) and "other" in list(l2)
    for i in range(6):                                                         # bounded
        codes_of(pd.DataFrame({"a": pd.Series(rng.integers(0, 5, n)).astype(str).astype(object)}))
    assert len(tl._factor_cache) <= tl._FACTOR_CACHE_MAX


def test_native_str_factorisation_equals_pandas():
    """csrc/host_ingest.c (libpilot_host.so): codes and labels in order of first appearance like
    pd.factorize(sort=False); ASCII, non-ASCII, empty strings, equal values in different objects, runs of one
    object; non-str cells make it step aside for pandas."""
    if tl._host_helper() is None:
        pytest.skip("libpilot_host.so not built")
    rng = np.random.default_rng(9)
    n = 70_000
    pools = [np.array([f"type {i}" for i in range(37)], dtype=object),
             np.array(["α-cell", "β", "naïve T", "", "x", "naive T", "日本"], dtype=object)]
    for pool in pools:
        shared = pool[rng.integers(0, len(pool), n)]                       # few objects, many references
        distinct = pd.Series(shared).astype(str).astype(object).to_numpy()  # every cell its own object
        runs = np.repeat(pool, n // len(pool) + 1)[:n]                      # long runs of one pointer
        for vals in (shared, distinct, runs):
            got = tl._factorize_str_native(np.ascontiguousarray(vals))
            assert got is not None
            wc, wl = pd.factorize(vals, sort=False)
            assert np.array_equal(got[0], wc) and list(got[1]) == list(wl)
    mixed = pools[0][rng.integers(0, 37, n)].copy()
    mixed[n // 2] = None
    assert tl._factorize_str_native(mixed) is None
    mixed[n // 2] = 3.5
    assert tl._factorize_str_native(mixed) is None
