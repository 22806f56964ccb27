"""GPU parity of stage 3 (all-pairs exact EMD and stabilised Sinkhorn) against the oracle, the
packed/partitioned layout, and size-independent properties at larger sizes."""
import numpy as np
import pytest
import torch

from oracle import pilot_oracle as po
from pilot_b200 import _lib, ops, pairs, synth

pytestmark = pytest.mark.gpu

EMD_RTOL = 1e-9      # BASELINE.json: 1e-9 relative in FP64 mode
SK_RTOL = 1e-9


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def oracle_emd_matrix(P, M):
    return po.emd_rows(P, M, 0, P.shape[0])


@pytest.mark.parametrize("K", [1, 2, 3, 10, 30, 32, 33, 40, 64])
def test_emd_pairs_match_oracle(K):
    S = 24
    P, M = synth.make_pairs(S, K, seed=300 + K)
    if K == 1:
        M = np.zeros((1, 1))
    rng = ops.make_range(S * S, _lib.PAIRS_FULL)
    out, status, piv = ops.emd_pairs(dev(P), dev(M), rng, want_info=True)
    got = out.cpu().numpy().reshape(S, S)
    want = oracle_emd_matrix(P, M)
    assert (status.cpu().numpy() == 0).all()
    np.testing.assert_allclose(got, want, rtol=EMD_RTOL, atol=1e-15)
    assert np.abs(np.diag(got)).max() <= 1e-15


@pytest.mark.parametrize("K", [3, 10, 30, 40, 64])
def test_emd_fp32_mode(K):
    """north_star's FP32 tier: the same solver on float data, within 1e-4 relative of the FP64 oracle."""
    S = 24
    P, M = synth.make_pairs(S, K, seed=300 + K)
    rng = ops.make_range(S * S, _lib.PAIRS_FULL)
    out, status, piv = ops.emd_pairs(dev(P), dev(M), rng, want_info=True, precision="f32")
    got = out.cpu().numpy().reshape(S, S)
    want = oracle_emd_matrix(P, M)
    assert (status.cpu().numpy() == 0).all()
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("scan_rows", ["1", "4", "64"])
def test_emd_scan_rows_do_not_change_the_optimum(scan_rows, monkeypatch):
    monkeypatch.setenv("PILOT_EMD_SCAN_ROWS", scan_rows)
    for K in (10, 40, 64):
        S = 12
        P, M = synth.make_pairs(S, K, seed=700 + K)
        out, status, piv = ops.emd_pairs(dev(P), dev(M), ops.make_range(S * S, _lib.PAIRS_FULL), want_info=True)
        assert (status.cpu().numpy() == 0).all()
        np.testing.assert_allclose(out.cpu().numpy().reshape(S, S), oracle_emd_matrix(P, M), rtol=EMD_RTOL, atol=1e-15)


@pytest.mark.parametrize("limit", ["0", "3"])
def test_emd_general_pivot_path(limit, monkeypatch):
    """Cycles longer than the one-lane-per-node path can take (forced here) use the general path."""
    monkeypatch.setenv("PILOT_EMD_FAST_LIMIT", limit)
    for K in (10, 40, 64):
        S = 12
        P, M = synth.make_pairs(S, K, seed=500 + K)
        out, status, piv = ops.emd_pairs(dev(P), dev(M), ops.make_range(S * S, _lib.PAIRS_FULL), want_info=True)
        assert (status.cpu().numpy() == 0).all()
        np.testing.assert_allclose(out.cpu().numpy().reshape(S, S), oracle_emd_matrix(P, M), rtol=EMD_RTOL, atol=1e-15)


def test_emd_degenerate_inputs():
    K = 12
    P, M = synth.make_pairs(6, K, seed=1)
    P[1] = P[0]                                   # identical samples
    P[2] = np.roll(P[0], 1)                       # permuted masses
    P[3, :4] = 0.0; P[3] /= P[3].sum()            # zero masses (POT drops them)
    P[4] = 1.0 / K                                # uniform
    P[5] = P[4]
    rng = ops.make_range(36, _lib.PAIRS_FULL)
    out, status, _ = ops.emd_pairs(dev(P), dev(M), rng, want_info=True)
    got = out.cpu().numpy().reshape(6, 6)
    want = oracle_emd_matrix(P, M)
    assert (status.cpu().numpy() == 0).all()
    np.testing.assert_allclose(got, want, rtol=EMD_RTOL, atol=1e-14)


@pytest.mark.parametrize("K,S", [(65, 10), (100, 10), (128, 8), (200, 6), (256, 5)])
def test_emd_more_than_64_types_general_kernel(K, S):
    """ot.emd2 has no limit on the number of cell types; beyond the bit-mask solver (K <= 64) the general network
    simplex of emd_general.cu takes over (Trajectory.py:511).  Same optimum as the oracle, zero diagonal."""
    P, M = synth.make_pairs(S, K, seed=900 + K)
    P[1, : K // 3] = 0.0
    P[1] /= P[1].sum()                                 # zero masses stay in the problem with supply 0
    P[2] = P[0]                                        # an identical pair
    rng = ops.make_range(S * S, _lib.PAIRS_FULL)
    out, status, piv = ops.emd_pairs(dev(P), dev(M), rng, want_info=True)
    got = out.cpu().numpy().reshape(S, S)
    want = oracle_emd_matrix(P, M)
    assert (status.cpu().numpy() == 0).all()
    np.testing.assert_allclose(got, want, rtol=EMD_RTOL, atol=1e-14)
    assert np.abs(np.diag(got)).max() <= 1e-14 and abs(got[0, 2]) <= 1e-14
    assert piv.cpu().numpy().max() > 0
    # the driver above it: symmetric cost -> upper triangle + mirror, as for K <= 64
    full = pairs.all_pairs(dev(P), dev(M), "unreg").cpu().numpy()
    np.testing.assert_allclose(full, want, rtol=EMD_RTOL, atol=1e-14)


def test_emd_general_kernel_agrees_with_the_bitmask_kernel_and_limits():
    """K = 64 padded with an empty 65th type goes through the general kernel and must reproduce the K = 64 result of
    the bit-mask kernel; more than 256 types raise."""
    S, K = 12, 64
    P, M = synth.make_pairs(S, K, seed=77)
    rng = ops.make_range(S * S, _lib.PAIRS_FULL)
    base = ops.emd_pairs(dev(P), dev(M), rng).cpu().numpy()
    P65 = np.concatenate([P, np.zeros((S, 1))], axis=1)
    M65 = np.ones((65, 65)) * 0.5
    M65[:64, :64] = M
    M65[64, 64] = 0.0
    padded = ops.emd_pairs(dev(P65), dev(M65), rng).cpu().numpy()
    np.testing.assert_allclose(padded, base, rtol=1e-12, atol=1e-15)
    with pytest.raises(_lib.PilotLibraryError):
        ops.emd_pairs(dev(np.full((2, 257), 1 / 257)), dev(np.zeros((257, 257))), ops.make_range(4, _lib.PAIRS_FULL))


def test_emd_nonmetric_cost_and_counts():
    # raw counts with equal totals (normalization=False) and an asymmetric, non-zero-diagonal cost
    rng_ = np.random.default_rng(3)
    K, S = 9, 10
    P = rng_.multinomial(500, np.ones(K) / K, size=S).astype(np.float64)
    M = rng_.random((K, K)) + 0.1
    got = pairs.all_pairs(dev(P), dev(M), "unreg").cpu().numpy()
    want = oracle_emd_matrix(P, M)
    np.testing.assert_allclose(got, want, rtol=EMD_RTOL, atol=1e-12)


@pytest.mark.parametrize("K,reg", [(10, 0.1), (30, 0.1), (64, 0.1), (64, 0.01), (40, 0.02), (5, 0.5), (30, 0.01), (32, 0.02), (16, 0.05), (48, 0.1), (33, 0.05), (20, 0.1), (56, 0.1), (50, 0.05)])
@pytest.mark.parametrize("algo", [0, 1, 3])
def test_sinkhorn_pairs_match_oracle(K, reg, algo):
    S = 12 if reg < 0.05 else 20
    P, M = synth.make_pairs(S, K, seed=400 + K)
    rng = ops.make_range(S * S, _lib.PAIRS_FULL)
    out, iters, absn, status = ops.sinkhorn_pairs(dev(P), dev(M), reg, rng, algo=algo, want_info=True)
    want, witers, wabs = po.sinkhorn_rows(P, M, reg, 0, S)
    got = out.cpu().numpy().reshape(S, S)
    np.testing.assert_array_equal(iters.cpu().numpy().reshape(S, S), witers)
    np.testing.assert_array_equal(absn.cpu().numpy().reshape(S, S), wabs)
    np.testing.assert_allclose(got, want, rtol=SK_RTOL, atol=1e-300)
    st = status.cpu().numpy()
    assert ((st == 0) | (st == 1)).all()


@pytest.mark.parametrize("K,reg", [(65, 0.1), (100, 0.1), (100, 0.02), (129, 0.1)])
def test_sinkhorn_more_than_64_types(K, reg):
    """K > 64 takes the reference-form kernel (one thread per row/column, roundup(K, 32) threads)."""
    S = 6
    P, M = synth.make_pairs(S, K, seed=900 + K)
    rng = ops.make_range(S * S, _lib.PAIRS_FULL)
    out, iters, absn, status = ops.sinkhorn_pairs(dev(P), dev(M), reg, rng, want_info=True)
    want, witers, wabs = po.sinkhorn_rows(P, M, reg, 0, S)
    np.testing.assert_array_equal(iters.cpu().numpy().reshape(S, S), witers)
    np.testing.assert_array_equal(absn.cpu().numpy().reshape(S, S), wabs)
    np.testing.assert_allclose(out.cpu().numpy().reshape(S, S), want, rtol=SK_RTOL, atol=1e-300)


@pytest.mark.parametrize("K", [12, 40])
@pytest.mark.parametrize("cap", ["0", "5", None])
def test_sinkhorn_redo_list_overflow(K, cap, monkeypatch):
    """More problems need the reference form than the redo list holds (capacity forced down here; 2^20 in
    production): the rest is found by the marker scan -- no output may stay NaN."""
    if cap is not None:
        monkeypatch.setenv("PILOT_SK_REDO_CAP", cap)
    S = 40
    P, M = synth.make_pairs(S, K, seed=31)
    P[::3, 2] = 0.0                       # zero masses -> POT's log(u) = -inf -> reference form
    P /= P.sum(axis=1, keepdims=True)
    rng = ops.make_range(S * S, _lib.PAIRS_FULL)
    out, iters, absn, status = ops.sinkhorn_pairs(dev(P), dev(M), 0.05, rng, want_info=True)
    ref = ops.sinkhorn_pairs(dev(P), dev(M), 0.05, rng, algo=1).cpu().numpy()
    got = out.cpu().numpy()
    assert (status.cpu().numpy() >= 0).all(), "a problem was left unsolved"
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    np.testing.assert_allclose(got, ref, rtol=1e-12, equal_nan=True)
    want, _, _ = po.sinkhorn_rows(P, M, 0.05, 0, 3)
    np.testing.assert_allclose(got.reshape(S, S)[:3], want, rtol=SK_RTOL, equal_nan=True)


@pytest.mark.parametrize("K,reg,S", [(10, 0.1, 20), (30, 0.1, 100), (64, 0.1, 40), (64, 0.01, 16), (40, 0.02, 16),
                                     (5, 0.5, 12), (30, 0.01, 16), (33, 0.05, 16), (64, 0.05, 24), (1, 0.1, 3)])
def test_sinkhorn_fp32_mode(K, reg, S):
    """north_star's FP32 tier: single-precision Sinkhorn (per-problem kernel in registers) within 1e-4 relative
    of the FP64 oracle, at the C2 (K=30, reg 0.1), C4 (K=64, reg 0.01) and C5 (K=64, reg 0.1) shapes and others.
    The iteration count may differ from the FP64 schedule only through the float stop threshold (5e-7)."""
    P, M = synth.make_pairs(S, K, seed=400 + K)
    if K == 1:
        M = np.zeros((1, 1))
    rng = ops.make_range(S * S, _lib.PAIRS_FULL)
    out, iters, absn, status = ops.sinkhorn_pairs(dev(P), dev(M), reg, rng, want_info=True, precision="f32")
    want, witers, wabs = po.sinkhorn_rows(P, M, reg, 0, S)
    got = out.cpu().numpy().reshape(S, S)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-7)
    it = iters.cpu().numpy().reshape(S, S)
    assert (it <= witers).all() and (it >= 1).all()
    assert (it % 20 == 1).all() or (it == 1000).any() or K == 1
    st = status.cpu().numpy()
    assert ((st == 0) | (st == 1)).all()


def test_sinkhorn_fp32_zero_mass_and_asymmetric_cost():
    K, S = 12, 6
    P, M = synth.make_pairs(S, K, seed=3)
    P[1, 3] = 0.0
    P[1] /= P[1].sum()
    rng = ops.make_range(S * S, _lib.PAIRS_FULL)
    a = ops.sinkhorn_pairs(dev(P), dev(M), 0.01, rng, precision="f32").cpu().numpy()
    want, _, _ = po.sinkhorn_rows(P, M, 0.01, 0, S)
    np.testing.assert_allclose(a, want.ravel(), rtol=1e-4, equal_nan=True)
    Ma = np.random.default_rng(3).random((K, K))
    Ma /= Ma.max()
    P2, _ = synth.make_pairs(S, K, seed=4)
    b = ops.sinkhorn_pairs(dev(P2), dev(Ma), 0.1, rng, precision="f32").cpu().numpy()
    want2, _, _ = po.sinkhorn_rows(P2, Ma, 0.1, 0, S)
    np.testing.assert_allclose(b, want2.ravel(), rtol=1e-4)


@pytest.mark.parametrize("K", [12, 64])
def test_sinkhorn_asymmetric_cost(K):
    """wasserstein_d accepts any cost matrix: a non-symmetric one keeps K0^T in shared memory."""
    S = 10
    P, _ = synth.make_pairs(S, K, seed=77)
    M = np.random.default_rng(3).random((K, K))
    M /= M.max()
    rng = ops.make_range(S * S, _lib.PAIRS_FULL)
    out, iters, absn, status = ops.sinkhorn_pairs(dev(P), dev(M), 0.1, rng, want_info=True)
    want, witers, wabs = po.sinkhorn_rows(P, M, 0.1, 0, S)
    np.testing.assert_array_equal(iters.cpu().numpy().reshape(S, S), witers)
    np.testing.assert_allclose(out.cpu().numpy().reshape(S, S), want, rtol=SK_RTOL)


def test_sinkhorn_zero_mass_goes_through_reference_form():
    # zero masses make POT's log(u) = -inf -> NaN roll-back; the batched kernel must hand these over
    K, S = 10, 4
    P, M = synth.make_pairs(S, K, seed=2)
    P[1, 3] = 0.0
    P[1] /= P[1].sum()
    rng = ops.make_range(S * S, _lib.PAIRS_FULL)
    a = ops.sinkhorn_pairs(dev(P), dev(M), 0.01, rng, algo=0).cpu().numpy()
    b = ops.sinkhorn_pairs(dev(P), dev(M), 0.01, rng, algo=1).cpu().numpy()
    want, _, _ = po.sinkhorn_rows(P, M, 0.01, 0, S)
    np.testing.assert_allclose(b, want.ravel(), rtol=1e-9, equal_nan=True)
    np.testing.assert_allclose(a, want.ravel(), rtol=1e-9, equal_nan=True)


@pytest.mark.parametrize("mode", [_lib.PAIRS_FULL, _lib.PAIRS_UPPER])
@pytest.mark.parametrize("nranks,block", [(1, None), (2, 7), (4, 5), (8, 64)])
def test_partition_invariance_and_unpack(mode, nranks, block):
    S, K = 23, 10
    P, M = synth.make_pairs(S, K, seed=11)
    Pd, Md = dev(P), dev(M)
    total = ops.n_pairs(S, mode)
    ref = pairs.all_pairs(Pd, Md, "unreg", symmetric=(mode == _lib.PAIRS_UPPER)).cpu().numpy()
    block = block or total
    chunk = max(1, pairs.range_count(total, block, nranks, 0))
    packed = torch.zeros((nranks * chunk,), dtype=torch.float64, device="cuda")
    for r in range(nranks):
        rng = _lib.PairRange(total=total, block=block, nranks=nranks, rank=r, mode=mode, reserved=0)
        n = _lib.range_count(rng)
        if n:
            ops.emd_pairs(Pd, Md, rng, out=packed[r * chunk:(r + 1) * chunk])
    rng0 = _lib.PairRange(total=total, block=block, nranks=nranks, rank=0, mode=mode, reserved=0)
    dense = ops.unpack_pairs(packed, chunk, S, rng0, 0.0).cpu().numpy()
    assert np.array_equal(dense, ref), "partitioned result must be bit-identical"
    want = oracle_emd_matrix(P, M)
    np.testing.assert_allclose(dense, want, rtol=EMD_RTOL, atol=1e-15)


def test_properties_at_scale():
    """Size-independent properties on a C5-shaped slice (K = 64): symmetry and zero diagonal of the
    exact EMD computed as ORDERED pairs, EMD <= Sinkhorn cost, determinism, Sinkhorn row marginal."""
    S, K = 160, 64
    P, M = synth.make_pairs(S, K, seed=5)
    Pd, Md = dev(P), dev(M)
    full = ops.emd_pairs(Pd, Md, ops.make_range(S * S, _lib.PAIRS_FULL)).cpu().numpy().reshape(S, S)
    assert np.abs(full - full.T).max() <= 1e-12
    assert np.abs(np.diag(full)).max() <= 1e-15
    tri = pairs.all_pairs(Pd, Md, "unreg").cpu().numpy()
    np.testing.assert_allclose(tri, full, rtol=1e-10, atol=1e-15)
    sk1 = pairs.all_pairs(Pd, Md, "reg", 0.1).cpu().numpy()
    sk2 = pairs.all_pairs(Pd, Md, "reg", 0.1).cpu().numpy()
    # which stragglers the DMMA panels hand to the warp-form tail kernel depends on timing, but the tail sums
    # in the panels' order (DMMA = FMA chain in ascending k): runs are bit-identical
    assert np.array_equal(sk1, sk2), "the Sinkhorn matrix must be bit-identical between runs"
    # ... and independent of how the pair space is cut: a row window solved alone gives the same bits
    win = ops.sinkhorn_pairs(Pd, Md, 0.1, _lib.PairRange(total=7 * S, block=7 * S, nranks=1, rank=0,
                                                         mode=_lib.PAIRS_FULL, reserved=0, first=40 * S)).cpu().numpy()
    assert np.array_equal(win.reshape(7, S), sk1[40:47])
    e1 = pairs.all_pairs(Pd, Md, "unreg").cpu().numpy()
    assert np.array_equal(tri, e1), "the exact EMD must be bit-identical between runs"
    off = ~np.eye(S, dtype=bool)
    assert (sk1[off] >= full[off] * (1 - 1e-9)).all()
    # spot-check 64 random entries against the oracle
    r = np.random.default_rng(0)
    for i, j in zip(r.integers(0, S, 64), r.integers(0, S, 64)):
        assert abs(tri[i, j] - po.emd2(P[i], P[j], M)) <= EMD_RTOL * max(tri[i, j], 1e-12)
        w = po.sinkhorn2(P[i], P[j], M, 0.1)
        assert abs(sk1[i, j] - w) <= SK_RTOL * w


@pytest.mark.parametrize("K,reg,S", [(64, 0.1, 100), (64, 0.01, 40), (40, 0.05, 60), (12, 0.1, 460), (48, 0.02, 40),
                                     (64, 0.1, 257)])  # 257: slots refilled and handed over at once (first product = c0)
def test_sinkhorn_bit_reproducible_with_tail_handover(K, reg, S):
    """Small batches end with most problems handed from the DMMA panels to the warp-form tail kernel at
    timing-dependent moments; the results must not depend on it: repeated runs, the panel-only solver of a
    different batch composition and a partitioned solve all give the same bits."""
    P, M = synth.make_pairs(S, K, seed=77 + K)
    Pd, Md = dev(P), dev(M)
    rng = ops.make_range(S * S, _lib.PAIRS_FULL)
    ref = ops.sinkhorn_pairs(Pd, Md, reg, rng).cpu().numpy()
    for _ in range(3):
        assert np.array_equal(ops.sinkhorn_pairs(Pd, Md, reg, rng).cpu().numpy(), ref)
    got = np.empty_like(ref)
    nranks, block = 3, 17
    for rank in range(nranks):
        r = _lib.PairRange(total=S * S, block=block, nranks=nranks, rank=rank, mode=_lib.PAIRS_FULL, reserved=0)
        out = ops.sinkhorn_pairs(Pd, Md, reg, r).cpu().numpy()
        idx = np.array([pairs.local_to_global(l, block, nranks, rank) for l in range(len(out))])
        got[idx] = out
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("mode", [_lib.PAIRS_FULL, _lib.PAIRS_UPPER])
def test_pair_range_windows(mode):
    """A window [first, first + total) of the pair space solves exactly the problems of that window."""
    S, K = 37, 12
    P, M = synth.make_pairs(S, K, seed=21)
    Pd, Md = dev(P), dev(M)
    total = ops.n_pairs(S, mode)
    full = ops.emd_pairs(Pd, Md, ops.make_range(total, mode)).cpu().numpy()
    for first, cnt, nranks in [(0, 5, 1), (17, 101, 1), (total - 9, 9, 1), (40, 300, 3)]:
        got = np.full(cnt, np.nan)
        for rank in range(nranks):
            block = 7 if nranks > 1 else cnt
            rng = _lib.PairRange(total=cnt, block=block, nranks=nranks, rank=rank, mode=mode, reserved=0, first=first)
            out = ops.emd_pairs(Pd, Md, rng).cpu().numpy()
            for l in range(len(out)):
                got[pairs.local_to_global(l, block, nranks, rank)] = out[l]
        assert np.array_equal(got, full[first:first + cnt])
        sk = ops.sinkhorn_pairs(Pd, Md, 0.1, _lib.PairRange(total=cnt, block=cnt, nranks=1, rank=0, mode=mode,
                                                            reserved=0, first=first)).cpu().numpy()
        for l in (0, cnt - 1):
            i, j = pairs.global_to_ij(first + l, S, mode)
            assert abs(sk[l] - po.sinkhorn2(P[i], P[j], M, 0.1)) <= SK_RTOL * sk[l]


@pytest.mark.parametrize("regularized", ["unreg", "reg"])
@pytest.mark.parametrize("n_bands", [1, 3, 7, 64])
def test_all_pairs_host_bands(regularized, n_bands):
    """The banded host pipeline returns the same matrix as the one-window device path, and its transpose."""
    S, K = 61, 9
    P, M = synth.make_pairs(S, K, seed=13)
    want = pairs.all_pairs(dev(P), dev(M), regularized, 0.1).cpu().numpy()
    got, got_T = pairs.all_pairs_host(dev(P), dev(M), regularized, 0.1, n_bands=n_bands, with_transpose=True)
    assert np.array_equal(got, want)
    assert np.array_equal(got_T, got.T) and got_T.flags.c_contiguous and got.flags.c_contiguous
    asym = np.random.default_rng(5).random((K, K))
    got2, got2_T = pairs.all_pairs_host(dev(P), dev(asym), regularized, 0.1, n_bands=n_bands, with_transpose=True)
    assert np.array_equal(got2_T, got2.T)
    assert np.array_equal(got2, pairs.all_pairs(dev(P), dev(asym), regularized, 0.1).cpu().numpy())


def test_wasserstein_d_big_matrix_band_pipeline():
    """S large enough for the band pipeline to kick in by itself (S*S*8 >= 256 MB): both containers, exact."""
    from pilot_b200 import tl
    S, K = 6000, 8
    P, M = synth.make_pairs(S, K, seed=17)
    rep = {f"s{i:05d}": P[i] for i in range(S)}
    EMD, df = tl.wasserstein_d(rep, M)
    assert EMD.shape == (S, S) and np.array_equal(EMD, EMD.T) and np.abs(np.diag(EMD)).max() == 0.0
    assert df.index.name == "sampleID" and list(df.columns[:3]) == ["s00000", "s00001", "s00002"]
    assert np.array_equal(df.to_numpy(), EMD.T)
    r = np.random.default_rng(2)
    for i, j in zip(r.integers(0, S, 300), r.integers(0, S, 300)):
        w = po.emd2(P[i], P[j], M)
        assert abs(EMD[i, j] - w) <= EMD_RTOL * max(w, 1e-300) + 1e-15
    SK, sdf = tl.wasserstein_d(rep, M, regularized="reg", reg=0.1)
    assert np.array_equal(sdf.to_numpy(), SK.T)
    for i, j in zip(r.integers(0, S, 100), r.integers(0, S, 100)):
        w = po.sinkhorn2(P[i], P[j], M, 0.1)
        assert abs(SK[i, j] - w) <= SK_RTOL * w
