"""CPU-side checks of the C-ABI library: it loads without a GPU, exports every symbol the
header declares, validates arguments, and its partition arithmetic equals the host statement."""
import ctypes
import os
import re

import numpy as np
import pytest

from pilot_b200 import _lib, pairs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "pilot_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pilot_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = header_functions()
    assert set(names) == set(_lib.EXPORTS)
    for n in names:
        assert hasattr(L, n), n
    assert L.pilot_abi_version() == _lib.ABI_VERSION


def test_argument_validation_without_gpu():
    L = _lib.lib()
    assert L.pilot_hist(None, None, 10, 0, 3, None, None, None, None) < 0
    assert b"pilot_hist" in L.pilot_last_error()
    assert L.pilot_cdist(None, 4, 4, 0, None, None, None, None) < 0
    r = _lib.PairRange(total=10, block=0, nranks=1, rank=0, mode=0, reserved=0)
    assert L.pilot_emd_pairs(1, 4, 3, 1, 0, ctypes.byref(r), 1, 1, None, None, 1, 256, None) < 0
    assert L.pilot_emd_pairs(1, 4, 65, 1, 0, ctypes.byref(r), 1, 1, None, None, 1, 1 << 20, None) < 0
    assert L.pilot_sinkhorn_pairs(1, 4, 3, 1, -1.0, 1000, 1e-9, 1e3, 20, ctypes.byref(r), 0, 1, 1, None, None, None,
                                  1, 1 << 30, None) < 0
    assert L.pilot_workspace_bytes(_lib.WS_SINKHORN, 0, 64, 0, 0) > (1 << 20)
    assert L.pilot_workspace_bytes(_lib.WS_MEDIAN, 0, 30, 0, 50) >= 30 * 50 * 2 * 256 * 4


@pytest.mark.parametrize("total,block,nranks", [(0, 7, 3), (1, 1, 1), (4950, 300, 4), (10000, 625, 8),
                                                (199990000, 4096, 8), (17, 5, 8), (64, 64, 2)])
def test_partition_matches_host_statement(total, block, nranks):
    counts = []
    for rank in range(nranks):
        r = _lib.PairRange(total=total, block=block, nranks=nranks, rank=rank, mode=0, reserved=0)
        c = _lib.range_count(r)
        assert c == pairs.range_count(total, block, nranks, rank)
        counts.append(c)
    assert sum(counts) == total
    assert counts[0] == max(counts)
    if total <= 20000:
        seen = np.zeros(total, dtype=int)
        for rank in range(nranks):
            for l in range(counts[rank]):
                g = pairs.local_to_global(l, block, nranks, rank)
                assert pairs.global_to_local(g, block, nranks) == (rank, l)
                seen[g] += 1
        assert (seen == 1).all()


def test_upper_triangle_indexing():
    for S in (2, 3, 10, 101):
        g = 0
        for i in range(S):
            for j in range(i + 1, S):
                assert pairs.global_to_ij(g, S, _lib.PAIRS_UPPER) == (i, j)
                g += 1
    assert pairs.global_to_ij(7, 5, _lib.PAIRS_FULL) == (1, 2)


def test_no_cpu_fallback():
    import torch
    from pilot_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.PilotLibraryError):
        ops.pipe_peak(0)
    with pytest.raises(_lib.PilotLibraryError):
        import pandas as pd
        from pilot_b200 import tl
        tl.Cluster_Representations(pd.DataFrame({"cell_type": ["a", "b"], "sampleID": ["x", "y"]}))
