import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(autouse=True)
def _in_tmp_cwd(tmp_path, monkeypatch):
    # the hot path mkdirs ./Results_PILOT/plots (Trajectory.py:159-164): keep it out of the repo
    monkeypatch.chdir(tmp_path)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=True)
    return {k: z[k] for k in z.files}


def golden_names():
    # the synthetic-cohort fixtures of make_golden.py (g6 is the real-data fixture of make_kidney_golden.py)
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and "kidney" not in f)


def build_golden_adata(case):
    """Re-create the cohort a golden fixture was generated from (see tests/golden/make_golden.py)."""
    from pilot_b200 import synth
    n, d, k, s, seed, labels, dtype = int(case[0]), int(case[1]), int(case[2]), int(case[3]), int(case[4]), \
        str(case[5]), str(case[6])
    data_type = str(case[11])
    X, obs = synth.make_cells(n, d, k, s, seed, dtype=np.dtype(dtype), labels=labels)
    if data_type == "scRNA":
        return synth.FakeAnnData(obs, obsm={"X_PCA": X})
    return synth.FakeAnnData(obs, X=X, var_names=[f"feat{i}" for i in range(d)])


def golden_kwargs(case):
    return dict(emb_matrix="X_PCA", clusters_col="cell_types", sample_col="sampleID", status="status",
                metric=str(case[7]), regularized=str(case[8]), reg=float(case[9]), regulizer=float(case[10]),
                data_type=str(case[11]))
