"""Generate the committed golden fixtures from the REFERENCE ITSELF.

Run in the build container (where /root/reference is mounted):
    python tests/golden/make_golden.py

For every case the reference's own functions (ast-extracted verbatim from
/root/reference/pilotpy/tools/Trajectory.py by oracle/ref_exec.py) are run on a seeded
synthetic cohort; their outputs are stored next to the generator parameters:
  * proportions, cell/sample order, cost matrix, real labels  -> produced by reference code alone
  * EMD matrix -> produced by the reference's wasserstein_d loop with the oracle's `ot` shim
    (POT is not installable here; see oracle/pilot_oracle.py for the pinning status)
The .npz files are small (a few KB each) and travel to the GPU box.
"""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pilot_oracle, ref_exec  # noqa: E402
from pilot_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name: (cells, dim, types, samples, seed, labels, dtype, metric, regularized, reg, regulizer, data_type)
CASES = {
    "g1_str_cosine_emd": (3000, 8, 6, 9, 11, "str", "float32", "cosine", "unreg", 0.1, 0.2, "scRNA"),
    "g2_cat_euclid_sinkhorn": (4000, 12, 10, 12, 12, "categorical", "float32", "euclidean", "reg", 0.1, 0.2, "scRNA"),
    "g3_int_f64_emd": (2500, 5, 7, 8, 13, "int", "float64", "cosine", "unreg", 0.1, 0.5, "scRNA"),
    "g4_str_cosine_sinkhorn_smallreg": (3000, 10, 12, 6, 14, "str", "float32", "cosine", "reg", 0.02, 0.2, "scRNA"),
    "g5_pathomics_emd": (2000, 6, 5, 7, 15, "str", "float32", "cosine", "unreg", 0.1, 0.2, "Pathomics"),
}


def build_adata(case):
    n, d, k, s, seed, labels, dtype, *_rest, data_type = case
    X, obs = synth.make_cells(n, d, k, s, seed, dtype=np.dtype(dtype), labels=labels)
    if data_type == "scRNA":
        return synth.FakeAnnData(obs, obsm={"X_PCA": X})
    return synth.FakeAnnData(obs, X=X, var_names=[f"feat{i}" for i in range(d)])


def main():
    ref = ref_exec.load(pilot_oracle.OtShim(numpy_sinkhorn=True))
    cwd = os.getcwd()
    os.chdir("/tmp")  # the reference mkdirs ./Results_PILOT/plots
    try:
        for name, case in CASES.items():
            n, d, k, s, seed, labels, dtype, metric, regularized, reg, regulizer, data_type = case
            adata = build_adata(case)
            ref.wasserstein_distance(adata, emb_matrix="X_PCA", clusters_col="cell_types", sample_col="sampleID",
                                     status="status", metric=metric, regulizer=regulizer, regularized=regularized,
                                     reg=reg, data_type=data_type)
            u = adata.uns
            props = u["proportions"]
            np.savez_compressed(
                os.path.join(HERE, name + ".npz"),
                case=np.array([str(x) for x in case], dtype=object),
                samples=np.array([str(x) for x in props.keys()], dtype=object),
                cells=np.array([str(x) for x in u["cost"].columns], dtype=object),
                props=np.stack([props[x] for x in props.keys()]),
                cost=u["cost"].to_numpy(),
                EMD=u["EMD"],
                EMD_df=u["EMD_df"].to_numpy(),
                real_labels=np.array([str(x) for x in u["real_labels"]], dtype=object),
            )
            print(name, "props", np.stack(list(props.values())).shape, "cost", u["cost"].shape, "EMD", u["EMD"].shape)
    finally:
        os.chdir(cwd)


if __name__ == "__main__":
    main()
