"""oracle/ref_exec.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Runs the reference's OWN source for the hot path without importing ``pilotpy``
(whose import chain needs scanpy/anndata/POT/... that are absent here): the six
hot-path functions are ast-extracted from
``/root/reference/pilotpy/tools/Trajectory.py`` (36-116, 146-164, 234-299,
377-523, 617-642) and exec'd with pandas / NumPy / SciPy in scope and an ``ot``
shim (oracle/pilot_oracle.py::OtShim) in place of POT.

No reference source is copied into this repository: the text is read from
``/root/reference`` at call time.  That directory exists only in the build
container, so this module is used (a) by ``tests/golden/make_golden.py`` to
generate the committed fixtures and (b) by ``-m "not gpu"`` tests that skip
when the reference is not mounted.  Nothing run on the GPU box imports it.
"""
from __future__ import annotations

import ast
import os
import types

REFERENCE_ROOT = os.environ.get("PILOT_REFERENCE_ROOT", "/root/reference")
_TRAJ = os.path.join(REFERENCE_ROOT, "pilotpy", "tools", "Trajectory.py")

HOT_PATH_FUNCTIONS = (
    "wasserstein_distance", "set_path_for_results",
    "extract_data_anno_scRNA_from_h5ad", "extract_data_anno_pathomics_from_h5ad",
    "Cluster_Representations", "cost_matrix", "wasserstein_d", "return_real_labels",
)


def available() -> bool:
    return os.path.isfile(_TRAJ)


def load(ot_module=None) -> types.SimpleNamespace:
    """Return a namespace holding the reference's hot-path functions, exec'd verbatim."""
    if not available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    import numpy as np
    import pandas as pd
    import scipy
    import scipy.spatial.distance  # noqa: F401  (the reference uses scipy.spatial.distance.pdist)

    if ot_module is None:
        from . import pilot_oracle
        ot_module = pilot_oracle.OtShim(numpy_sinkhorn=True)
    with open(_TRAJ, "r", encoding="utf-8") as fh:
        src = fh.read()
    tree = ast.parse(src, filename=_TRAJ)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in HOT_PATH_FUNCTIONS]
    missing = set(HOT_PATH_FUNCTIONS) - {n.name for n in keep}
    if missing:
        raise RuntimeError(f"reference functions not found: {sorted(missing)}")
    mod = ast.Module(body=keep, type_ignores=[])
    ns = {"pd": pd, "np": np, "scipy": scipy, "os": os, "ot": ot_module, "__name__": "pilot_reference_exec"}
    exec(compile(mod, _TRAJ, "exec"), ns)
    return types.SimpleNamespace(**{k: ns[k] for k in HOT_PATH_FUNCTIONS}, _globals=ns)
