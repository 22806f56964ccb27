"""Host-side profile of the end-to-end call (pilot_b200.tl.wasserstein_distance on C2 inputs):
where the milliseconds outside the H2D copy and the kernels go."""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pilot_b200 import synth, tl

n, d, k, s, seed = synth.CONFIGS["c2"]
X, obs = synth.make_cells(n, d, k, s, seed, labels="categorical")
pinned = torch.empty(X.shape, dtype=torch.float32, pin_memory=True)
Xp = pinned.numpy(); Xp[...] = X
kw = dict(emb_matrix="X_PCA", clusters_col="cell_types", sample_col="sampleID", status="status", regularized="reg", reg=0.1)
def step():
    adata = synth.FakeAnnData(obs, obsm={"X_PCA": Xp})
    tl.wasserstein_distance(adata, **kw)
    return adata
for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    step()
torch.cuda.synchronize()
print("ms per call", (time.perf_counter() - t0) * 100)
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
