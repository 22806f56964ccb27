"""Timing of the two consumer kernels on a resident S x S matrix (CUDA events; the Gram DGEMM is torch.mm):
    python profiles/prof_consumers.py [S] [cpu_S]
Row kNN (k = 64, pilotpy.pl.trajectory's diffusion map) and silhouette (cosine, Sil_computing), next to
scikit-learn on the host cores at a smaller size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pilot_b200 import ops

S = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
cpu_S = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
rng = np.random.default_rng(0)
A = rng.random((S, 24))
E = torch.from_numpy(np.abs(A[:, None, :8].sum(-1) - A[None, :, :8].sum(-1)) + 0.0).cuda()   # a distance-like matrix
E = (E + E.t()) / 2
E /= E.max()
labels = rng.integers(0, 4, size=S)
hbm = 6538.3


def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best


gram = torch.mm(E, E.t())
ms_gemm = t(lambda: torch.mm(E, E.t()))
ms_knn = t(lambda: ops.knn_rows(E, 64)) - ms_gemm
ms_sil = t(lambda: ops.silhouette_rows(E, labels, "cosine")) - ms_gemm
ms_silp = t(lambda: ops.silhouette_rows(E, labels, "precomputed"))
b = 8.0 * S * S
print(f"S={S}: Gram DGEMM {ms_gemm:.1f} ms ({2.0 * S**3 / ms_gemm / 1e9:.1f} TFLOP/s)")
print(f"  knn_rows (k=64) kernel part   {ms_knn:8.2f} ms  {b / ms_knn / 1e6:7.1f} GB/s ({b / ms_knn / 1e6 / hbm:.2f} of HBM peak)")
print(f"  silhouette cosine kernel part {ms_sil:8.2f} ms  {b / ms_sil / 1e6:7.1f} GB/s ({b / ms_sil / 1e6 / hbm:.2f})")
print(f"  silhouette precomputed        {ms_silp:8.2f} ms  {b / ms_silp / 1e6:7.1f} GB/s ({b / ms_silp / 1e6 / hbm:.2f})")
from sklearn import metrics
from sklearn.neighbors import NearestNeighbors
Eh = E[:cpu_S, :cpu_S].cpu().numpy()
t0 = time.perf_counter(); metrics.silhouette_score(Eh, labels[:cpu_S], metric="cosine"); t1 = time.perf_counter()
NearestNeighbors(n_neighbors=64).fit(Eh).kneighbors(Eh); t2 = time.perf_counter()
print(f"scikit-learn on {os.cpu_count()} host cores at S={cpu_S}: silhouette {1e3 * (t1 - t0):.0f} ms, kNN {1e3 * (t2 - t1):.0f} ms "
      f"(cost grows ~S^3: x{(S / cpu_S) ** 3:.0f} at S={S})")
