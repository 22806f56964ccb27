"""Stage timing of the cells path (hist, props, median, cdist) on a BASELINE cohort, CUDA events, device resident:
    python profiles/prof_stages.py [c2|c3|c4]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pilot_b200 import ops, synth, tl

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
n, d, k, s, seed = synth.CONFIGS[cfg]
X, obs = synth.make_cells(n, d, k, s, seed, labels="categorical")
annot = obs[["cell_types", "sampleID", "status"]].copy(); annot.columns = ["cell_type", "sampleID", "status"]
lab = tl._Labels(annot, "cell_type", "sampleID")
Xd = torch.from_numpy(X).cuda()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best, out
ms_h, (counts_raw, _, _) = t(lambda: ops.hist(lab.ct_dev, lab.sm_dev, lab.K_raw, lab.S_raw))
ms_p, _ = t(lambda: ops.props_finalize(counts_raw, lab.perm_k_dev, lab.perm_s_dev, lab.n, 0.2, True))
ms_m, (cent, cent64) = t(lambda: ops.centroid_median(Xd, lab.ct_dev, lab.K_raw))
ms_c, _ = t(lambda: ops.cdist(cent64, "cosine"))
hbm = 6538.3
bm = n * d * X.itemsize + 4 * n
print(f"{cfg}: n={n} D={d} K={k} S={s}")
print(f"hist   {ms_h*1e3:8.1f} us  {8*n/ms_h/1e6:8.1f} GB/s ({8*n/ms_h/1e6/hbm:.3f} of HBM peak)")
print(f"props  {ms_p*1e3:8.1f} us")
print(f"median {ms_m*1e3:8.1f} us  {bm/ms_m/1e6:8.1f} GB/s ({bm/ms_m/1e6/hbm:.3f} of HBM peak), fallbacks {ops.median_fallbacks(k, d)}")
print(f"cdist  {ms_c*1e3:8.1f} us")
ref = np.stack([np.nanmedian(X[lab.ct_dev.cpu().numpy() == c], axis=0) for c in range(lab.K_raw)])
print("median bit-exact vs np.nanmedian:", np.array_equal(ref, cent.cpu().numpy()))
