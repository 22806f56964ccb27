"""EMD kernel timing sweep on a C5-shaped slice (run on the GPU box):
    python profiles/prof_emd_sweep.py [K] [S] [rows]   -> one line per PILOT_EMD_SCAN_ROWS value."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pilot_b200 import _lib, ops, synth

K = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
rows = int(sys.argv[3]) if len(sys.argv) > 3 else 250
P, M = synth.make_pairs(S, K, seed=5)
Pd, Md = torch.from_numpy(P).cuda(), torch.from_numpy(M).cuda()
n = rows * S
for prec in ("f64", "f32"):
    for R in ([4, 8, 16, 32, 64] if K > 32 else [4, 8, 16, 32]):
        if R > K:
            continue
        os.environ["PILOT_EMD_SCAN_ROWS"] = str(R)
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = ops.emd_pairs(Pd, Md, ops.make_range(n, _lib.PAIRS_UPPER), want_info=True, precision=prec)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"emd {prec} K={K} problems={n} scan_rows={R}: {best:.2f} ms = {n / best / 1e3:.2f} M pairs/s, "
              f"mean pivots {out[2].float().mean().item():.1f}, bad status {(out[1] != 0).sum().item()}", flush=True)
