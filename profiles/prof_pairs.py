"""Profiling driver for the two pair kernels (run under ncu on the GPU box).
    python profiles/prof_pairs.py sinkhorn|emd|sinkhorn_small [K] [S] [rows] [reg] [algo]
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pilot_b200 import _lib, ops, synth

what = sys.argv[1] if len(sys.argv) > 1 else "sinkhorn"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 64
S = int(sys.argv[3]) if len(sys.argv) > 3 else 4000
rows = int(sys.argv[4]) if len(sys.argv) > 4 else 24
reg = float(sys.argv[5]) if len(sys.argv) > 5 else 0.1
algo = int(sys.argv[6]) if len(sys.argv) > 6 else 0
prec = sys.argv[7] if len(sys.argv) > 7 else "f64"
P, M = synth.make_pairs(S, K, seed=5)
Pd, Md = torch.from_numpy(P).cuda(), torch.from_numpy(M).cuda()
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if what == "sinkhorn":
        rng = ops.make_range(rows * S, _lib.PAIRS_FULL)
        out = ops.sinkhorn_pairs(Pd, Md, reg, rng, algo=algo, want_info=True, precision=prec)
        info = f"mean iters {out[1].float().mean().item():.1f}"
    else:
        rng = ops.make_range(rows * S, _lib.PAIRS_UPPER)
        out = ops.emd_pairs(Pd, Md, rng, want_info=True, precision=prec)
        info = f"mean pivots {out[2].float().mean().item():.1f}"
    e1.record(); torch.cuda.synchronize()
    print(what, prec, "algo", algo, "K", K, "problems", rows * S, "ms", e0.elapsed_time(e1), info)
