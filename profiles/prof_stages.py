"""Profiling driver for stages 1-2 at a BASELINE shape (run under ncu on the GPU box).
    python profiles/prof_stages.py [c2|c3|c4] [reps]
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pilot_b200 import ops, synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n, d, k, s, seed = synth.CONFIGS[cfg]
rng = np.random.default_rng(seed)
X = torch.from_numpy(rng.normal(size=(n, d)).astype(np.float32)).cuda()
ct = torch.from_numpy(rng.integers(0, k, n).astype(np.int32)).cuda()
sm = torch.from_numpy(rng.integers(0, s, n).astype(np.int32)).cuda()
for rep in range(reps):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    counts, f1, f2 = ops.hist(ct, sm, k, s)
    e[1].record()
    cent, cent64 = ops.centroid_median(X, ct, k)
    e[2].record()
    torch.cuda.synchronize()
    print(cfg, "hist ms", e[0].elapsed_time(e[1]), "median ms", e[1].elapsed_time(e[2]), "fallbacks", ops.median_fallbacks(k, d))
