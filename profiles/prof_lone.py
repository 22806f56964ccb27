import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pilot_b200 import _lib, ops, synth
K = int(sys.argv[1]) if len(sys.argv) > 1 else 30
reg = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
P, M = synth.make_pairs(100, K, seed=2)
Pd, Md = torch.from_numpy(P).cuda(), torch.from_numpy(M).cuda()
rng = ops.make_range(10000, _lib.PAIRS_FULL)
for rep in range(2):
    out = ops.sinkhorn_pairs(Pd, Md, reg, rng, want_info=True)
    torch.cuda.synchronize()
print(out[1].max().item(), out[1].float().mean().item())
