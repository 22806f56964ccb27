"""Per-iteration latency of the batched Sinkhorn kernel: a handful of problems that run to the
iteration cap (reg = 0.01) -> time / 1000 = latency of one iteration of a lone warp."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pilot_b200 import _lib, ops, synth
algo = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for K in (12, 30, 64):
    for S in (2, 8, 40):
        P, M = synth.make_pairs(S, K, seed=5)
        Pd, Md = torch.from_numpy(P).cuda(), torch.from_numpy(M).cuda()
        rng = ops.make_range(S * S, _lib.PAIRS_FULL)
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = ops.sinkhorn_pairs(Pd, Md, 0.01, rng, algo=algo, want_info=True); e1.record(); torch.cuda.synchronize()
        it = out[1]
        print(f"algo={algo} K={K} problems={S*S} max_iters={it.max().item()} mean={it.float().mean().item():.0f} ms={e0.elapsed_time(e1):.3f} -> {e0.elapsed_time(e1)*1e3/it.max().item():.2f} us/iter")
