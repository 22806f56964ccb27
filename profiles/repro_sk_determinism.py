import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pilot_b200 import _lib, ops, pairs, synth
S, K, reg = 257, 64, 0.1
P, M = synth.make_pairs(S, K, seed=60 + K)
Pd, Md = torch.from_numpy(P).cuda(), torch.from_numpy(M).cuda()
total = S * S
ref, it, ab, st = [x.cpu().numpy() for x in ops.sinkhorn_pairs(Pd, Md, reg, ops.make_range(total, _lib.PAIRS_FULL), want_info=True)]
for trial in range(3):
    again = ops.sinkhorn_pairs(Pd, Md, reg, ops.make_range(total, _lib.PAIRS_FULL)).cpu().numpy()
    print("repeat identical:", np.array_equal(again, ref), int((again != ref).sum()))
nranks = 2; block = pairs.choose_block(total, nranks)
got = np.empty_like(ref)
for rank in range(nranks):
    r = _lib.PairRange(total=total, block=block, nranks=nranks, rank=rank, mode=_lib.PAIRS_FULL, reserved=0)
    out = ops.sinkhorn_pairs(Pd, Md, reg, r).cpu().numpy()
    idx = np.array([pairs.local_to_global(l, block, nranks, rank) for l in range(len(out))])
    got[idx] = out
bad = np.nonzero(got != ref)[0]
print("2 virtual ranks: differing", len(bad), "of", total, "block", block)
for g in bad[:10]:
    print(g, divmod(int(g), S), it[g], ab[g], st[g], ref[g], got[g], (got[g]-ref[g])/ref[g])
algo3 = ops.sinkhorn_pairs(Pd, Md, reg, ops.make_range(total, _lib.PAIRS_FULL), algo=3).cpu().numpy()
print("algo3 vs algo0 differing:", int((algo3 != ref).sum()))
print("iters of differing:", np.unique(it[bad], return_counts=True))
