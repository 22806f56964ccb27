"""Worker of tests/test_gpu_multirank.py (launched under torchrun, one rank per GPU, NCCL): the matrices
assembled from the ranks' shares with the NCCL all-gather must equal a single-rank solve of the same
problems entry by entry, BIT FOR BIT, for both solvers -- for the device path
(pairs.all_pairs), the banded host path (pairs.all_pairs_host) and the public tl.wasserstein_d."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pilot_oracle as po  # noqa: E402  (checker)
from pilot_b200 import pairs, synth, tl  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    worst = 0.0
    for S, K, reg in ((301, 30, 0.1), (257, 64, 0.1), (120, 40, 0.02), (97, 10, 0.1)):
        P, M = synth.make_pairs(S, K, seed=60 + K)
        Pd, Md = torch.from_numpy(P).cuda(), torch.from_numpy(M).cuda()
        for regularized in ("unreg", "reg"):
            multi = pairs.all_pairs(Pd, Md, regularized, reg)
            single = pairs.all_pairs(Pd, Md, regularized, reg, single_rank=True)
            host, host_T = pairs.all_pairs_host(Pd, Md, regularized, reg, n_bands=5, with_transpose=True)
            rel = float(((multi - single).abs() / single.abs().clamp_min(1e-300)).max().item())
            worst = max(worst, rel)
            assert torch.equal(multi, single), \
                f"{regularized}: {world} ranks and 1 rank differ (max rel {rel}, S={S}, K={K})"
            assert np.array_equal(host, single.cpu().numpy())
            assert np.array_equal(host_T, host.T)
            # every rank holds the same matrix
            ref = multi.clone()
            dist.broadcast(ref, 0)
            assert torch.equal(ref, multi), "ranks disagree after the all-gather"
        rep = {f"s{i}": P[i] for i in range(S)}
        EMD, df = tl.wasserstein_d(rep, M)
        if rank == 0:
            r = np.random.default_rng(S)
            for i, j in zip(r.integers(0, S, 40), r.integers(0, S, 40)):
                w = po.emd2(P[i], P[j], M)
                assert abs(EMD[i, j] - w) <= 1e-9 * max(w, 1e-300) + 1e-15
            assert np.array_equal(df.to_numpy(), EMD.T)
    # cells path: every rank uploads only its row slice of the embedding (pinned -> side stream, pageable ->
    # blocking), the slices meet in an in-place NCCL all-gather; results must equal the unsharded upload
    import tempfile
    os.chdir(tempfile.mkdtemp())
    X, obs = synth.make_cells(90_001, 24, 9, 15, seed=9, labels="categorical")
    pinned = torch.empty(X.shape, dtype=torch.float32, pin_memory=True)
    pinned.numpy()[...] = X
    outs = []
    for thresh, emb in (("0", X), ("0", pinned.numpy()), (str(1 << 40), X)):
        os.environ["PILOT_SHARD_H2D_MIN_BYTES"] = thresh
        adata = synth.FakeAnnData(obs, obsm={"X_PCA": emb})
        tl.wasserstein_distance(adata, regularized="reg", reg=0.1)
        outs.append((adata.uns["cost"].to_numpy(), adata.uns["EMD"]))
    for which, (c, e) in zip(("pageable", "pinned"), outs[:2]):
        dc, de = np.abs(c - outs[2][0]).max(), np.abs(e - outs[2][1]).max()
        assert np.array_equal(c, outs[2][0]) and np.array_equal(e, outs[2][1]), \
            f"sharded upload ({which}) changed the result: max |d cost| {dc}, max |d EMD| {de}"
    dist.barrier()
    if rank == 0:
        print(f"MULTIRANK OK world={world} worst_rel_diff={worst:.3e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
