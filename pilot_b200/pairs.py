"""All-pairs OT on one or several GPUs.

The pair space is partitioned over the ranks of the current ``torch.distributed``
process group exactly as ``pilot_pair_range`` (include/pilot_b200.h) describes:
blocks of consecutive problems dealt round-robin.  Every rank solves its blocks
with the CUDA kernels, the packed results are assembled with ONE all-gather
(NCCL over NVLink on a B200 box) and unpacked (and mirrored for the symmetric
exact-EMD case) into the dense S x S matrix on every rank.

Replaces the double loop of the reference's ``wasserstein_d``
(/root/reference/pilotpy/tools/Trajectory.py:505-515).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib, ops
from ._lib import PairRange


def world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


# ---- host-side statement of the partition (mirrors common.cuh; used for sizing and tests) ----
def range_count(total: int, block: int, nranks: int, rank: int) -> int:
    if total <= 0:
        return 0
    nblocks = -(-total // block)
    mine = nblocks // nranks + (1 if (nblocks % nranks) > rank else 0)
    if mine == 0:
        return 0
    last_block = (mine - 1) * nranks + rank
    cnt = mine * block
    if last_block == nblocks - 1:
        cnt -= nblocks * block - total
    return cnt


def local_to_global(l: int, block: int, nranks: int, rank: int) -> int:
    lb, off = divmod(l, block)
    return (lb * nranks + rank) * block + off


def global_to_local(g: int, block: int, nranks: int) -> Tuple[int, int]:
    b, off = divmod(g, block)
    return b % nranks, (b // nranks) * block + off


def global_to_ij(g: int, S: int, mode: int) -> Tuple[int, int]:
    if mode == _lib.PAIRS_FULL:
        return divmod(g, S)
    i = 0
    # rows before i hold i*(2S-i-1)/2 entries
    lo, hi = 0, S - 2
    while lo < hi:
        mid = (lo + hi + 1) // 2
        if mid * (2 * S - mid - 1) // 2 <= g:
            lo = mid
        else:
            hi = mid - 1
    i = lo
    return i, g - i * (2 * S - i - 1) // 2 + i + 1


def choose_block(total: int, nranks: int) -> int:
    """>= 4 blocks per rank when possible, at most 4096 problems per block."""
    if nranks <= 1:
        return max(1, total)
    return max(1, min(4096, -(-total // (nranks * 4))))


def gather_packed(local: torch.Tensor, chunk: int, group=None) -> torch.Tensor:
    """All-gather equally padded per-rank chunks -> [nranks * chunk]."""
    nranks, _ = world(group)
    if nranks == 1:
        return local
    assert local.numel() == chunk
    out = torch.empty((nranks * chunk,), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return out


def cost_is_symmetric_metric_like(cost: torch.Tensor) -> bool:
    """True when mirroring the exact-EMD matrix is parity-safe: cost symmetric, zero
    diagonal and non-negative, so EMD(a,b) == EMD(b,a) and EMD(a,a) == 0."""
    c = cost
    return bool((c == c.t()).all().item() and (torch.diagonal(c) == 0).all().item() and (c >= 0).all().item())


def all_pairs(props: torch.Tensor, cost_norm: torch.Tensor, regularized="unreg", reg: float = 0.1,
              group=None, algo: int = 0, symmetric: Optional[bool] = None, want_info: bool = False):
    """Dense S x S matrix EMD[i, j] = OT(props[i], props[j]) on every rank.

    regularized == "unreg" -> exact EMD, anything else -> stabilised Sinkhorn
    (the string comparison is the reference's, Trajectory.py:507).
    """
    S, K = props.shape
    nranks, rank = world(group)
    exact = isinstance(regularized, str) and regularized == "unreg"
    if exact:
        if symmetric is None:
            symmetric = cost_is_symmetric_metric_like(cost_norm)
        mode = _lib.PAIRS_UPPER if symmetric else _lib.PAIRS_FULL
    else:
        mode = _lib.PAIRS_FULL  # Sinkhorn is neither symmetric nor zero on the diagonal
    total = ops.n_pairs(S, mode)
    block = choose_block(total, nranks)
    rng = PairRange(total=total, block=block, nranks=nranks, rank=rank, mode=mode, reserved=0)
    chunk = max(1, range_count(total, block, nranks, 0))  # rank 0 always holds the largest share
    packed = torch.zeros((chunk,), dtype=torch.float64, device=props.device) if nranks > 1 else None
    info = None
    if exact:
        res = ops.emd_pairs(props, cost_norm, rng, want_info=want_info, out=packed)
    else:
        res = ops.sinkhorn_pairs(props, cost_norm, reg, rng, algo=algo, want_info=want_info, out=packed)
    if want_info:
        info = res[1:]
        res = res[0]
    if nranks > 1:
        gathered = gather_packed(packed, chunk, group)
    else:
        gathered = res
        chunk = max(1, total)
    if total == 0:
        dense = torch.zeros((S, S), dtype=torch.float64, device=props.device)
    else:
        dense = ops.unpack_pairs(gathered.contiguous(), chunk, S, rng, 0.0)
    if want_info:
        return dense, info
    return dense
