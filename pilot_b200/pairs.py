"""All-pairs OT on one or several GPUs.

The pair space is partitioned over the ranks of the current ``torch.distributed``
process group exactly as ``pilot_pair_range`` (include/pilot_b200.h) describes:
blocks of consecutive problems dealt round-robin.  Every rank solves its blocks
with the CUDA kernels, the packed results are assembled with ONE all-gather
(NCCL over NVLink on a B200 box) and unpacked (and mirrored for the symmetric
exact-EMD case) into the dense S x S matrix on every rank.

``all_pairs``      -> dense matrix on the device (one window over the whole pair space)
``all_pairs_host`` -> dense matrix in host memory: the pair space is cut into bands of whole
                      matrix rows; band b's rows cross PCIe on a copy stream while band b+1
                      is being solved, so the 3.2 GB of a 20 000-sample matrix cost no time
                      on top of the kernels.

Replaces the double loop of the reference's ``wasserstein_d``
(/root/reference/pilotpy/tools/Trajectory.py:505-515).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
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


def row_start(i: int, S: int, mode: int) -> int:
    """Linear index of the first problem of matrix row i (i == S: one past the end)."""
    return i * S if mode == _lib.PAIRS_FULL else i * (2 * S - i - 1) // 2


def band_rows(S: int, mode: int, n_bands: int) -> List[int]:
    """Row boundaries r_0 = 0 < r_1 < ... < r_B = S of bands holding about the same number of problems
    (upper-triangle rows get shorter, so later bands take more rows)."""
    total = row_start(S, S, mode) if mode == _lib.PAIRS_FULL else S * (S - 1) // 2
    n_bands = max(1, min(n_bands, S))
    edges = [0]
    for b in range(1, n_bands):
        target = total * b // n_bands
        lo, hi = edges[-1], S
        while lo < hi:  # first row whose start is >= target
            mid = (lo + hi) // 2
            if row_start(mid, S, mode) >= target:
                hi = mid
            else:
                lo = mid + 1
        if lo > edges[-1] and lo < S:
            edges.append(lo)
    edges.append(S)
    return edges


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


def cost_is_symmetric_metric_like(cost) -> bool:
    """True when mirroring the exact-EMD matrix is parity-safe: cost symmetric, zero
    diagonal and non-negative, so EMD(a,b) == EMD(b,a) and EMD(a,a) == 0.  A host array is
    checked on the host; a device tensor costs ONE read-back."""
    if isinstance(cost, np.ndarray):
        return bool((cost == cost.T).all() and (np.diagonal(cost) == 0).all() and (cost >= 0).all())
    c = cost
    return bool(((c == c.t()).all() & (torch.diagonal(c) == 0).all() & (c >= 0).all()).item())


def _mode_for(regularized, cost_norm, symmetric: Optional[bool]) -> Tuple[bool, int]:
    exact = isinstance(regularized, str) and regularized == "unreg"
    if exact:
        if symmetric is None:
            symmetric = cost_is_symmetric_metric_like(cost_norm)
        return True, (_lib.PAIRS_UPPER if symmetric else _lib.PAIRS_FULL)
    return False, _lib.PAIRS_FULL  # Sinkhorn is neither symmetric nor zero on the diagonal


def _solve_window(props, cost_norm, exact: bool, reg: float, mode: int, first: int, total: int, nranks: int,
                  rank: int, group, algo: int, precision, dense: torch.Tensor, want_info: bool = False,
                  marks: Optional[list] = None):
    """Solve the window [first, first + total) of the pair space over the ranks, all-gather, unpack into
    `dense`.  Everything is asynchronous on the current stream."""
    S = props.shape[0]
    block = choose_block(total, nranks)
    rng = PairRange(total=total, block=block, nranks=nranks, rank=rank, mode=mode, reserved=0, first=first)
    chunk = max(1, range_count(total, block, nranks, 0))  # rank 0 always holds the largest share
    packed = torch.zeros((chunk,), dtype=torch.float64, device=props.device) if nranks > 1 else None

    def mark(name):
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    mark("start")
    if exact:
        res = ops.emd_pairs(props, cost_norm, rng, want_info=want_info, out=packed, precision=precision)
    else:
        res = ops.sinkhorn_pairs(props, cost_norm, reg, rng, algo=algo, want_info=want_info, out=packed,
                                 precision=precision)
    mark("solve")
    info = None
    if want_info:
        info = res[1:]
        res = res[0]
    if nranks > 1:
        gathered = gather_packed(packed, chunk, group)
    else:
        gathered = res
        chunk = max(1, total)
    mark("gather")
    ops.unpack_pairs(gathered.contiguous(), chunk, S, rng, 0.0, dense=dense)
    mark("unpack")
    return info


def all_pairs(props: torch.Tensor, cost_norm: torch.Tensor, regularized="unreg", reg: float = 0.1,
              group=None, algo: int = 0, symmetric: Optional[bool] = None, want_info: bool = False,
              precision="f64", single_rank: bool = False, marks: Optional[list] = None):
    """Dense S x S matrix EMD[i, j] = OT(props[i], props[j]) on every rank (device tensor).

    regularized == "unreg" -> exact EMD, anything else -> stabilised Sinkhorn
    (the string comparison is the reference's, Trajectory.py:507).
    single_rank=True solves everything on this rank, whatever the process group (parity checks).
    marks: optional list receiving (name, cuda event) pairs around solve / gather / unpack.
    """
    S, K = props.shape
    nranks, rank = (1, 0) if single_rank else world(group)
    exact, mode = _mode_for(regularized, cost_norm, symmetric)
    total = ops.n_pairs(S, mode)
    if total == 0:
        dense = torch.zeros((S, S), dtype=torch.float64, device=props.device)
        return (dense, None) if want_info else dense
    dense = torch.empty((S, S), dtype=torch.float64, device=props.device)
    info = _solve_window(props, cost_norm, exact, reg, mode, 0, total, nranks, rank, group, algo, precision, dense,
                         want_info, marks)
    if want_info:
        return dense, info
    return dense


_copy_streams = {}


def _copy_stream(dev: torch.device) -> "torch.cuda.Stream":
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _copy_streams:
        _copy_streams[idx] = torch.cuda.Stream(device=dev)
    return _copy_streams[idx]


def all_pairs_host(props: torch.Tensor, cost_norm: torch.Tensor, regularized="unreg", reg: float = 0.1,
                   group=None, algo: int = 0, symmetric: Optional[bool] = None, precision="f64",
                   n_bands: Optional[int] = None, with_transpose: bool = False
                   ) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """The dense S x S matrix in HOST memory on every rank, and optionally an independent ndarray holding
    its transpose (the reference hands out the ndarray and a DataFrame of the transpose, Trajectory.py:518).

    The pair space is cut into bands of whole matrix rows holding equal numbers of problems.  All bands are
    enqueued at once (solve -> all-gather -> unpack, asynchronous); the host then walks the bands and copies
    band b's rows out on a copy stream as soon as its unpack has finished, i.e. while the kernels of the
    later bands run.  With the upper-triangle mode a band's rows are final once it is unpacked (their
    left part was mirrored in by the earlier bands) and the transpose is the matrix itself; otherwise the
    transpose is formed on the device after the last band and copied out then.
    """
    S, K = props.shape
    nranks, rank = world(group)
    exact, mode = _mode_for(regularized, cost_norm, symmetric)
    total = ops.n_pairs(S, mode)
    out = np.empty((S, S), dtype=np.float64)
    out_T = np.empty((S, S), dtype=np.float64) if with_transpose else None
    if S == 0 or total == 0:
        out[...] = 0.0
        if out_T is not None:
            out_T[...] = 0.0
        return out, out_T
    dev = props.device
    if n_bands is None:
        # small matrices: one band (the copy takes microseconds); big ones: ~256 MB of rows per band
        n_bands = max(1, min(64, (S * S * 8) >> 28))
    edges = band_rows(S, mode, n_bands)
    dense = torch.empty((S, S), dtype=torch.float64, device=dev)
    main = torch.cuda.current_stream(dev)
    side = _copy_stream(dev)
    done = []
    for b in range(len(edges) - 1):
        first = row_start(edges[b], S, mode)
        cnt = row_start(edges[b + 1], S, mode) - first
        if cnt > 0:
            _solve_window(props, cost_norm, exact, reg, mode, first, cnt, nranks, rank, group, algo, precision, dense)
        else:  # the last row of the upper triangle holds no pair: only its diagonal entry
            dense[edges[b]:edges[b + 1]].diagonal(offset=edges[b]).zero_()
        ev = torch.cuda.Event()
        ev.record(main)
        done.append(ev)
    dense_T = None
    if with_transpose and mode == _lib.PAIRS_FULL:
        dense_T = dense.t().contiguous()
        ev = torch.cuda.Event()
        ev.record(main)
        done.append(ev)
    host, host_T = torch.from_numpy(out), (torch.from_numpy(out_T) if with_transpose else None)
    with torch.cuda.stream(side):
        for b in range(len(edges) - 1):
            side.wait_event(done[b])
            r0, r1 = edges[b], edges[b + 1]
            host[r0:r1].copy_(dense[r0:r1])  # pageable destination: returns when the rows have arrived
            if with_transpose and dense_T is None:
                host_T[r0:r1].copy_(dense[r0:r1])  # symmetric by construction (mirrored entries)
        if dense_T is not None:
            side.wait_event(done[-1])
            for r0 in range(0, S, 2048):
                host_T[r0:r0 + 2048].copy_(dense_T[r0:r0 + 2048])
            dense_T.record_stream(side)
    dense.record_stream(side)
    return out, out_T
