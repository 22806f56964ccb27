"""N > 1 host logic on CPU (gloo, world_size 2): partition, equal-chunk all-gather and the packed
layout the unpack kernel consumes.  The CUDA kernels are replaced by a deterministic marker value
per problem so the test exercises only what runs on the host."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pilot_b200 import _lib, pairs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, S, mode, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert pairs.world() == (world, rank)
        total = S * S if mode == _lib.PAIRS_FULL else S * (S - 1) // 2
        block = pairs.choose_block(total, world)
        n_local = pairs.range_count(total, block, world, rank)
        chunk = max(1, pairs.range_count(total, block, world, 0))
        packed = torch.zeros(chunk, dtype=torch.float64)
        for l in range(n_local):
            g = pairs.local_to_global(l, block, world, rank)
            i, j = pairs.global_to_ij(g, S, mode)
            packed[l] = 1000.0 * i + j          # what a kernel would have produced for (i, j)
        gathered = pairs.gather_packed(packed, chunk)
        assert gathered.numel() == world * chunk
        dense = np.full((S, S), -1.0)
        for g in range(total):                   # host statement of pilot_unpack_pairs
            owner, l = pairs.global_to_local(g, block, world)
            i, j = pairs.global_to_ij(g, S, mode)
            dense[i, j] = gathered[owner * chunk + l].item()
            if mode == _lib.PAIRS_UPPER:
                dense[j, i] = dense[i, j]
        if mode == _lib.PAIRS_UPPER:
            np.fill_diagonal(dense, 0.0)
        q.put((rank, dense))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("S,mode", [(13, _lib.PAIRS_FULL), (13, _lib.PAIRS_UPPER), (2, _lib.PAIRS_UPPER)])
def test_two_rank_gather_layout(S, mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, S, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    i, j = np.meshgrid(np.arange(S), np.arange(S), indexing="ij")
    if mode == _lib.PAIRS_FULL:
        want = 1000.0 * i + j
    else:
        want = np.where(i < j, 1000.0 * i + j, 1000.0 * j + i)
        np.fill_diagonal(want, 0.0)
    for r in range(2):
        np.testing.assert_array_equal(res[r], want)
