"""CPU coverage of the k-sharded multi-GPU path (world_size 2 and 3, gloo): the host-side sharding
(outerspace_b200/distributed.py) and the algorithmic claim the CUDA path relies on -- merging the
shards' partial products per output row in ascending source order reproduces the single-device
k-ordered result BIT FOR BIT.  The exchange itself is emulated with gloo object collectives and a
numpy merge (the oracle is the checker); the NCCL exchange is covered by tests/test_gpu_dist.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import operands, oracle_spgemm, rand_sparse
from outerspace_b200 import distributed as osd
from outerspace_b200.formats import CSRMatrix


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _row_partials(a_g: CSRMatrix, b_g: CSRMatrix, rows: int):
    """Per output row: list of (col, val) in the order multiplyPhase appends them (k ascending, then B-row order)."""
    out = [[] for _ in range(rows)]
    ap, bp = a_g.pos.astype(np.int64), b_g.pos.astype(np.int64)
    for i in range(a_g.NRow()):
        for e in range(ap[i], ap[i + 1]):
            k, av = int(a_g.data["idx"][e]), a_g.data["val"][e]
            for t in range(bp[k], bp[k + 1]):
                out[i].append((int(b_g.data["idx"][t]), np.float32(av * b_g.data["val"][t])))
    return out


def _merge(parts):
    """Stable sort by column + left fold in arrival order with separately rounded fp32 adds."""
    order = sorted(range(len(parts)), key=lambda p: (parts[p][0], p))
    cols, vals = [], []
    for p in order:
        c, v = parts[p]
        if cols and cols[-1] == c:
            vals[-1] = np.float32(vals[-1] + v)
        else:
            cols.append(c); vals.append(np.float32(v))
    return cols, vals


def _worker(rank, world, port, seed, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(seed)                      # same matrices on every rank
        m, k, n = 57, 41, 33
        A, B = rand_sparse(rng, m, k, 0.2), rand_sparse(rng, k, n, 0.25)
        a_csc, a_csr, b_csr = operands(A, B)
        ranges = osd.k_ranges(a_csr, b_csr, k, world)
        k0, k1 = ranges[rank]
        a_g, b_g = osd.shard_operands(a_csr, b_csr, k0, k1)
        mine = _row_partials(a_g, b_g, m)
        # "all-to-allv": every rank publishes its per-row partials, owners pick their rows in source order
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        r0, r1 = osd.row_block(m, world, rank)
        want, _ = oracle_spgemm(a_csc, b_csr, rows_override=m)
        ok = True
        for i in range(r0, r1):
            parts = [p for src in range(world) for p in gathered[src][i]]
            cols, vals = _merge(parts)
            lo, hi = int(want.pos[i]), int(want.pos[i + 1])
            ok &= cols == list(want.data["idx"][lo:hi])
            ok &= np.array_equal(np.array(vals, np.float32).view(np.uint32), want.data["val"][lo:hi].view(np.uint32))
        flags = [None] * world
        dist.all_gather_object(flags, bool(ok))
        if rank == 0:
            q.put((all(flags), ranges))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_merge_is_bit_exact(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 11 + world, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, ranges = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
    assert ranges[0][0] == 0 and ranges[-1][1] == 41
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))


def test_k_ranges_balance_and_shards_partition_the_operands():
    rng = np.random.default_rng(3)
    A, B = rand_sparse(rng, 300, 200, 0.05), rand_sparse(rng, 200, 150, 0.05)
    _, a_csr, b_csr = operands(A, B)
    for world in (1, 2, 4, 8):
        ranges = osd.k_ranges(a_csr, b_csr, 200, world)
        assert ranges[0][0] == 0 and ranges[-1][1] == 200
        nnz_a = nnz_b = 0
        work = []
        for k0, k1 in ranges:
            a_g, b_g = osd.shard_operands(a_csr, b_csr, k0, k1)
            assert a_g.NRow() == 300 and b_g.NRow() == k1 - k0
            assert a_g.nnz == 0 or int(a_g.data["idx"].max()) < k1 - k0
            nnz_a += a_g.nnz; nnz_b += b_g.nnz
            nnzc = np.bincount(a_g.data["idx"], minlength=k1 - k0)
            work.append(int((nnzc * np.diff(b_g.pos.astype(np.int64))).sum()))
        assert nnz_a == a_csr.nnz and nnz_b == b_csr.nnz
        if world > 1:
            assert max(work) <= 1.5 * (sum(work) / world) + 50
    assert osd.row_block(10, 4, 0) == (0, 2) and osd.row_block(10, 4, 3) == (7, 10)
