"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: query partition and the id-sharded
gather + merge layout.  The CUDA engine is replaced by the oracle here (test infrastructure)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from hannoy_b200.sharded import pad_topk, partition_queries, shard_of

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_queries_covers_batch_contiguously():
    for nq in (0, 1, 7, 10_000, 100_003):
        for world in (1, 2, 4, 8):
            spans = [partition_queries(nq, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == nq
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _numpy_merge(ids, dist):
    world, nq, k = ids.shape
    o_ids = np.full((nq, k), 0xFFFFFFFF, np.uint32)
    o_dist = np.full((nq, k), np.inf, np.float32)
    lens = np.zeros(nq, np.uint32)
    for i in range(nq):
        keys = [(int(dist[p, i, j:j + 1].view(np.uint32)[0]), int(ids[p, i, j])) for p in range(world) for j in range(k)
                if ids[p, i, j] != 0xFFFFFFFF]
        keys.sort()
        keys = keys[:k]
        lens[i] = len(keys)
        for j, (b, idv) in enumerate(keys):
            o_ids[i, j] = idv
            o_dist[i, j] = np.array([b], np.uint32).view(np.float32)[0]
    return o_ids, o_dist, lens


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oracle.oracle import OracleDb
    from hannoy_b200.sharded import ShardedSearcher
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    n, dims, k = 600, 24, 10
    x = rng.normal(0, 1, (n, dims)).astype(np.float32)
    ids = np.arange(n, dtype=np.uint32) * 3 + 1
    q = rng.normal(0, 1, (16, dims)).astype(np.float32)
    mine = shard_of(ids, world) == rank
    db = OracleDb("euclidean", dims)
    db.add_items(ids[mine], x[mine])
    db.build(M=8, M0=16, ef_construction=64, seed=rank)

    def local_search(qq, count, ef):
        a, b, c, _ = db.search_by_vector(qq, count, ef=ef)
        return a, b, c

    s = ShardedSearcher(local_search=local_search, merge=_numpy_merge)
    got = s.search(q, k, 600)  # ef = n per shard -> exact per shard -> merged result is the exact global top-k
    full = OracleDb("euclidean", dims)
    full.add_items(ids, x)
    full.build(M=8, M0=16)
    w_ids, w_dist = full.exact_knn(q, k)
    ok = bool(np.array_equal(got[0], w_ids) and np.array_equal(got[1].view(np.uint32), w_dist.view(np.uint32)) and np.all(got[2] == k))
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gather_merge_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0] and ret[1]


def test_pad_topk():
    ids = np.arange(6, dtype=np.uint32).reshape(2, 3)
    dist = np.ones((2, 3), np.float32)
    i2, d2 = pad_topk(ids, dist, np.array([2, 0xFFFFFFFF], np.uint32), 3)
    assert i2[0, 2] == 0xFFFFFFFF and np.isinf(d2[0, 2]) and np.all(i2[1] == 0xFFFFFFFF)
