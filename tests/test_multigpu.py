"""Multi-GPU paths on real devices (needs >= 2 GPUs; skipped on a 1-GPU box): replicas with a partitioned query
batch, and the id-sharded search with the NCCL all-gather + device merge, both against the oracle."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, ret):
    try:
        _worker_body(rank, world, port, ret)
    except Exception:
        import traceback
        ret[rank] = traceback.format_exc()
        raise


def _worker_body(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import hannoy_b200 as hb
    from hannoy_b200.sharded import ShardedSearcher, partition_queries, shard_of
    from oracle.oracle import OracleDb
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    rng = np.random.default_rng(0)
    n, dims, k, ef = 4000, 64, 10, 64
    x = rng.normal(0, 1, (n, dims)).astype(np.float32)
    ids = (np.arange(n, dtype=np.uint32) * 7 + 3)
    q = rng.normal(0, 1, (257, dims)).astype(np.float32)
    ok = True
    why = []
    # ---- replicas: full graph on every rank, contiguous query slices, no collective ----
    full = OracleDb("cosine", dims)
    full.add_items(ids, x)
    full.build(M=16, M0=32, ef_construction=64, seed=5, n_threads=4)
    rd = hb.Reader.from_arrays("cosine", dims, full.ids(), full.rows(), full.headers(), full.layers(), full.entry_points,
                               full.max_level, device=rank)
    a, b = partition_queries(len(q), world, rank)
    got = rd.nns(k).ef_search(ef).by_vectors_raw(q[a:b])
    want = full.search_by_vector(q[a:b], k, ef=ef)
    ok &= bool(np.array_equal(got[0], want[0]) and np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32)) and np.array_equal(got[2], want[2]))
    if not ok:
        why.append("replica slice differs from the oracle")
    # ---- id shards: shard s = ids with id % world == s, one graph per shard; all-gather + merge on the device ----
    shards = []
    for s_ in range(world):
        m = shard_of(ids, world) == s_
        db = OracleDb("cosine", dims)
        db.add_items(ids[m], x[m])
        db.build(M=16, M0=32, ef_construction=64, seed=9 + s_, n_threads=1)  # every rank must build the SAME shard graphs
        shards.append(db)
    mine = shards[rank]
    rs = hb.Reader.from_arrays("cosine", dims, mine.ids(), mine.rows(), mine.headers(), mine.layers(), mine.entry_points,
                               mine.max_level, index=rank, device=rank)
    got = ShardedSearcher(reader=rs, device=rank).search(q, k, ef)
    # expected: the reference reader on each shard index, merged by (distance bits, id)
    parts = [db.search_by_vector(q, k, ef=ef) for db in shards]
    for i in range(len(q)):
        keys = sorted((int(p[1][i, j:j + 1].view(np.uint32)[0]), int(p[0][i, j])) for p in parts for j in range(int(p[2][i])))[:k]
        ok &= int(got[2][i]) == len(keys)
        ok &= [int(v) for v in got[0][i, :len(keys)]] == [kk[1] for kk in keys]
        ok &= [int(v) for v in got[1][i, :len(keys)].view(np.uint32)] == [kk[0] for kk in keys]
        if not ok and len(why) < 3:
            why.append(f"shard merge query {i}: got {got[0][i].tolist()} len {int(got[2][i])}, want {[kk[1] for kk in keys]}")
    # ---- the same search with the all-gather fused into the kernel epilogue (peer memory, no NCCL in the data path) ----
    ss = ShardedSearcher(reader=rs, device=rank).connect_fused(nq_cap=len(q), k_cap=k)
    d_q = torch.from_numpy(q).to(torch.device("cuda", rank))
    for rep in range(3):   # several epochs: exercises the double-buffered exchange
        f_ids, f_dist, f_len = ss.search_device_fused(d_q, k, ef)
    torch.cuda.synchronize()
    same = (np.array_equal(f_ids.cpu().numpy().view(np.uint32), got[0]) and np.array_equal(f_dist.cpu().numpy().view(np.uint32), got[1].view(np.uint32))
            and np.array_equal(f_len.cpu().numpy().view(np.uint32), got[2]))
    if not same:
        ok = False
        why.append("fused peer-memory exchange differs from the NCCL all-gather path")
    dist.barrier()
    ss.close_fused()
    ret[rank] = True if ok else "; ".join(why)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")
def test_replicas_and_id_shards_world2_nccl():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
    assert all(p.exitcode == 0 for p in procs), dict(ret)
    assert ret[0] is True and ret[1] is True, dict(ret)
