"""Round-2 GPU parity cases: layer-0 lists longer than 32 (the reference's (24,48) / (32,64) builds, python.rs:280, and
the M0 = 768 shape of its fuzzer, src/tests/fuzz.rs:79-146), batches smaller than the resident warps (helper warps share
the rows of one query), replicas (one call, batch partitioned over devices), the device builder at M0 > 32 and its item
set on a database with pending updates — all against the CPU oracle, through the C-ABI."""
import ctypes as C

import numpy as np
import pytest

from helpers import assert_counters_same, assert_same, make_db, make_vectors, open_reader_arrays, open_reader_kv
from oracle.oracle import OracleDb
import hannoy_b200 as hb
from hannoy_b200 import _lib as L

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("metric,n,dims,kind,M,M0,efc", [
    ("cosine", 6000, 96, "clustered", 24, 48, 100),                  # two adjacency lines, second one partly filled
    ("euclidean", 5000, 128, "uniform", 32, 64, 100),                # two full lines
    ("euclidean", 3000, 20, "uniform", 24, 48, 64),                  # one lane per row (SSE order)
    ("binary quantized cosine", 5000, 512, "clustered", 32, 64, 100),
    ("hamming", 4000, 256, "uniform", 24, 48, 64),
    ("manhattan", 2000, 32, "uniform", 32, 64, 64),
    ("cosine", 1000, 32, "uniform", 16, 768, 32),                    # fuzz.rs: DIM 32, NUMEL 1000, M 16, M0 768, efC 32 -> CSR path
    ("euclidean", 2500, 64, "clustered", 48, 200, 200),              # lists of up to 200: CSR path, several chunks per expansion
])
def test_wide_layer0_lists_match_oracle(metric, n, dims, kind, M, M0, efc):
    db, x = make_db(metric, n, dims, seed=n + M0, kind=kind, M=M, M0=M0, efc=efc, n_threads=4)
    max_deg = int(np.diff(db.layers()[0][0]).max())
    assert max_deg > 32, "the case must exercise lists longer than one line"
    q = make_vectors(96, dims, seed=17, kind=kind)
    q[:6] = x[:6]
    for rd, route in ((open_reader_arrays(db, metric), "arrays"), (open_reader_kv(db, metric, index=1), "kv")):
        for count, ef in [(10, 64), (1, 1), (10, 10), (100, 100), (20, 300)]:
            want = db.search_by_vector(q, count, ef=max(ef, count), counters=True, n_threads=4)
            got = rd.nns(count).ef_search(ef).by_vectors_raw(q, counters=True)
            assert_same(got, want, f"{metric} M0={M0} (max degree {max_deg}) {route} k={count} ef={ef}")
            assert_counters_same(got[3], want[3], f"{metric} M0={M0} {route} k={count} ef={ef}")
        items = np.array([0, 3, n - 1, n + 5], np.uint32)
        assert_same(rd.nns(10).ef_search(50).by_items_raw(items), db.search_by_item(items, 10, ef=50), f"by_item {metric} M0={M0}")
        cand = np.arange(0, n, 2, dtype=np.uint32)
        want = db.search_by_vector(q, 10, ef=64, candidates=cand, linear_below=0, counters=True, n_threads=4)
        got = rd.nns(10).ef_search(64).candidates(cand).linear_below(0).by_vectors_raw(q, counters=True)
        assert_same(got, want, f"filtered {metric} M0={M0}")
        assert_counters_same(got[3], want[3], f"filtered {metric} M0={M0}")
        for polls in (1, 4, 30):   # cancellation polls are counted per pop, not per chunk of a long list
            want = db.search_by_vector(q[:16], 10, ef=64, cancel_after=polls, counters=True)
            got = rd.nns(10).ef_search(64).with_cancellation(polls).by_vectors_raw(q[:16], counters=True)
            got = (got[0], got[1], got[2] & 0x7fffffff, got[3])   # bit 31 = did_cancel
            assert_same(got, want, f"cancel after {polls} polls, {metric} M0={M0}")
            assert_counters_same(got[3], want[3], f"cancel after {polls} polls, {metric} M0={M0}")


@pytest.mark.parametrize("metric,dims", [("cosine", 768), ("euclidean", 128), ("cosine", 100)])
def test_small_batches_share_rows_between_warps(metric, dims):
    """Batches below the number of resident warps leave warps without a query: they gather rows for the others (search.cu
    TeamShared).  Every batch size from a single query up must give the oracle's ids, distance bits and counters, and the
    same as with the helpers switched off."""
    n = 6000
    db, x = make_db(metric, n, dims, seed=dims, kind="clustered", n_threads=8)
    rd = open_reader_arrays(db, metric)
    q = make_vectors(700, dims, seed=23, kind="clustered")
    want = db.search_by_vector(q, 10, ef=128, counters=True, n_threads=8)
    lib = L.lib()
    for nq in (1, 2, 3, 7, 33, 150, 445, 700):
        for team in (1, 0):
            assert lib.hb_tune(b"team", team) == L.HB_OK
            got = rd.nns(10).ef_search(128).by_vectors_raw(q[:nq], counters=True)
            assert_same(got, tuple(w[:nq] for w in want), f"{metric} d={dims} nq={nq} team={team}")
            assert_counters_same(got[3], want[3][:nq], f"{metric} d={dims} nq={nq} team={team}")
    lib.hb_tune(b"team", 1)
    # many short calls back to back from several threads: helpers of one launch never leak into the next
    import threading
    errs = []

    def worker(t):
        try:
            for i in range(40):
                a = (t * 40 + i) % 600
                g = rd.nns(10).ef_search(128).by_vectors_raw(q[a:a + 1 + (i % 5)])
                assert_same(g, tuple(w[a:a + 1 + (i % 5)] for w in want), f"thread {t} call {i}")
        except Exception as e:   # noqa: BLE001
            errs.append(repr(e))

    ths = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert not errs, errs[:2]
    # by_item and the large-ef (global heaps) pass go through the same row sharing
    items = np.arange(0, 60, dtype=np.uint32)
    assert_same(rd.nns(10).ef_search(64).by_items_raw(items), db.search_by_item(items, 10, ef=64, n_threads=8), "by_item small batch")
    assert_same(rd.nns(10).ef_search(3000).by_vectors_raw(q[:5]), db.search_by_vector(q[:5], 10, ef=3000), "ef=3000 small batch")


def test_replicas_partition_one_batch():
    """hb_index_replicate: ONE by_vectors call, contiguous nq / n_devices slices, one host thread + stream per device.
    On a single-GPU box the copies share device 0 (same code path: separate buffers, streams, workspaces)."""
    import torch
    n_gpu = torch.cuda.device_count()
    db, x = make_db("cosine", 5000, 96, seed=31, kind="clustered", n_threads=4)
    rd = open_reader_arrays(db, "cosine")
    extra = [d % n_gpu for d in range(1, 4)] if n_gpu > 1 else [0, 0]
    rd.replicate(extra)
    assert rd.devices() == [0] + extra
    q = make_vectors(1001, 96, seed=4, kind="clustered")
    want = db.search_by_vector(q, 10, ef=64, counters=True, n_threads=4)
    got = rd.nns(10).ef_search(64).by_vectors_raw(q, counters=True)
    assert_same(got, want, "replicated by_vectors")
    assert_counters_same(got[3], want[3], "replicated by_vectors")
    for nq in (1, 2, 3, 5):
        assert_same(rd.nns(10).ef_search(64).by_vectors_raw(q[:nq]), tuple(w[:nq] for w in want), f"replicated nq={nq}")
    items = np.array([0, 9, 4999, 7000, 12, 13, 14], np.uint32)
    assert_same(rd.nns(5).ef_search(40).by_items_raw(items), db.search_by_item(items, 5, ef=40), "replicated by_items")
    cand = np.arange(0, 5000, 7, dtype=np.uint32)
    want_c = db.search_by_vector(q[:200], 10, ef=64, candidates=cand, linear_below=0)
    assert_same(rd.nns(10).ef_search(64).candidates(cand).linear_below(0).by_vectors_raw(q[:200]), want_c, "replicated + candidates")
    tok = hb.CancelToken(0)
    g = rd.nns(10).ef_search(64).with_cancellation(tok).by_vectors_raw(q[:50])       # a token pins the batch to its device
    assert_same(g, tuple(w[:50] for w in want), "replicated + idle cancel token")
    with pytest.raises(hb.HannoyError):
        rd.nns(10).by_vectors_raw(np.zeros((4, 95), np.float32))


@pytest.mark.parametrize("metric,dims,n,M,M0", [("cosine", 96, 12000, 24, 48), ("euclidean", 64, 12000, 32, 64),
                                                ("binary quantized cosine", 512, 8000, 24, 48)])
def test_device_build_with_wide_layer0(metric, dims, n, M, M0):
    """hb_index_build_graph at the reference's larger instantiations ((24,48), (32,64), python.rs:280): valid graph,
    degrees within M0 / M and really above 32, the oracle reader agrees bit for bit on it, recall on par with the CPU build."""
    from test_gpu_build import _decode, _recall
    ids = np.arange(n, dtype=np.uint32) * 3 + 1
    ref, x = make_db(metric, n, dims, seed=n + M0, kind="clustered", ids=ids, M=M, M0=M0, efc=100, n_threads=8)
    stats = {}
    rd = hb.Reader.build(metric, dims, ids, ref.rows(), ref.headers(), M=M, M0=M0, ef_construction=100, seed=3, index=1, stats=stats)
    meta, links, _ = _decode(rd.export_kv(with_items=False), 1)
    deg0 = np.array([len(links[(int(i), 0)]) for i in ids])
    assert deg0.max() <= M0 and deg0.max() > 32 and (deg0 > 0).all()
    assert all(len(nb) <= M for (item, layer), nb in links.items() if layer > 0)
    assert all(item not in nb for (item, layer), nb in links.items())
    cpu = OracleDb(metric, dims)
    cpu.add_items(ids, x)
    for (item, layer), nb in links.items():
        cpu.set_links(item, layer, nb)
    cpu.set_entry_points(meta["eps"], meta["max_level"])
    q = make_vectors(300, dims, seed=6, kind="clustered")
    want = cpu.search_by_vector(q, 10, ef=64, counters=True, n_threads=8)
    got = rd.nns(10).ef_search(64).by_vectors_raw(q, counters=True)
    assert_same(got, want, f"device-built {metric} M0={M0}")
    assert_counters_same(got[3], want[3], f"device-built {metric} M0={M0}")
    gt, _ = hb.exact_knn(rd, q, 10)
    c = ref.search_by_vector(q, 10, ef=64, n_threads=8)
    r_gpu, r_cpu = _recall(got[0], got[2], gt), _recall(c[0], c[2], gt)
    print(f"{metric} M0={M0}: recall@10 ef=64: device build {r_gpu:.4f}, cpu build {r_cpu:.4f}, max degree {deg0.max()}")
    assert r_gpu >= r_cpu - 0.03


def test_rebuild_takes_the_items_present_not_the_stale_metadata():
    """A built database with pending updates (writer.rs:539-553: the item set of a build is (updated | indexed) - deleted,
    i.e. the Item nodes present): items added after the last build carry no entry in metadata.items, deleted ones have
    lost their Item node.  Reader::open refuses it (NeedBuild); hb_index_build_graph builds over exactly the nodes present."""
    n, dims, metric = 3000, 32, "euclidean"
    ids = np.arange(n, dtype=np.uint32) * 2
    src, x = make_db(metric, n, dims, seed=5, kind="clustered", ids=ids)
    pairs = [(bytes(k), bytes(v)) for k, v in src.export_kv(0)]
    deleted = set(int(i) for i in ids[10:40])
    new_ids = np.arange(1, 201, 2, dtype=np.uint32)          # odd ids: not in the stale metadata
    new_x = make_vectors(len(new_ids), dims, seed=77, kind="clustered")
    extra = OracleDb(metric, dims)
    extra.add_items(new_ids, new_x)
    new_pairs = [(bytes(k), bytes(v)) for k, v in extra.export_kv(0) if k[2] == 3]
    db_pairs = [(k, v) for k, v in pairs if not (k[2] == 3 and int.from_bytes(k[3:7], "big") in deleted)] + new_pairs
    db_pairs += [(bytes([0, 0, 1]) + int(i).to_bytes(4, "big") + b"\0", b"\1") for i in sorted(deleted)]      # UpdateStatus::Removed
    db_pairs += [(bytes([0, 0, 1]) + int(i).to_bytes(4, "big") + b"\0", b"\0") for i in new_ids]             # UpdateStatus::Updated
    db_pairs.sort(key=lambda kv: kv[0])
    with pytest.raises(hb.NeedBuild):
        hb.Reader.open(db_pairs, 0, metric)
    lib = L.lib()
    h = C.c_void_p()
    assert lib.hb_index_begin(0, 0, C.byref(h)) == L.HB_OK
    for k, v in db_pairs:
        assert lib.hb_index_push_kv(h, k, len(k), v, len(v)) == L.HB_OK, lib.hb_last_error()
    opts = L.BuildOpts(16, 32, 100, 1.0, 5, 0, 0)
    assert lib.hb_index_build_graph(h, C.byref(opts), 0, None) == L.HB_OK, lib.hb_last_error()
    assert lib.hb_index_finalize(h, 0) == L.HB_OK, lib.hb_last_error()
    rd = hb.Reader(h, hb.Euclidean, 0, 0)
    expect = sorted((set(int(i) for i in ids) - deleted) | set(int(i) for i in new_ids))
    assert rd.item_ids().tolist() == expect
    got = rd.nns(1).ef_search(32).by_vectors_raw(new_x[:50])
    assert (got[0][:, 0] == new_ids[:50]).all() and (got[1][:, 0] == 0).all()      # the new items are found as themselves
    got = rd.nns(10).ef_search(200).by_vectors_raw(x[10:40])
    assert not (set(got[0].ravel().tolist()) & deleted)                              # the deleted ones never come back
    opts_bad = L.BuildOpts(16, 32, 100, 1.0, 5, 0, dims + 1)
    h3 = C.c_void_p()
    assert lib.hb_index_begin(0, 0, C.byref(h3)) == L.HB_OK
    for k, v in db_pairs[:50]:
        lib.hb_index_push_kv(h3, k, len(k), v, len(v))
    assert lib.hb_index_build_graph(h3, C.byref(opts_bad), 0, None) == L.HB_EDIM
    lib.hb_index_free(h3)


def test_from_arrays_rejects_inconsistent_csr():
    db, x = make_db("euclidean", 200, 16, seed=3)
    layers = [(o.copy(), b.copy()) for o, b in db.layers()]
    off, nbr = layers[0]
    bad_off = off.copy(); bad_off[5], bad_off[6] = bad_off[6] + 3, bad_off[5]
    with pytest.raises(hb.HannoyError):
        hb.Reader.from_arrays("euclidean", 16, db.ids(), db.rows(), db.headers(), [(bad_off, nbr)] + layers[1:], db.entry_points, db.max_level)
    bad_nbr = nbr.copy()
    a, b = int(off[7]), int(off[8])
    assert b - a >= 2
    bad_nbr[a], bad_nbr[a + 1] = bad_nbr[a + 1], bad_nbr[a]
    with pytest.raises(hb.HannoyError):
        hb.Reader.from_arrays("euclidean", 16, db.ids(), db.rows(), db.headers(), [(off, bad_nbr)] + layers[1:], db.entry_points, db.max_level)
    bad0 = off.copy(); bad0[0] = 1
    with pytest.raises(hb.HannoyError):
        hb.Reader.from_arrays("euclidean", 16, db.ids(), db.rows(), db.headers(), [(bad0, nbr)] + layers[1:], db.entry_points, db.max_level)


def test_sharded_nccl_path_pads_short_shards():
    """A shard with fewer than `count` items contributes fewer than `count` hits: the NCCL route must pad them with the
    merge sentinel (the kernel zero-fills past out_len), or (id 0, distance 0) entries win the merge (world of one rank
    here; the 2-GPU test covers the exchange itself)."""
    import torch
    import torch.distributed as dist
    from hannoy_b200.sharded import ShardedSearcher
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29647", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        n, dims, k = 6, 40, 10
        db, x = make_db("euclidean", n, dims, seed=2, ids=np.arange(n, dtype=np.uint32) + 100)
        rd = open_reader_arrays(db, "euclidean")
        q = make_vectors(20, dims, seed=9)
        ss = ShardedSearcher(reader=rd, device=0)
        ids_, dist_, lens_ = ss.search_device(torch.from_numpy(q).cuda(), k, 32)
        torch.cuda.synchronize()
        want = db.search_by_vector(q, k, ef=32)
        got = (ids_.cpu().numpy().view(np.uint32), dist_.cpu().numpy(), lens_.cpu().numpy().view(np.uint32))
        assert (got[2] == n).all()
        assert_same(got, want, "short shard through the NCCL route")
    finally:
        dist.destroy_process_group()
