"""GPU parity: the CUDA engine, called through the C-ABI, against the CPU oracle on the same graphs.
Bit-exact ids AND distance bits for every metric (the kernels reproduce the reference's summation order)."""
import numpy as np
import pytest

from helpers import assert_counters_same, assert_same, make_db, make_vectors, open_reader_arrays, open_reader_kv

pytestmark = pytest.mark.gpu

CASES = [
    # metric, n, dims, kind
    ("euclidean", 3000, 128, "uniform"),
    ("euclidean", 1500, 100, "clustered"),   # AVX main part + 4-element scalar tail
    ("euclidean", 800, 20, "uniform"),       # SSE path
    ("euclidean", 500, 7, "uniform"),        # scalar path
    ("cosine", 2000, 96, "clustered"),
    ("cosine", 1000, 45, "uniform"),
    ("cosine", 600, 24, "uniform"),
    ("cosine", 400, 3, "uniform"),
    ("manhattan", 1200, 64, "uniform"),
    ("manhattan", 500, 10, "int"),
    ("hamming", 2000, 256, "uniform"),
    ("hamming", 700, 70, "uniform"),
    ("binary quantized cosine", 2000, 1024, "clustered"),
    ("binary quantized cosine", 900, 100, "uniform"),
    ("binary quantized euclidean", 1000, 192, "uniform"),
    ("binary quantized manhattan", 1000, 65, "uniform"),
]


@pytest.mark.parametrize("metric,n,dims,kind", CASES)
def test_by_vector_matches_oracle(metric, n, dims, kind):
    db, x = make_db(metric, n, dims, seed=n + dims, kind=kind)
    rd = open_reader_arrays(db, metric)
    q = make_vectors(64, dims, seed=7, kind=kind)
    q[:8] = x[:8]  # self queries
    for count, ef in [(10, 64), (1, 1), (10, 10), (100, 100), (5, 200)]:
        want = db.search_by_vector(q, count, ef=max(ef, count), counters=True)
        got = rd.nns(count).ef_search(ef).by_vectors_raw(q, counters=True)
        assert_same(got, want, f"{metric} n={n} d={dims} k={count} ef={ef}")
        assert_counters_same(got[3], want[3], f"{metric} k={count} ef={ef}")


@pytest.mark.parametrize("metric,n,dims,kind", CASES[::3])
def test_by_item_matches_oracle(metric, n, dims, kind):
    db, x = make_db(metric, n, dims, seed=n + dims + 1, kind=kind)
    rd = open_reader_arrays(db, metric)
    items = np.array([0, 1, 5, n - 1, n + 10, 17], np.uint32)  # n+10 is absent -> None
    for count, ef in [(10, 64), (3, 3)]:
        want = db.search_by_item(items, count, ef=max(ef, count), counters=True)
        got = rd.nns(count).ef_search(ef).by_items_raw(items, counters=True)
        assert_same(got, want, f"by_item {metric}")
        assert got[2][4] == 0xFFFFFFFF
        for i in (0, 1, 2, 3, 5):
            assert items[i] not in got[0][i, :got[2][i]]


@pytest.mark.parametrize("metric", ["euclidean", "cosine", "hamming", "binary quantized cosine"])
def test_kv_route_equals_array_route(metric):
    n, dims = 700, 40
    ids = np.sort(np.random.default_rng(3).choice(2_000_000, n, replace=False)).astype(np.uint32)
    ids[-1] = 0xFFFFFFFF  # sparse ids up to u32::MAX (src/tests/writer.rs:88-107)
    db, x = make_db(metric, n, dims, seed=11, ids=ids)
    ra = open_reader_arrays(db, metric)
    rk = open_reader_kv(db, metric, index=3)
    q = make_vectors(32, dims, seed=5)
    want = db.search_by_vector(q, 10, ef=50)
    assert_same(ra.nns(10).ef_search(50).by_vectors_raw(q), want, "arrays")
    assert_same(rk.nns(10).ef_search(50).by_vectors_raw(q), want, "kv")
    assert rk.n_items() == n and rk.dimensions() == dims and rk.version() == (0, 1, 3)
    assert np.array_equal(rk.item_ids(), ids)


def test_config1_shape():
    """BASELINE config 1: 10k x 128 Euclidean, M=16/M0=32, efC=100, 1k queries, top-10, ef=64."""
    db, x = make_db("euclidean", 10000, 128, seed=1, n_threads=8)
    rd = open_reader_arrays(db, "euclidean")
    q = make_vectors(1000, 128, seed=2)
    want = db.search_by_vector(q, 10, ef=64, counters=True, n_threads=8)
    got = rd.nns(10).ef_search(64).by_vectors_raw(q, counters=True)
    assert_same(got, want, "config 1")
    assert_counters_same(got[3], want[3], "config 1")


def test_id_sharded_search_merge_on_one_device():
    """The id-sharded path with both shards on one GPU: per-shard device-resident search into padded
    buffers, k-way merge kernel, against `reference reader on each shard index, merged by (distance bits, id)`."""
    import torch
    import hannoy_b200 as hb
    from hannoy_b200.sharded import shard_of
    rng = np.random.default_rng(0)
    n, dims, k, ef, world = 3000, 48, 10, 40, 3
    x = rng.normal(0, 1, (n, dims)).astype(np.float32)
    ids = np.arange(n, dtype=np.uint32) * 5 + 2
    q = rng.normal(0, 1, (130, dims)).astype(np.float32)
    dev = torch.device("cuda", 0)
    d_q = torch.from_numpy(q).to(dev)
    nq = len(q)
    g_ids = torch.full((world, nq, k), -1, dtype=torch.int32, device=dev)
    g_dist = torch.full((world, nq, k), float("inf"), dtype=torch.float32, device=dev)
    lens = torch.empty((nq,), dtype=torch.int32, device=dev)
    parts, readers = [], []
    from oracle.oracle import OracleDb
    stream = torch.cuda.current_stream().cuda_stream
    for s in range(world):
        m = shard_of(ids, world) == s
        db = OracleDb("euclidean", dims)
        db.add_items(ids[m], x[m])
        db.build(M=8, M0=16, ef_construction=40, seed=s)
        rd = hb.Reader.from_arrays("euclidean", dims, db.ids(), db.rows(), db.headers(), db.layers(), db.entry_points, db.max_level, index=s)
        readers.append(rd)
        rd.search_device(d_q.data_ptr(), nq, k, ef, g_ids[s].data_ptr(), g_dist[s].data_ptr(), lens.data_ptr(), None, stream)
        parts.append(db.search_by_vector(q, k, ef=ef))
        torch.cuda.synchronize()
        assert np.array_equal(g_ids[s].cpu().numpy().view(np.uint32), parts[-1][0]), f"shard {s}: device-resident search differs"
    o_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    o_dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
    o_len = torch.empty((nq,), dtype=torch.int32, device=dev)
    hb.merge_topk_device(0, g_ids.data_ptr(), g_dist.data_ptr(), world, nq, k, o_ids.data_ptr(), o_dist.data_ptr(), o_len.data_ptr(), stream)
    torch.cuda.synchronize()
    gi, gd, gl = o_ids.cpu().numpy().view(np.uint32), o_dist.cpu().numpy().view(np.uint32), o_len.cpu().numpy()
    for i in range(nq):
        keys = sorted((int(p[1][i, j:j + 1].view(np.uint32)[0]), int(p[0][i, j])) for p in parts for j in range(int(p[2][i])))[:k]
        assert gl[i] == len(keys)
        assert gi[i, :len(keys)].tolist() == [kk[1] for kk in keys], f"query {i}"
        assert gd[i, :len(keys)].tolist() == [kk[0] for kk in keys], f"query {i}"


def test_concurrent_calls_threads_and_streams():
    """`Reader: Send + Sync` (an hb_index is immutable after finalize): host calls from several threads and
    device-resident calls on several streams may overlap; every call gets a private workspace."""
    import threading
    import torch
    db, x = make_db("cosine", 4000, 96, seed=21, kind="clustered", n_threads=4)
    rd = open_reader_arrays(db, "cosine")
    qs = [make_vectors(700, 96, seed=100 + i, kind="clustered") for i in range(4)]
    want = [db.search_by_vector(q, 10, ef=64, n_threads=4) for q in qs]
    got = [None] * 4

    def worker(i):
        for _ in range(3):
            got[i] = rd.nns(10).ef_search(64).by_vectors_raw(qs[i])

    ths = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for i in range(4):
        assert_same(got[i], want[i], f"thread {i}")
    # device-resident API on four streams, all in flight together
    dev = torch.device("cuda", 0)
    streams = [torch.cuda.Stream(dev) for _ in range(4)]
    bufs = []
    for i, st in enumerate(streams):
        d_q = torch.from_numpy(qs[i]).to(dev)
        d_ids = torch.empty((700, 10), dtype=torch.int32, device=dev)
        d_dist = torch.empty((700, 10), dtype=torch.float32, device=dev)
        d_len = torch.empty((700,), dtype=torch.int32, device=dev)
        bufs.append((d_q, d_ids, d_dist, d_len))
    torch.cuda.synchronize()
    for rep in range(3):
        for i, st in enumerate(streams):
            d_q, d_ids, d_dist, d_len = bufs[i]
            rd.search_device(d_q.data_ptr(), 700, 10, 64, d_ids.data_ptr(), d_dist.data_ptr(), d_len.data_ptr(), None, st.cuda_stream)
    torch.cuda.synchronize()
    for i in range(4):
        d_q, d_ids, d_dist, d_len = bufs[i]
        g = (d_ids.cpu().numpy().view(np.uint32), d_dist.cpu().numpy(), d_len.cpu().numpy().view(np.uint32))
        assert_same(g, want[i], f"stream {i}")


@pytest.mark.parametrize("metric,n,dims,kind", [("cosine", 5000, 64, "clustered"), ("binary quantized cosine", 6000, 1024, "clustered"),
                                                ("euclidean", 4000, 20, "uniform"), ("hamming", 4000, 128, "uniform")])
def test_large_ef_heaps(metric, n, dims, kind):
    """ef beyond 256 entries: the heaps are merged tile by tile in place (sorted.cuh merge_batch_large), still in shared
    memory up to what fits, in global memory (pass 1) beyond — same ids, distance bits and traversal counters."""
    db, x = make_db(metric, n, dims, seed=n + dims + 5, kind=kind, n_threads=4)
    rd = open_reader_arrays(db, metric)
    q = make_vectors(48, dims, seed=3, kind=kind)
    for count, ef in [(100, 300), (100, 800), (10, 257), (300, 300), (50, 2000), (1000, 1000), (10, n)]:
        want = db.search_by_vector(q, count, ef=max(ef, count), counters=True, n_threads=4)
        got = rd.nns(count).ef_search(ef).by_vectors_raw(q, counters=True)
        assert_same(got, want, f"{metric} k={count} ef={ef}")
        assert_counters_same(got[3], want[3], f"{metric} k={count} ef={ef}")
    cand = np.arange(0, n, 3, dtype=np.uint32)
    want = db.search_by_vector(q, 100, ef=600, candidates=cand, linear_below=0, counters=True, n_threads=4)
    got = rd.nns(100).ef_search(600).candidates(cand).linear_below(0).by_vectors_raw(q, counters=True)
    assert_same(got, want, f"{metric} filtered, ef=600")
    assert_counters_same(got[3], want[3], f"{metric} filtered, ef=600")
    items = np.array([0, 11, n - 1], np.uint32)
    assert_same(rd.nns(100).ef_search(700).by_items_raw(items), db.search_by_item(items, 100, ef=700), f"{metric} by_item ef=700")
