"""Graph construction on the device: hb_index_build_graph, i.e. HannoyBuilder::build (writer.rs:521-603 ->
hnsw.rs:122-216) run in batches on the GPU.  The reference's own build is parallel and scheduling-dependent, so there is
no bit-level oracle for the links; what is checked is what makes a graph a valid hannoy graph and a good one:
  * structure: one Links node per (item, layer <= its level), degrees within M0 / M, no self links, links stay inside their
    layer, the entry points are exactly the items of the top level, the level histogram follows the reference's law;
  * the exported pairs are the reference encoding: the CPU oracle reader opens them and the CUDA reader agrees with it
    bit for bit on the device-built graph;
  * quality: recall@10 at equal ef_search is on par with the restated sequential builder's graph."""
import numpy as np
import pytest

from helpers import assert_same, make_db, make_vectors
from oracle import oracle as O
from oracle.oracle import OracleDb
import hannoy_b200 as hb

pytestmark = pytest.mark.gpu


def _decode(kv, index):
    """-> (metadata dict, {(item, layer): [ids]}, n_item_nodes)"""
    links, meta, n_items = {}, None, 0
    for k, v in kv:
        assert len(k) == 8 and int.from_bytes(k[:2], "big") == index
        mode, item, layer = k[2], int.from_bytes(k[3:7], "big"), k[7]
        if mode == 0 and item == 0:
            name_end = v.index(b"\0")
            dims = int.from_bytes(v[name_end + 1:name_end + 5], "big")
            size = int.from_bytes(v[name_end + 5:name_end + 9], "big")
            items = O.roaring_deserialize(v[name_end + 9:name_end + 9 + size])
            rest = v[name_end + 9 + size:]
            meta = dict(distance=v[:name_end].decode(), dims=dims, items=items, eps=np.frombuffer(rest[:-1], np.uint32), max_level=rest[-1])
        elif mode == 2:
            assert v[0] == 1
            links[(item, layer)] = O.roaring_deserialize(v[1:])
        elif mode == 3:
            n_items += 1
    return meta, links, n_items


def _recall(ids, lens, gt):
    return float(np.mean([len(set(ids[i, :lens[i]].tolist()) & set(gt[i].tolist())) / gt.shape[1] for i in range(len(gt))]))


@pytest.mark.parametrize("metric,dims,n,M,M0", [("euclidean", 64, 20000, 16, 32), ("cosine", 96, 12000, 16, 32),
                                                ("binary quantized cosine", 512, 12000, 16, 32), ("manhattan", 24, 6000, 8, 16),
                                                ("hamming", 256, 6000, 12, 24)])
def test_device_built_graph_is_valid_and_as_good_as_the_cpu_build(metric, dims, n, M, M0):
    ids = (np.arange(n, dtype=np.uint32) * 2 + 7)
    ref, x = make_db(metric, n, dims, seed=n + dims, kind="clustered", ids=ids, M=M, M0=M0, efc=100, n_threads=8)
    stats = {}
    rd = hb.Reader.build(metric, dims, ids, ref.rows(), ref.headers(), M=M, M0=M0, ef_construction=100, seed=7, index=3, stats=stats)
    assert rd.n_items() == n and stats["items"] == n and stats["batches"] >= 1
    kv = rd.export_kv(with_items=True)
    meta, links, n_item_nodes = _decode(kv, 3)
    assert [k for k, _ in kv] == sorted(k for k, _ in kv)                       # LMDB key order
    assert meta["distance"] == metric and meta["dims"] == dims and np.array_equal(meta["items"], ids) and n_item_nodes == n
    L = meta["max_level"]
    assert L == stats["max_level"] == rd.max_level()
    # levels: P(level >= l) = M^-l (hnsw.rs:94-110): the layer populations shrink by about M per layer
    level = {}
    for (item, layer) in links:
        level[item] = max(level.get(item, 0), layer)
    assert set(level) == set(ids.tolist())                                       # every item has its layer-0 node
    pop = [sum(1 for v in level.values() if v >= l) for l in range(L + 1)]
    assert pop[0] == n and abs(pop[1] - n / M) < 6 * (n / M) ** 0.5 + 2
    for (item, layer), nb in links.items():
        assert layer <= level[item]
        assert len(nb) <= (M0 if layer == 0 else M), (item, layer, len(nb))
        assert item not in nb
        assert all(level[int(t)] >= layer for t in nb)                            # links stay inside their layer
    for item, lv in level.items():                                               # add_in_layers_below: a node on every layer below
        assert all((item, l) in links for l in range(lv + 1))
    assert sorted(meta["eps"].tolist()) == sorted(i for i, lv in level.items() if lv == L)   # hnsw.rs:268-279
    deg0 = np.array([len(links[(int(i), 0)]) for i in ids])
    assert (deg0 > 0).all() and deg0.mean() > 0.3 * M0

    # the CPU reader on the exported graph == the CUDA reader on it (ids, distance bits, traversal counters)
    cpu = OracleDb(metric, dims)
    cpu.add_items(ids, x)
    for (item, layer), nb in links.items():
        cpu.set_links(item, layer, nb)
    cpu.set_entry_points(meta["eps"], L)
    q = make_vectors(300, dims, seed=5, kind="clustered")
    for count, ef in [(10, 64), (50, 50)]:
        want = cpu.search_by_vector(q, count, ef=ef, counters=True, n_threads=8)
        got = rd.nns(count).ef_search(ef).by_vectors_raw(q, counters=True)
        assert_same(got, want, f"device-built graph {metric}")
        assert np.array_equal(got[3][:, :6], want[3][:, :6])
    rk = hb.Reader.open(kv, 3, metric)                                           # ... and through the reference encoding
    assert_same(rk.nns(10).ef_search(64).by_vectors_raw(q), cpu.search_by_vector(q, 10, ef=64, n_threads=8), "re-opened from exported pairs")

    # quality: recall@10 against exact k-NN, device build vs the restated CPU build, same M / M0 / ef_construction
    gt, _ = hb.exact_knn(rd, q, 10)
    for ef in (32, 64):
        g = rd.nns(10).ef_search(ef).by_vectors_raw(q)
        c = ref.search_by_vector(q, 10, ef=ef, n_threads=8)
        r_gpu, r_cpu = _recall(g[0], g[2], gt), _recall(c[0], c[2], gt)
        print(f"{metric}: recall@10 ef={ef}: device build {r_gpu:.4f}, cpu build {r_cpu:.4f}, batches {stats['batches']}")
        assert r_gpu >= r_cpu - 0.03, (metric, ef, r_gpu, r_cpu)


def test_build_small_and_degenerate_indexes():
    for n in (1, 2, 5, 40):
        x = make_vectors(n, 16, seed=n)
        db = OracleDb("euclidean", 16)
        db.add_items(np.arange(n, dtype=np.uint32), x)
        rd = hb.Reader.build("euclidean", 16, np.arange(n, dtype=np.uint32), db.rows(), db.headers(), seed=n)
        ids_, dist_, lens_ = rd.nns(min(n, 10)).ef_search(64).by_vectors_raw(x)
        assert (lens_ == min(n, 10)).all()
        assert (ids_[:, 0] == np.arange(n)).all() and (dist_[:, 0] == 0).all()      # self query (writer.rs:282-295)
        full = rd.nns(n).ef_search(max(n, 1)).by_vectors_raw(x[:1])
        assert sorted(full[0][0, :full[2][0]].tolist()) == list(range(n))           # everything reachable with ef = n (reader.rs:80-110)
    rd = hb.Reader.build("cosine", 8, np.zeros(0, np.uint32), np.zeros((0, 8), np.float32), None)
    assert rd.n_items() == 0 and rd.nns(5).by_vectors_raw(np.ones((2, 8), np.float32))[2].tolist() == [0, 0]


def test_build_from_the_writers_items_and_write_the_links_back():
    """The integration path: the database holds what `Writer::add_item` left (Item nodes, metadata without a graph, the
    `Updated` stones that make Reader::open fail with NeedBuild); the items are snapshotted through push_kv, the graph is
    built on the device, and the Links / metadata pairs to `put` back are exported.  The CPU reader then opens them."""
    import ctypes as C
    from hannoy_b200 import _lib as L
    n, dims, metric = 4000, 48, "euclidean"
    ids = np.sort(np.random.default_rng(3).choice(1 << 20, n, replace=False)).astype(np.uint32)
    src, x = make_db(metric, n, dims, seed=9, kind="clustered", ids=ids, build=True)
    lib = L.lib()
    pairs = [(bytes(k), bytes(v)) for k, v in src.export_kv(2)]
    items_only = [(k, v) for k, v in pairs if k[2] == 3 or (k[2] == 0)]            # Item nodes + metadata/version
    items_only += [(bytes([0, 2, 1]) + int(i).to_bytes(4, "big") + b"\0", b"") for i in ids[:5]]   # Updated stones (update_status.rs)
    h = C.c_void_p()
    assert lib.hb_index_begin(0, 2, C.byref(h)) == L.HB_OK
    for k, v in items_only:
        assert lib.hb_index_push_kv(h, k, len(k), v, len(v)) == L.HB_OK, lib.hb_last_error()
    opts = L.BuildOpts(16, 32, 100, 1.0, 11, 0, 0)
    st = np.zeros(8, np.uint64)
    assert lib.hb_index_build_graph(h, C.byref(opts), 0, st.ctypes.data_as(C.c_void_p)) == L.HB_OK, lib.hb_last_error()
    # a database that was never built: no metadata pair at all, the dimensions come from the caller
    h2 = C.c_void_p()
    assert lib.hb_index_begin(0, 2, C.byref(h2)) == L.HB_OK
    for k, v in pairs:
        if k[2] == 3:
            assert lib.hb_index_push_kv(h2, k, len(k), v, len(v)) == L.HB_OK
    assert lib.hb_index_build_graph(h2, C.byref(opts), 0, None) == L.HB_EMISSING_METADATA
    opts2 = L.BuildOpts(16, 32, 100, 1.0, 11, 0, dims)
    assert lib.hb_index_build_graph(h2, C.byref(opts2), 0, None) == L.HB_OK, lib.hb_last_error()
    assert lib.hb_index_finalize(h2, 0) == L.HB_OK
    rd2 = hb.Reader(h2, hb.Euclidean, 2, 0)
    assert rd2.n_items() == n and rd2.dimensions() == dims
    assert int(st[2]) == n
    assert lib.hb_index_finalize(h, 0) == L.HB_OK, lib.hb_last_error()
    rd = hb.Reader(h, hb.Euclidean, 2, 0)
    kv = rd.export_kv(with_items=False)
    meta, links, n_item_nodes = _decode(kv, 2)
    assert n_item_nodes == 0 and np.array_equal(meta["items"], ids) and len(links) >= n
    cpu = OracleDb(metric, dims)
    cpu.add_items(ids, x)
    for (item, layer), nb in links.items():
        cpu.set_links(item, layer, nb)
    cpu.set_entry_points(meta["eps"], meta["max_level"])
    q = make_vectors(200, dims, seed=4, kind="clustered")
    want = cpu.search_by_vector(q, 10, ef=64, counters=True, n_threads=4)
    got = rd.nns(10).ef_search(64).by_vectors_raw(q, counters=True)
    assert_same(got, want, "built from pushed items")
    assert np.array_equal(got[3][:, :6], want[3][:, :6])
    gt, _ = hb.exact_knn(rd, q, 10)
    assert _recall(got[0], got[2], gt) > 0.8
    g2 = rd2.nns(10).ef_search(64).by_vectors_raw(q)
    assert_same(g2, got, "same seed, same items: the device build is deterministic")
