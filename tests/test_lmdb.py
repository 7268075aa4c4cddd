"""Native LMDB ingestion: Reader::open straight from a data.mdb, walked by the product's own read-only
B+tree walker (hannoy_b200/csrc/lmdb_walk.cpp) — no liblmdb, no heed.  The files are produced by tests/lmdb_writer.py
(liblmdb is absent from this image; see its header for what that means for format pinning)."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from helpers import assert_same, make_db, make_vectors
from lmdb_writer import LmdbWriter
from hannoy_b200 import _lib as L
import hannoy_b200 as hb


def _scan(path, db_name=None, prefix=b""):
    got = []

    def cb(user, k, kl, v, vl):
        got.append((bytes(k[:kl]), bytes(v[:vl])))
        return 0

    txn = C.c_uint64()
    st = L.lib().hb_lmdb_scan(os.fsencode(path), db_name.encode() if db_name else None, prefix, len(prefix), L.KV_VISIT(cb), None,
                              C.byref(txn))
    return st, got, txn.value


def _random_pairs(n, seed, big_every=0, klen=8):
    rng = np.random.default_rng(seed)
    keys = sorted({bytes(rng.integers(0, 256, klen, dtype=np.uint8)) for _ in range(n)})
    pairs = []
    for i, k in enumerate(keys):
        ln = int(rng.integers(0, 300))
        if big_every and i % big_every == 0:
            ln = int(rng.integers(2000, 20000))  # forces overflow pages
        pairs.append((k, bytes(rng.integers(0, 256, ln, dtype=np.uint8))))
    return pairs


@pytest.mark.parametrize("n,psize,fill,big_every", [(0, 4096, 1.0, 0), (1, 4096, 1.0, 0), (40, 4096, 1.0, 0), (3000, 4096, 1.0, 7),
                                                    (3000, 4096, 0.55, 0), (20000, 4096, 0.7, 50), (2500, 16384, 0.8, 5),
                                                    (1500, 512, 1.0, 9)])
def test_scan_returns_every_pair_in_key_order(tmp_path, n, psize, fill, big_every):
    pairs = _random_pairs(n, seed=n + psize, big_every=big_every)
    w = LmdbWriter(psize=psize, fill=fill)
    w.put_unnamed(pairs)
    w.save(str(tmp_path / "env"), txnid=11)
    st, got, txn = _scan(str(tmp_path / "env"))
    assert st == L.HB_OK, L.lib().hb_last_error()
    assert txn == 11
    assert got == pairs


def test_scan_prefix_range_named_databases_meta_choice_and_nosubdir(tmp_path):
    pairs = _random_pairs(6000, seed=5, big_every=11)
    w = LmdbWriter(fill=0.8, junk_branch_key0=True)  # a reader must never look at a branch page's first key
    w.put_named("vectors", pairs)
    other = _random_pairs(300, seed=6)
    w.put_named("documents", other)
    w.put_named("vectors-old", [])
    fn = w.save(str(tmp_path / "env.mdb"), txnid=3, newest_meta=0, nosubdir=True)  # meta page 0 is the newer one
    for prefix in (b"", pairs[0][0][:1], pairs[3000][0][:2], pairs[-1][0][:2], b"\x00", b"\xff\xff"):
        st, got, txn = _scan(fn, "vectors", prefix)
        assert st == L.HB_OK, L.lib().hb_last_error()
        assert txn == 3
        assert got == [p for p in pairs if p[0].startswith(prefix)]
    st, got, _ = _scan(fn, "documents")
    assert st == L.HB_OK and got == other
    st, got, _ = _scan(fn, "vectors-old")
    assert st == L.HB_OK and got == []
    st, got, _ = _scan(fn, "nope")
    assert st == L.HB_EINVAL and b"no database named" in L.lib().hb_last_error()
    # the unnamed database of this environment holds the three sub-database records
    st, got, _ = _scan(fn)
    assert st == L.HB_OK and [k for k, _ in got] == [b"documents", b"vectors", b"vectors-old"]


def test_stale_meta_page_is_not_read(tmp_path):
    """The meta page with the smaller txnid describes the previous commit; only the newest is used."""
    old = _random_pairs(50, seed=1)
    new = _random_pairs(80, seed=2)
    w = LmdbWriter()
    stale = w.build_tree(old)
    w.put_unnamed(new)
    for newest in (0, 1):
        w.save(str(tmp_path / f"e{newest}"), txnid=100, newest_meta=newest, stale_main=stale)
        st, got, txn = _scan(str(tmp_path / f"e{newest}"))
        assert st == L.HB_OK and txn == 100 and got == new


def test_corrupt_files_are_rejected_not_crashed_on(tmp_path):
    pairs = _random_pairs(4000, seed=9, big_every=13)
    w = LmdbWriter()
    w.put_unnamed(pairs)
    fn = w.save(str(tmp_path / "env"))
    blob = bytearray(open(fn, "rb").read())
    lib = L.lib()

    def run(mut):
        b = bytearray(blob)
        mut(b)
        p = str(tmp_path / "bad.mdb")
        open(p, "wb").write(b)
        return _scan(p)[0]

    assert run(lambda b: b.__setitem__(slice(16, 20), b"\0\0\0\0")) == L.HB_EFORMAT            # magic of meta 0
    assert b"not an LMDB data file" in lib.hb_last_error()
    assert run(lambda b: b.__delitem__(slice(len(b) // 2, len(b)))) == L.HB_EFORMAT            # truncated file
    root = w.main.root

    def self_loop(b):  # first child of the root points back at the root
        off = root * 4096
        ptr0 = struct.unpack_from("<H", b, off + 16)[0]
        struct.pack_into("<HHH", b, off + ptr0, root & 0xffff, (root >> 16) & 0xffff, 0)
    assert run(self_loop) == L.HB_EFORMAT

    def wild_ptr(b):
        struct.pack_into("<H", b, root * 4096 + 16, 4095)
    assert run(wild_ptr) == L.HB_EFORMAT

    def bad_flags(b):
        struct.pack_into("<H", b, root * 4096 + 10, 0)
    assert run(bad_flags) == L.HB_EFORMAT
    assert _scan(str(tmp_path / "missing"))[0] == L.HB_EINVAL
    open(str(tmp_path / "tiny"), "wb").write(b"abc")
    assert _scan(str(tmp_path / "tiny"))[0] == L.HB_EFORMAT


def test_commit_during_the_walk_is_detected(tmp_path):
    """No reader-table slot is taken, so a commit landing while the file is walked invalidates the snapshot: the scan
    reports HB_ESTATE instead of returning possibly recycled pages."""
    pairs = _random_pairs(500, seed=3)
    w = LmdbWriter()
    w.put_unnamed(pairs)
    fn = w.save(str(tmp_path / "env"), txnid=20, newest_meta=1)
    seen = []

    def cb(user, k, kl, v, vl):
        seen.append(1)
        if len(seen) == 100:  # a writer commits txn 21 into meta page 0
            with open(fn, "r+b") as f:
                f.seek(16 + 24 + 96 + 8)
                f.write(struct.pack("<Q", 21))
        return 0

    st = L.lib().hb_lmdb_scan(os.fsencode(fn), None, b"", 0, L.KV_VISIT(cb), None, None)
    assert st == L.HB_ESTATE and b"committed to during the snapshot" in L.lib().hb_last_error()
    assert len(seen) == len(pairs)
    # a visitor may stop the scan
    fn = w.save(str(tmp_path / "env2"), txnid=20)
    st = L.lib().hb_lmdb_scan(os.fsencode(fn), None, b"", 0, L.KV_VISIT(lambda *a: 1), None, None)
    assert st == L.HB_ESTATE


def _write_env(tmp_path, dbs, name=None, **kw):
    """dbs = {index: OracleDb}: all indexes share one LMDB database, like hannoy's u16 index prefix allows."""
    pairs = []
    for index, db in dbs.items():
        pairs += [(bytes(k), bytes(v)) for k, v in db.export_kv(index)]
    pairs.sort()
    w = LmdbWriter(**kw)
    if name:
        w.put_named(name, pairs)
        w.put_named("zz-other", _random_pairs(100, seed=1))
    else:
        w.put_unnamed(pairs)
    return w.save(str(tmp_path / "env"))


@pytest.mark.parametrize("metric,dims", [("cosine", 768), ("euclidean", 40), ("hamming", 70), ("binary quantized cosine", 1024)])
def test_push_lmdb_decodes_the_index(tmp_path, metric, dims):
    """hb_index_push_lmdb + the host half of finalize reproduce what the Writer stored (768-d f32 items are 3 077-byte
    values: every one of them lives on overflow pages)."""
    n = 400
    ids = np.sort(np.random.default_rng(8).choice(1 << 28, n, replace=False)).astype(np.uint32)
    db, x = make_db(metric, n, dims, seed=2, ids=ids)
    db2, _ = make_db(metric, 60, dims, seed=3)
    _write_env(tmp_path, {0: db2, 3: db, 4: db2, 0x0103: db2}, name="vecs", fill=0.75)
    lib = L.lib()
    mid = hb.reader._distance_of(metric).ID
    h = C.c_void_p()
    assert lib.hb_index_begin(mid, 3, C.byref(h)) == L.HB_OK
    npairs = C.c_uint64()
    st = lib.hb_index_push_lmdb(h, os.fsencode(str(tmp_path / "env")), b"vecs", C.byref(npairs))
    assert st == L.HB_OK, lib.hb_last_error()
    assert npairs.value == len(db.export_kv(3))          # only index 3's key range was visited
    st = lib.hb_index_finalize(h, 0)
    assert st in (L.HB_OK, L.HB_ECUDA), lib.hb_last_error()
    assert lib.hb_index_n_items(h) == n and lib.hb_index_dimensions(h) == dims
    assert lib.hb_index_max_level(h) == db.max_level and lib.hb_index_n_entry_points(h) == len(db.entry_points)
    got = np.zeros(n, np.uint32)
    lib.hb_index_item_ids(h, got.ctypes.data_as(C.c_void_p), n)
    assert np.array_equal(got, ids)
    v = np.zeros(dims, np.float32)
    for s in (0, 200, n - 1):
        assert lib.hb_index_item_vector(h, int(ids[s]), v.ctypes.data_as(C.c_void_p)) == L.HB_OK
        if metric in ("euclidean", "cosine"):
            assert np.array_equal(v, x[s])
        elif metric == "hamming":
            assert np.array_equal(v, (x[s] > 0).astype(np.float32))
        else:
            assert np.array_equal(v, np.where(np.signbit(x[s]), -1.0, 1.0).astype(np.float32))
    lib.hb_index_free(h)
    # Reader::open errors surface through the path route too
    h = C.c_void_p()
    assert lib.hb_index_open_lmdb(os.fsencode(str(tmp_path / "env")), b"vecs", mid, 9, 0, C.byref(h)) == L.HB_EMISSING_METADATA
    assert lib.hb_index_open_lmdb(os.fsencode(str(tmp_path / "env")), b"vecs", (mid + 1) % 7, 3, 0, C.byref(h)) == L.HB_EUNMATCHING_DISTANCE
    with pytest.raises(hb.MissingMetadata):
        hb.Reader.open_path(str(tmp_path / "env"), 9, metric, db_name="vecs")


@pytest.mark.gpu
@pytest.mark.parametrize("metric,dims,named", [("cosine", 768, True), ("euclidean", 128, False), ("hamming", 256, False),
                                                ("binary quantized cosine", 1024, True)])
def test_open_path_search_equals_oracle(tmp_path, metric, dims, named):
    n = 2500
    ids = np.sort(np.random.default_rng(5).choice(1 << 26, n, replace=False)).astype(np.uint32)
    db, x = make_db(metric, n, dims, seed=12, kind="clustered", ids=ids)
    other, _ = make_db(metric, 100, dims, seed=13)
    path = _write_env(tmp_path, {6: other, 7: db, 8: other}, name="hannoy" if named else None)
    rd = hb.Reader.open_path(path if not named else os.path.dirname(path), 7, metric, db_name="hannoy" if named else None)
    assert rd.n_items() == n and rd.dimensions() == dims
    q = make_vectors(96, dims, seed=3, kind="clustered")
    q[:6] = x[:6]
    for count, ef in [(10, 64), (100, 100)]:
        want = db.search_by_vector(q, count, ef=ef, counters=True)
        got = rd.nns(count).ef_search(ef).by_vectors_raw(q, counters=True)
        assert_same(got, want, f"lmdb route {metric}")
        assert np.array_equal(got[3][:, :6], want[3][:, :6])
    items = np.array([ids[0], ids[17], ids[-1], ids[-1] + 1], np.uint32)
    assert_same(rd.nns(5).by_items_raw(items), db.search_by_item(items, 5, ef=100), "lmdb route by_item")


# ---- flat-file snapshot cache ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric,dims", [("cosine", 100), ("hamming", 130), ("binary quantized euclidean", 64)])
def test_snapshot_file_roundtrip_on_the_host(tmp_path, metric, dims):
    """hb_index_save after the host half of finalize, hb_index_load into a fresh index: same ids / vectors / entry points /
    version; corrupt, truncated, foreign-distance and foreign-index files are refused."""
    n = 500
    ids = np.sort(np.random.default_rng(1).choice(1 << 24, n, replace=False)).astype(np.uint32)
    db, x = make_db(metric, n, dims, seed=4, ids=ids)
    lib = L.lib()
    mid = hb.reader._distance_of(metric).ID
    h = C.c_void_p()
    assert lib.hb_index_begin(mid, 2, C.byref(h)) == L.HB_OK
    fn = os.fsencode(str(tmp_path / "snap.hb"))
    for k, v in db.export_kv(2):
        assert lib.hb_index_push_kv(h, bytes(k), len(k), bytes(v), len(v)) == L.HB_OK
    assert lib.hb_index_save(h, fn) == L.HB_ESTATE                 # nothing decoded yet
    assert lib.hb_index_finalize(h, 0) in (L.HB_OK, L.HB_ECUDA)    # the host half runs either way
    assert lib.hb_index_save(h, fn) == L.HB_OK, lib.hb_last_error()
    g = C.c_void_p()
    assert lib.hb_index_begin(mid, 2, C.byref(g)) == L.HB_OK
    assert lib.hb_index_load(g, fn) == L.HB_OK, lib.hb_last_error()
    assert lib.hb_index_load(g, fn) == L.HB_ESTATE                 # not empty any more
    assert lib.hb_index_finalize(g, 0) in (L.HB_OK, L.HB_ECUDA)
    for a in ("hb_index_n_items", "hb_index_dimensions", "hb_index_max_level", "hb_index_n_entry_points"):
        assert getattr(lib, a)(g) == getattr(lib, a)(h), a
    ga, ha = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    lib.hb_index_item_ids(g, ga.ctypes.data_as(C.c_void_p), n)
    lib.hb_index_item_ids(h, ha.ctypes.data_as(C.c_void_p), n)
    assert np.array_equal(ga, ha) and np.array_equal(ga, ids)
    va, vb = np.zeros(dims, np.float32), np.zeros(dims, np.float32)
    for s in range(0, n, 37):
        assert lib.hb_index_item_vector(g, int(ids[s]), va.ctypes.data_as(C.c_void_p)) == L.HB_OK
        assert lib.hb_index_item_vector(h, int(ids[s]), vb.ctypes.data_as(C.c_void_p)) == L.HB_OK
        assert np.array_equal(va, vb)
    ver = [C.c_uint32() for _ in range(3)]
    assert lib.hb_index_version(g, *[C.byref(v) for v in ver]) == L.HB_OK and [v.value for v in ver] == [0, 1, 3]
    # saving the loaded index reproduces the file byte for byte
    fn2 = os.fsencode(str(tmp_path / "snap2.hb"))
    assert lib.hb_index_save(g, fn2) == L.HB_OK
    blob = open(fn, "rb").read()
    assert open(fn2, "rb").read() == blob
    lib.hb_index_free(g)
    lib.hb_index_free(h)

    def load(blob_, metric_id=mid, index=2):
        p = str(tmp_path / "x.hb")
        open(p, "wb").write(blob_)
        t = C.c_void_p()
        assert lib.hb_index_begin(metric_id, index, C.byref(t)) == L.HB_OK
        st = lib.hb_index_load(t, os.fsencode(p))
        lib.hb_index_free(t)
        return st
    assert load(blob) == L.HB_OK
    assert load(blob[:-9]) == L.HB_EFORMAT                                  # truncated
    flipped = bytearray(blob); flipped[len(blob) // 2] ^= 0x40
    assert load(bytes(flipped)) == L.HB_EFORMAT                             # checksum
    assert load(b"not a snapshot at all") == L.HB_EFORMAT
    assert load(blob, metric_id=(mid + 1) % 7) == L.HB_EUNMATCHING_DISTANCE
    assert load(blob, index=3) == L.HB_EINVAL


@pytest.mark.gpu
def test_snapshot_file_search_equals_oracle(tmp_path):
    db, x = make_db("cosine", 3000, 96, seed=6, kind="clustered")
    path = _write_env(tmp_path, {0: db})
    rd = hb.Reader.open_path(path, 0, "cosine")
    rd.save(str(tmp_path / "c.hb"))
    rd2 = hb.Reader.load(str(tmp_path / "c.hb"), 0, hb.Cosine)
    q = make_vectors(64, 96, seed=2, kind="clustered")
    want = db.search_by_vector(q, 10, ef=80, counters=True)
    for r in (rd, rd2):
        got = r.nns(10).ef_search(80).by_vectors_raw(q, counters=True)
        assert_same(got, want, "snapshot file")
        assert np.array_equal(got[3][:, :6], want[3][:, :6])
    with pytest.raises(hb.UnmatchingDistance):
        hb.Reader.load(str(tmp_path / "c.hb"), 0, hb.Euclidean)


# ---- randomized trees (hypothesis) ------------------------------------------------------------------------------------------
from hypothesis import given, settings, strategies as hst


@settings(max_examples=40, deadline=None)
@given(n=hst.integers(0, 1500), psize=hst.sampled_from([512, 1024, 4096, 8192, 65536]), fill=hst.floats(0.3, 1.0),
       big_every=hst.integers(0, 12), klen=hst.integers(1, 24), named=hst.booleans(), newest=hst.integers(0, 1), seed=hst.integers(0, 1 << 30),
       prefix_len=hst.integers(0, 3))
def test_scan_randomized_trees(tmp_path_factory, n, psize, fill, big_every, klen, named, newest, seed, prefix_len):
    """Any key length, page size, fill factor, overflow density, database naming and meta-page order: a prefix scan returns
    exactly the pairs whose key starts with the prefix, in key order."""
    pairs = _random_pairs(n, seed=seed, big_every=big_every, klen=klen)
    # keys longer than what fits a branch page next to another key are not valid LMDB keys (mdb_env_get_maxkeysize ~ psize / 8)
    if klen > psize // 8:
        return
    w = LmdbWriter(psize=psize, fill=fill)
    if named:
        w.put_named("db", pairs)
    else:
        w.put_unnamed(pairs)
    d = tmp_path_factory.mktemp("env")
    w.save(str(d), txnid=5, newest_meta=newest)
    prefix = pairs[len(pairs) // 2][0][:prefix_len] if pairs else b"\x01\x02\x03"[:prefix_len]
    st, got, txn = _scan(str(d), "db" if named else None, prefix)
    assert st == L.HB_OK, L.lib().hb_last_error()
    assert txn == 5
    assert got == [p for p in pairs if p[0].startswith(prefix)]


@settings(max_examples=60, deadline=None)
@given(seed=hst.integers(0, 1 << 30), n_flips=hst.integers(1, 40), region=hst.sampled_from(["meta", "any"]))
def test_corrupted_data_files_never_fault(tmp_path_factory, seed, n_flips, region):
    """Random byte damage anywhere in data.mdb (or concentrated in the meta pages): the walker answers with a status —
    every page number, node offset and size is bounds-checked — and the index decoder behind it never faults either."""
    rng = np.random.default_rng(seed)
    db, _ = make_db("euclidean", 120, 24, seed=1)
    d = tmp_path_factory.mktemp("env")
    fn = _write_env(d, {3: db}, fill=0.8)
    blob = bytearray(open(fn, "rb").read())
    hi = 2 * 4096 if region == "meta" else len(blob)
    for _ in range(n_flips):
        blob[int(rng.integers(0, hi))] = int(rng.integers(0, 256))
    open(fn, "wb").write(blob)
    st, got, _ = _scan(fn)
    assert st in (L.HB_OK, L.HB_EFORMAT, L.HB_EINVAL, L.HB_ESTATE)
    lib = L.lib()
    h = C.c_void_p()
    assert lib.hb_index_begin(0, 3, C.byref(h)) == L.HB_OK
    st = lib.hb_index_push_lmdb(h, os.fsencode(fn), None, None)
    assert st in (L.HB_OK, L.HB_EFORMAT, L.HB_EINVAL, L.HB_ESTATE)
    if st == L.HB_OK:
        assert lib.hb_index_finalize(h, 0) in (L.HB_OK, L.HB_ECUDA, L.HB_EFORMAT, L.HB_EMISSING_METADATA, L.HB_EUNMATCHING_DISTANCE, L.HB_ENEED_BUILD, L.HB_EINVAL)
    lib.hb_index_free(h)


def test_build_from_path_reads_the_items_then_needs_a_device(tmp_path):
    """Reader.build_from_path on an environment that holds what Writer::add_item left (items, Updated stones, no graph):
    the walker and the decoder do their part on the host; the build itself is device code (no CPU path)."""
    db, x = make_db("cosine", 300, 40, seed=6, build=True)
    pairs = [(bytes(k), bytes(v)) for k, v in db.export_kv(0)]
    never_built = sorted([(k, v) for k, v in pairs if k[2] == 3] + [(bytes([0, 0, 1]) + int(i).to_bytes(4, "big") + b"\0", b"") for i in range(0, 300, 7)])
    w = LmdbWriter(fill=0.7)
    w.put_named("vectors", never_built)
    w.save(str(tmp_path / "env"))
    import torch
    if torch.cuda.is_available():
        st = {}
        rd = hb.Reader.build_from_path(str(tmp_path / "env"), 0, hb.Cosine, db_name="vectors", dimensions=40, stats=st)
        assert rd.n_items() == 300 and st["items"] == 300
        q = make_vectors(20, 40, seed=1)
        ids, dist, lens = rd.nns(10).ef_search(64).by_vectors_raw(x[:20])
        assert (ids[:, 0] == np.arange(20)).all()
    else:
        with pytest.raises(hb.MissingMetadata):                   # never built and no dimensions given
            hb.Reader.build_from_path(str(tmp_path / "env"), 0, hb.Cosine, db_name="vectors")
        with pytest.raises(hb.HannoyError) as e:
            hb.Reader.build_from_path(str(tmp_path / "env"), 0, hb.Cosine, db_name="vectors", dimensions=40)
        assert e.value.status == L.HB_ECUDA and "no CUDA device" in str(e.value)
        with pytest.raises(hb.MissingMetadata):                   # the plain reader refuses it: nothing was ever built, reader.rs:390-393
            hb.Reader.open_path(str(tmp_path / "env"), 0, hb.Cosine, db_name="vectors")
        # items added after a build: metadata + graph + Updated stones -> NeedBuild for the reader (reader.rs:407-416)
        w2 = LmdbWriter()
        w2.put_unnamed(sorted(set(pairs) | {p for p in never_built if p[0][2] == 1}))
        w2.save(str(tmp_path / "env2"))
        with pytest.raises(hb.NeedBuild):
            hb.Reader.open_path(str(tmp_path / "env2"), 0, hb.Cosine)
