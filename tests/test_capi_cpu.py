"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/hannoy_b200.h declares,
decodes the reference's KV encoding, applies the Reader::open checks, and refuses to compute without a GPU
(no CPU fallback).  No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import make_db
from hannoy_b200 import _lib as L
import hannoy_b200 as hb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "hannoy_b200.h")).read()
    declared = set(re.findall(r"\b(hb_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)
    dev_hdr = open(os.path.join(ROOT, "include", "hannoy_b200_dev.h")).read()
    dev_declared = set(re.findall(r"\b(hb_[a-z_0-9]+)\s*\(", dev_hdr))
    assert dev_declared == set(L.DEV_EXPORTS), dev_declared ^ set(L.DEV_EXPORTS)
    lib = C.CDLL(L.SO_PATH)
    for name in declared | dev_declared:
        assert hasattr(lib, name), name


def test_metric_names_are_the_reference_strings():
    names = ["euclidean", "cosine", "manhattan", "hamming", "binary quantized cosine", "binary quantized euclidean",
             "binary quantized manhattan"]
    for i, n in enumerate(names):
        assert L.lib().hb_metric_name(i).decode() == n
        assert L.lib().hb_metric_from_name(n.encode()) == i
    assert L.lib().hb_metric_from_name(b"dot") == -1


def _begin(metric_id, index=0):
    h = C.c_void_p()
    assert L.lib().hb_index_begin(metric_id, index, C.byref(h)) == L.HB_OK
    return h


def _push_all(h, kv):
    for k, v in kv:
        st = L.lib().hb_index_push_kv(h, k, len(k), v, len(v))
        assert st == L.HB_OK, L.lib().hb_last_error()


def test_open_checks_missing_metadata_unmatching_distance_need_build():
    db, _ = make_db("cosine", 50, 12, seed=1)
    kv = db.export_kv(0)
    lib = L.lib()
    # MissingMetadata — reader.rs:390-393
    h = _begin(1)
    _push_all(h, [p for p in kv if p[0][2] != 0])
    assert lib.hb_index_finalize(h, 0) == L.HB_EMISSING_METADATA
    lib.hb_index_free(h)
    # UnmatchingDistance — reader.rs:400-405
    h = _begin(0)
    _push_all(h, kv)
    assert lib.hb_index_finalize(h, 0) == L.HB_EUNMATCHING_DISTANCE
    assert b"unmatching distance" in lib.hb_last_error()
    lib.hb_index_free(h)
    # NeedBuild: an Updated stone is present — reader.rs:407-416
    h = _begin(1)
    _push_all(h, kv + [(bytes([0, 0, 1, 0, 0, 0, 9, 0]), bytes([0]))])
    assert lib.hb_index_finalize(h, 0) == L.HB_ENEED_BUILD
    lib.hb_index_free(h)
    # pairs of another index are ignored -> metadata missing for index 5
    h = _begin(1, index=5)
    _push_all(h, kv)
    assert lib.hb_index_finalize(h, 0) == L.HB_EMISSING_METADATA
    lib.hb_index_free(h)


def test_bad_pairs_are_rejected():
    lib = L.lib()
    h = _begin(0)
    assert lib.hb_index_push_kv(h, b"short", 5, b"x", 1) == L.HB_EFORMAT
    assert lib.hb_index_push_kv(h, bytes([0, 0, 9, 0, 0, 0, 0, 0]), 8, b"x", 1) == L.HB_EFORMAT   # unknown NodeMode
    assert lib.hb_index_push_kv(h, bytes([0, 0, 2, 0, 0, 0, 0, 0]), 8, bytes([1, 9, 9, 9]), 4) == L.HB_EFORMAT  # bad roaring
    assert lib.hb_index_push_kv(h, bytes([0, 0, 3, 0, 0, 0, 0, 0]), 8, bytes([0, 0, 0, 0, 0, 1, 2]), 7) == L.HB_EFORMAT  # SizeMismatch
    lib.hb_index_free(h)


@pytest.mark.parametrize("metric", ["euclidean", "cosine", "hamming", "binary quantized cosine"])
def test_kv_decode_matches_what_was_written(metric):
    """push_kv + the host half of finalize reproduce ids / dims / entry points / vectors; without a GPU finalize then
    stops with HB_ECUDA and search refuses to run."""
    n, dims = 300, 70
    ids = np.sort(np.random.default_rng(4).choice(1 << 30, n, replace=False)).astype(np.uint32)
    db, x = make_db(metric, n, dims, seed=3, ids=ids)
    lib = L.lib()
    mid = hb.reader._distance_of(metric).ID
    h = _begin(mid, index=2)
    _push_all(h, db.export_kv(2))
    st = lib.hb_index_finalize(h, 0)
    if _has_gpu():
        assert st == L.HB_OK
    else:
        assert st == L.HB_ECUDA, "the product must fail loudly without a CUDA device"
        assert b"no CUDA device" in lib.hb_last_error()
    assert lib.hb_index_n_items(h) == n and lib.hb_index_dimensions(h) == dims
    assert lib.hb_index_max_level(h) == db.max_level
    assert lib.hb_index_n_entry_points(h) == len(db.entry_points)
    got = np.zeros(n, np.uint32)
    lib.hb_index_item_ids(h, got.ctypes.data_as(C.c_void_p), n)
    assert np.array_equal(got, ids)
    v = np.zeros(dims, np.float32)
    for s in (0, 17, n - 1):
        assert lib.hb_index_item_vector(h, int(ids[s]), v.ctypes.data_as(C.c_void_p)) == L.HB_OK
        if metric in ("euclidean", "cosine"):
            assert np.array_equal(v, x[s])
        elif metric == "hamming":
            assert np.array_equal(v, (x[s] > 0).astype(np.float32))
        else:
            assert np.array_equal(v, np.where(np.signbit(x[s]), -1.0, 1.0).astype(np.float32))
    assert lib.hb_index_contains_item(h, int(ids[3])) == 1 and lib.hb_index_contains_item(h, int(ids[3]) + 1) in (0, 1)
    if not _has_gpu():
        out = np.zeros(10, np.uint32)
        q = np.zeros(dims, np.float32)
        st = lib.hb_search_by_vector(h, q.ctypes.data_as(C.c_void_p), 1, dims, 10, 100, None, out.ctypes.data_as(C.c_void_p),
                                     out.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), None)
        assert st == L.HB_ESTATE  # not finalized -> no search, and no CPU path to fall back to
    lib.hb_index_free(h)


@pytest.mark.parametrize("order", ["sorted", "reversed", "shuffled"])
def test_push_kv_accepts_any_order(order):
    """Sorted (LMDB cursor) order sees the metadata first and writes rows straight into their slots; any other order
    is staged and resolved at finalize.  Same snapshot either way."""
    n, dims = 200, 33
    db, x = make_db("cosine", n, dims, seed=5)
    kv = sorted((bytes(k), bytes(v)) for k, v in db.export_kv(1))
    if order == "reversed":
        kv = kv[::-1]
    elif order == "shuffled":
        kv = [kv[i] for i in np.random.default_rng(0).permutation(len(kv))]
    lib = L.lib()
    h = _begin(1, index=1)
    _push_all(h, kv)
    assert lib.hb_index_finalize(h, 0) in (L.HB_OK, L.HB_ECUDA)
    assert lib.hb_index_n_items(h) == n
    v = np.zeros(dims, np.float32)
    for s in range(0, n, 7):
        assert lib.hb_index_item_vector(h, s, v.ctypes.data_as(C.c_void_p)) == L.HB_OK
        assert np.array_equal(v, x[s])
    lib.hb_index_free(h)


def test_python_mirror_surface():
    """The host mirror keeps the reference names: Reader.open/nns, QueryBuilder.ef_search/candidates/linear_below(_ratio)/
    by_vector/by_item, Searched.into_nns/did_cancel, and the defaults of reader.rs:23-32."""
    for name in ("open", "nns", "dimensions", "n_entrypoints", "n_items", "item_ids", "index", "version", "item_vector",
                 "is_empty", "contains_item"):
        assert hasattr(hb.Reader, name)
    for name in ("ef_search", "candidates", "linear_below", "linear_below_ratio", "by_vector", "by_item",
                 "by_vector_with_cancellation", "by_item_with_cancellation"):
        assert hasattr(hb.QueryBuilder, name)
    qb = hb.QueryBuilder(None, 150)
    assert qb.ef == 100 and qb._linear_below == 1000 and qb._linear_below_ratio == 1.0
    assert qb.ef_search(20).ef == 150        # ef_search stores max(ef, count) — reader.rs:217-220
    assert qb.ef_search(300).ef == 300
    with pytest.raises(AssertionError):
        qb.linear_below_ratio(1.5)
    s = hb.Searched([(1, 0.5)], False)
    assert s.into_nns() == [(1, 0.5)] and s.did_cancel() is False
    assert [d.name() for d in (hb.Euclidean, hb.Cosine, hb.Manhattan, hb.Hamming)] == ["euclidean", "cosine", "manhattan", "hamming"]


def _roaring_with_runs(ids):
    """RoaringFormatSpec, cookie 12347: every container run-encoded (what `optimize()`d bitmaps and other writers emit;
    roaring-rs reads them too)."""
    import struct
    ids = sorted(set(int(i) for i in ids))
    if not ids:
        return struct.pack("<II", 12346, 0)
    groups = {}
    for i in ids:
        groups.setdefault(i >> 16, []).append(i & 0xffff)
    keys = sorted(groups)
    n = len(keys)
    out = struct.pack("<I", 12347 | ((n - 1) << 16)) + bytes([0xff] * ((n + 7) // 8))
    bodies = []
    for k in keys:
        lo = groups[k]
        runs, s, prev = [], lo[0], lo[0]
        for v in lo[1:]:
            if v != prev + 1:
                runs.append((s, prev - s))
                s = v
            prev = v
        runs.append((s, prev - s))
        out += struct.pack("<HH", k, len(lo) - 1)
        bodies.append(struct.pack("<H", len(runs)) + b"".join(struct.pack("<HH", a, b) for a, b in runs))
    if n >= 4:  # NO_OFFSET_THRESHOLD
        pos = len(out) + 4 * n
        for b in bodies:
            out += struct.pack("<I", pos)
            pos += len(b)
    return out + b"".join(bodies)


def test_kv_decode_bitmap_and_run_containers():
    """A dense index of more than 4096 ids per 64K chunk stores `items` as a Roaring BITMAP container (every real index
    does); Links written by a run-optimising encoder use RUN containers.  Both decode to the same snapshot."""
    n, dims = 9000, 4
    ids = np.concatenate([np.arange(n - 3, dtype=np.uint32), np.array([70000, 70001, 200000], np.uint32)])
    db, x = make_db("euclidean", n, dims, seed=2, ids=ids, M=6, M0=12, efc=24, n_threads=4)
    kv = [(bytes(k), bytes(v)) for k, v in db.export_kv(0)]
    meta = [v for k, v in kv if k[2] == 0 and k[3:7] == b"\0\0\0\0"][0]
    roaring = meta[len(b"euclidean") + 1 + 8:]
    assert roaring[:4] == (12346).to_bytes(4, "little")   # no-run cookie; first container card 8997 > 4096 -> bitmap
    from oracle import oracle as O
    lib = L.lib()
    handles = []
    for rewrite in (False, True):
        h = _begin(0)
        for k, v in kv:
            if rewrite and k[2] == 2:      # Links: [1][roaring] -> same set, run containers
                v = b"\x01" + _roaring_with_runs(O.roaring_deserialize(v[1:]))
            if rewrite and k[2] == 0 and k[3:7] == b"\0\0\0\0":  # metadata `items`, run-encoded (3 containers)
                name_end = v.index(b"\0") + 1
                items_size = int.from_bytes(v[name_end + 4:name_end + 8], "big")
                r = _roaring_with_runs(ids)
                v = v[:name_end + 4] + len(r).to_bytes(4, "big") + r + v[name_end + 8 + items_size:]
            assert lib.hb_index_push_kv(h, k, len(k), v, len(v)) == L.HB_OK, lib.hb_last_error()
        assert lib.hb_index_finalize(h, 0) in (L.HB_OK, L.HB_ECUDA), lib.hb_last_error()
        got = np.zeros(n, np.uint32)
        assert lib.hb_index_n_items(h) == n
        lib.hb_index_item_ids(h, got.ctypes.data_as(C.c_void_p), n)
        assert np.array_equal(got, ids)
        handles.append(h)
    # identical snapshots: the cache files are byte-identical
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        blobs = []
        for i, h in enumerate(handles):
            fn = os.path.join(d, f"s{i}.hb").encode()
            assert lib.hb_index_save(h, fn) == L.HB_OK
            blobs.append(open(fn, "rb").read())
            lib.hb_index_free(h)
        assert blobs[0] == blobs[1]
    # 5 containers -> the run cookie carries an offset header
    many = np.array([1, 2, 3, 65536 + 9, 3 * 65536 + 1, 3 * 65536 + 2, 9 * 65536, 11 * 65536 + 5], np.uint32)
    assert np.array_equal(O.roaring_deserialize(_roaring_with_runs(many)), many)


@pytest.mark.parametrize("metric,dims,n,dense", [("cosine", 33, 600, False), ("hamming", 70, 300, False),
                                                 ("binary quantized cosine", 100, 5000, False), ("euclidean", 8, 9000, True)])
def test_export_kv_writes_the_reference_encoding(metric, dims, n, dense):
    """hb_index_export_kv (metadata with array / bitmap containers, version, one Links node per (item, layer), items) is
    byte for byte what the oracle's independent encoder of the reference format writes for the same graph — and feeding
    it back through push_kv reproduces the snapshot."""
    ids = np.arange(n, dtype=np.uint32) if dense else np.sort(np.random.default_rng(1).choice(1 << 22, n, replace=False)).astype(np.uint32)
    db, x = make_db(metric, n, dims, seed=3, ids=ids, M=8, M0=16, efc=32, n_threads=4)
    lib = L.lib()
    d = hb.reader._distance_of(metric)
    h = _begin(d.ID, index=5)
    layers = db.layers()
    offs = [np.ascontiguousarray(o, dtype=np.uint64) for o, _ in layers]
    nbrs = [np.ascontiguousarray(b, dtype=np.uint32) for _, b in layers]
    off_pp = (C.c_void_p * len(layers))(*[o.ctypes.data for o in offs])
    nbr_pp = (C.c_void_p * len(layers))(*[b.ctypes.data for b in nbrs])
    rows, hdr = np.ascontiguousarray(db.rows()), np.ascontiguousarray(db.headers(), dtype=np.float32)
    eps = np.ascontiguousarray(db.entry_points, dtype=np.uint32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.hb_index_from_arrays(h, dims, vp(ids), n, vp(rows), vp(hdr), len(layers), C.cast(off_pp, C.c_void_p),
                                    C.cast(nbr_pp, C.c_void_p), vp(eps), len(eps), db.max_level) == L.HB_OK, lib.hb_last_error()
    out = []

    def cb(u, k, kl, v, vl):
        out.append((bytes(k[:kl]), bytes(v[:vl])))
        return 0
    assert lib.hb_index_export_kv(h, 1, L.KV_VISIT(cb), None) == L.HB_OK, lib.hb_last_error()
    want = [(bytes(k), bytes(v)) for k, v in db.export_kv(5)]
    assert out == want
    assert lib.hb_index_export_kv(h, 1, L.KV_VISIT(lambda *a: 1), None) == L.HB_ESTATE   # a visitor may stop it
    g = _begin(d.ID, index=5)
    _push_all(g, out)
    assert lib.hb_index_finalize(g, 0) in (L.HB_OK, L.HB_ECUDA)
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        fa, fb = os.path.join(tmp, "a.hb").encode(), os.path.join(tmp, "b.hb").encode()
        assert lib.hb_index_save(h, fa) == L.HB_OK and lib.hb_index_save(g, fb) == L.HB_OK
        assert open(fa, "rb").read() == open(fb, "rb").read()
    if not _has_gpu():   # the builder itself is device code: no CPU path
        e = _begin(d.ID)
        assert lib.hb_index_from_arrays(e, dims, vp(ids), n, vp(rows), vp(hdr), 0, None, None, None, 0, 0) == L.HB_OK
        assert lib.hb_index_build_graph(e, None, 0, None) == L.HB_ECUDA
        lib.hb_index_free(e)
    lib.hb_index_free(g)
    lib.hb_index_free(h)


# ---- hostile input never crashes the decoders ---------------------------------------------------------------------------------
from hypothesis import given, settings, strategies as hst


@settings(max_examples=300, deadline=None)
@given(mode=hst.sampled_from([0, 1, 2, 3]), item=hst.integers(0, 3), layer=hst.integers(0, 3), val=hst.binary(min_size=0, max_size=96),
       metric=hst.sampled_from([0, 1, 3, 4]))
def test_push_kv_survives_arbitrary_values(mode, item, layer, val, metric):
    """Any byte string as a value: HB_OK or HB_EFORMAT, never a fault (metadata / version / roaring / item decoders)."""
    lib = L.lib()
    h = _begin(metric, index=1)
    key = bytes([0, 1, mode]) + item.to_bytes(4, "big") + bytes([layer])
    st = lib.hb_index_push_kv(h, key, len(key), val, len(val))
    assert st in (L.HB_OK, L.HB_EFORMAT), st
    # a plausible roaring prefix followed by garbage exercises the container bounds checks
    r = (12346).to_bytes(4, "little") + (3).to_bytes(4, "little") + val
    st = lib.hb_index_push_kv(h, bytes([0, 1, 2, 0, 0, 0, 5, 0]), 8, b"\x01" + r, 1 + len(r))
    assert st in (L.HB_OK, L.HB_EFORMAT)
    r = (12347 | (2 << 16)).to_bytes(4, "little") + b"\x07" + val
    st = lib.hb_index_push_kv(h, bytes([0, 1, 2, 0, 0, 0, 6, 0]), 8, b"\x01" + r, 1 + len(r))
    assert st in (L.HB_OK, L.HB_EFORMAT)
    assert lib.hb_index_finalize(h, 0) in (L.HB_OK, L.HB_ECUDA, L.HB_EFORMAT, L.HB_EMISSING_METADATA, L.HB_EUNMATCHING_DISTANCE, L.HB_ENEED_BUILD, L.HB_EINVAL)
    lib.hb_index_free(h)
