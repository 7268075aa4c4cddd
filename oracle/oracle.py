"""ctypes binding of the CPU oracle (oracle/hannoy_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package `hannoy_b200` never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

METRICS = {
    "euclidean": 0, "cosine": 1, "manhattan": 2, "hamming": 3,
    "binary quantized cosine": 4, "binary quantized euclidean": 5, "binary quantized manhattan": 6,
}
FLAG_FALLBACK, FLAG_LINEAR, FLAG_CANCELLED = 1, 2, 8


def build(force=False):
    src = os.path.join(_HERE, "hannoy_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        vp, u32, u64, f32, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_float, C.c_int
        sig = {
            "orc_db_new": (vp, [i32, u32]), "orc_db_free": (None, [vp]),
            "orc_db_add_items": (i32, [vp, vp, u64, vp]), "orc_db_add_rows": (i32, [vp, vp, u64, vp]),
            "orc_db_build": (i32, [vp, u32, u32, u32, f32, u64, i32]),
            "orc_db_set_links": (i32, [vp, u32, u32, vp, u32]),
            "orc_db_set_entry_points": (i32, [vp, vp, u32, u32]),
            "orc_db_set_csr": (i32, [vp, u32, vp, vp, u64]),
            "orc_db_n_items": (u64, [vp]), "orc_db_row_bytes": (u64, [vp]), "orc_db_max_level": (u32, [vp]),
            "orc_db_n_entry_points": (u32, [vp]), "orc_db_get_entry_points": (None, [vp, vp]),
            "orc_db_get_ids": (None, [vp, vp]), "orc_db_get_rows": (None, [vp, vp]),
            "orc_db_get_headers": (None, [vp, vp]), "orc_db_n_layers": (u32, [vp]),
            "orc_db_layer_nnz": (u64, [vp, u32]), "orc_db_get_layer": (None, [vp, u32, vp, vp]),
            "orc_search_by_vector": (i32, [vp, vp, u64, u32, u32, vp, u64, i32, u32, f32, vp, vp, vp, vp, i32]),
            "orc_search_by_item": (i32, [vp, vp, u64, u32, u32, vp, u64, i32, u32, f32, vp, vp, vp, vp, i32]),
            "orc_exact_knn": (i32, [vp, vp, u64, u32, vp, vp, i32]), "orc_set_cancel_after": (None, [u64]),
            "orc_dot_product": (f32, [vp, vp, u64]), "orc_euclidean": (f32, [vp, vp, u64]),
            "orc_dot_scalar": (f32, [vp, vp, u64]), "orc_euclid_scalar": (f32, [vp, vp, u64]),
            "orc_dot_sse": (f32, [vp, vp, u64]), "orc_euclid_sse": (f32, [vp, vp, u64]),
            "orc_quantize": (None, [i32, vp, u64, vp]), "orc_distance": (f32, [i32, vp, vp, u32]),
            "orc_ordered_float_cmp": (i32, [f32, f32]),
            "orc_roaring_serialize": (u64, [vp, u64, vp, u64]), "orc_roaring_deserialize": (C.c_int64, [vp, u64, vp, u64]),
            "orc_db_export_kv": (u64, [vp, C.c_uint16, vp, u64]),
            "orc_metric_name": (C.c_char_p, [i32]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


class OracleDb:
    """One hannoy index held in memory, built and searched by the restated reference algorithm."""

    def __init__(self, metric, dims):
        self.metric = METRICS[metric] if isinstance(metric, str) else int(metric)
        self.dims = int(dims)
        self.h = lib().orc_db_new(self.metric, self.dims)
        if not self.h:
            raise ValueError("bad metric")

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_db_free(self.h)
            self.h = None

    # Writer::add_item
    def add_items(self, ids, vecs):
        ids, vecs = _u32(ids), _f32(vecs).reshape(-1, self.dims)
        assert len(ids) == len(vecs)
        lib().orc_db_add_items(self.h, _p(ids), len(ids), _p(vecs))

    def add_rows(self, ids, rows):
        ids = _u32(ids)
        rows = np.ascontiguousarray(rows)
        assert rows.nbytes == len(ids) * self.row_bytes
        lib().orc_db_add_rows(self.h, _p(ids), len(ids), _p(rows))

    # Writer::builder(rng).ef_construction(efc).alpha(a).build::<M, M0>()
    def build(self, M=16, M0=32, ef_construction=100, alpha=1.0, seed=42, n_threads=1):
        rc = lib().orc_db_build(self.h, M, M0, ef_construction, alpha, seed, n_threads)
        if rc:
            raise ValueError("bad build parameters")

    def set_links(self, item, level, nbrs):
        nbrs = _u32(nbrs)
        lib().orc_db_set_links(self.h, item, level, _p(nbrs), len(nbrs))

    def set_entry_points(self, eps, max_level):
        eps = _u32(eps)
        if lib().orc_db_set_entry_points(self.h, _p(eps), len(eps), max_level):
            raise ValueError("entry point not in db")

    @property
    def n_items(self):
        return lib().orc_db_n_items(self.h)

    @property
    def row_bytes(self):
        return lib().orc_db_row_bytes(self.h)

    @property
    def max_level(self):
        return lib().orc_db_max_level(self.h)

    @property
    def entry_points(self):
        out = np.zeros(lib().orc_db_n_entry_points(self.h), np.uint32)
        lib().orc_db_get_entry_points(self.h, _p(out))
        return out

    def ids(self):
        out = np.zeros(self.n_items, np.uint32)
        lib().orc_db_get_ids(self.h, _p(out))
        return out

    def rows(self):
        """Encoded rows in slot (ascending id) order: f32 [n, dims] or u64 [n, words]."""
        n = self.n_items
        if self.metric >= 3:
            out = np.zeros((n, self.row_bytes // 8), np.uint64)
        else:
            out = np.zeros((n, self.dims), np.float32)
        lib().orc_db_get_rows(self.h, _p(out))
        return out

    def headers(self):
        out = np.zeros(self.n_items, np.float32)
        lib().orc_db_get_headers(self.h, _p(out))
        return out

    def layers(self):
        """[(offsets u64[n+1], neighbour ITEM IDS u32[nnz])] per level."""
        res = []
        n = self.n_items
        for l in range(lib().orc_db_n_layers(self.h)):
            off = np.zeros(n + 1, np.uint64)
            nbr = np.zeros(lib().orc_db_layer_nnz(self.h, l), np.uint32)
            lib().orc_db_get_layer(self.h, l, _p(off), _p(nbr))
            res.append((off, nbr))
        return res

    def _search(self, fn, qarr, nq, count, ef, candidates, linear_below, linear_below_ratio, n_threads, counters, cancel_after=0):
        lib().orc_set_cancel_after(cancel_after)  # cancel_fn = "true from its cancel_after-th call on" (0 = never)
        ids = np.zeros((nq, count), np.uint32)
        dist = np.zeros((nq, count), np.float32)
        lens = np.zeros(nq, np.uint32)
        ctr = np.zeros((nq, 8), np.uint64) if counters else None
        cand = _u32(candidates) if candidates is not None else None
        fn(self.h, _p(qarr), nq, count, ef, _p(cand), 0 if cand is None else len(cand), int(cand is not None),
           linear_below, linear_below_ratio, _p(ids), _p(dist), _p(lens), _p(ctr), n_threads)
        return ids, dist, lens, ctr

    def search_by_vector(self, q, count, ef=100, candidates=None, linear_below=1000, linear_below_ratio=1.0,
                         n_threads=1, counters=False, cancel_after=0):
        """reader.nns(count)[.ef_search()].by_vector for each row of q. `ef` is the raw
        QueryBuilder.ef field (callers apply ef_search's max(ef, count) themselves)."""
        q = _f32(q).reshape(-1, self.dims)
        return self._search(lib().orc_search_by_vector, q, len(q), count, ef, candidates, linear_below,
                            linear_below_ratio, n_threads, counters, cancel_after)

    def search_by_item(self, items, count, ef=100, candidates=None, linear_below=1000, linear_below_ratio=1.0,
                       n_threads=1, counters=False, cancel_after=0):
        """by_item for each id; lens == 0xFFFFFFFF encodes `None`."""
        items = _u32(items)
        return self._search(lib().orc_search_by_item, items, len(items), count, ef, candidates, linear_below,
                            linear_below_ratio, n_threads, counters, cancel_after)

    def exact_knn(self, q, k, n_threads=1):
        q = _f32(q).reshape(-1, self.dims)
        ids = np.zeros((len(q), k), np.uint32)
        dist = np.zeros((len(q), k), np.float32)
        lib().orc_exact_knn(self.h, _p(q), len(q), k, _p(ids), _p(dist), n_threads)
        return ids, dist

    def export_kv(self, index=0):
        """[(key bytes, value bytes)] in LMDB key order, in the reference's on-disk encoding."""
        n = lib().orc_db_export_kv(self.h, index, None, 0)
        buf = np.zeros(n, np.uint8)
        lib().orc_db_export_kv(self.h, index, _p(buf), n)
        raw = buf.tobytes()
        out, pos = [], 0
        while pos < len(raw):
            kl = int.from_bytes(raw[pos:pos + 4], "little"); pos += 4
            k = raw[pos:pos + kl]; pos += kl
            vl = int.from_bytes(raw[pos:pos + 4], "little"); pos += 4
            v = raw[pos:pos + vl]; pos += vl
            out.append((k, v))
        return out


def dot_product(a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().orc_dot_product(_p(a), _p(b), len(a)))


def euclidean(a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().orc_euclidean(_p(a), _p(b), len(a)))


def distance(metric, a, b):
    a, b = _f32(a), _f32(b)
    m = METRICS[metric] if isinstance(metric, str) else metric
    return float(lib().orc_distance(m, _p(a), _p(b), len(a)))


def quantize(v, binary_codec):
    v = _f32(v)
    out = np.zeros((len(v) + 63) // 64, np.uint64)
    lib().orc_quantize(int(binary_codec), _p(v), len(v), _p(out))
    return out


def roaring_serialize(sorted_ids):
    a = _u32(sorted_ids)
    n = lib().orc_roaring_serialize(_p(a), len(a), None, 0)
    out = np.zeros(n, np.uint8)
    lib().orc_roaring_serialize(_p(a), len(a), _p(out), n)
    return out.tobytes()


def roaring_deserialize(b):
    buf = np.frombuffer(b, np.uint8)
    n = lib().orc_roaring_deserialize(_p(buf), len(buf), None, 0)
    if n < 0:
        raise ValueError("bad roaring bytes")
    out = np.zeros(n, np.uint32)
    lib().orc_roaring_deserialize(_p(buf), len(buf), _p(out), n)
    return out
