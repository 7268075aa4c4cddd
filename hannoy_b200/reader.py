"""Reader / QueryBuilder / Searched / Distance — the reference's search API (src/reader.rs:36-261,
374-431,545-620; src/distance/mod.rs:26-48) over the C-ABI.

Differences from the Rust API, all forced by the snapshot design:
  * `Reader.open(kv_pairs, index, distance)` takes an iterable of raw `(key, value)` byte pairs — what a
    heed cursor over the LMDB read transaction yields — instead of `(&RoTxn, index, Database<D>)`;
    the transaction is read once, queries need no `rtxn`.
  * batched `by_vectors` / `by_items` are added; `by_vector` / `by_item` are the single-query forms.
  * cancellation closures cannot cross the ABI as such: a `CancelToken` (a device flag any host thread may trip) or a
    poll count stand in for `cancel_fn`; `Searched.did_cancel()` reports it like the reference.
  * `Reader.replicate(devices)` copies the snapshot to further GPUs; every batched call then partitions its batch.
"""
import ctypes as C
import os

import numpy as np

from . import _lib as L

DEFAULT_EF_SEARCH = 100                     # reader.rs:23
DEFAULT_LINEAR_SCAN_THRESHOLD = 1000        # reader.rs:29
DEFAULT_LINEAR_SCAN_THRESHOLD_RATIO = 1.00  # reader.rs:32


# ---- errors (src/error.rs) -----------------------------------------------------------------------
class HannoyError(Exception):
    def __init__(self, msg, status=None):
        super().__init__(msg)
        self.status = status


class InvalidVecDimension(HannoyError):
    pass


class MissingMetadata(HannoyError):
    pass


class UnmatchingDistance(HannoyError):
    pass


class NeedBuild(HannoyError):
    pass


_ERR = {L.HB_EDIM: InvalidVecDimension, L.HB_EMISSING_METADATA: MissingMetadata,
        L.HB_EUNMATCHING_DISTANCE: UnmatchingDistance, L.HB_ENEED_BUILD: NeedBuild}


def _check(status):
    if status != L.HB_OK:
        msg = L.lib().hb_last_error().decode("utf-8", "replace")
        raise _ERR.get(status, HannoyError)(msg or f"hb_status {status}", status)


# ---- distances (src/distance/*.rs) --------------------------------------------------------------------
class Distance:
    """`trait Distance` as far as search needs it: the name stored in the metadata, and the codec."""
    ID = -1

    @classmethod
    def name(cls):
        return L.lib().hb_metric_name(cls.ID).decode()

    @classmethod
    def is_binary(cls):
        return cls.ID >= 3


class Euclidean(Distance):
    ID = 0


class Cosine(Distance):
    ID = 1


class Manhattan(Distance):
    ID = 2


class Hamming(Distance):
    ID = 3


class BinaryQuantizedCosine(Distance):
    ID = 4


class BinaryQuantizedEuclidean(Distance):
    ID = 5


class BinaryQuantizedManhattan(Distance):
    ID = 6


def _distance_of(d):
    if isinstance(d, type) and issubclass(d, Distance):
        return d
    if isinstance(d, str):
        for cls in Distance.__subclasses__():
            if cls.name() == d:
                return cls
    if isinstance(d, int):
        for cls in Distance.__subclasses__():
            if cls.ID == d:
                return cls
    raise ValueError(f"unknown distance {d!r}")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Searched:
    """reader.rs:36-57"""
    __slots__ = ("nns", "did_cancel_")

    def __init__(self, nns, did_cancel=False):
        self.nns = nns
        self.did_cancel_ = did_cancel

    def did_cancel(self):
        return self.did_cancel_

    def into_nns(self):
        return self.nns

    def __repr__(self):
        return f"Searched(nns={self.nns!r}, did_cancel={self.did_cancel_})"


class CancelToken:
    """What a `cancel_fn` closure becomes on this side of the ABI (reader.rs:91-188): a flag on the device that any host
    thread may trip while searches carrying it are in flight (hb_cancel_token_*)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        _check(L.lib().hb_cancel_token_create(device, C.byref(self._h)))

    def cancel(self):
        _check(L.lib().hb_cancel_token_cancel(self._h))

    def reset(self):
        _check(L.lib().hb_cancel_token_reset(self._h))

    def is_cancelled(self):
        return bool(L.lib().hb_cancel_token_is_cancelled(self._h))

    def close(self):
        if getattr(self, "_h", None):
            L.lib().hb_cancel_token_free(self._h)
            self._h = None

    def __del__(self):
        self.close()


class QueryBuilder:
    """reader.rs:60-261"""

    def __init__(self, reader, count):
        self.reader = reader
        self.count = int(count)
        self.ef = DEFAULT_EF_SEARCH
        self._candidates = None
        self._linear_below = DEFAULT_LINEAR_SCAN_THRESHOLD
        self._linear_below_ratio = DEFAULT_LINEAR_SCAN_THRESHOLD_RATIO
        self._cancel = None        # CancelToken
        self._cancel_after = 0     # deterministic cancel_fn: true from its N-th call on

    def ef_search(self, ef):  # reader.rs:217-220
        self.ef = max(int(ef), self.count)
        return self

    def candidates(self, candidates):  # reader.rs:200-203
        self._candidates = np.ascontiguousarray(np.asarray(list(candidates) if not isinstance(candidates, np.ndarray) else candidates, dtype=np.uint32))
        return self

    def linear_below(self, threshold):  # reader.rs:234-237
        self._linear_below = int(threshold)
        return self

    def linear_below_ratio(self, ratio):  # reader.rs:252-260
        assert 0.0 <= ratio <= 1.0, "linear scan threshold ratio must be between 0.0 and 1.0"
        self._linear_below_ratio = float(ratio)
        return self

    def _opts(self):
        o = L.QueryOpts()
        if self._candidates is not None:
            o.candidates = self._candidates.ctypes.data if len(self._candidates) else None
            o.n_candidates = len(self._candidates)
            o.has_candidates = 1
        o.linear_below = self._linear_below
        o.linear_below_ratio = self._linear_below_ratio
        o.cancel = self._cancel._h if self._cancel is not None else None
        o.cancel_after_polls = self._cancel_after
        return o

    # -- batched forms (raw arrays) --
    def by_vectors_raw(self, vectors, counters=False):
        """-> (ids [nq,count] u32, dist [nq,count] f32, len [nq] u32[, counters [nq,8] u64])"""
        r = self.reader
        q = np.ascontiguousarray(vectors, dtype=np.float32)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        nq, dims = q.shape
        ids = np.zeros((nq, self.count), np.uint32)
        dist = np.zeros((nq, self.count), np.float32)
        lens = np.zeros(nq, np.uint32)
        ctr = np.zeros((nq, L.HB_N_CTR), np.uint64) if counters else None
        o = self._opts()
        _check(L.lib().hb_search_by_vector(r._h, _ptr(q), nq, dims, self.count, self.ef, C.byref(o), _ptr(ids),
                                           _ptr(dist), _ptr(lens), _ptr(ctr)))
        return (ids, dist, lens, ctr) if counters else (ids, dist, lens)

    def by_items_raw(self, items, counters=False):
        """-> like by_vectors_raw; len == 0xFFFFFFFF encodes `None`"""
        r = self.reader
        it = np.ascontiguousarray(items, dtype=np.uint32).reshape(-1)
        nq = len(it)
        ids = np.zeros((nq, self.count), np.uint32)
        dist = np.zeros((nq, self.count), np.float32)
        lens = np.zeros(nq, np.uint32)
        ctr = np.zeros((nq, L.HB_N_CTR), np.uint64) if counters else None
        o = self._opts()
        _check(L.lib().hb_search_by_item(r._h, _ptr(it), nq, self.count, self.ef, C.byref(o), _ptr(ids), _ptr(dist),
                                         _ptr(lens), _ptr(ctr)))
        return (ids, dist, lens, ctr) if counters else (ids, dist, lens)

    @staticmethod
    def _searched(ids, dist, n):
        n = int(n)
        return Searched([(int(ids[j]), float(dist[j])) for j in range(n & 0x7fffffff)], bool(n & L.LEN_CANCELLED))

    def by_vectors(self, vectors):
        ids, dist, lens = self.by_vectors_raw(vectors)
        return [self._searched(ids[i], dist[i], lens[i]) for i in range(len(lens))]

    def by_items(self, items):
        ids, dist, lens = self.by_items_raw(items)
        return [None if lens[i] == 0xFFFFFFFF else self._searched(ids[i], dist[i], lens[i]) for i in range(len(lens))]

    # -- the reference's single-query forms --
    def by_vector(self, vector, rtxn=None):  # reader.rs:132-148
        v = np.asarray(vector, dtype=np.float32).reshape(-1)
        if len(v) != self.reader.dimensions():
            raise InvalidVecDimension(
                f"Invalid vector dimensions. Got {len(v)} but expected {self.reader.dimensions()}", L.HB_EDIM)
        return self.by_vectors(v.reshape(1, -1))[0]

    def by_item(self, item, rtxn=None):  # reader.rs:81-89
        return self.by_items([item])[0]

    def with_cancellation(self, cancel_fn):
        """The batched spelling of *_with_cancellation: every query of the next by_vectors / by_items call polls it."""
        self._cancel, self._cancel_after = None, 0
        if isinstance(cancel_fn, CancelToken):
            self._cancel = cancel_fn
        elif isinstance(cancel_fn, int) and not isinstance(cancel_fn, bool):
            self._cancel_after = int(cancel_fn)
        elif cancel_fn is not None:
            raise TypeError("pass a CancelToken, a poll count, or use by_*_with_cancellation for a callable")
        return self

    def _run_cancellable(self, cancel_fn, run):
        """reader.rs:108-118,167-188.  `cancel_fn` is a CancelToken, an int N (the closure that turns true at its N-th
        call: polled exactly where the reference polls), or any callable — evaluated by a watcher thread on the host
        while the kernels run, tripping a token when it first returns True."""
        if not callable(cancel_fn):
            return run(self.with_cancellation(cancel_fn))
        import threading
        tok = CancelToken(self.reader._device)
        done = threading.Event()

        def watch():
            while not done.is_set():
                if cancel_fn():
                    tok.cancel()
                    return
                done.wait(50e-6)

        t = threading.Thread(target=watch, daemon=True)
        t.start()
        try:
            return run(self.with_cancellation(tok))
        finally:
            done.set()
            t.join()
            self._cancel = None
            tok.close()

    def by_vector_with_cancellation(self, vector, cancel_fn):
        return self._run_cancellable(cancel_fn, lambda qb: qb.by_vector(vector))

    def by_item_with_cancellation(self, item, cancel_fn):
        return self._run_cancellable(cancel_fn, lambda qb: qb.by_item(item))


class Reader:
    """reader.rs:374-620.  Holds the HBM-resident snapshot of one hannoy index."""

    def __init__(self, handle, distance, index, device=0):
        self._h = handle
        self._distance = distance
        self._index = index
        self._device = device

    def __del__(self):
        self.close()

    def close(self):
        if getattr(self, "_h", None):
            L.lib().hb_index_free(self._h)
            self._h = None

    @classmethod
    def open(cls, kv_pairs, index, distance, device=0):
        """Reader::open (reader.rs:387-431): `kv_pairs` = iterable of raw LMDB (key, value) byte pairs."""
        d = _distance_of(distance)
        lib = L.lib()
        h = C.c_void_p()
        _check(lib.hb_index_begin(d.ID, index, C.byref(h)))
        try:
            for k, v in kv_pairs:
                _check(lib.hb_index_push_kv(h, bytes(k), len(k), bytes(v), len(v)))
            _check(lib.hb_index_finalize(h, device))
        except Exception:
            lib.hb_index_free(h)
            raise
        return cls(h, d, index, device)

    @classmethod
    def open_path(cls, path, index, distance, db_name=None, device=0):
        """Reader::open from an LMDB environment on disk (directory holding data.mdb, or the data file): the library
        walks the B+tree itself (hb_index_open_lmdb) — the `(&RoTxn, index, Database<D>)` triple of reader.rs:387
        becomes `(path, db_name)`; `db_name=None` is the unnamed database."""
        d = _distance_of(distance)
        h = C.c_void_p()
        _check(L.lib().hb_index_open_lmdb(os.fsencode(path), db_name.encode() if db_name else None, d.ID, index, device,
                                          C.byref(h)))
        return cls(h, d, index, device)

    @classmethod
    def build(cls, distance, dims, ids, rows, hdr, M=16, M0=32, ef_construction=100, alpha=1.0, seed=42, batch_max=0, index=0,
              device=0, stats=None):
        """HannoyBuilder::build on the device (hb_index_build_graph) over items given as flat arrays (rows / hdr as in
        from_arrays), then Reader::open on the result.  `stats`: optional dict filled with batches / launches / max_level."""
        d = _distance_of(distance)
        lib = L.lib()
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        rows = np.ascontiguousarray(rows)
        assert rows.dtype == (np.uint64 if d.is_binary() else np.float32)
        hdr_a = None if hdr is None else np.ascontiguousarray(hdr, dtype=np.float32)
        h = C.c_void_p()
        _check(lib.hb_index_begin(d.ID, index, C.byref(h)))
        try:
            none_pp = (C.c_void_p * 1)()
            _check(lib.hb_index_from_arrays(h, dims, _ptr(ids), len(ids), _ptr(rows), _ptr(hdr_a), 0, C.cast(none_pp, C.c_void_p),
                                            C.cast(none_pp, C.c_void_p), None, 0, 0))
            o = L.BuildOpts(M, M0, ef_construction, alpha, seed, batch_max, 0)
            st = np.zeros(8, np.uint64)
            _check(lib.hb_index_build_graph(h, C.byref(o), device, _ptr(st)))
            if stats is not None:
                stats.update(batches=int(st[0]), launches=int(st[1]), items=int(st[2]), max_level=int(st[3]), dropped=int(st[4]), cut=int(st[5]), kernels_ms=int(st[6]), build_call_ms=int(st[7]))
            _check(lib.hb_index_finalize(h, device))
        except Exception:
            lib.hb_index_free(h)
            raise
        return cls(h, d, index, device)

    @classmethod
    def build_from_path(cls, path, index, distance, db_name=None, dimensions=0, M=16, M0=32, ef_construction=100, alpha=1.0, seed=42,
                        device=0, stats=None):
        """The items `Writer::add_item` left in an LMDB environment on disk (read by the built-in walker), graph built on the
        device.  `dimensions` is only needed for a database that was never built (no metadata pair yet).  `export_kv(False)`
        on the result yields the Metadata / Links pairs to put back."""
        d = _distance_of(distance)
        lib = L.lib()
        h = C.c_void_p()
        _check(lib.hb_index_begin(d.ID, index, C.byref(h)))
        try:
            _check(lib.hb_index_push_lmdb(h, os.fsencode(path), db_name.encode() if db_name else None, None))
            o = L.BuildOpts(M, M0, ef_construction, alpha, seed, 0, dimensions)
            st = np.zeros(8, np.uint64)
            _check(lib.hb_index_build_graph(h, C.byref(o), device, _ptr(st)))
            if stats is not None:
                stats.update(batches=int(st[0]), launches=int(st[1]), items=int(st[2]), max_level=int(st[3]), dropped=int(st[4]), cut=int(st[5]),
                             kernels_ms=int(st[6]), build_call_ms=int(st[7]))
            _check(lib.hb_index_finalize(h, device))
        except Exception:
            lib.hb_index_free(h)
            raise
        return cls(h, d, index, device)

    def export_kv(self, with_items=True):
        """[(key, value)] in LMDB key order, in the encodings Writer::build writes (hb_index_export_kv)."""
        out = []

        def cb(user, k, kl, v, vl):
            out.append((bytes(k[:kl]), bytes(v[:vl])))
            return 0

        _check(L.lib().hb_index_export_kv(self._h, int(with_items), L.KV_VISIT(cb), None))
        return out

    def save(self, path):
        """Write the decoded snapshot to a flat cache file (hb_index_save); `Reader.load` brings it back without LMDB."""
        _check(L.lib().hb_index_save(self._h, os.fsencode(path)))

    @classmethod
    def load(cls, path, index, distance, device=0):
        d = _distance_of(distance)
        lib = L.lib()
        h = C.c_void_p()
        _check(lib.hb_index_begin(d.ID, index, C.byref(h)))
        try:
            _check(lib.hb_index_load(h, os.fsencode(path)))
            _check(lib.hb_index_finalize(h, device))
        except Exception:
            lib.hb_index_free(h)
            raise
        return cls(h, d, index, device)

    @classmethod
    def from_arrays(cls, distance, dims, ids, rows, hdr, layers, entry_points, max_level, index=0, device=0):
        """Flat-array route (bench / tests): layers = [(offsets u64[n+1], neighbour item ids u32[nnz])]."""
        d = _distance_of(distance)
        lib = L.lib()
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        n = len(ids)
        rows = np.ascontiguousarray(rows)
        assert rows.dtype == (np.uint64 if d.is_binary() else np.float32)
        hdr_a = None if hdr is None else np.ascontiguousarray(hdr, dtype=np.float32)
        offs = [np.ascontiguousarray(o, dtype=np.uint64) for o, _ in layers]
        nbrs = [np.ascontiguousarray(b, dtype=np.uint32) for _, b in layers]
        for o in offs:
            assert len(o) == n + 1
        nl = len(layers)
        off_pp = (C.c_void_p * max(nl, 1))(*[o.ctypes.data for o in offs])
        nbr_pp = (C.c_void_p * max(nl, 1))(*[b.ctypes.data for b in nbrs])
        eps = np.ascontiguousarray(entry_points, dtype=np.uint32)
        h = C.c_void_p()
        _check(lib.hb_index_begin(d.ID, index, C.byref(h)))
        try:
            _check(lib.hb_index_from_arrays(h, dims, _ptr(ids), n, _ptr(rows), _ptr(hdr_a), nl,
                                            C.cast(off_pp, C.c_void_p), C.cast(nbr_pp, C.c_void_p), _ptr(eps), len(eps),
                                            max_level))
            _check(lib.hb_index_finalize(h, device))
        except Exception:
            lib.hb_index_free(h)
            raise
        return cls(h, d, index, device)

    def replicate(self, devices):
        """Copy the snapshot to further GPUs (hb_index_replicate): one `by_vectors` / `by_items` call then searches
        contiguous nq / n_devices slices of its batch on every device at once."""
        devs = np.ascontiguousarray(list(devices), dtype=np.int32)
        _check(L.lib().hb_index_replicate(self._h, _ptr(devs), len(devs)))
        return self

    def devices(self):
        n = L.lib().hb_index_n_devices(self._h)
        return [L.lib().hb_index_device(self._h, i) for i in range(n)]

    # accessors — reader.rs:545-606
    def dimensions(self):
        return L.lib().hb_index_dimensions(self._h)

    def n_entrypoints(self):
        return L.lib().hb_index_n_entry_points(self._h)

    def n_items(self):
        return L.lib().hb_index_n_items(self._h)

    def max_level(self):
        return L.lib().hb_index_max_level(self._h)

    def item_ids(self):
        n = self.n_items()
        out = np.zeros(n, np.uint32)
        L.lib().hb_index_item_ids(self._h, _ptr(out), n)
        return out

    def index(self):
        return self._index

    def entry_points(self):
        n = L.lib().hb_index_entry_points(self._h, None, 0)
        out = np.zeros(n, np.uint32)
        L.lib().hb_index_entry_points(self._h, _ptr(out), n)
        return out

    def layers(self):
        """[(offsets u64[n+1], neighbour item ids u32[nnz])] per layer — the shape `from_arrays` takes (hb_index_layer_csr)."""
        out = []
        n = self.n_items()
        for l in range(L.lib().hb_index_n_layers(self._h)):
            nnz = C.c_uint64()
            _check(L.lib().hb_index_layer_csr(self._h, l, None, None, 0, C.byref(nnz)))
            off = np.zeros(n + 1, np.uint64)
            nbr = np.zeros(nnz.value, np.uint32)
            _check(L.lib().hb_index_layer_csr(self._h, l, _ptr(off), _ptr(nbr), nnz.value, None))
            out.append((off, nbr))
        return out

    def version(self):
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(L.lib().hb_index_version(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return (a.value, b.value, c.value)

    def is_empty(self):
        return self.n_items() == 0

    def contains_item(self, item):
        return bool(L.lib().hb_index_contains_item(self._h, item))

    def item_vector(self, item):
        if not self.contains_item(item):
            return None
        out = np.zeros(self.dimensions(), np.float32)
        _check(L.lib().hb_index_item_vector(self._h, item, _ptr(out)))
        return out

    def nns(self, count):  # reader.rs:611-620
        return QueryBuilder(self, count)

    # device-resident batched search (timing the kernels alone); torch tensors or raw pointers
    def search_device(self, d_q_ptr, nq, count, ef, d_ids_ptr, d_dist_ptr, d_len_ptr, d_ctr_ptr=None, stream=0):
        _check(L.lib().hb_search_by_vector_device(self._h, d_q_ptr, nq, count, ef, d_ids_ptr, d_dist_ptr, d_len_ptr,
                                                  d_ctr_ptr, stream))


def exact_knn(reader, vectors, k):
    """Exact k-NN over all items in the index metric (recall ground truth)."""
    q = np.ascontiguousarray(vectors, dtype=np.float32)
    if q.ndim == 1:
        q = q.reshape(1, -1)
    ids = np.zeros((len(q), k), np.uint32)
    dist = np.zeros((len(q), k), np.float32)
    _check(L.lib().hb_exact_knn(reader._h, _ptr(q), len(q), q.shape[1], k, _ptr(ids), _ptr(dist)))
    return ids, dist


def merge_topk_device(device, d_ids_ptr, d_dist_ptr, n_parts, nq, k, d_out_ids_ptr, d_out_dist_ptr, d_out_len_ptr=None, stream=0):
    _check(L.lib().hb_merge_topk_device(device, d_ids_ptr, d_dist_ptr, n_parts, nq, k, d_out_ids_ptr, d_out_dist_ptr,
                                        d_out_len_ptr, stream))
