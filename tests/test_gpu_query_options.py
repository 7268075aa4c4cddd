"""GPU parity of the rest of the QueryBuilder surface, through the C-ABI, against the oracle:
candidates filter on the graph walk (reader.rs:200-203,322,355-359), linear scan below the thresholds
(reader.rs:622-640,668-711), by_item with candidates (reader.rs:826-842), the exhaustive fallback over unseen items
(reader.rs:771-795,865-889) and *_with_cancellation (reader.rs:91-188,330,684,749-764)."""
import numpy as np
import pytest

from helpers import assert_counters_same, assert_same, make_db, make_vectors, open_reader_arrays, open_reader_kv
from oracle import oracle as O
from oracle.oracle import OracleDb
from hannoy_b200 import _lib as L
import hannoy_b200 as hb

pytestmark = pytest.mark.gpu

METRICS = [("cosine", 96), ("euclidean", 20), ("manhattan", 33), ("hamming", 200), ("binary quantized cosine", 512)]


def _split(got):
    """(ids, dist, len low bits, counters), cancelled mask"""
    lens = got[2]
    none = lens == 0xFFFFFFFF
    canc = (~none) & ((lens >> 31) == 1)
    clean = np.where(none, lens, lens & 0x7FFFFFFF).astype(np.uint32)
    return (got[0], got[1], clean) + tuple(got[3:]), canc


@pytest.mark.parametrize("metric,dims", METRICS)
def test_candidates_filter_graph_walk_and_linear_scan(metric, dims):
    n = 3000
    ids = (np.arange(n, dtype=np.uint32) * 3 + 1)
    db, x = make_db(metric, n, dims, seed=dims, kind="clustered", ids=ids)
    rd = open_reader_arrays(db, metric)
    q = make_vectors(48, dims, seed=9, kind="clustered")
    rng = np.random.default_rng(1)
    for n_cand, lb, ratio in [(1500, 1000, 1.0),   # above the threshold: graph walk with a filter
                              (600, 1000, 1.0),    # below it: linear scan
                              (600, 1000, 0.1),    # ratio vetoes the scan (600/3000 > 0.1): graph walk
                              (600, 0, 1.0),       # linear_below = 0 disables the scan
                              (40, 1000, 1.0), (5, 1000, 1.0), (1, 1000, 1.0)]:
        cand = rng.choice(ids, n_cand, replace=False).astype(np.uint32)
        cand_x = np.concatenate([cand, np.array([0, 2, 5, 4_000_000], np.uint32)])  # ids that are not in the index
        rng.shuffle(cand_x)
        for count, ef in [(10, 64), (100, 100), (3, 1)]:
            want = db.search_by_vector(q, count, ef=max(ef, count), candidates=cand_x, linear_below=lb, linear_below_ratio=ratio,
                                       counters=True)
            got = rd.nns(count).ef_search(ef).candidates(cand_x).linear_below(lb).linear_below_ratio(ratio).by_vectors_raw(q, counters=True)
            what = f"{metric} cand={n_cand} lb={lb} ratio={ratio} k={count} ef={ef}"
            assert_same(got, want, what)
            assert_counters_same(got[3], want[3], what)
            assert np.array_equal(got[3][:, 6] & 3, want[3][:, 6] & 3), what  # FALLBACK / LINEAR flags
            valid = np.arange(count)[None, :] < got[2][:, None]
            assert set(got[0][valid].tolist()) <= set(cand.tolist())
            assert not got[0][~valid].any() and not got[1][~valid].any()   # nothing stale past out_len
    # disjoint candidates -> [] (reader.rs:654-656); empty bitmap likewise
    for c in (np.array([0, 2], np.uint32), np.zeros(0, np.uint32)):
        ids_, dist_, lens_ = rd.nns(10).candidates(c).by_vectors_raw(q)
        assert np.all(lens_ == 0)


@pytest.mark.parametrize("metric,dims", METRICS[:2] + METRICS[3:4])
def test_by_item_with_candidates(metric, dims):
    n = 2000
    db, x = make_db(metric, n, dims, seed=3 + dims, kind="clustered")
    rd = open_reader_kv(db, metric, index=1)
    rng = np.random.default_rng(2)
    items = np.array([0, 7, 99, n - 1, n + 5, 1234], np.uint32)
    for n_cand, lb in [(1200, 1000), (300, 1000), (300, 0), (8, 1000)]:
        cand = rng.choice(n, n_cand, replace=False).astype(np.uint32)
        cand[:3] = [0, 7, 99]  # the query items themselves are candidates: they must still be excluded on the graph path
        cand = np.unique(cand)
        for count, ef in [(10, 50), (30, 30)]:
            want = db.search_by_item(items, count, ef=max(ef, count), candidates=cand, linear_below=lb, counters=True)
            got = rd.nns(count).ef_search(ef).candidates(cand).linear_below(lb).by_items_raw(items, counters=True)
            what = f"by_item {metric} cand={n_cand} lb={lb} k={count}"
            assert_same(got, want, what)
            assert_counters_same(got[3], want[3], what)
            assert got[2][4] == 0xFFFFFFFF


def _islands_db(metric, dims, n_islands=40, size=6, seed=0):
    """Disconnected layer 0: rings of `size` items with no edge between rings, so a walk sees one ring and the
    reference falls back to seeding from every unseen item (reader.rs:771-795)."""
    n = n_islands * size
    x = make_vectors(n, dims, seed=seed)
    db = OracleDb(metric, dims)
    db.add_items(np.arange(n, dtype=np.uint32), x)
    for g in range(n_islands):
        for j in range(size):
            i = g * size + j
            db.set_links(i, 0, sorted({g * size + (j + 1) % size, g * size + (j - 1) % size}))
    for g in (3, 17):  # a tiny top layer
        db.set_links(g * size, 1, [])
    db.set_links(3 * size, 1, [17 * size])
    db.set_links(17 * size, 1, [3 * size])
    db.set_entry_points([3 * size], 1)
    return db, x


@pytest.mark.parametrize("metric,dims", METRICS)
def test_exhaustive_fallback(metric, dims):
    db, x = _islands_db(metric, dims, seed=dims)
    rd = open_reader_arrays(db, metric)
    q = make_vectors(40, dims, seed=5)
    for count, ef in [(10, 10), (10, 64), (50, 50), (240, 240), (7, 100)]:
        want = db.search_by_vector(q, count, ef=max(ef, count), counters=True)
        got = rd.nns(count).ef_search(ef).by_vectors_raw(q, counters=True)
        what = f"fallback {metric} k={count} ef={ef}"
        assert (want[3][:, 6] & O.FLAG_FALLBACK).all()
        assert_same(got, want, what)
        assert_counters_same(got[3], want[3], what)
        assert np.array_equal(got[3][:, 6] & 3, want[3][:, 6] & 3)
    items = np.array([0, 13, 100, 239], np.uint32)
    for count in (4, 10, 30):
        want = db.search_by_item(items, count, ef=count, counters=True)
        got = rd.nns(count).ef_search(count).by_items_raw(items, counters=True)
        assert_same(got, want, f"fallback by_item {metric} k={count}")
        assert_counters_same(got[3], want[3], f"fallback by_item {metric} k={count}")
    # candidates + fallback: only some islands hold candidates
    cand = np.arange(60, 200, dtype=np.uint32)
    want = db.search_by_vector(q, 20, ef=20, candidates=cand, linear_below=0, counters=True)
    got = rd.nns(20).ef_search(20).candidates(cand).linear_below(0).by_vectors_raw(q, counters=True)
    assert_same(got, want, f"fallback + candidates {metric}")
    assert_counters_same(got[3], want[3], f"fallback + candidates {metric}")


# ---- cancellation -----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric,dims", METRICS)
def test_cancel_after_n_polls_equals_reference(metric, dims):
    """cancel_fn = "true from its N-th call on": same Cancelled(..) prefix results, same did_cancel, same traversal
    counters as the reference's loop (polled before every layer-0 pop; the descent is never cancelled)."""
    n = 2500
    db, x = make_db(metric, n, dims, seed=dims + 1, kind="clustered")
    rd = open_reader_arrays(db, metric)
    q = make_vectors(40, dims, seed=4, kind="clustered")
    for count, ef in [(10, 64), (100, 100)]:
        for after in (1, 2, 3, 8, 40, 90, 140, 100000):
            want = db.search_by_vector(q, count, ef=ef, counters=True, cancel_after=after)
            got, canc = _split(rd.nns(count).ef_search(ef).with_cancellation(after).by_vectors_raw(q, counters=True))
            what = f"cancel {metric} k={count} ef={ef} after={after}"
            assert_same(got, want, what)
            assert_counters_same(got[3], want[3], what)
            assert np.array_equal(canc, (want[3][:, 6] & O.FLAG_CANCELLED) != 0), what
            assert np.array_equal((got[3][:, 6] & L.FLAG_CANCELLED) != 0, canc)
        assert not canc.any()  # 100000 polls are never reached
    items = np.array([0, 5, 77, n + 1], np.uint32)
    for after in (1, 4, 60):
        want = db.search_by_item(items, 10, ef=64, counters=True, cancel_after=after)
        got, canc = _split(rd.nns(10).ef_search(64).with_cancellation(after).by_items_raw(items, counters=True))
        assert_same(got, want, f"cancel by_item {metric} after={after}")
        assert np.array_equal(canc, (want[3][:, 6] & O.FLAG_CANCELLED) != 0)
        assert got[2][3] == 0xFFFFFFFF


def test_cancel_in_linear_scan_and_fallback():
    db, x = make_db("cosine", 1500, 64, seed=8)
    rd = open_reader_arrays(db, "cosine")
    q = make_vectors(16, 64, seed=2)
    cand = np.concatenate([np.arange(100, 400, dtype=np.uint32), np.array([5000, 6000], np.uint32)])
    for after in (1, 2, 50, 302, 303, 400):  # 302 candidate ids -> call 302 is the last one made
        want = db.search_by_vector(q, 10, candidates=cand, counters=True, cancel_after=after)
        got, canc = _split(rd.nns(10).candidates(cand).with_cancellation(after).by_vectors_raw(q, counters=True))
        assert (want[3][:, 6] & O.FLAG_LINEAR).all()
        assert_same(got, want, f"linear cancel after={after}")
        assert_counters_same(got[3], want[3], f"linear cancel after={after}")
        assert np.array_equal(canc, (want[3][:, 6] & O.FLAG_CANCELLED) != 0), after
    assert not canc.any()
    # fallback: the closure is shared by every visit of the query; only the interrupted visit's heap comes back
    db, x = _islands_db("euclidean", 24, seed=3)
    rd = open_reader_arrays(db, "euclidean")
    q = make_vectors(24, 24, seed=6)
    for after in (1, 3, 7, 8, 9, 15, 30, 200, 5000):
        want = db.search_by_vector(q, 30, ef=30, counters=True, cancel_after=after)
        got, canc = _split(rd.nns(30).ef_search(30).with_cancellation(after).by_vectors_raw(q, counters=True))
        assert_same(got, want, f"fallback cancel after={after}")
        assert_counters_same(got[3], want[3], f"fallback cancel after={after}")
        assert np.array_equal(canc, (want[3][:, 6] & O.FLAG_CANCELLED) != 0), after


def test_cancel_token():
    db, x = make_db("cosine", 3000, 96, seed=12, kind="clustered")
    rd = open_reader_arrays(db, "cosine")
    q = make_vectors(64, 96, seed=1, kind="clustered")
    full = db.search_by_vector(q, 10, ef=64)
    tok = hb.CancelToken(0)
    qb = rd.nns(10).ef_search(64).with_cancellation(tok)
    got, canc = _split(qb.by_vectors_raw(q))                      # token not tripped: a plain search
    assert_same(got, full, "token idle")
    assert not canc.any()
    tok.cancel()
    assert tok.is_cancelled()
    got, canc = _split(qb.by_vectors_raw(q))                      # tripped before the call: the first poll of every query
    assert canc.all()                                             # cancels it == the closure `|| true`
    assert_same(got, db.search_by_vector(q, 10, ef=64, cancel_after=1), "token tripped")
    tok.reset()
    got, canc = _split(qb.by_vectors_raw(q))
    assert not canc.any()
    assert_same(got, full, "token reset")
    # the reference's single-query spellings; a callable is evaluated by a host watcher thread
    s = rd.nns(10).ef_search(64).by_vector_with_cancellation(q[0], lambda: False)
    assert not s.did_cancel() and [i for i, _ in s.into_nns()] == full[0][0, :full[2][0]].tolist()
    s = rd.nns(10).ef_search(64).by_vector_with_cancellation(q[0], 1)
    assert s.did_cancel() and len(s.into_nns()) >= 1
    s = rd.nns(10).by_item_with_cancellation(5, tok)
    assert not s.did_cancel() and len(s.into_nns()) == 10
    assert rd.nns(10).by_item_with_cancellation(999999, 1) is None
    # tripping the token from another thread while a long batch runs ends it early: every query either completed with
    # the exact answer or reports did_cancel
    import threading
    big = make_vectors(20000, 96, seed=3, kind="clustered")
    t = threading.Timer(0.002, tok.cancel)
    t.start()
    got, canc = _split(rd.nns(10).ef_search(400).with_cancellation(tok).by_vectors_raw(big))
    t.join()
    done = ~canc
    if done.any():
        idx = np.nonzero(done)[0][:200]
        want = db.search_by_vector(big[idx], 10, ef=400)
        assert_same((got[0][idx], got[1][idx], got[2][idx]), want, "queries that finished before the trip")
    tok.close()
