"""Parity at (or near) BASELINE.json's sizes, through properties that do not need the oracle to walk the whole
batch: batch-composition independence (a query's answer does not depend on what else is in the batch), self-query
=> self at distance 0, ascending (distance bits, id) order without duplicates, agreement with the exact k-NN kernel
where the graph search is exhaustive, plus a bit-exact oracle comparison on a bounded sample of the batch."""
import os

import numpy as np
import pytest

from helpers import assert_counters_same, assert_same, open_reader_arrays
from oracle.oracle import OracleDb

pytestmark = pytest.mark.gpu
THREADS = len(os.sched_getaffinity(0))


def clustered(n, dims, seed, nc=256, sigma=0.35):
    rng = np.random.default_rng(seed)
    centers = np.random.default_rng(77).normal(0, 1, (nc, dims)).astype(np.float32)
    out = np.empty((n, dims), np.float32)
    step = 100_000
    for s in range(0, n, step):
        m = min(step, n - s)
        out[s:s + m] = centers[rng.integers(0, nc, m)] + sigma * rng.standard_normal((m, dims), dtype=np.float32)
    return out


def check_structure(ids, dist, lens, k, n_items):
    assert np.all(lens == k)
    bits = dist.view(np.uint32).astype(np.uint64)
    keys = (bits << np.uint64(32)) | ids.astype(np.uint64)
    assert np.all(keys[:, 1:] > keys[:, :-1]), "results must be strictly ascending by (distance bits, id)"
    assert ids.max() < n_items


def run_case(metric, n, dims, nq, k, ef, n_sample):
    x = clustered(n, dims, 1)
    q = clustered(nq, dims, 2)
    db = OracleDb(metric, dims)
    db.add_items(np.arange(n, dtype=np.uint32), x)
    db.build(M=16, M0=32, ef_construction=100, seed=42, n_threads=THREADS)
    rd = open_reader_arrays(db, metric)
    qb = rd.nns(k).ef_search(ef)
    ids, dist, lens = qb.by_vectors_raw(q)
    check_structure(ids, dist, lens, k, n)
    # the batch split in three uneven parts gives the same answers, query by query
    cuts = [0, nq // 7, nq // 2 + 3, nq]
    for a, b in zip(cuts, cuts[1:]):
        pi, pd, pl = qb.by_vectors_raw(q[a:b])
        assert np.array_equal(pi, ids[a:b]) and np.array_equal(pd.view(np.uint32), dist[a:b].view(np.uint32)) and np.array_equal(pl, lens[a:b])
    # bounded oracle sample: ids, distance bits and traversal counters
    sel = np.linspace(0, nq - 1, n_sample).astype(np.int64)
    want = db.search_by_vector(q[sel], k, ef=max(ef, k), counters=True, n_threads=THREADS)
    got = qb.by_vectors_raw(q[sel], counters=True)
    assert_same(got, want, f"{metric} {n}x{dims}")
    assert_counters_same(got[3], want[3], f"{metric} {n}x{dims}")
    assert np.array_equal(got[0], ids[sel])
    # self queries: the item itself comes first, at the metric's zero (the walk is approximate at this scale: a few
    # self-queries may stop in another basin, exactly as the reference's would — the oracle sample above pins that)
    items = np.linspace(0, n - 1, 512).astype(np.int64)
    si, sd, sl = rd.nns(1).ef_search(ef).by_vectors_raw(x[items])
    zero = sd[:, 0] == 0.0 if metric in ("euclidean", "hamming", "binary quantized euclidean", "binary quantized manhattan") else np.abs(sd[:, 0]) < 1e-6
    assert np.mean(zero) > 0.95, np.mean(zero)
    if metric in ("euclidean", "cosine"):
        assert np.all(si[zero, 0] == items[zero])          # a float self-query that reaches distance 0 found itself
    # by_item at scale: never returns the item, same structure
    bi, bd, bl = rd.nns(k).ef_search(ef).by_items_raw(items[:256].astype(np.uint32))
    check_structure(bi, bd, bl, k, n)
    assert not np.any(bi == items[:256, None])
    return rd, db, q, ids


def test_config2_full_size_1m_x_128_euclidean():
    """BASELINE config 2 shape: 1M x 128 f32 Euclidean, 10k-query batch, top-10."""
    import hannoy_b200 as hb
    rd, db, q, ids = run_case("euclidean", 1_000_000, 128, 10_000, 10, 64, 300)
    gt, _ = hb.exact_knn(rd, q[:500], 10)
    recall = np.mean([len(set(ids[i]) & set(gt[i])) / 10 for i in range(500)])
    wide, _, _ = rd.nns(10).ef_search(256).by_vectors_raw(q[:500])
    recall_wide = np.mean([len(set(wide[i]) & set(gt[i])) / 10 for i in range(500)])
    # recall against the exact k-NN kernel: sane at ef = 64 and not worse with a wider beam (the oracle sample in
    # run_case already pins it to the reference's, query by query)
    assert recall > 0.5 and recall_wide >= recall - 0.01, (recall, recall_wide)


def test_config4_scaled_300k_x_1024_bq_cosine_top100():
    """BASELINE config 4 shape, scaled to 300k items: 1024-d BinaryQuantizedCosine codes, 20k-query batch, top-100
    (ef >= 100, reader.rs:217-220): ids must be bit-exact, ties included."""
    run_case("binary quantized cosine", 300_000, 1024, 20_000, 100, 100, 300)
