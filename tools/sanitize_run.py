#!/usr/bin/env python
"""A small pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): graph search
in full and tiny batches (helper warps), by_item, candidates + linear scan, cancellation, large-ef global heaps, wide
adjacency, the binary kernel, the device graph builder, the exact k-NN scan and its tensor-core shortlist, the top-k merge.
Sizes are tiny: the tools slow kernels down by one to two orders of magnitude.  Results are still checked against the oracle.

  compute-sanitizer --tool memcheck  python tools/sanitize_run.py
  compute-sanitizer --tool racecheck python tools/sanitize_run.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import assert_same, make_db, make_vectors, open_reader_arrays  # noqa: E402
import hannoy_b200 as hb  # noqa: E402
from hannoy_b200 import _lib as L  # noqa: E402


def main():
    lib = L.lib()
    # f32 ring kernel: full and tiny batches (team), by_item, filter, linear scan, cancellation, ef beyond shared memory
    db, x = make_db("cosine", 1500, 96, seed=1, kind="clustered")
    rd = open_reader_arrays(db, "cosine")
    q = make_vectors(40, 96, seed=2, kind="clustered")
    for nq in (40, 3, 1):
        assert_same(rd.nns(10).ef_search(48).by_vectors_raw(q[:nq]), db.search_by_vector(q[:nq], 10, ef=48), f"cosine nq={nq}")
    items = np.array([0, 7, 1499, 5000], np.uint32)
    assert_same(rd.nns(5).ef_search(20).by_items_raw(items), db.search_by_item(items, 5, ef=20), "by_item")
    cand = np.arange(0, 1500, 3, dtype=np.uint32)
    assert_same(rd.nns(5).ef_search(20).candidates(cand).linear_below(0).by_vectors_raw(q[:8]),
                db.search_by_vector(q[:8], 5, ef=20, candidates=cand, linear_below=0), "filtered")
    assert_same(rd.nns(5).candidates(cand[:50]).by_vectors_raw(q[:8]), db.search_by_vector(q[:8], 5, candidates=cand[:50]), "linear scan")
    g = rd.nns(5).ef_search(20).with_cancellation(3).by_vectors_raw(q[:8])
    assert_same((g[0], g[1], g[2] & 0x7fffffff), db.search_by_vector(q[:8], 5, ef=20, cancel_after=3), "cancelled")
    assert_same(rd.nns(10).ef_search(1500).by_vectors_raw(q[:4]), db.search_by_vector(q[:4], 10, ef=1500), "global heaps")
    # wide layer-0 lists (two adjacency lines) and the CSR path
    for M, M0 in ((24, 48), (16, 200)):
        dbw, _ = make_db("euclidean", 1200, 64, seed=M0, M=M, M0=M0, efc=64)
        rw = open_reader_arrays(dbw, "euclidean")
        qw = make_vectors(12, 64, seed=4)
        assert_same(rw.nns(10).ef_search(40).by_vectors_raw(qw), dbw.search_by_vector(qw, 10, ef=40), f"M0={M0}")
    # one lane per row, binary codes
    for metric, dims in (("euclidean", 20), ("binary quantized cosine", 256), ("hamming", 128)):
        d2, _ = make_db(metric, 1000, dims, seed=dims)
        r2 = open_reader_arrays(d2, metric)
        q2 = make_vectors(10, dims, seed=6)
        assert_same(r2.nns(10).ef_search(32).by_vectors_raw(q2), d2.search_by_vector(q2, 10, ef=32), metric)
    # device graph builder (search + prune/link + apply-reverse kernels), M0 = 32 and 48
    for M, M0 in ((16, 32), (24, 48)):
        xb = make_vectors(1500, 48, seed=9, kind="clustered")
        from oracle.oracle import OracleDb
        ob = OracleDb("cosine", 48)
        ob.add_items(np.arange(1500, dtype=np.uint32), xb)
        rb = hb.Reader.build("cosine", 48, np.arange(1500, dtype=np.uint32), xb, ob.headers(), M=M, M0=M0, ef_construction=48, seed=3)
        got = rb.nns(1).ef_search(32).by_vectors_raw(xb[:50])
        assert (got[0][:, 0] == np.arange(50)).all(), "device-built graph: self queries"
    # exact k-NN: scan and tensor-core shortlist
    lib.hb_tune(b"exact_tc_min_pairs", 0)
    xe = make_vectors(1300, 96, seed=11, kind="clustered")
    oe = OracleDb("cosine", 96)
    oe.add_items(np.arange(1300, dtype=np.uint32), xe)
    re_ = hb.Reader.from_arrays("cosine", 96, oe.ids(), oe.rows(), oe.headers(), [(np.zeros(1301, np.uint64), np.zeros(0, np.uint32))], np.array([0], np.uint32), 0)
    qe = make_vectors(130, 96, seed=12, kind="clustered")
    a = hb.exact_knn(re_, qe, 10)
    lib.hb_tune(b"exact_tc", 0)
    b = hb.exact_knn(re_, qe, 10)
    lib.hb_tune(b"exact_tc", 1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)), "exact k-NN: tensor-core path != scan"
    print("sanitize_run: all checks passed")


if __name__ == "__main__":
    main()
