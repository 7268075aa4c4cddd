#!/usr/bin/env python
"""Generates tests/golden/search_golden.npz: for every metric a small seeded database (vectors, the graph the
restated builder made for it, entry points), seeded queries, and the oracle's answers (ids, distance bit
patterns, lengths, traversal counters) for by_vector and by_item at two (count, ef) settings.

The reference itself cannot produce these (Rust, no toolchain in this image — DESIGN.md §2); the fixture pins the
oracle against drift (tests/test_golden.py, CPU) and gives the CUDA engine a committed, oracle-independent target
(-m gpu).  Regenerate with `python tests/golden/make_search_golden.py` only when the oracle is changed on purpose.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.oracle import OracleDb  # noqa: E402

CASES = [  # metric, n, dims
    ("euclidean", 160, 64), ("euclidean", 120, 37), ("cosine", 160, 96), ("cosine", 120, 20), ("manhattan", 120, 24),
    ("hamming", 160, 128), ("binary quantized cosine", 160, 200), ("binary quantized euclidean", 120, 64),
    ("binary quantized manhattan", 120, 70),
]
SETTINGS = [(10, 48), (3, 3)]


def main():
    out = {}
    for ci, (metric, n, dims) in enumerate(CASES):
        rng = np.random.default_rng(1000 + ci)
        x = rng.normal(0, 1, (n, dims)).astype(np.float32)
        ids = np.sort(rng.choice(5 * n, n, replace=False)).astype(np.uint32)
        q = rng.normal(0, 1, (24, dims)).astype(np.float32)
        q[:4] = x[:4]
        db = OracleDb(metric, dims)
        db.add_items(ids, x)
        db.build(M=8, M0=16, ef_construction=40, seed=7 + ci, n_threads=1)
        p = f"c{ci}_"
        out[p + "metric"] = np.array(metric)
        out[p + "dims"] = np.array(dims)
        out[p + "ids"] = ids
        out[p + "x"] = x
        out[p + "q"] = q
        out[p + "eps"] = db.entry_points
        out[p + "max_level"] = np.array(db.max_level)
        for l, (off, nbr) in enumerate(db.layers()):
            out[p + f"off{l}"] = off.astype(np.uint32)
            out[p + f"nbr{l}"] = nbr
        out[p + "n_layers"] = np.array(len(db.layers()))
        items = np.concatenate([ids[:6], [ids[-1] + 1]]).astype(np.uint32)   # the last one is absent
        out[p + "items"] = items
        for si, (count, ef) in enumerate(SETTINGS):
            i, d, l, c = db.search_by_vector(q, count, ef=max(ef, count), counters=True)
            out[p + f"v{si}_ids"], out[p + f"v{si}_dbits"], out[p + f"v{si}_len"], out[p + f"v{si}_ctr"] = i, d.view(np.uint32), l, c[:, :6].astype(np.uint32)
            i, d, l, c = db.search_by_item(items, count, ef=max(ef, count), counters=True)
            out[p + f"i{si}_ids"], out[p + f"i{si}_dbits"], out[p + f"i{si}_len"] = i, d.view(np.uint32), l
    np.savez_compressed(os.path.join(HERE, "search_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "search_golden.npz"), os.path.getsize(os.path.join(HERE, "search_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
