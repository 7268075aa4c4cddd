#!/usr/bin/env python
"""Generates tests/golden/options_golden.npz from the databases of search_golden.npz: the oracle's answers for the rest of
the QueryBuilder surface — candidates on the graph walk, linear scan, by_item with candidates, and cancellation after N
polls.  Pins the oracle against drift (tests/test_golden.py); regenerate only when the oracle is changed on purpose."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_golden import N_CASES, load_case, oracle_db  # noqa: E402

CANCEL_AFTER = (1, 3, 20)


def answers(db, g, ci):
    rng = np.random.default_rng(500 + ci)
    ids = g["ids"]
    cand_big = np.sort(rng.choice(ids, len(ids) // 2, replace=False)).astype(np.uint32)
    cand_small = np.sort(rng.choice(ids, 12, replace=False)).astype(np.uint32)
    out = {"cand_big": cand_big, "cand_small": cand_small}

    def put(key, r):
        out[key + "_ids"], out[key + "_dbits"], out[key + "_len"], out[key + "_ctr"] = r[0], r[1].view(np.uint32), r[2], r[3][:, :7].astype(np.uint32)
    put("walk", db.search_by_vector(g["q"], 10, ef=48, candidates=cand_big, linear_below=0, counters=True))
    put("linear", db.search_by_vector(g["q"], 10, ef=48, candidates=cand_small, counters=True))
    put("item", db.search_by_item(g["items"], 5, ef=32, candidates=cand_big, linear_below=0, counters=True))
    for a in CANCEL_AFTER:
        put(f"cancel{a}", db.search_by_vector(g["q"], 10, ef=48, counters=True, cancel_after=a))
    return out


def main():
    out = {}
    for ci in range(N_CASES):
        g = load_case(ci)
        for k, v in answers(oracle_db(g), g, ci).items():
            out[f"c{ci}_{k}"] = v
    path = os.path.join(HERE, "options_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
