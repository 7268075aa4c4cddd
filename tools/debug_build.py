"""dev: one small device graph build (run under compute-sanitizer on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import make_vectors
from oracle.oracle import OracleDb
import hannoy_b200 as hb
n, dims = int(sys.argv[1]) if len(sys.argv) > 1 else 3000, 64
x = make_vectors(n, dims, seed=1, kind="clustered")
db = OracleDb("euclidean", dims); db.add_items(np.arange(n, dtype=np.uint32), x)
stats = {}
rd = hb.Reader.build("euclidean", dims, np.arange(n, dtype=np.uint32), db.rows(), db.headers(), seed=3, stats=stats)
print("built", stats)
ids, dist, lens = rd.nns(10).ef_search(64).by_vectors_raw(x[:100])
print("self hits", int((ids[:, 0] == np.arange(100)).sum()))
