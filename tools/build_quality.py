"""Dev tool (GPU box): recall of device-built vs CPU-built graphs, per metric, with the build statistics.

  python tools/build_quality.py [more]
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import make_db, make_vectors
import hannoy_b200 as hb

def recall(ids, lens, gt):
    return float(np.mean([len(set(ids[i, :lens[i]].tolist()) & set(gt[i].tolist())) / gt.shape[1] for i in range(len(gt))]))

CASES = [("euclidean", 64, 20000), ("cosine", 96, 12000), ("binary quantized cosine", 512, 12000)]
if len(sys.argv) > 1:
    CASES = [("binary quantized cosine", 512, 12000), ("cosine", 96, 12000)]
for metric, dims, n in CASES:
    ids = np.arange(n, dtype=np.uint32)
    ref, x = make_db(metric, n, dims, seed=n + dims, kind="clustered", ids=ids, efc=100, n_threads=8)
    q = make_vectors(300, dims, seed=5, kind="clustered")
    gt = None
    for bm in (0,):
        st = {}
        t = time.time()
        rd = hb.Reader.build(metric, dims, ids, ref.rows(), ref.headers(), seed=7, batch_max=bm, stats=st)
        dt = time.time() - t
        if gt is None:
            gt, _ = hb.exact_knn(rd, q, 10)
        out = []
        for ef in (32, 64, 128):
            g = rd.nns(10).ef_search(ef).by_vectors_raw(q)
            out.append(round(recall(g[0], g[2], gt), 4))
        kv = rd.export_kv(with_items=False)
        degs = [len(v) for k, v in kv if k[2] == 2 and k[7] == 0]
        print(metric, "batch_max", bm, st, f"{dt:.2f}s", "recall@10 ef32/64/128:", out, "mean L0 links bytes", round(float(np.mean(degs)), 1), flush=True)
    out = []
    for ef in (32, 64, 128):
        c = ref.search_by_vector(q, 10, ef=ef, n_threads=8)
        out.append(round(recall(c[0], c[2], gt), 4))
    off, nbr = ref.layers()[0]
    print(metric, "cpu build recall:", out, "mean L0 degree", round(len(nbr) / n, 2), flush=True)
