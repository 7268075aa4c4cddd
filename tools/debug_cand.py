"""dev: compact diff of candidate-filtered searches, CUDA vs oracle (run on the GPU box)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import make_db, make_vectors, open_reader_arrays

def diff(tag, got, want):
    gi, gd, gl, gc = got; wi, wd, wl, wc = want
    bad = []
    for i in range(len(gl)):
        n = 0 if wl[i] == 0xFFFFFFFF else int(wl[i])
        same_len = gl[i] == wl[i]
        same_ids = same_len and np.array_equal(gi[i, :n], wi[i, :n])
        same_d = same_len and np.array_equal(gd[i, :n].view(np.uint32), wd[i, :n].view(np.uint32))
        same_c = np.array_equal(gc[i, :6], wc[i, :6])
        if not (same_len and same_ids and same_d and same_c):
            bad.append((i, same_len, same_ids, same_d, same_c, gc[i, :7].tolist(), int(gl[i]), wc[i, :7].tolist(), int(wl[i])))
    print(f"{tag}: {len(bad)}/{len(gl)} differ")
    for b in bad[:6]:
        print("   q%d len_ok=%s ids_ok=%s dist_ok=%s ctr_ok=%s\n      got ctr %s len %d\n      want ctr %s len %d" % b)

for metric, dims in [("euclidean", 20), ("hamming", 200)]:
    n = 3000
    ids = (np.arange(n, dtype=np.uint32) * 3 + 1)
    db, x = make_db(metric, n, dims, seed=dims, kind="clustered", ids=ids)
    rd = open_reader_arrays(db, metric)
    q = make_vectors(48, dims, seed=9, kind="clustered")
    rng = np.random.default_rng(1)
    for n_cand, lb, ratio in [(1500, 1000, 1.0), (600, 1000, 1.0), (600, 1000, 0.1), (600, 0, 1.0), (40, 1000, 1.0), (5, 1000, 1.0), (1, 1000, 1.0)]:
        cand = rng.choice(ids, n_cand, replace=False).astype(np.uint32)
        for extra in (True,):
            cand_x = np.concatenate([cand, np.array([0, 2, 5, 4_000_000], np.uint32)]) if extra else cand
            for count, ef in [(10, 64), (100, 100), (3, 1)]:
                want = db.search_by_vector(q, count, ef=max(ef, count), candidates=cand_x, linear_below=lb, linear_below_ratio=ratio, counters=True)
                got = rd.nns(count).ef_search(ef).candidates(cand_x).linear_below(lb).linear_below_ratio(ratio).by_vectors_raw(q, counters=True)
                diff(f"{metric} cand={n_cand} extra={extra} lb={lb} ratio={ratio} k={count} ef={ef}", got, want)
    items = np.array([0, 7, 99, n - 1, n + 5, 1234], np.uint32) * 3 + 1
    for n_cand, lb in [(1200, 1000), (300, 1000), (300, 0), (8, 1000)]:
        cand = rng.choice(ids, n_cand, replace=False).astype(np.uint32)
        for count, ef in [(10, 50), (30, 30)]:
            want = db.search_by_item(items, count, ef=max(ef, count), candidates=cand, linear_below=lb, counters=True)
            got = rd.nns(count).ef_search(ef).candidates(cand).linear_below(lb).by_items_raw(items, counters=True)
            diff(f"by_item {metric} cand={n_cand} lb={lb} k={count} ef={ef}", got, want)
