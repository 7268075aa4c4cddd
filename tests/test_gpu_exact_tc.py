"""Exact k-NN through the tensor-core shortlist (exact_tc.cu: tf32 tcgen05 GEMM -> provably sufficient candidates -> bit-exact
re-rank) against the full CUDA-core scan and the CPU oracle: identical ids and distance bits, including ties, duplicates
(candidate overflow -> the scan takes the query over), ragged sizes and padded dimensions."""
import numpy as np
import pytest

from helpers import make_db, make_vectors
import hannoy_b200 as hb
from hannoy_b200 import _lib as L

pytestmark = pytest.mark.gpu


def _reader_without_graph(metric, x, ids=None):
    from oracle.oracle import OracleDb
    n, dims = x.shape
    ids = np.arange(n, dtype=np.uint32) if ids is None else ids
    db = OracleDb(metric, dims)
    db.add_items(ids, x)
    off = np.zeros(n + 1, np.uint64)
    rd = hb.Reader.from_arrays(metric, dims, db.ids(), db.rows(), db.headers(), [(off, np.zeros(0, np.uint32))], np.array([ids[0]], np.uint32), 0)
    return rd, db


def _both(rd, q, k):
    lib = L.lib()
    assert lib.hb_tune(b"exact_tc", 1) == L.HB_OK
    a = hb.exact_knn(rd, q, k)
    assert lib.hb_tune(b"exact_tc", 0) == L.HB_OK
    b = hb.exact_knn(rd, q, k)
    lib.hb_tune(b"exact_tc", 1)
    return a, b


def _lowrank(n, dims, seed, nc=64, r=16, noise=0.02):
    rng = np.random.default_rng(seed)
    A = rng.normal(0, 1, (r, dims)).astype(np.float32) / np.sqrt(r)
    C = rng.normal(0, 1, (nc, r)).astype(np.float32)
    z = C[rng.integers(0, nc, n)] + rng.normal(0, 1, (n, r)).astype(np.float32)
    x = z @ A + noise * rng.normal(0, 1, (n, dims)).astype(np.float32)
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


@pytest.mark.parametrize("metric,n,dims,nq,k,gen", [
    ("cosine", 50_000, 768, 300, 10, "lowrank"),         # config-3 shaped
    ("cosine", 33_333, 100, 130, 100, "lowrank"),        # padded dimensions (3 AVX blocks + 4-element tail), ragged n and nq
    ("euclidean", 60_000, 128, 257, 10, "sift"),         # config-2 shaped: large norms, wide error band
    ("euclidean", 20_000, 96, 64, 128, "uniform"),
    ("cosine", 40_000, 256, 1, 10, "lowrank"),           # a single query
    ("cosine", 20_000, 64, 200, 7, "dups"),              # heavy duplication: ties by slot, candidate overflow -> scan fallback
])
def test_tensor_core_shortlist_equals_full_scan(metric, n, dims, nq, k, gen):
    rng = np.random.default_rng(n + dims)
    if gen == "lowrank":
        x = _lowrank(n, dims, 1)
        q = _lowrank(nq, dims, 2)
    elif gen == "sift":
        x = np.clip(np.round(27 + 21.5 * rng.normal(0, 1, (n, 8)).astype(np.float32) @ rng.normal(0, 1, (8, dims)).astype(np.float32) / 2.8 + 5.7 * rng.normal(0, 1, (n, dims))), 0, 255).astype(np.float32)
        q = x[rng.integers(0, n, nq)] + rng.integers(-3, 4, (nq, dims)).astype(np.float32)
    elif gen == "uniform":
        x = rng.uniform(-1, 1, (n, dims)).astype(np.float32)
        q = rng.uniform(-1, 1, (nq, dims)).astype(np.float32)
    else:   # 40 distinct vectors, each repeated n / 40 times
        base = _lowrank(40, dims, 3)
        x = base[rng.integers(0, 40, n)]
        q = base[rng.integers(0, 40, nq)] + 1e-4 * rng.normal(0, 1, (nq, dims)).astype(np.float32)
    q[0] = x[n // 2]   # a self query
    ids = np.arange(n, dtype=np.uint32) * 3 + 5
    rd, db = _reader_without_graph(metric, x, ids)
    (ti, td), (si, sd) = _both(rd, q, k)
    assert np.array_equal(ti, si), f"ids differ for {np.nonzero((ti != si).any(1))[0][:5]}"
    assert np.array_equal(td.view(np.uint32), sd.view(np.uint32))
    oi, od = db.exact_knn(q[:32], k, n_threads=4)
    assert np.array_equal(ti[:32], oi) and np.array_equal(td[:32].view(np.uint32), od.view(np.uint32))


def test_tensor_core_path_declines_what_it_cannot_prove():
    """Zero vectors (cosine.rs:44-55 returns 0.0 when the norms vanish), tiny indexes, other metrics: the call still
    answers, through the scan."""
    rng = np.random.default_rng(5)
    x = _lowrank(30_000, 96, 7)
    x[100] = 0.0
    q = _lowrank(200, 96, 8)
    q[3] = 0.0
    rd, db = _reader_without_graph("cosine", x)
    (ti, td), (si, sd) = _both(rd, q, 10)
    assert np.array_equal(ti, si) and np.array_equal(td.view(np.uint32), sd.view(np.uint32))
    x2 = rng.normal(0, 1, (30_000, 64)).astype(np.float32)
    q2 = rng.normal(0, 1, (150, 64)).astype(np.float32)
    q2[5] = 0.0   # a degenerate query is handed to the scan, the rest of the batch stays on the tensor cores
    rd2, db2 = _reader_without_graph("cosine", x2)
    (ti, td), (si, sd) = _both(rd2, q2, 10)
    assert np.array_equal(ti, si) and np.array_equal(td.view(np.uint32), sd.view(np.uint32))
    rd3, db3 = _reader_without_graph("manhattan", x2[:5000])
    (ti, td), (si, sd) = _both(rd3, q2[:20], 10)
    assert np.array_equal(ti, si)
