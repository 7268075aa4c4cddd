"""Shared test helpers: make an oracle db (the stand-in for a hannoy-built LMDB), open the CUDA Reader on
the same graph through either ingestion route, and compare results."""
import numpy as np

from oracle.oracle import OracleDb

METRIC_NAMES = ["euclidean", "cosine", "manhattan", "hamming", "binary quantized cosine",
                "binary quantized euclidean", "binary quantized manhattan"]


def make_vectors(n, dims, seed, kind="uniform"):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.uniform(-1, 1, (n, dims)).astype(np.float32)
    if kind == "clustered":
        nc = max(2, n // 50)
        centers = rng.normal(0, 1, (nc, dims)).astype(np.float32)
        x = centers[rng.integers(0, nc, n)] + 0.3 * rng.normal(0, 1, (n, dims)).astype(np.float32)
        return x.astype(np.float32)
    if kind == "int":
        return rng.integers(-3, 4, (n, dims)).astype(np.float32)
    raise ValueError(kind)


def make_db(metric, n, dims, seed=0, kind="uniform", M=16, M0=32, efc=100, ids=None, n_threads=1, build=True):
    db = OracleDb(metric, dims)
    x = make_vectors(n, dims, seed, kind)
    if ids is None:
        ids = np.arange(n, dtype=np.uint32)
    if n:
        db.add_items(ids, x)
    if build:
        db.build(M=M, M0=M0, ef_construction=efc, seed=seed + 42, n_threads=n_threads)
    return db, x


def open_reader_arrays(db, metric, device=0):
    import hannoy_b200 as hb
    return hb.Reader.from_arrays(metric, db.dims, db.ids(), db.rows(), db.headers(), db.layers(),
                                 db.entry_points, db.max_level, device=device)


def open_reader_kv(db, metric, index=0, device=0):
    import hannoy_b200 as hb
    return hb.Reader.open(db.export_kv(index), index, metric, device=device)


def assert_same(got, want, what=""):
    gi, gd, gl = got[:3]
    wi, wd, wl = want[:3]
    assert np.array_equal(gl, wl), f"{what}: result lengths differ: {gl[:10]} vs {wl[:10]}"
    for i in range(len(gl)):
        n = 0 if gl[i] == 0xFFFFFFFF else int(gl[i])
        assert np.array_equal(gi[i, :n], wi[i, :n]), f"{what}: ids differ for query {i}: {gi[i,:n]} vs {wi[i,:n]}"
        assert np.array_equal(gd[i, :n].view(np.uint32), wd[i, :n].view(np.uint32)), \
            f"{what}: distance bits differ for query {i}: {gd[i,:n]} vs {wd[i,:n]}"


def assert_counters_same(got_ctr, want_ctr, what=""):
    # dist evals, expansions, adjacency entries per layer class must match the oracle's traversal exactly
    assert np.array_equal(got_ctr[:, :6], want_ctr[:, :6]), f"{what}: traversal counters differ"
