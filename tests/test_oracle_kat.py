"""Pins the CPU oracle to every known-answer vector, fixture and property the reference's own tests hold
for the search path.  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle.oracle import OracleDb


def bits_str(words):
    return ["{:08b}".format(b) for b in np.asarray(words, np.uint64).tobytes()]


# ---- quantizer bit patterns: src/unaligned_vector/binary_quantized_test.rs:11-27,100-167 -----------------
def test_bq_from_slice_kat():
    v = [0.1, 0.2, -0.3, 0.4, -0.5, 0.6, -0.7, 0.8, -0.9]
    assert bits_str(O.quantize(v, False)) == ["10101011"] + ["00000000"] * 7


def test_bq_smol_kat():
    assert bits_str(O.quantize([-1.0, 2.0, -3.0, 4.0, 5.0], False)) == ["00011010"] + ["00000000"] * 7


def test_bq_large_kat():
    v = [-1.0 if (n % 3 == 0 or n % 5 == 0) else 1.0 for n in range(100)]
    want = ["10010110", "01101001", "11001011", "10110100", "01100101", "11011010", "00110010", "01101101",
            "10011001", "10110110", "01001100", "01011011", "00000110", "00000000", "00000000", "00000000"]
    assert bits_str(O.quantize(v, False)) == want


@pytest.mark.parametrize("n", [1, 5, 9, 63, 64])
def test_quantized_len_is_word_multiple(n):  # unaligned_binary_quantized_iter_size: len() == 64
    assert len(O.quantize(np.ones(n), False)) * 64 == 64
    assert len(O.quantize(np.ones(n), True)) * 64 == 64


def test_binary_codec_is_strictly_positive():  # src/unaligned_vector/binary.rs:80-94
    v = np.array([0.0, -0.0, 1e-30, -1e-30, 2.0, -2.0, np.inf, -np.inf], np.float32)
    assert bits_str(O.quantize(v, True))[0] == "01010100"
    # BinaryQuantized: is_sign_positive -> +0.0 is a 1, -0.0 a 0 (binary_quantized.rs:84-87)
    assert bits_str(O.quantize(v, False))[0] == "01010101"


# ---- SIMD == scalar, exact, on integer-valued vectors: simple_avx.rs:112-153, simple_sse.rs:112-151 ----
def _avx_vectors():
    base = [float(x) for x in range(10, 26)]
    v1 = base * 4 + [26., 27., 28., 29., 30., 31.]
    v2 = [float(x) for x in range(40, 56)] + base * 3 + [56., 57., 58., 59., 60., 61.]
    return np.array(v1, np.float32), np.array(v2, np.float32)


def test_spaces_avx_equals_scalar():
    v1, v2 = _avx_vectors()
    L = O.lib()
    assert O.euclidean(v1, v2) == L.orc_euclid_scalar(O._p(v1), O._p(v2), len(v1))
    assert O.dot_product(v1, v2) == L.orc_dot_scalar(O._p(v1), O._p(v2), len(v1))
    assert O.euclidean(v1, v2) == float(((v1.astype(np.float64) - v2) ** 2).sum())


def test_spaces_sse_equals_scalar():
    v1 = np.array([float(x) for x in range(10, 26)] + [26., 27., 28., 29., 30., 31.], np.float32)
    v2 = np.array([float(x) for x in range(40, 56)] + [56., 57., 58., 59., 60., 61.], np.float32)
    L = O.lib()
    assert L.orc_euclid_sse(O._p(v1), O._p(v2), 22) == L.orc_euclid_scalar(O._p(v1), O._p(v2), 22)
    assert L.orc_dot_sse(O._p(v1), O._p(v2), 22) == L.orc_dot_scalar(O._p(v1), O._p(v2), 22)
    assert O.euclidean(v1, v2) == L.orc_euclid_sse(O._p(v1), O._p(v2), 22)  # 16 <= n < 32 dispatches to SSE


def test_avx_lane_order_matches_independent_numpy_model():
    """Re-derive simple_avx.rs:85-110 with numpy float32 ops (fma emulated in float64, exact for f32 products)."""
    rng = np.random.default_rng(0)
    for n in (32, 64, 100, 128, 768, 777):
        a = rng.normal(0, 1, n).astype(np.float32)
        b = rng.normal(0, 1, n).astype(np.float32)
        m = n - n % 32
        acc = np.zeros(32, np.float32)
        for i in range(0, m, 32):
            acc = (a[i:i + 32].astype(np.float64) * b[i:i + 32].astype(np.float64) + acc.astype(np.float64)).astype(np.float32)
        hs = []
        for k in range(4):
            x = acc[8 * k:8 * k + 8]
            r = (x[4:] + x[:4]).astype(np.float32)
            hs.append(np.float32(np.float32(r[0] + r[2]) + np.float32(r[1] + r[3])))
        res = np.float32(np.float32(np.float32(hs[0] + hs[1]) + hs[2]) + hs[3])
        for i in range(m, n):
            res = np.float32(res + np.float32(a[i] * b[i]))
        assert O.dot_product(a, b) == float(res), n


# ---- OrderedFloat: src/ordered_float.rs:37-46 ---------------------------------------------------------------
def test_ordered_float_bits_order_is_numeric_order_on_nonnegatives():
    rng = np.random.default_rng(1)
    hi = rng.uniform(0, 3e38, 2000).astype(np.float32)
    lo = (hi * rng.uniform(0, 1, 2000)).astype(np.float32)
    for u, l in zip(hi, lo):
        if u != l:
            assert O.lib().orc_ordered_float_cmp(float(u), float(l)) == 1


# ---- distance formulas on hand-computable inputs (src/distance/*.rs) --------------------------------------------
def test_distance_formulas():
    a = np.array([1, 0, 0, 0], np.float32)
    b = np.array([0, 1, 0, 0], np.float32)
    assert O.distance("euclidean", a, b) == 2.0            # squared L2, no sqrt
    assert O.distance("cosine", a, b) == 0.5
    assert O.distance("cosine", a, a) == 0.0
    assert O.distance("cosine", a, -a) == 1.0
    assert O.distance("cosine", a, 0 * a) == 0.0           # pn*qn <= EPSILON -> 0
    assert O.distance("manhattan", a, b) == 2.0
    assert O.distance("hamming", a, b) == 2.0 / 64.0       # popcount / padded length
    assert O.distance("binary quantized euclidean", a - 0.5, b - 0.5) == 8.0
    assert O.distance("binary quantized manhattan", a - 0.5, b - 0.5) == 4.0
    assert O.distance("binary quantized cosine", a - 0.5, b - 0.5) == (1 - (64 - 4) / 64.0) / 2


# ---- tests/test_basic.py:8-34 -------------------------------------------------------------------------------------
def test_hamming_one_hot_kat():
    db = OracleDb("hamming", 3)
    db.add_items([0, 1, 2], [[1, 0, 0], [0, 1, 0], [0, 0, 1]])
    db.build(M=4, M0=8, ef_construction=10)
    ids, dist, lens, _ = db.search_by_vector([[0, 1, 0]], 2, ef=200)
    assert lens[0] == 2 and ids[0, 0] == 1 and dist[0, 0] == 0.0


# ---- src/tests/writer.rs:517-547 — emptied index answers [] -------------------------------------------------------
def test_empty_index_returns_nothing():
    db = OracleDb("euclidean", 2)
    db.build()
    ids, dist, lens, _ = db.search_by_vector([[0.0, 1.0]], 10)
    assert lens[0] == 0
    _, _, lens, _ = db.search_by_item([0], 10)
    assert lens[0] == 0xFFFFFFFF


# ---- src/tests/writer.rs:282-295,358-371 — self query returns self at ~0 --------------------------------------------
@pytest.mark.parametrize("metric", ["cosine", "binary quantized cosine"])
def test_self_query_returns_self(metric):
    rng = np.random.default_rng(2)
    x = rng.uniform(-1, 1, (100, 1025)).astype(np.float32)
    db = OracleDb(metric, 1025)
    db.add_items(np.arange(100), x)
    db.build(M=16, M0=32)
    ids, dist, lens, _ = db.search_by_vector(x, 1)
    assert np.all(lens == 1)
    if metric == "cosine":
        assert np.array_equal(ids[:, 0], np.arange(100))
    assert np.all(np.abs(dist[:, 0]) < 1e-6)


# ---- src/tests/reader.rs:82-111 — all items are reachable (M = M0 = 6, ef = n) ------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 7, 100, 1500])
def test_all_items_are_reachable(n):
    rng = np.random.default_rng(n)
    x = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
    db = OracleDb("cosine", 8)
    db.add_items(np.arange(n), x)
    db.build(M=6, M0=6, n_threads=1)
    ids, dist, lens, _ = db.search_by_vector(np.zeros((1, 8), np.float32), n, ef=n)
    assert lens[0] == n and sorted(ids[0].tolist()) == list(range(n))


# ---- src/tests/reader.rs:41-78 — candidates honoured (linear scan path) -------------------------------------------------
def test_search_on_candidates_has_right_num():
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, (1000, 768)).astype(np.float32)
    db = OracleDb("cosine", 768)
    db.add_items(np.arange(1000), x)
    db.build(n_threads=4)
    for c in range(3):
        cand = rng.choice(1000, 10, replace=False).astype(np.uint32)
        ids, dist, lens, ctr = db.search_by_vector(rng.uniform(-1, 1, (1, 768)), 10, candidates=cand, counters=True)
        assert lens[0] == 10 and set(ids[0].tolist()) == set(cand.tolist())
        assert ctr[0, 6] & O.FLAG_LINEAR
    # graph path with a filter (linear_below = 0 disables the scan)
    cand = rng.choice(1000, 300, replace=False).astype(np.uint32)
    ids, dist, lens, ctr = db.search_by_vector(rng.uniform(-1, 1, (4, 768)), 10, candidates=cand, linear_below=0, counters=True)
    assert np.all(lens == 10) and not (ctr[:, 6] & O.FLAG_LINEAR).any()
    assert set(ids.ravel().tolist()) <= set(cand.tolist())


# ---- src/tests/reader.rs:113-143 — by_item excludes the item; None when absent ------------------------------------------
def test_by_item_excludes_item_and_none_if_absent():
    rng = np.random.default_rng(6)
    x = rng.uniform(-1, 1, (200, 16)).astype(np.float32)
    db = OracleDb("euclidean", 16)
    db.add_items(np.arange(200), x)
    db.build()
    ids, dist, lens, _ = db.search_by_item(np.arange(200), 10)
    for i in range(200):
        assert lens[i] == 10 and i not in ids[i].tolist()
    _, _, lens, _ = db.search_by_item([12345], 10)
    assert lens[0] == 0xFFFFFFFF


# ---- golden graph topology: src/tests/writer.rs:388-412 (6 points on a line, M = M0 = 3) ---------------------------------
def line_graph_db():
    db = OracleDb("euclidean", 2)
    db.add_items(np.arange(6), [[i, 0.0] for i in range(6)])
    l0 = {0: [1, 2], 1: [0, 2], 2: [0, 1, 3], 3: [2, 4], 4: [3, 5], 5: [4]}
    l1 = {0: [2], 2: [0, 3], 3: [2]}
    for i, nb in l0.items():
        db.set_links(i, 0, nb)
    for i, nb in l1.items():
        db.set_links(i, 1, nb)
    db.set_entry_points([0, 2, 3], 1)
    return db


def test_golden_line_graph_search_is_hand_checkable():
    db = line_graph_db()
    # query x=4.4: top layer {0,2,3} -> closest 3; layer 0 from 3 with ef=3 -> {4, 5, 3}
    ids, dist, lens, ctr = db.search_by_vector([[4.4, 0.0]], 3, ef=3, counters=True)
    assert ids[0].tolist() == [4, 5, 3]
    np.testing.assert_allclose(dist[0], [0.16, 0.36, 1.96], rtol=1e-6)
    assert ctr[0, 0] == 3  # three entry points evaluated on the top layer, nothing else to visit
    # exhaustive: ef = n returns every point in distance order
    ids, _, lens, _ = db.search_by_vector([[-1.0, 0.0]], 6, ef=6)
    assert ids[0].tolist() == [0, 1, 2, 3, 4, 5]
    # ties: x=2.5 is equidistant from 2 and 3 -> smaller id first (drain_asc on (bits, id))
    ids, dist, _, _ = db.search_by_vector([[2.5, 0.0]], 2, ef=6)
    assert ids[0].tolist() == [2, 3] and dist[0, 0] == dist[0, 1]


def test_multi_entry_point_top_layer_grows_past_ef():
    """reader.rs:315-325: entry points are pushed unconditionally, so with 3 entry points and ef=1 the top-layer
    result holds 3 entries; peek_min picks the closest."""
    db = line_graph_db()
    ids, _, _, _ = db.search_by_vector([[0.2, 0.0]], 1, ef=1)
    assert ids[0, 0] == 0


# ---- Roaring portable format + key layout -----------------------------------------------------------------------------------
def test_roaring_known_bytes_and_roundtrip():
    # RoaringFormatSpec: cookie 12346, container count, (key, card-1), offsets, sorted u16 values
    b = O.roaring_serialize([1, 2, 3, 65536 + 7])
    want = (12346).to_bytes(4, "little") + (2).to_bytes(4, "little") + bytes([0, 0, 2, 0, 1, 0, 0, 0]) + \
        (24).to_bytes(4, "little") + (30).to_bytes(4, "little") + bytes([1, 0, 2, 0, 3, 0, 7, 0])
    assert b == want
    rng = np.random.default_rng(3)
    for n in (0, 1, 100, 5000, 70000):
        ids = np.unique(rng.integers(0, 1 << 20, n).astype(np.uint32))
        if n == 70000:
            ids = np.unique(np.concatenate([ids, np.arange(5000, dtype=np.uint32) + (3 << 16)]))  # a bitmap container
        assert np.array_equal(O.roaring_deserialize(O.roaring_serialize(ids)), ids)


def test_roaring_run_container_is_decoded():
    # cookie 12347 | (n-1)<<16, run flag bitmap, (key, card-1), one run [10, 10+5]
    b = ((12347) | (0 << 16)).to_bytes(4, "little") + bytes([1]) + bytes([0, 0, 5, 0]) + (1).to_bytes(2, "little") + \
        (10).to_bytes(2, "little") + (5).to_bytes(2, "little")
    assert O.roaring_deserialize(b).tolist() == [10, 11, 12, 13, 14, 15]


def test_kv_export_layout():
    db = line_graph_db()
    kv = db.export_kv(index=7)
    keys = [k for k, _ in kv]
    assert all(len(k) == 8 for k in keys)                       # src/key.rs:129-162
    assert keys == sorted(keys)                                 # LMDB order
    assert keys[0] == bytes([0, 7, 0, 0, 0, 0, 0, 0])           # metadata: mode 0, item 0
    assert keys[1] == bytes([0, 7, 0, 0, 0, 0, 1, 0])           # version: mode 0, item 1
    links = [k for k in keys if k[2] == 2]
    items = [k for k in keys if k[2] == 3]
    assert len(links) == 9 and len(items) == 6
    assert max(links) < min(items)                              # NodeId ordering: links(u32::MAX, 0) < item(0)
    meta = kv[0][1]
    assert meta.startswith(b"euclidean\x00") and meta[10:14] == (2).to_bytes(4, "big") and meta[-1] == 1
    assert kv[1][1] == (0).to_bytes(4, "big") + (1).to_bytes(4, "big") + (3).to_bytes(4, "big")
    item0 = dict(kv)[bytes([0, 7, 3, 0, 0, 0, 0, 0])]
    assert item0[0] == 0 and len(item0) == 1 + 4 + 8            # tag, header f32, two f32
    db_h = OracleDb("hamming", 70)
    db_h.add_items([5], [np.ones(70)])
    db_h.build()
    item = [v for k, v in db_h.export_kv() if k[2] == 3][0]
    assert len(item) == 1 + 8 + 16                              # NodeHeaderHamming{idx: usize} is 8 bytes
