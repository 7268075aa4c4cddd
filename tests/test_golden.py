"""Committed fixtures (tests/golden/):
  * search_golden.npz       — seeded databases + graphs + queries + expected answers for all seven metrics
                              (made by tests/golden/make_search_golden.py);
  * ref_snapshot_graphs.json — graph topologies built by the REFERENCE Writer, parsed out of its insta snapshots
                              (src/tests/snapshots/*.snap) by tests/golden/make_reference_fixtures.py.
CPU tests check the oracle against them; the -m gpu tests check the CUDA engine (through the C-ABI) against the
same committed answers, independently of the oracle code on the box."""
import json
import os

import numpy as np
import pytest

from oracle.oracle import OracleDb

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SETTINGS = [(10, 48), (3, 3)]
N_CASES = 9


def load_case(ci):
    z = np.load(os.path.join(GOLDEN, "search_golden.npz"))
    p = f"c{ci}_"
    g = {k[len(p):]: z[k] for k in z.files if k.startswith(p)}
    g["metric"] = str(g["metric"])
    g["dims"] = int(g["dims"])
    g["layers"] = [(g[f"off{l}"].astype(np.uint64), g[f"nbr{l}"]) for l in range(int(g["n_layers"]))]
    return g


def oracle_db(g):
    db = OracleDb(g["metric"], g["dims"])
    db.add_items(g["ids"], g["x"])
    for l, (off, nbr) in enumerate(g["layers"]):
        for s, item in enumerate(g["ids"]):
            if off[s + 1] > off[s] or l == 0:
                db.set_links(int(item), l, nbr[int(off[s]):int(off[s + 1])])
    db.set_entry_points(g["eps"], int(g["max_level"]))
    return db


def check(got, g, key, what):
    ids, dist, lens = got[:3]
    assert np.array_equal(lens, g[key + "_len"]), what
    for i, n in enumerate(lens):
        n = 0 if n == 0xFFFFFFFF else int(n)
        assert np.array_equal(ids[i, :n], g[key + "_ids"][i, :n]), f"{what}: ids of query {i}"
        assert np.array_equal(dist[i, :n].view(np.uint32), g[key + "_dbits"][i, :n]), f"{what}: distance bits of query {i}"


@pytest.mark.parametrize("ci", range(N_CASES))
def test_oracle_reproduces_committed_answers(ci):
    g = load_case(ci)
    db = oracle_db(g)
    for si, (count, ef) in enumerate(SETTINGS):
        got = db.search_by_vector(g["q"], count, ef=max(ef, count), counters=True)
        check(got, g, f"v{si}", f"{g['metric']} by_vector k={count}")
        assert np.array_equal(got[3][:, :6].astype(np.uint32), g[f"v{si}_ctr"])
        check(db.search_by_item(g["items"], count, ef=max(ef, count)), g, f"i{si}", f"{g['metric']} by_item k={count}")


@pytest.mark.gpu
@pytest.mark.parametrize("ci", range(N_CASES))
def test_cuda_reproduces_committed_answers(ci):
    import hannoy_b200 as hb
    g = load_case(ci)
    binary = g["metric"] not in ("euclidean", "cosine", "manhattan")
    db = oracle_db(g)  # only used to encode rows / headers the way the Writer stores them
    rd = hb.Reader.from_arrays(g["metric"], g["dims"], g["ids"], db.rows(), db.headers(), g["layers"], g["eps"],
                               int(g["max_level"]))
    assert (db.rows().dtype == np.uint64) == binary
    for si, (count, ef) in enumerate(SETTINGS):
        got = rd.nns(count).ef_search(ef).by_vectors_raw(g["q"], counters=True)
        check(got, g, f"v{si}", f"{g['metric']} by_vector k={count}")
        assert np.array_equal(got[3][:, :6].astype(np.uint32), g[f"v{si}_ctr"])
        check(rd.nns(count).ef_search(ef).by_items_raw(g["items"]), g, f"i{si}", f"{g['metric']} by_item k={count}")


# ---- reference-built topologies --------------------------------------------------------------------------------
def ref_graphs():
    return json.load(open(os.path.join(GOLDEN, "ref_snapshot_graphs.json")))["graphs"]


def db_on_ref_graph(g, metric, seed):
    md = g["metadata"]
    db = OracleDb(metric, md["dimensions"])
    rng = np.random.default_rng(seed)
    items = np.array(md["items"], np.uint32)
    x = rng.uniform(-1, 1, (len(items), md["dimensions"])).astype(np.float32)
    db.add_items(items, x)
    for item, layer, nb in g["links"]:
        db.set_links(item, layer, nb)
    db.set_entry_points(md["entry_points"], md["max_level"])
    return db, x


@pytest.mark.parametrize("gi", range(4))
def test_reference_built_graphs_satisfy_reader_invariants(gi):
    """Reader::assert_validity (reader.rs:904-948) on the reference's own graphs, through the oracle, plus the
    all_items_are_reachable property (src/tests/reader.rs:82-111): nns(n).ef_search(n) returns every item."""
    g = ref_graphs()[gi]
    md = g["metadata"]
    items = set(md["items"])
    seen_l0 = set()
    for item, layer, nb in g["links"]:
        assert item in items and all(t in items for t in nb)
        assert layer <= md["max_level"]
        if layer == 0:
            seen_l0.add(item)
    assert seen_l0 == items and all(e in items for e in md["entry_points"])
    db, x = db_on_ref_graph(g, md["distance"], 5)
    n = len(items)
    ids, _, lens, _ = db.search_by_vector(np.zeros((1, md["dimensions"]), np.float32), n, ef=n)
    assert lens[0] == n and set(ids[0].tolist()) == items


@pytest.mark.gpu
@pytest.mark.parametrize("gi", range(4))
@pytest.mark.parametrize("metric", ["euclidean", "cosine", "hamming"])
def test_cuda_on_reference_built_graphs(gi, metric):
    """The CUDA engine on topologies produced by the reference Writer (7 levels, M = M0 = 3), both ingestion
    routes, against the oracle; and the reachability property on the device."""
    import hannoy_b200 as hb
    g = ref_graphs()[gi]
    md = g["metadata"]
    db, x = db_on_ref_graph(g, metric, 11 + gi)
    n = len(md["items"])
    rk = hb.Reader.open(db.export_kv(2), 2, metric)
    ra = hb.Reader.from_arrays(metric, md["dimensions"], db.ids(), db.rows(), db.headers(), db.layers(), db.entry_points,
                               db.max_level)
    q = np.random.default_rng(3).uniform(-1, 1, (16, md["dimensions"])).astype(np.float32)
    for count, ef in [(5, 5), (10, 30), (n, n)]:
        want = db.search_by_vector(q, count, ef=max(ef, count))
        for rd in (rk, ra):
            got = rd.nns(count).ef_search(ef).by_vectors_raw(q)
            assert np.array_equal(got[2], want[2])
            for i, m in enumerate(got[2]):
                assert np.array_equal(got[0][i, :m], want[0][i, :m])
                assert np.array_equal(got[1][i, :m].view(np.uint32), want[1][i, :m].view(np.uint32))
    ids, _, lens = rk.nns(n).ef_search(n).by_vectors_raw(np.zeros((1, md["dimensions"]), np.float32))
    assert lens[0] == n and set(ids[0].tolist()) == set(md["items"])


# ---- the rest of the QueryBuilder surface: candidates, linear scan, by_item + candidates, cancellation -------------------------
@pytest.mark.parametrize("ci", range(N_CASES))
def test_oracle_reproduces_committed_option_answers(ci):
    """options_golden.npz (tests/golden/make_options_golden.py): the oracle's filtered / linear / cancelled answers are pinned
    like the plain ones, so a change to the oracle's visit() or cancel handling cannot go unnoticed."""
    import sys
    sys.path.insert(0, GOLDEN)
    from make_options_golden import answers
    g = load_case(ci)
    z = np.load(os.path.join(GOLDEN, "options_golden.npz"))
    got = answers(oracle_db(g), g, ci)
    keys = [k for k in z.files if k.startswith(f"c{ci}_")]
    assert len(keys) == len(got) and len(keys) > 20
    for k in keys:
        assert np.array_equal(z[k], got[k[len(f"c{ci}_"):]]), k
    # the fixture exercises what it claims to: a linear scan, a filtered walk, cancelled and completed queries
    assert (z[f"c{ci}_linear_ctr"][:, 6] & 2).all() and not (z[f"c{ci}_walk_ctr"][:, 6] & 2).any()
    assert (z[f"c{ci}_cancel1_ctr"][:, 6] & 8).all() and (z[f"c{ci}_cancel1_len"] <= 10).all()
    assert set(z[f"c{ci}_walk_ids"][z[f"c{ci}_walk_len"] > 0, 0].tolist()) <= set(z[f"c{ci}_cand_big"].tolist())


@pytest.mark.gpu
@pytest.mark.parametrize("ci", range(N_CASES))
def test_cuda_reproduces_committed_option_answers(ci):
    """The CUDA engine against the same committed answers (ids, distance bits, lengths, traversal counters, LINEAR /
    CANCELLED flags), independently of the oracle code on the box."""
    import hannoy_b200 as hb
    g = load_case(ci)
    z = np.load(os.path.join(GOLDEN, "options_golden.npz"))
    p = f"c{ci}_"
    db = oracle_db(g)  # only used to encode rows / headers the way the Writer stores them
    rd = hb.Reader.from_arrays(g["metric"], g["dims"], g["ids"], db.rows(), db.headers(), g["layers"], g["eps"], int(g["max_level"]))

    def same(got, key):
        ids, dist, lens, ctr = got
        canc = (lens != 0xFFFFFFFF) & ((lens >> 31) == 1)
        clean = np.where(lens == 0xFFFFFFFF, lens, lens & 0x7FFFFFFF).astype(np.uint32)
        assert np.array_equal(clean, z[p + key + "_len"]), key
        for i, n in enumerate(clean):
            n = 0 if n == 0xFFFFFFFF else int(n)
            assert np.array_equal(ids[i, :n], z[p + key + "_ids"][i, :n]), (key, i)
            assert np.array_equal(dist[i, :n].view(np.uint32), z[p + key + "_dbits"][i, :n]), (key, i)
        want = z[p + key + "_ctr"]
        assert np.array_equal(ctr[:, :6].astype(np.uint32), want[:, :6]), key
        assert np.array_equal((ctr[:, 6] & 2) != 0, (want[:, 6] & 2) != 0), key      # LINEAR
        assert np.array_equal(canc, (want[:, 6] & 8) != 0), key                       # CANCELLED == Searched::did_cancel

    same(rd.nns(10).ef_search(48).candidates(z[p + "cand_big"]).linear_below(0).by_vectors_raw(g["q"], counters=True), "walk")
    same(rd.nns(10).ef_search(48).candidates(z[p + "cand_small"]).by_vectors_raw(g["q"], counters=True), "linear")
    same(rd.nns(5).ef_search(32).candidates(z[p + "cand_big"]).linear_below(0).by_items_raw(g["items"], counters=True), "item")
    for a in (1, 3, 20):
        same(rd.nns(10).ef_search(48).with_cancellation(a).by_vectors_raw(g["q"], counters=True), f"cancel{a}")
