#!/usr/bin/env python
"""BASELINE config 4 at FULL size on one B200 (GPU box): 10M x 1024 BinaryQuantizedCosine codes, graph built on the device
(hb_index_build_graph — the CPU builder needs tens of minutes for this), 100k-query batches, top-100.  Reports build time,
recall@100 against the exact k-NN kernel in the quantized metric, device-resident and end-to-end QPS, and size-independent
properties of the results (sorted, unique, in range, self queries found), and ORACLE PARITY at full size: the device-built
graph is handed to the CPU oracle as CSR (hb_index_layer_csr) and 256 queries are compared bit for bit (ids, distance bits,
traversal counters) at the ef the recall rule picked.

  python tools/c4_full.py [--n-items 10000000] [--nq 100000] [--out file.json]
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


quantize_bq = bench.quantize_bq


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-items", type=int, default=10_000_000)
    ap.add_argument("--nq", type=int, default=100_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    import hannoy_b200 as hb
    dev = torch.device("cuda", 0)
    n, dims, k, nq = args.n_items, 1024, 100, args.nq
    log = lambda m: print(f"[c4-full] {m}", file=sys.stderr, flush=True)
    t = time.time()
    codes = np.empty((n, dims // 64), np.uint64)
    chunk = 500_000
    for s in range(0, n, chunk):   # same generator family as bench.py's c4s ("lowrank"), one seed per chunk
        m = min(chunk, n - s)
        x = bench.gen_vectors("lowrank", m, dims, 7000 + s // chunk, dev)
        codes[s:s + m] = quantize_bq(x)
        del x
    q = bench.gen_vectors("lowrank", nq, dims, 8, dev)
    q_host = q.cpu().numpy()
    hdr = np.full(n, np.float32(np.sqrt(np.float32(dims))), np.float32)   # new_header: sqrt(bq_dot(v, v)) = sqrt(1024)
    log(f"{n} items generated and quantized in {time.time() - t:.1f}s")
    st = {}
    t = time.time()
    rd = hb.Reader.build("binary quantized cosine", dims, np.arange(n, dtype=np.uint32), codes, hdr, M=16, M0=32, ef_construction=100,
                         seed=42, stats=st)
    t_build = time.time() - t
    log(f"device build + upload in {t_build:.1f}s {st}")
    n_gt = 1000
    t = time.time()
    gt, _ = hb.exact_knn(rd, q_host[:n_gt], k)
    log(f"exact k-NN (quantized metric) for {n_gt} queries in {time.time() - t:.1f}s")
    sweep, ef_pick = {}, None
    for ef in (100, 200, 400, 800):
        ids, dd, lens = rd.nns(k).ef_search(ef).by_vectors_raw(q_host[:n_gt])
        sweep[ef] = round(bench.recall_at_k(ids, lens, gt, k), 4)
        if ef_pick is None and sweep[ef] >= 0.95:
            ef_pick = ef
            break
    if ef_pick is None:
        ef_pick = max(sweep)
    log(f"recall@{k} sweep {sweep} -> ef_search={ef_pick}")
    # oracle parity at full size: the CPU reader on the very same graph
    t = time.time()
    ids_all = np.arange(n, dtype=np.uint32)
    db = bench.oracle_from_reader(rd, "binary quantized cosine", dims, codes, ids_all)
    n_par = 256
    want = db.search_by_vector(q_host[:n_par], k, ef=ef_pick, n_threads=len(os.sched_getaffinity(0)), counters=True)
    got = rd.nns(k).ef_search(ef_pick).by_vectors_raw(q_host[:n_par], counters=True)
    parity_ok = bool(np.array_equal(got[2], want[2]) and np.array_equal(got[0], want[0]) and np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32))
                     and np.array_equal(got[3][:, :6], want[3][:, :6]))
    log(f"oracle parity on {n_par} queries at ef={ef_pick}: {'bit-exact' if parity_ok else 'MISMATCH'} ({time.time() - t:.1f}s)")
    del db
    # properties on the full batch
    ids, dd, lens = rd.nns(k).ef_search(ef_pick).by_vectors_raw(q_host)
    props = dict(all_full=bool((lens == k).all()), sorted=bool((np.diff(dd.view(np.uint32).astype(np.int64), axis=1) >= 0).all()),
                 in_range=bool((ids < n).all()), unique=bool(all(len(set(r.tolist())) == k for r in ids[:2000])))
    # self queries: the item's own code is at distance 0 (or tied with duplicates of the same code)
    self_ids = np.arange(0, n, n // 1000, dtype=np.uint32)[:1000]
    si, sd, sl = rd.nns(1).ef_search(ef_pick).by_items_raw(self_ids)   # by_item excludes the item itself: nearest OTHER item
    props["by_item_excludes_self"] = bool((si[:, 0] != self_ids).all())
    log(f"properties {props}")
    # device-resident timing
    dq = q.contiguous()
    d_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    d_dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
    d_len = torch.empty((nq,), dtype=torch.int32, device=dev)
    d_ctr = torch.zeros((nq, 8), dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()

    def step(ctr=False):
        rd.search_device(dq.data_ptr(), nq, k, ef_pick, d_ids.data_ptr(), d_dist.data_ptr(), d_len.data_ptr(), d_ctr.data_ptr() if ctr else None,
                         stream.cuda_stream)
    step(ctr=True)
    torch.cuda.synchronize()
    ctr = d_ctr.cpu().numpy().astype(np.uint64)
    w = dict(metric="binary quantized cosine", dims=dims)
    alg, vec = bench.algorithmic_bytes(ctr, w)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    t = time.time()
    for _ in range(2):
        rd.nns(k).ef_search(ef_pick).by_vectors_raw(q_host)
    e2e_ms = (time.time() - t) / 2 * 1e3
    res = dict(workload=f"{n} x {dims} BinaryQuantizedCosine codes, {nq}-query batch, top-{k} (BASELINE config 4, full size)", n_items=n, dims=dims,
               batch_queries=nq, k=k, M=16, M0=32, ef_construction=100, builder="device (hb_index_build_graph)", device_build_s=round(t_build, 1),
               device_build_stats=st, ef_search=ef_pick, recall_at_k=sweep[ef_pick], recall_sweep=sweep, ms_per_step=round(ms, 3),
               qps_device_resident=round(nq / ms * 1e3, 1), qps_e2e_host_buffers=round(nq / e2e_ms * 1e3, 1), algorithmic_gbs=round(alg / ms / 1e6, 1),
               dist_evals_per_query=float(ctr[:, :2].sum() / nq), properties=props,
               parity_vs_oracle=f"bit-exact (ids, distance bits, traversal counters; {n_par} queries at ef={ef_pick}, oracle on the exported device-built graph)" if parity_ok else "MISMATCH")
    print(json.dumps(res))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
