#!/usr/bin/env python
"""Dev tool (GPU box): time the search kernel alone on one workload under several tunable settings, check
parity against the oracle once, and (with HB_LIB_VARIANT=phases) print the per-phase cycle breakdown.

  python tools/dev_sweep.py --workload c3 --ef 128 --sweep "ring_bytes=12288,24576;blocks_per_sm=2,3"
"""
import argparse
import itertools
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--n-items", type=int, default=0)
    ap.add_argument("--ef", type=int, default=128)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--sweep", default="")
    ap.add_argument("--parity", type=int, default=256)
    ap.add_argument("--out", default="")
    ap.add_argument("--trace", default="", help="HB_TRACE builds: save the event trace of one warp over one step (.npy)")
    ap.add_argument("--nq", type=int, default=0, help="override the batch size")
    ap.add_argument("--nq-list", default="", help="also time these batch sizes (prefixes of the batch), comma separated")
    ap.add_argument("--device-build", action="store_true", help="build the graph on the device (seconds) instead of with the oracle builder")
    args = ap.parse_args()
    import torch
    import hannoy_b200 as hb
    from hannoy_b200 import _lib
    L = _lib.lib()
    w = dict(bench.WORKLOADS[args.workload])
    if args.n_items:
        w["n"] = args.n_items
    if args.nq:
        w["nq"] = args.nq
    dev = torch.device("cuda", 0)
    threads = len(os.sched_getaffinity(0))
    log = lambda m: print(f"[sweep] {m}", file=sys.stderr, flush=True)
    binary = "binary" in w["metric"] or w["metric"] == "hamming"
    q = bench.gen_vectors(w["gen"], w["nq"], w["dims"], w["seed"] + 1, dev)
    q_host = q.cpu().numpy()
    if binary:
        # the same items bench.run_other_workload builds: quantized on the GPU chunk by chunk, graph built on the device
        ids = np.arange(w["n"], dtype=np.uint32)
        rows = np.empty((w["n"], w["dims"] // 64), np.uint64)
        chunk = 250_000
        for s0 in range(0, w["n"], chunk):
            m = min(chunk, w["n"] - s0)
            rows[s0:s0 + m] = bench.quantize_bq(bench.gen_vectors(w["gen"], m, w["dims"], w["seed"] * 1000 + s0 // chunk, dev))
        hdr = np.full(w["n"], np.float32(np.sqrt(np.float32(w["dims"]))), np.float32)
        rd = hb.Reader.build(w["metric"], w["dims"], ids, rows, hdr, M=16, M0=32, ef_construction=100, seed=42)
        db = bench.oracle_from_reader(rd, w["metric"], w["dims"], rows, ids)
    else:
        x = bench.gen_vectors(w["gen"], w["n"], w["dims"], w["seed"], dev)
        x_host = x.cpu().numpy()
        del x
        if args.device_build:
            from oracle.oracle import OracleDb
            ids = np.arange(w["n"], dtype=np.uint32)
            db = OracleDb(w["metric"], w["dims"])
            db.add_items(ids, x_host)
            rd = hb.Reader.build(w["metric"], w["dims"], ids, x_host, db.headers() if w["metric"] == "cosine" else None, M=16, M0=32, ef_construction=100, seed=42)
            from oracle import oracle as O
            for l, (off, nbr) in enumerate(rd.layers()):
                O.lib().orc_db_set_csr(db.h, l, O._p(off), O._p(nbr), len(nbr))
            db.set_entry_points(rd.entry_points(), rd.max_level())
        else:
            db = bench.build_or_load_graph(w, x_host, dev.type, threads, log)
            rd = hb.Reader.from_arrays(w["metric"], w["dims"], db.ids(), db.rows(), db.headers(), db.layers(), db.entry_points,
                                       db.max_level, device=0)
    k, nq = w["k"], w["nq"]
    ef_raw = max(args.ef, k)
    dq = q.contiguous()
    d_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    d_dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
    d_len = torch.empty((nq,), dtype=torch.int32, device=dev)
    d_ctr = torch.zeros((nq, 8), dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()

    def step(ctr=False):
        rd.search_device(dq.data_ptr(), nq, k, ef_raw, d_ids.data_ptr(), d_dist.data_ptr(), d_len.data_ptr(),
                         d_ctr.data_ptr() if ctr else None, stream.cuda_stream)

    keys, vals = [], []
    for part in filter(None, args.sweep.split(";")):
        kname, vs = part.split("=")
        keys.append(kname)
        vals.append([int(v) for v in vs.split(",")])
    combos = list(itertools.product(*vals)) if keys else [()]
    want = None
    results = []
    nq_full = nq
    for nq, combo in [(n_, c_) for n_ in [nq_full] + [int(v) for v in args.nq_list.split(",") if v] for c_ in combos]:
        for kname, v in zip(keys, combo):
            L.hb_tune(kname.encode(), v)
        step(ctr=True)
        torch.cuda.synchronize()
        ctr = d_ctr.cpu().numpy().astype(np.uint64)[:nq]
        alg, vec = bench.algorithmic_bytes(ctr, w)
        n_par = min(args.parity, nq)
        if want is None or len(want[2]) < n_par:
            want = db.search_by_vector(q_host[:n_par], k, ef=ef_raw, n_threads=threads, counters=True)
        ids = d_ids[:n_par].cpu().numpy().view(np.uint32)
        dd = d_dist[:n_par].cpu().numpy()
        ln = d_len[:n_par].cpu().numpy().view(np.uint32)
        ok = bool(np.array_equal(ln, want[2][:n_par]) and np.array_equal(ids, want[0][:n_par]) and np.array_equal(dd.view(np.uint32), want[1][:n_par].view(np.uint32)))
        ok_ctr = bool(np.array_equal(ctr[:n_par, :6], want[3][:n_par, :6].astype(np.uint64))) if len(want) > 3 else None
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ph0 = np.zeros(16, np.uint64)
        L.hb_debug_phases(ph0.ctypes.data)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        ph = np.zeros(16, np.uint64)
        L.hb_debug_phases(ph.ctypes.data)
        r = dict(nq=nq, tune=dict(zip(keys, combo)), ms=round(ms, 4), qps=round(nq / ms * 1e3), gbs=round(alg / ms / 1e6, 1), parity=ok, counters=ok_ctr,
                 slow=int((ctr[:, 6] & 4).astype(bool).sum()), evals_per_q=float(ctr[:, :2].sum() / nq), exp_per_q=float(ctr[:, 2:4].sum() / nq))
        if ph.sum():
            tot = float(ph[7]) or 1.0
            names = ["stage", "upper", "adj", "vis", "rows", "heap", "tail", "total", "post", "collect", "accept", "decide"]
            r["phase_frac"] = {n: round(float(p) / tot, 3) for n, p in zip(names, ph)}
            r["cycles_per_query"] = round(tot / (nq * args.steps))
        if args.trace:
            buf = np.zeros(1 << 18, np.uint64)
            L.hb_debug_trace(buf.ctypes.data, len(buf))   # reset
            step()
            torch.cuda.synchronize()
            n = L.hb_debug_trace(buf.ctypes.data, len(buf))
            np.save(args.trace, buf[:n])
            r["trace_events"] = int(n)
            # mean cycles between consecutive events, by (previous event -> event) pair
            names = {1: "qstart", 2: "pop", 3: "adj", 4: "vis", 5: "posted", 6: "rowwait", 7: "group", 8: "heap", 9: "qend", 10: "l0"}
            ev = (buf[:n] & 0xff).astype(int)
            ts = (buf[:n] >> 8).astype(np.int64)
            pairs = {}
            for i in range(1, n):
                key = f"{names.get(ev[i - 1], ev[i - 1])}->{names.get(ev[i], ev[i])}"
                pairs.setdefault(key, []).append(int(ts[i] - ts[i - 1]))
            r["trace_pairs"] = {k: {"n": len(v), "mean": round(float(np.mean(v))), "p50": int(np.median(v))} for k, v in sorted(pairs.items(), key=lambda kv: -sum(kv[1]))[:16]}
        results.append(r)
        print(json.dumps(r), flush=True)
    if args.out:
        json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
