#!/usr/bin/env python
"""Dev tool (GPU box): device graph build vs the restated CPU builder on a bench workload: build time and recall@k at
the workload's ef values on the same queries.

  python tools/build_bench.py --workload c3 [--n-items N] [--batch-max B]
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--n-items", type=int, default=0)
    ap.add_argument("--batch-max", type=int, default=0)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    import hannoy_b200 as hb
    from oracle.oracle import OracleDb
    w = dict(bench.WORKLOADS[args.workload])
    if args.n_items:
        w["n"] = args.n_items
    dev = torch.device("cuda", 0)
    threads = len(os.sched_getaffinity(0))
    log = lambda m: print(f"[build-bench] {m}", file=sys.stderr, flush=True)
    x = bench.gen_vectors(w["gen"], w["n"], w["dims"], w["seed"], dev)
    q = bench.gen_vectors(w["gen"], 2000, w["dims"], w["seed"] + 1, dev)
    x_host, q_host = x.cpu().numpy(), q.cpu().numpy()
    del x
    db = OracleDb(w["metric"], w["dims"])
    ids = np.arange(w["n"], dtype=np.uint32)
    db.add_items(ids, x_host)
    rows, hdr = db.rows(), db.headers()
    st = {}
    t = time.time()
    rd = hb.Reader.build(w["metric"], w["dims"], ids, rows, hdr, M=bench.M, M0=bench.M0, ef_construction=bench.EFC, alpha=bench.ALPHA,
                         seed=42, batch_max=args.batch_max, stats=st)
    t_gpu = time.time() - t
    log(f"device build + upload: {t_gpu:.1f}s {st}")
    k = w["k"]
    gt, _ = hb.exact_knn(rd, q_host, k)
    res = dict(workload=w["desc"], n_items=w["n"], dims=w["dims"], metric=w["metric"], M=bench.M, M0=bench.M0, ef_construction=bench.EFC,
               device_build_s=round(t_gpu, 2), device_build_stats=st, recall_device_build={}, host_threads=threads)
    for ef in w["efs"]:
        got = rd.nns(k).ef_search(ef).by_vectors_raw(q_host)
        res["recall_device_build"][ef] = round(bench.recall_at_k(got[0], got[2], gt, k), 4)
    log(f"recall on the device-built graph: {res['recall_device_build']}")
    if not args.skip_cpu:
        t = time.time()
        db.build(M=bench.M, M0=bench.M0, ef_construction=bench.EFC, alpha=bench.ALPHA, seed=42, n_threads=threads)
        res["cpu_build_s"] = round(time.time() - t, 2)
        log(f"restated CPU builder on {threads} threads: {res['cpu_build_s']}s")
        res["recall_cpu_build"] = {}
        for ef in w["efs"]:
            c = db.search_by_vector(q_host, k, ef=ef, n_threads=threads)
            res["recall_cpu_build"][ef] = round(bench.recall_at_k(c[0], c[2], gt, k), 4)
        log(f"recall on the CPU-built graph: {res['recall_cpu_build']}")
        res["speedup"] = round(res["cpu_build_s"] / res["device_build_s"], 2)
    print(json.dumps(res))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
