#!/usr/bin/env python
"""Exact k-NN ground truth: tensor-core shortlist (exact_tc.cu) vs the full CUDA-core scan on a bench workload's items
(no graph needed).  Wall time of hb_exact_knn (query upload, staging, GEMM, re-rank, download), equality of the two
answers, useful TFLOP/s of the GEMM (2 * nq * n * dims / time).  Dev tool (GPU box).

  python tools/exact_bench.py [--workload c3] [--nq 2000] [--out file.json]
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--nq", type=int, default=2000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    import hannoy_b200 as hb
    from hannoy_b200 import _lib as L
    from oracle.oracle import OracleDb
    dev = torch.device("cuda", 0)
    w = dict(bench.WORKLOADS[args.workload])
    x = bench.gen_vectors(w["gen"], w["n"], w["dims"], w["seed"], dev).cpu().numpy()
    q = bench.gen_vectors(w["gen"], args.nq, w["dims"], w["seed"] + 1, dev).cpu().numpy()
    ids = np.arange(w["n"], dtype=np.uint32)
    db = OracleDb(w["metric"], w["dims"])
    db.add_items(ids, x)
    off = np.zeros(w["n"] + 1, np.uint64)
    rd = hb.Reader.from_arrays(w["metric"], w["dims"], ids, x, db.headers(), [(off, np.zeros(0, np.uint32))], ids[:1], 0)
    res = {"workload": w["desc"], "nq": args.nq, "k": args.k}
    out = {}
    for tc in (1, 0):
        L.lib().hb_tune(b"exact_tc", tc)
        hb.exact_knn(rd, q[:256], args.k)   # warm-up (module load, allocations)
        ts = []
        for _ in range(3):
            t = time.perf_counter()
            out[tc] = hb.exact_knn(rd, q, args.k)
            ts.append(time.perf_counter() - t)
        best = min(ts)
        key = "tensor_core_shortlist" if tc else "cuda_core_scan"
        res[key] = {"seconds": round(best, 4), "useful_tflops": round(2.0 * args.nq * w["n"] * w["dims"] / best / 1e12, 1)}
    L.lib().hb_tune(b"exact_tc", 1)
    res["identical_ids_and_distance_bits"] = bool(np.array_equal(out[1][0], out[0][0]) and np.array_equal(out[1][1].view(np.uint32), out[0][1].view(np.uint32)))
    peaks = bench.load_peaks()
    if peaks.get("bf16_tflops"):
        res["tf32_peak_estimate_tflops"] = round(peaks["bf16_tflops"] / 2, 1)   # dense tf32 = half the bf16 rate
    print(json.dumps(res))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
