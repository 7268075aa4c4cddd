#!/usr/bin/env python
"""A/B of the helper-warp row sharing (hb_tune "team") on device-built C3 / C2 graphs: device-resident ms per launch for
batch sizes from one query to the full batch, team on and off, same box, same graph.  Dev tool (GPU box).

  python tools/ab_team.py [--workload c3] [--out file.json]
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--ef", type=int, default=128)
    ap.add_argument("--out", default="")
    ap.add_argument("--tunes", default="team=1;team=0")
    ap.add_argument("--nqs", default="10000,5000,2500,1250,600,148,16,1")
    args = ap.parse_args()
    import torch
    import hannoy_b200 as hb
    from hannoy_b200 import _lib as L
    from oracle.oracle import OracleDb
    dev = torch.device("cuda", 0)
    w = dict(bench.WORKLOADS[args.workload])
    x = bench.gen_vectors(w["gen"], w["n"], w["dims"], w["seed"], dev).cpu().numpy()
    q = bench.gen_vectors(w["gen"], w["nq"], w["dims"], w["seed"] + 1, dev)
    ids = np.arange(w["n"], dtype=np.uint32)
    hdr = None
    if w["metric"] == "cosine":
        db = OracleDb(w["metric"], w["dims"]); db.add_items(ids, x); hdr = db.headers(); del db
    t = time.time()
    rd = hb.Reader.build(w["metric"], w["dims"], ids, x, hdr, M=16, M0=32, ef_construction=100, seed=42)
    print(f"[ab] built in {time.time() - t:.1f}s", file=sys.stderr)
    k = w["k"]
    res = {}
    for tune in args.tunes.split(";"):
        for kv in tune.split(","):
            key, val = kv.split("=")
            L.lib().hb_tune(key.encode(), int(val))
        row = {}
        for nq in [int(v) for v in args.nqs.split(",")]:
            steps = 20 if nq >= 1000 else 50
            ms, ctr, _ = bench.time_device_steps(rd, q[:nq].contiguous(), nq, k, args.ef, steps, 5, dev)
            row[nq] = round(ms, 4)
        res[tune] = row
        print(f"[ab] {tune}: {row}", file=sys.stderr)
    print(json.dumps({"workload": args.workload, "ef": args.ef, "ms_per_launch": res}))
    if args.out:
        json.dump({"workload": args.workload, "ef": args.ef, "ms_per_launch": res}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
