#!/usr/bin/env python
"""Write the vectors and queries of a bench.py workload as raw row-major f32 files for baseline/rust/bench_qps.

  python tools/dump_workload.py c3 x.f32 q.f32      (generates with torch on cuda:0 if present, else on the CPU)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    name, xf, qf = sys.argv[1:4]
    w = bench.WORKLOADS[name]
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    bench.gen_vectors(w["gen"], w["n"], w["dims"], w["seed"], dev).cpu().numpy().tofile(xf)
    bench.gen_vectors(w["gen"], w["nq"], w["dims"], w["seed"] + 1, dev).cpu().numpy().tofile(qf)
    print(f"{name}: {w['n']} x {w['dims']} -> {xf}, {w['nq']} queries -> {qf} (generated on {dev})")


if __name__ == "__main__":
    main()
