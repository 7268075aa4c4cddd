#!/usr/bin/env python
"""One of bench.py's secondary workloads (c1 / c2 / c4s) on its own: `python tools/other_workload.py c4s [steps]` prints the
sub-record bench.py embeds (device-built graph, ef by the recall rule, QPS, roofline, parity, CPU baseline).  Dev tool."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402

name = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
rec = bench.run_other_workload(name, dev, 0, len(os.sched_getaffinity(0)), lambda m: print(f"[other] {m}", file=sys.stderr), steps=steps)
print(json.dumps(rec))
