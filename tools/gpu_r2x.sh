#!/bin/bash
# run X: last check of the final library — all GPU tests, the driver's bench command, config 4 at full size
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2x
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; tail -2 $O/gpu_tests.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -c 250 $O/bench.json; echo
timeout 400 python tools/c4_full.py --out $O/c4_full.json > $O/c4_full.log 2>&1; tail -1 $O/c4_full.log | cut -c1-200
