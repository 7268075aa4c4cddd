#!/bin/bash
# run S: binary kernel without the speculative prefetch — timing, ncu capture (DRAM traffic), config 4 at full size, parity tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2s
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_query_options.py -m gpu -q 2>&1 | tail -2
timeout 300 python tools/dev_sweep.py --workload c4s --ef 200 --steps 20 --device-build > $O/c4s.log 2>&1; grep -h '^{' $O/c4s.log | cut -c1-150
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 4 -c 1 -f -o $O/c4s_ncu python tools/dev_sweep.py --workload c4s --ef 200 --steps 2 --device-build > $O/c4s_ncu.log 2>&1
timeout 900 python tools/c4_full.py --out $O/c4_full.json > $O/c4_full.log 2>&1; tail -1 $O/c4_full.log | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -c 200 $O/bench.json
