#!/bin/bash
# run F: tests, speculative prefetch in the binary kernel A/B, the driver's bench command (N=1) with all extras, ncu capture of C3 for traffic.json
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2f
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -3 $O/gpu_tests.log
run() { # variant workload ef extra...
  local v=$1 w=$2 ef=$3; shift 3
  HB_LIB_VARIANT=$v timeout 300 python tools/dev_sweep.py --workload $w --ef $ef --steps 10 --device-build "$@" > $O/${w}_${v:-prod}.log 2>&1
  echo "== $w ${v:-prod}"; grep -h '^{' $O/${w}_${v:-prod}.log | cut -c1-200
}
for v in "" nospecbin; do run "$v" c4s 200; done
run "" c2 128
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err ) 2> $O/bench.time
tail -3 $O/bench.time; tail -c 600 $O/bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 6 -c 1 -f -o $O/c3_ncu python bench.py --steps 2 --warmup 3 --ef 128 --extras 0 > $O/c3_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --ef 128 --extras 0 > $O/launches.log 2>&1
ls -la $O
