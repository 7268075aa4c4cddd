#!/bin/bash
# run W (reproducible builds): visited set read + conditional reduction in the reader (builder keeps the atomic); all GPU tests; C4s DRAM traffic
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2w
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; tail -2 $O/gpu_tests.log
run() { # variant workload ef extra...
  local v=$1 w=$2 ef=$3; shift 3
  HB_LIB_VARIANT=$v timeout 600 python tools/dev_sweep.py --workload $w --ef $ef --steps 20 --device-build "$@" > $O/${w}_${v:-prod}.log 2>&1
  echo "== $w ${v:-prod}"; grep -h '^{' $O/${w}_${v:-prod}.log | cut -c1-125
}
run "" c4s 200; run visatom c4s 200
run "" c2 128; run "" c3 128 --nq-list 1250,1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:hnsw_search_kernel -s 4 -c 1 --csv --log-file $O/c4s_traffic.csv python tools/dev_sweep.py --workload c4s --ef 200 --steps 2 --device-build > $O/c4s_traffic.log 2>&1
tail -4 $O/c4s_traffic.csv | cut -c1-300
