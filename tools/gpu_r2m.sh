#!/bin/bash
# run M: merge fast path for tiles without new keys, binary kernel at 4 CTAs/SM when shared memory limits it, row groups of 8 for C2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2m
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_query_options.py -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -3 $O/gpu_tests.log
run() { # variant workload ef extra...
  local v=$1 w=$2 ef=$3; shift 3
  HB_LIB_VARIANT=$v timeout 600 python tools/dev_sweep.py --workload $w --ef $ef --steps 10 --device-build "$@" > $O/${w}_${v:-prod}.log 2>&1
  echo "== $w ${v:-prod}"; grep -h '^{' $O/${w}_${v:-prod}.log | cut -c1-150
}
run "" c4 800 --nq 20000 --parity 64 --sweep "bin_wide=1,0"
run "" c4s 200
run "" c3 128 --nq-list 2500,1250,1
run "" c2 128
run rg8 c2 128
