#!/bin/bash
# run H: helpers shared chunk by chunk A/B on small batches
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2h
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -3 $O/gpu_tests.log
run() { # variant workload ef extra...
  local v=$1 w=$2 ef=$3; shift 3
  HB_LIB_VARIANT=$v timeout 300 python tools/dev_sweep.py --workload $w --ef $ef --steps 20 --device-build "$@" > $O/${w}_${v:-prod}.log 2>&1
  echo "== $w ${v:-prod}"; grep -h '^{' $O/${w}_${v:-prod}.log | cut -c1-120
}
run "" c3 128 --nq-list 5000,2500,1500,1250,1000,888,1
run noshare c3 128 --nq-list 5000,2500,1500,1250,1000,888,1
run "" c2 128 --nq-list 1250
run noshare c2 128 --nq-list 1250
