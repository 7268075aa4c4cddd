#!/bin/bash
# run G: large-ef parity test, prefetch dedupe A/B, small-batch scan with and without helpers
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2g
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -5 $O/gpu_tests.log
run() { # variant workload ef extra...
  local v=$1 w=$2 ef=$3; shift 3
  HB_LIB_VARIANT=$v timeout 300 python tools/dev_sweep.py --workload $w --ef $ef --steps 10 --device-build "$@" > $O/${w}_${v:-prod}.log 2>&1
  echo "== $w ${v:-prod}"; grep -h '^{' $O/${w}_${v:-prod}.log | cut -c1-200
}
for v in "" nodedupe; do run "$v" c2 128; run "$v" c4s 200; done
run "" c3 128 --nq-list 2500,1250,888,444,148,1 --sweep "team=1,0"
run nodedupe c3 128 --nq-list 1250,1
