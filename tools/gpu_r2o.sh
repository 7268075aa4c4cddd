#!/bin/bash
# run O: upper-layer rows kept in L2 (evict_last) A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2o
mkdir -p $O
run() { # variant workload ef extra...
  local v=$1 w=$2 ef=$3; shift 3
  HB_LIB_VARIANT=$v timeout 600 python tools/dev_sweep.py --workload $w --ef $ef --steps 20 --device-build "$@" > $O/${w}_${v:-prod}.log 2>&1
  echo "== $w ${v:-prod}"; grep -h '^{' $O/${w}_${v:-prod}.log | cut -c1-150
}
run "" c2 128
run nokeep c2 128
run "" c3 128 --nq-list 1250,1
run nokeep c3 128 --nq-list 1250,1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
